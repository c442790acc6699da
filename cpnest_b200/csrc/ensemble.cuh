// Ensemble statistics behind EnsembleEigenVector (cpnest/proposal.py:306-323): mean, np.cov
// (unbiased; D == 1 uses np.var, ddof = 0, proposal.py:314-318) and the symmetric
// eigen-decomposition (np.linalg.eigh) of the live ensemble, on the device.
// Covariance: fixed number of blocks, each reduces a contiguous slice of points into a D x D
// partial; partials are summed in a fixed order (deterministic, no atomics).
// Eigen: cyclic Jacobi with round-robin (parallel) ordering in one block.
#pragma once
#include <math.h>
#include <stdint.h>

namespace cpn {

// (the number of slices of the ensemble, i.e. of blocks of the partial kernels, is api.cu stat_blocks(nlive))
constexpr int kStatThreads = 256;

__global__ void __launch_bounds__(kStatThreads) k_mean_partial(const double* __restrict__ x, int n, int D, int Dp,
                                                               double* __restrict__ partial /*[blocks][D]*/) {
    // D <= kStatThreads.  thread t: dimension t % D, point lane t / D of this block's slice.
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(n, p0 + per);
    __shared__ double sh[kStatThreads];
    const int L = kStatThreads / D;
    const int d = (int)threadIdx.x % D, l = (int)threadIdx.x / D;
    double acc = 0.0;
    if (l < L)
        for (int i = p0 + l; i < p1; i += L) acc += x[(size_t)i * Dp + d];
    sh[threadIdx.x] = acc;
    __syncthreads();
    if ((int)threadIdx.x < D) {
        double s = 0.0;
        for (int k = 0; k < L; ++k) s += sh[k * D + threadIdx.x];
        partial[(size_t)blockIdx.x * D + threadIdx.x] = s;
    }
}

__global__ void k_mean_final(const double* __restrict__ partial, int nb, int n, int D, double* __restrict__ mean) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[(size_t)b * D + d];
    mean[d] = s / (double)n;
}

// Partial covariance sums of a slice of the ensemble on the FP64 tensor cores: C = sum_p xc_p xc_p^T with xc = x - mean,
// as 8x8 output tiles of mma.sync.m8n8k4.f64 (DMMA: IEEE fp64 multiply-add, four points per instruction).  Per tile
// (ti, tj) and group of four points p..p+3, lane l = 4 r + k supplies A[r][k] = xc[p+k][8 ti + r] and
// B[k][r] = xc[p+k][8 tj + r] and accumulates C[r][2k], C[r][2k+1] (the PTX fragment layout of m8n8k4).  The centred
// points of a stage live in shared memory with a leading dimension LD = 4 or 12 (mod 16), which makes the fragment loads
// conflict-free, zero-padded to whole tiles and whole groups of four.  A warp owns up to kMmaTilesPerWarp tiles per pass
// over the points; blocks, stages and the order of accumulation are fixed, so the sums are reproducible.
constexpr int kMmaStage = 32;          // points staged in shared memory per step (multiple of 4)
constexpr int kMmaTilesPerWarp = 8;
__host__ __device__ inline int cov_mma_ld(int D) {
    int ld = ((D + 7) / 8) * 8;
    while ((ld & 15) != 4 && (ld & 15) != 12) ++ld;
    return ld;
}
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(kStatThreads) k_cov_partial(const double* __restrict__ x, const double* __restrict__ mean,
                                                              int n, int D, int Dp, double* __restrict__ partial /*[blocks][D*D]*/) {
    extern __shared__ double tile[];   // [kMmaStage][LD] centred points
    const int LD = cov_mma_ld(D);
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(n, p0 + per);
    const int T8 = (D + 7) >> 3, ntile = T8 * T8;
    const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31, nwarp = kStatThreads >> 5;
    const int r = lane >> 2, k = lane & 3;
    for (int tbase = 0; tbase < ntile; tbase += nwarp * kMmaTilesPerWarp) {
        double c0[kMmaTilesPerWarp], c1[kMmaTilesPerWarp];
        int oa[kMmaTilesPerWarp], ob[kMmaTilesPerWarp];        // column offsets 8 ti, 8 tj of the warp's tiles
#pragma unroll
        for (int q = 0; q < kMmaTilesPerWarp; ++q) {
            const int t = tbase + q * nwarp + warp;
            c0[q] = c1[q] = 0.0;
            oa[q] = (t < ntile) ? 8 * (t / T8) : 0;
            ob[q] = (t < ntile) ? 8 * (t % T8) : 0;
        }
        for (int s0 = p0; s0 < p1; s0 += kMmaStage) {
            const int ns = min(kMmaStage, p1 - s0), ns4 = (ns + 3) & ~3;
            for (int i = threadIdx.x; i < ns4 * LD; i += kStatThreads) {
                const int pi = i / LD, d = i % LD;
                tile[i] = (pi < ns && d < D) ? x[(size_t)(s0 + pi) * Dp + d] - mean[d] : 0.0;
            }
            __syncthreads();
            for (int pp = 0; pp < ns4; pp += 4) {
                const double* row = tile + (pp + k) * LD + r;
#pragma unroll
                for (int q = 0; q < kMmaTilesPerWarp; ++q) dmma_m8n8k4(c0[q], c1[q], row[oa[q]], row[ob[q]]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < kMmaTilesPerWarp; ++q) {
            const int t = tbase + q * nwarp + warp;
            if (t >= ntile) continue;
            const int i = oa[q] + r, j = ob[q] + 2 * k;
            double* out = partial + (size_t)blockIdx.x * D * D + (size_t)i * D;
            if (i < D && j < D) out[j] = c0[q];
            if (i < D && j + 1 < D) out[j + 1] = c1[q];
        }
    }
}

__global__ void k_cov_final(const double* __restrict__ partial, int nb, int n, int D, double* __restrict__ cov) {
    const int ent = blockIdx.x * blockDim.x + threadIdx.x;
    if (ent >= D * D) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[(size_t)b * D * D + ent];
    // np.cov: 1/(n-1); the D == 1 branch of the reference uses np.var: 1/n (proposal.py:314-318)
    cov[ent] = s / (double)((D == 1) ? n : n - 1);
}

// Cyclic two-sided Jacobi with round-robin (parallel) ordering: n/2 disjoint rotations per round,
// n-1 rounds per sweep, one block of kJacobiThreads threads.
// A round applies all of its rotations at once: with J the product of the round's disjoint plane
// rotations, A' = J^T A J is computed element-wise from the OLD matrix (every entry needs the four
// entries at the crossing of its row pair and its column pair) into a second buffer, so a round costs
// two block barriers.  A, A', V, V' live in shared memory with an odd leading dimension when they fit
// (D <= 80), otherwise in global scratch.  max_sweeps bounds the sweeps: V is a product of plane rotations and stays orthonormal
// whatever their number, and the EnsembleEigenVector proposal is symmetric for ANY orthonormal basis and scales - an unconverged
// decomposition costs proposal efficiency in proportion to what is left off the diagonal (relative 1e-5 after 5 sweeps from the
// identity at D = 50), not correctness.  On exit eigval is ascending (np.linalg.eigh order),
// eig_vec[i][0..D) = i-th eigenvector (a ROW here), eig_scale[i] = sqrt|eigval[i]| (proposal.py:339).
constexpr int kJacobiThreads = 1024;

__global__ void __launch_bounds__(kJacobiThreads) k_jacobi_eigh(const double* __restrict__ C, double* scratch, int use_smem,
                                                                int D, int Dp, const double* __restrict__ prev_vec, int max_sweeps,
                                                                double* __restrict__ eigval,
                                                                double* __restrict__ eig_vec, double* __restrict__ eig_scale,
                                                                int* __restrict__ sweeps_out) {
    extern __shared__ double jsm[];
    __shared__ double alpha[256], beta[256];     // row/column i becomes alpha_i * (i) + beta_i * (partner_i)
    __shared__ int partner[256];
    __shared__ double red[kJacobiThreads];
    __shared__ int converged;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int LD = D | 1;
    double* base = use_smem ? jsm : scratch;
    double* A = base;
    double* A2 = base + (size_t)D * LD;
    double* V = base + 2 * (size_t)D * LD;
    double* V2 = base + 3 * (size_t)D * LD;
    const int n = (D + 1) & ~1;            // even number of players; index D (if any) is a bye
    const int half = n / 2;
    // Warm start: the live ensemble changes little between two refreshes, so in the basis of the previous eigenvectors
    // (prev_vec, rows) the covariance is nearly diagonal: A = Vp^T C Vp, V = Vp, a few sweeps instead of ~9 from the identity.
    // Used when the previous rows are orthonormal to 1e-6 (a buffer that was never filled is all zeros): the decision is a
    // function of the buffers alone, so a run resumed from a checkpoint makes the same one.
    double dev = 0.0;
    if (prev_vec != nullptr) {
        for (int w = tid; w < D * D; w += nt) {
            const int i = w / D, j = w % D;
            if (j < i) continue;
            double dot = 0.0;
            for (int k = 0; k < D; ++k) dot = fma(prev_vec[(size_t)i * Dp + k], prev_vec[(size_t)j * Dp + k], dot);
            dev = fmax(dev, fabs(dot - (i == j ? 1.0 : 0.0)));
        }
    } else {
        dev = 1.0;
    }
    red[tid] = dev;
    __syncthreads();
    for (int s = nt >> 1; s > 0; s >>= 1) { if (tid < s) red[tid] = fmax(red[tid], red[tid + s]); __syncthreads(); }
    const bool warm = red[0] < 1e-6;
    __syncthreads();
    if (warm) {
        for (int w = tid; w < D * D; w += nt) {           // V[k][i] = component k of eigenvector i;  A2 = C V
            const int k = w / D, i = w % D;
            V[k * LD + i] = prev_vec[(size_t)i * Dp + k];
        }
        __syncthreads();
        for (int w = tid; w < D * D; w += nt) {
            const int k = w / D, j = w % D;
            double t = 0.0;
            for (int l = 0; l < D; ++l) t = fma(C[k * D + l], V[l * LD + j], t);
            A2[k * LD + j] = t;
        }
        __syncthreads();
        for (int w = tid; w < D * D; w += nt) {           // A = V^T (C V), computed for i <= j and mirrored
            const int i = w / D, j = w % D;
            if (j < i) continue;
            double t = 0.0;
            for (int k = 0; k < D; ++k) t = fma(V[k * LD + i], A2[k * LD + j], t);
            A[i * LD + j] = t;
            A[j * LD + i] = t;
        }
    } else {
        for (int i = tid; i < D * D; i += nt) {
            const int r = i / D, c = i % D;
            A[r * LD + c] = C[i];
            V[r * LD + c] = (r == c) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        // convergence: off-diagonal Frobenius norm against the diagonal
        double off = 0.0, dia = 0.0;
        {
            const int di = nt / D, dj = nt % D;
            for (int w = tid, i = tid / D, j = tid % D; w < D * D; w += nt) {
                const double a = A[i * LD + j];
                if (i == j) dia += a * a; else off += a * a;
                i += di; j += dj;
                if (j >= D) { j -= D; ++i; }
            }
        }
        red[tid] = off;
        __syncthreads();
        for (int s = nt >> 1; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
        off = red[0];
        __syncthreads();
        red[tid] = dia;
        __syncthreads();
        for (int s = nt >> 1; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
        dia = red[0];
        if (tid == 0) converged = (off <= 1e-24 * dia) || (D == 1);     // |off|_F <= 1e-12 |diag|_F
        __syncthreads();
        if (converged) break;
        for (int r = 0; r < n - 1; ++r) {
            if (tid < half) {
                const int i = (r + tid) % (n - 1);
                const int j = (tid == 0) ? n - 1 : (r + n - 1 - tid) % (n - 1);
                const int p = min(i, j), q = max(i, j);
                double c = 1.0, s = 0.0;
                if (q < D) {
                    const double apq = A[p * LD + q];
                    if (apq != 0.0) {
                        const double theta = (A[q * LD + q] - A[p * LD + p]) / (2.0 * apq);
                        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                    }
                    // J = [[c, s], [-s, c]] on (p, q): row'_p = c row_p - s row_q ; row'_q = c row_q + s row_p
                    alpha[p] = c; beta[p] = -s; partner[p] = q;
                    alpha[q] = c; beta[q] = s;  partner[q] = p;
                } else if (p < D) {
                    alpha[p] = 1.0; beta[p] = 0.0; partner[p] = p;       // paired with the bye
                }
            }
            __syncthreads();
            // (element w = i D + j, w = tid, tid + nt, ...: the pair (i, j) is advanced without divisions)
            const int di = nt / D, dj = nt % D;
            for (int w = tid, i = tid / D, j = tid % D; w < D * D; w += nt) {
                const int pi = partner[i], pj = partner[j];
                const double ai = alpha[i], bi = beta[i], aj = alpha[j], bj = beta[j];
                const double t0 = ai * A[i * LD + j] + bi * A[pi * LD + j];
                const double t1 = ai * A[i * LD + pj] + bi * A[pi * LD + pj];
                A2[i * LD + j] = aj * t0 + bj * t1;
                V2[i * LD + j] = aj * V[i * LD + j] + bj * V[i * LD + pj];
                i += di; j += dj;
                if (j >= D) { j -= D; ++i; }
            }
            __syncthreads();
            double* t = A; A = A2; A2 = t;
            t = V; V = V2; V2 = t;
        }
    }
    if (tid == 0 && sweeps_out != nullptr) *sweeps_out = sweep;
    // ascending order by counting, ties by index
    for (int i = tid; i < D; i += nt) {
        const double li = A[i * LD + i];
        int rank = 0;
        for (int j = 0; j < D; ++j) {
            const double lj = A[j * LD + j];
            rank += (lj < li) || (lj == li && j < i);
        }
        eigval[rank] = li;
        eig_scale[rank] = sqrt(fabs(li));
        for (int k = 0; k < Dp; ++k) eig_vec[(size_t)rank * Dp + k] = (k < D) ? V[k * LD + i] : 0.0;
    }
}

}  // namespace cpn
