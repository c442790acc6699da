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

constexpr int kStatBlocks = 128;
constexpr int kStatThreads = 256;
constexpr int kCovChunk = 16;      // covariance entries per thread per pass
constexpr int kCovTile = 32;       // points staged in shared memory per step

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

__global__ void __launch_bounds__(kStatThreads) k_cov_partial(const double* __restrict__ x, const double* __restrict__ mean,
                                                              int n, int D, int Dp, double* __restrict__ partial /*[blocks][D*D]*/) {
    extern __shared__ double tile[];   // [kCovTile][D] centred points
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(n, p0 + per);
    const int nent = D * D;
    for (int ebase = 0; ebase < nent; ebase += kStatThreads * kCovChunk) {
        double acc[kCovChunk];
        int ra[kCovChunk], rb[kCovChunk];
#pragma unroll
        for (int k = 0; k < kCovChunk; ++k) {
            acc[k] = 0.0;
            const int ent = ebase + k * kStatThreads + threadIdx.x;
            ra[k] = (ent < nent) ? ent / D : -1;
            rb[k] = (ent < nent) ? ent % D : 0;
        }
        for (int t0 = p0; t0 < p1; t0 += kCovTile) {
            const int nt = min(kCovTile, p1 - t0);
            for (int i = threadIdx.x; i < nt * D; i += kStatThreads) {
                const int pi = i / D, d = i % D;
                tile[pi * D + d] = x[(size_t)(t0 + pi) * Dp + d] - mean[d];
            }
            __syncthreads();
            for (int pi = 0; pi < nt; ++pi) {
                const double* row = tile + pi * D;
#pragma unroll
                for (int k = 0; k < kCovChunk; ++k)
                    if (ra[k] >= 0) acc[k] = fma(row[ra[k]], row[rb[k]], acc[k]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < kCovChunk; ++k) {
            const int ent = ebase + k * kStatThreads + threadIdx.x;
            if (ent < nent) partial[(size_t)blockIdx.x * nent + ent] = acc[k];
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

// Cyclic Jacobi, round-robin ordering: n/2 disjoint rotations per round, n-1 rounds per sweep.
// The input matrix Ain (D x D, row-major) is copied into the working matrix A, which together with
// the accumulated rotations V lives in shared memory when 2*D*D doubles fit (D <= 100), otherwise in
// the global scratch buffers.  On exit eigval is ascending (np.linalg.eigh order),
// eig_vec[i][0..D) = i-th eigenvector (a ROW here), eig_scale[i] = sqrt|eigval[i]| (proposal.py:339).
__global__ void __launch_bounds__(256) k_jacobi_eigh(const double* __restrict__ Ain, double* Aglob, double* Vglob,
                                                     int use_smem, int D, int Dp, double* __restrict__ eigval,
                                                     double* __restrict__ eig_vec, double* __restrict__ eig_scale) {
    extern __shared__ double jsm[];
    __shared__ double cs[128], sn[128];
    __shared__ int pp[128], qq[128];
    __shared__ double red[256];
    __shared__ int converged;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* A = use_smem ? jsm : Aglob;
    double* V = use_smem ? jsm + D * D : Vglob;
    const int n = (D + 1) & ~1;            // even number of players; index D (if any) is a bye
    const int half = n / 2;
    for (int i = tid; i < D * D; i += nt) {
        A[i] = Ain[i];
        V[i] = ((i / D) == (i % D)) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int sweep = 0; sweep < 30; ++sweep) {
        // convergence: off-diagonal Frobenius norm against the diagonal
        double off = 0.0, dia = 0.0;
        for (int i = tid; i < D * D; i += nt) {
            const double a = A[i];
            if ((i / D) == (i % D)) dia += a * a; else off += a * a;
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
        if (tid == 0) converged = (off <= 1e-25 * dia) || (D == 1);
        __syncthreads();
        if (converged) break;
        for (int r = 0; r < n - 1; ++r) {
            if (tid < half) {
                int i = (r + tid) % (n - 1);
                int j = (tid == 0) ? n - 1 : (r + n - 1 - tid) % (n - 1);
                int p = min(i, j), q = max(i, j);
                double c = 1.0, s = 0.0;
                if (q < D) {
                    const double apq = A[p * D + q];
                    if (apq != 0.0) {
                        const double theta = (A[q * D + q] - A[p * D + p]) / (2.0 * apq);
                        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                    }
                } else {
                    p = -1;
                }
                pp[tid] = p; qq[tid] = q; cs[tid] = c; sn[tid] = s;
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int w = tid; w < half * D; w += nt) {
                const int k = w / D, m = w % D;
                const int p = pp[k], q = qq[k];
                if (p >= 0) {
                    const double c = cs[k], s = sn[k];
                    const double ap = A[p * D + m], aq = A[q * D + m];
                    A[p * D + m] = c * ap - s * aq;
                    A[q * D + m] = s * ap + c * aq;
                }
            }
            __syncthreads();
            // columns: A <- A J, V <- V J
            for (int w = tid; w < half * D; w += nt) {
                const int k = w / D, m = w % D;
                const int p = pp[k], q = qq[k];
                if (p >= 0) {
                    const double c = cs[k], s = sn[k];
                    const double ap = A[m * D + p], aq = A[m * D + q];
                    A[m * D + p] = c * ap - s * aq;
                    A[m * D + q] = s * ap + c * aq;
                    const double vp = V[m * D + p], vq = V[m * D + q];
                    V[m * D + p] = c * vp - s * vq;
                    V[m * D + q] = s * vp + c * vq;
                }
            }
            __syncthreads();
        }
    }
    // ascending order by counting, ties by index
    for (int i = tid; i < D; i += nt) {
        const double li = A[i * D + i];
        int rank = 0;
        for (int j = 0; j < D; ++j) {
            const double lj = A[j * D + j];
            rank += (lj < li) || (lj == li && j < i);
        }
        eigval[rank] = li;
        eig_scale[rank] = sqrt(fabs(li));
        for (int k = 0; k < Dp; ++k) eig_vec[(size_t)rank * Dp + k] = (k < D) ? V[k * D + i] : 0.0;
    }
}

}  // namespace cpn
