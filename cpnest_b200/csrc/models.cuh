// Device functors for the Model contract (cpnest/model.py:20-33, 55-82): strict box test,
// log_prior, log_likelihood.  One functor per model the reference ships (tests/ and examples/);
// the Python callables of those models remain the parity oracle (tests/test_gpu_models.py).
//
// A functor sees the point distributed over the G lanes of its chain group: coordinate d is
// x[d / G] of lane d % G.  log_likelihood/log_prior return the same value on every lane.
#pragma once
#include "sysinc.cuh"
#include "group.cuh"

namespace cpn {

#define CPN_NEG_INF (__longlong_as_double(0xfff0000000000000LL))
#define CPN_HALF_LOG_2PI 0.91893853320467274178

// one FP64 tensor-core product C[8x8] += A[8x4] B[4x8]: lane 4 r + k supplies A[r][k] and B[k][r], holds C[r][2k], C[r][2k+1]
__device__ __forceinline__ void dmma_f64(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int kScratchRows = 2;   // shared-memory scratch per chain of functors with kScratch: kScratchRows * Dp doubles

struct ModelCtx {
    const double* blob;   // functor parameters (device)
    int D;                // dimension
    double* scratch;      // per-chain shared-memory scratch, kScratchRows * Dp doubles (functors with kScratch)
    double par[4];        // functor parameters kept in registers by Model::prepare (read once per chain, not per step)
    uint32_t vmsk[8];     // per slot e: all ones if coordinate lane + e*G exists (< D), else 0 (set by set_valid_mask)
    // team mode (MH production kernel, functors with kTeam > 1): `team` warps of the block evolve the SAME chain in
    // lock step (same draws, same decisions) and split the functor's data sum; team_buf[team] is block-shared
    int team = 1;
    int team_rank = 0;
    double* team_buf = nullptr;
    // block mode (MH production kernel, functors with block_coop<G>): the chains of a block evaluate the functor together
    // on data the block keeps in shared memory (Model::block_doubles(D) doubles of the dynamic array, filled by
    // Model::block_prepare)
    int bp[4] = {0, 0, 0, 0};   // per-thread constants of the block form (set by block_prepare)
    int bo = 0;
};

template <int G, int E>
__device__ __forceinline__ void set_valid_mask(const Group<G>& g, ModelCtx& m) {
#pragma unroll
    for (int e = 0; e < E; ++e) m.vmsk[e] = (g.lane + e * G < m.D) ? 0xffffffffu : 0u;
}
// v if slot e holds a coordinate, else +0.0: two bit-ANDs with a register mask
__device__ __forceinline__ double masked(const ModelCtx& m, int e, double v) {
    return __hiloint2double(__double2hiint(v) & (int)m.vmsk[e], __double2loint(v) & (int)m.vmsk[e]);
}

struct ModelBase {
    // warps per chain in the MH production kernel when a chain owns a whole warp (G == 32): > 1 for functors whose
    // cost is a long data sum, so that a batch of a few hundred chains still fills the SMs
    static constexpr int kTeam = 1;
    // true for the lane-group widths at which the functor offers log_likelihood_block: a collective of the whole block
    // (every thread calls it at every step, the block's warps stay in lock step; mh_kernel.cuh)
    template <int G>
    static constexpr bool block_coop = false;
    static __host__ __device__ __forceinline__ int block_doubles(int) { return 0; }
    static __device__ __forceinline__ void prepare(ModelCtx&) {}
    // true if the functor's value does not change when the slots beyond D of a point hold +0.0 and are NOT masked: the
    // MH kernel then skips the masks whenever the rows it gathers are zero-padded up to G*E doubles (mh_kernel.cuh)
    static __device__ __forceinline__ bool pad_ok(const ModelCtx&) { return false; }
    // called on the context of the mask-free loop: what pad_ok has established may be written down as constants
    static __device__ __forceinline__ void pad_specialise(ModelCtx&) {}
};

template <int G, int E>
__device__ __forceinline__ bool valid_dim(const Group<G>& g, int e, int D) { return g.lane + e * G < D; }

// ---- flat prior, model.py:66-82 -------------------------------------------------------------
struct FlatPrior : ModelBase {
    static constexpr bool kFlatPrior = true;
    template <int G, int E>
    static __device__ __forceinline__ double log_prior(const Group<G>&, const double (&)[E], const ModelCtx&) {
        return 0.0;
    }
    // Model.force (model.py:97-106): gradient of the potential -log_prior; zero for the flat prior, as in every
    // model the reference ships (tests/test_2d_hmc.py:22-24, examples/test_50d.py:25-28, examples/rosenbrock.py:24-26)
    template <int G, int E>
    static __device__ __forceinline__ void grad_potential(const Group<G>&, const double (&)[E], const ModelCtx&, double (&gv)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e) gv[e] = 0.0;
    }
};

// Gradients of log_likelihood (one component per coordinate, laid out like the point) give the normal of the
// iso-likelihood surface the constrained leapfrog reflects off; the reference estimates that normal from
// per-dimension spline fits over the ensemble (proposal.py:417-486), the device functors provide it analytically.

// ---- iid Gaussian: examples/test_50d.py:18-19 (sum_i -x_i^2/2 - log(2pi)/2), tests/test_2d.py:15-16,
//      tests/test_1d.py:20-21 and tests/test_half_gaussian.py:24-25 (scipy norm.logpdf) ----------------
struct GaussIid : FlatPrior {
    static constexpr bool kScratch = false;
    static constexpr bool kCheap = true;    // evaluated every step: cheaper than asking whether any chain needs it
    // blob = {mu, 1/sigma, c}, c = -log(sigma) - log(2pi)/2.  sum_i [c - ((x_i - mu)/sigma)^2 / 2] is evaluated as
    // D c - (sum_i (x_i - mu)^2) / (2 sigma^2): one subtraction and one FMA per coordinate in the chain kernels' inner loop
    static __device__ __forceinline__ void prepare(ModelCtx& m) {
        const double inv_sigma = m.blob[1];
        m.par[0] = m.blob[0];
        m.par[1] = -0.5 * inv_sigma * inv_sigma;
        m.par[2] = (double)m.D * m.blob[2];
        m.par[3] = inv_sigma * inv_sigma;
    }
    // a zero slot adds (0 - mu)^2 to the sum: nothing exactly when mu == 0
    static __device__ __forceinline__ bool pad_ok(const ModelCtx& m) { return m.par[0] == 0.0; }
    static __device__ __forceinline__ void pad_specialise(ModelCtx& m) { m.par[0] = 0.0; }   // x - 0.0 folds away
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        const double mu = m.par[0];
        double acc[2] = {0.0, 0.0};               // two partial sums: half the dependent FMA chain
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const double t = masked(m, e, x[e] - mu);
            acc[e & 1] = fma(t, t, acc[e & 1]);
        }
        return fma(m.par[1], g.sum(acc[0] + acc[1]), m.par[2]);
    }
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>&, const double (&x)[E], const ModelCtx& m, double (&gr)[E]) {
        const double mu = m.par[0], inv_var = m.par[3];
#pragma unroll
        for (int e = 0; e < E; ++e) gr[e] = -(x[e] - mu) * inv_var;
    }
};

// ---- correlated Gaussian N(0, Sigma), Sigma = L L^T: logL = -|L^-1 x|^2/2 - sum log L_ii - D/2 log 2pi
//      (not in the reference).  blob = {logdet, Linv[D][D] row-major lower, LinvT[D][S]}: the second
//      copy is the transpose, LinvT[c][r] = Linv[r][c] (zero for r < c and for the padding rows D..S-1, S = D rounded up
//      to 32), so that for a fixed column the lanes of a group, which own consecutive rows, read consecutive doubles -------
struct GaussCorr : FlatPrior {
    static constexpr bool kScratch = true;
    static constexpr bool kCheap = false;
    // y = Linv x as a group mat-vec: lane l owns the rows l + G e; the point is broadcast column by column through the
    // chain's shared-memory scratch.  Columns are walked in blocks of G: in block b the slots e > b take every column,
    // slot e == b meets the diagonal block (the zeros above the diagonal are stored), slots e < b are done - D(D+G)/2
    // multiply-adds on E independent accumulators per lane instead of a dependent chain of D per row.
    // row stride of the transposed copy: D rounded up to 32, the rows D.. of every column are zero (a lane whose row
    // does not exist multiplies zeros)
    static __host__ __device__ __forceinline__ int ldt(int D) { return (D + 31) & ~31; }
    template <int G, int E>
    static __device__ __forceinline__ void whiten(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&y)[E]) {
        const int D = m.D, S = ldt(D);
        const double* Lt = m.blob + 1 + (size_t)D * D + g.lane;      // + c*S + G*e
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            if (d < D) m.scratch[d] = x[e];
            y[e] = 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < E; ++b) {
            const int nc = min(G, D - G * b);                          // columns of this block (<= 0: none left)
            const double* col = Lt + (size_t)(G * b) * S;
            const double* xs = m.scratch + G * b;
#pragma unroll 4
            for (int c = 0; c < nc; ++c) {
                const double xc = xs[c];
#pragma unroll
                for (int e = b; e < E; ++e)
                    if (G * e < S) y[e] = fma(__ldg(col + G * e), xc, y[e]);
                col += S;
            }
        }
        __syncwarp();
    }
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        double y[E];
        whiten<G, E>(g, x, m, y);
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) acc = fma(y[e], y[e], acc);
        return -0.5 * g.sum(acc) - m.blob[0] - m.D * CPN_HALF_LOG_2PI;
    }
    // ---- block form: the eight 16-lane chains of a 128-thread block are the eight rows of FP64 tensor-core products.
    // Y^T[8 x 56] = X^T[8 x 52] Linv^T[52 x 56] in 4 x 8 tiles of Linv^T (lower triangle of Linv only: the block of output
    // coordinates i meets the blocks of input coordinates j <= 2 i + 1), one mma.sync.m8n8k4.f64 (DMMA) per tile: the
    // matrix is read once per EIGHT chains, from shared memory, where the per-lane form above issues one load per
    // multiply-add.  Tile (i, j) is stored in fragment order, lane 4 n + k holds Linv[8 i + n][4 j + k]; the points sit in a
    // table X[chain][kXs] (kXs = 4 mod 16: the A fragments, lane 4 c + k <- X[c][4 j + k], load without bank conflicts).
    // Warp w owns the output blocks w and nrb - 1 - w (a short and a long one: 16 products per warp at D = 50), which share
    // the A fragments; even and odd input blocks go to separate accumulators; every output block has an even number of tiles
    // (a zero tile is appended).  A lane ends up with two output coordinates per block of ONE chain (lane >> 2): their squares
    // are summed in the lane, over the four lanes of the chain by two shuffles, over the warps in a fixed order through shared
    // memory: every lane of a chain gets the bit-identical value.  All shared-memory traffic goes through the kernel's
    // dynamic array by offset (a generic pointer costs a window conversion per access).
    template <int G>
    static constexpr bool block_coop = (G == 16);
    static constexpr int kXs = 68;
    static __host__ __device__ __forceinline__ int ntile_row(int i, int nkb) {       // tiles of output block i (even)
        const int n = (2 * i + 2 < nkb) ? 2 * i + 2 : nkb;
        return (n + 1) & ~1;
    }
    static __host__ __device__ __forceinline__ int tile_base(int i, int nkb) {       // tiles of the output blocks before i
        int t = 0;
        for (int q = 0; q < i; ++q) t += ntile_row(q, nkb);
        return t;
    }
    static __host__ __device__ __forceinline__ int block_doubles(int D) {
        return D <= 64 ? tile_base((D + 7) >> 3, (D + 3) >> 2) * 32 + 8 * kXs + 32 : 0;
    }
    // called once by every thread of the block before the first step: fills the tile table and the thread's constants
    // bp = {shared-window byte address of its B fragments of the short output block, of the long one, of its A fragments, input
    // blocks of the short | long output block << 8}; bo = address of the point table
    static __device__ __forceinline__ void block_prepare(ModelCtx& m, int buf_off) {
        extern __shared__ double cpn_dyn_smem[];
        const int D = m.D, nrb = (D + 7) >> 3, nkb = (D + 3) >> 2;
        const double* Linv = m.blob + 1;
        double* tiles = cpn_dyn_smem + buf_off;
        const int ntile = tile_base(nrb, nkb);
        for (int i = 0, t0 = 0; i < nrb; ++i) {
            const int nj = ntile_row(i, nkb);
            for (int q = (int)threadIdx.x; q < nj * 32; q += (int)blockDim.x) {
                const int j = q >> 5, l = q & 31, r = 8 * i + (l >> 2), c = 4 * j + (l & 3);
                tiles[(t0 + j) * 32 + l] = (r < D && c <= r) ? __ldg(Linv + (size_t)r * D + c) : 0.0;
            }
            t0 += nj;
        }
        for (int q = (int)threadIdx.x; q < 8 * kXs + 32; q += (int)blockDim.x) tiles[ntile * 32 + q] = 0.0;
        const int lane = (int)threadIdx.x & 31, w = (int)threadIdx.x >> 5;
        const int iA = w, iB = nrb - 1 - w;               // short and long output block of this warp (iA > iB: none)
        int nA = 0, nB = 0;
        if (iA <= iB) {
            nA = (iA == iB) ? 0 : ntile_row(iA, nkb);
            nB = ntile_row(iB, nkb);
        }
        // (as shared-window byte addresses: log_likelihood_block reads and writes through them)
        const int sb = (int)(uint32_t)__cvta_generic_to_shared(cpn_dyn_smem);
        m.bp[0] = sb + 8 * (buf_off + tile_base(iA < nrb ? iA : 0, nkb) * 32 + lane);
        m.bp[1] = sb + 8 * (buf_off + tile_base(iB >= 0 ? iB : 0, nkb) * 32 + lane);
        m.bp[2] = sb + 8 * (buf_off + ntile * 32 + (lane >> 2) * kXs + (lane & 3));
        m.bp[3] = nA | (nB << 8);
        m.bo = sb + 8 * (buf_off + ntile * 32);
        __syncthreads();
    }
    // shared-memory accesses of the block form by 32-bit shared-window address (a pointer derived from the dynamic array is
    // generic: every loop rebuilt the window base from SR_CgaCtaId, a quarter of the functor's instructions)
    static __device__ __forceinline__ double lds64(uint32_t a) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ double lds64_256(uint32_t a) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1+256];" : "=d"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ double lds64_32(uint32_t a) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1+32];" : "=d"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood_block(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        static_assert(G == 16, "eight chains per 128-thread block");
        const int D = m.D;
        const int lane = (int)threadIdx.x & 31, w = (int)threadIdx.x >> 5, chain = (int)threadIdx.x >> 4;
        const uint32_t aX = (uint32_t)m.bo, aP = aX + 8u * 8u * kXs;               // point table, partial sums [4 warps][8 chains]
        uint32_t aA = (uint32_t)m.bp[0], aB = (uint32_t)m.bp[1], ax = (uint32_t)m.bp[2];
        const int nA = m.bp[3] & 0xff, nB = m.bp[3] >> 8;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            if (d < D) sts64(aX + 8u * (uint32_t)(chain * kXs + d), x[e]);
        }
        __syncthreads();
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        int j = 0;
        for (; j < nA; j += 2) {                          // both output blocks, which share the A fragments
            const double x0 = lds64(ax), x1 = lds64_32(ax);
            dmma_f64(b0, b1, x0, lds64(aB));
            dmma_f64(a0, a1, x0, lds64(aA));
            dmma_f64(b2, b3, x1, lds64_256(aB));
            dmma_f64(a0, a1, x1, lds64_256(aA));
            ax += 64u; aA += 512u; aB += 512u;
        }
        for (; j < nB; j += 2) {                          // the long output block alone (even and odd input blocks apart)
            dmma_f64(b0, b1, lds64(ax), lds64(aB));
            dmma_f64(b2, b3, lds64_32(ax), lds64_256(aB));
            ax += 64u; aB += 512u;
        }
        b0 += b2; b1 += b3;
        double q = fma(a0, a0, a1 * a1) + fma(b0, b0, b1 * b1);
        // the chain lives in lane >> 2: sum over its four lanes
        q += __shfl_xor_sync(FULL, q, 1);
        q += __shfl_xor_sync(FULL, q, 2);
        if ((lane & 3) == 0) sts64(aP + 8u * (uint32_t)(w * 8 + (lane >> 2)), q);
        __syncthreads();
        const uint32_t ac = aP + 8u * (uint32_t)chain;
        const double s = (lds64(ac) + lds64(ac + 64u)) + (lds64(ac + 128u) + lds64(ac + 192u));
        return -0.5 * s - m.blob[0] - m.D * CPN_HALF_LOG_2PI;
    }
    // grad = -Linv^T (Linv x): the second product walks the rows of Linv^T = columns of Linv, again consecutive in the
    // lanes when read from the row-major copy transposed, i.e. from Linv itself: (Linv^T y)_c = sum_{r >= c} Linv[r][c] y_r
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&gr)[E]) {
        const int D = m.D, Dp = (D + 3) & ~3;
        const double* Linv = m.blob + 1;
        double y[E];
        whiten<G, E>(g, x, m, y);
        double* ys = m.scratch + Dp;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            if (d < D) ys[d] = y[e];
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int c = g.lane + e * G;
            double v = 0.0;
            if (c < D)
                for (int r = c; r < D; ++r) v = fma(__ldg(Linv + (size_t)r * D + c), ys[r], v);
            gr[e] = -v;
        }
        __syncwarp();
    }
};

// ---- two-component isotropic Gaussian mixture in PARAMETER space (BASELINE.json config 5 names a "20-D Gaussian
//      mixture", which the reference does not contain: examples/gaussianmixture.py is 5-D over a data vector):
//      L(x) = w N(x; m1 1, s1^2 I) + (1 - w) N(x; m2 1, s2^2 I);  blob = {w, m1, s1, m2, s2};  Z = 1/volume(box) ------
struct GaussMixND : FlatPrior {
    static constexpr bool kScratch = false;
    static constexpr bool kCheap = true;
    static __device__ __forceinline__ void prepare(ModelCtx& m) {
        m.par[0] = m.blob[1]; m.par[1] = 1.0 / m.blob[2]; m.par[2] = m.blob[3]; m.par[3] = 1.0 / m.blob[4];
    }
    template <int G, int E>
    static __device__ __forceinline__ void components(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double& l1, double& l2) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const double t1 = (x[e] - m.par[0]) * m.par[1], t2 = (x[e] - m.par[2]) * m.par[3];
            a += masked(m, e, t1 * t1);
            b += masked(m, e, t2 * t2);
        }
        a = g.sum(a);
        b = g.sum(b);
        const double w = m.blob[0];
        l1 = log(w) - 0.5 * a + m.D * (log(m.par[1]) - CPN_HALF_LOG_2PI);
        l2 = log(1.0 - w) - 0.5 * b + m.D * (log(m.par[3]) - CPN_HALF_LOG_2PI);
    }
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        double l1, l2;
        components<G, E>(g, x, m, l1, l2);
        const double hi = fmax(l1, l2), lo = fmin(l1, l2);
        return hi + log1p(exp(lo - hi));
    }
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&gr)[E]) {
        double l1, l2;
        components<G, E>(g, x, m, l1, l2);
        const double hi = fmax(l1, l2), lae = hi + log1p(exp(fmin(l1, l2) - hi));
        const double r1 = exp(l1 - lae), r2 = exp(l2 - lae);
#pragma unroll
        for (int e = 0; e < E; ++e)
            gr[e] = masked(m, e, -r1 * (x[e] - m.par[0]) * m.par[1] * m.par[1] - r2 * (x[e] - m.par[2]) * m.par[3] * m.par[3]);
    }
};

// ---- eggbox: examples/eggbox.py:19-23  (2 + prod cos(x_i/2))^5 ------------------------------
struct Eggbox : FlatPrior {
    static constexpr bool kScratch = false;
    static constexpr bool kCheap = true;
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        double p = 1.0;
#pragma unroll
        for (int e = 0; e < E; ++e) p *= valid_dim<G, E>(g, e, m.D) ? cos(x[e] / 2.0) : 1.0;
        p = g.prod(p);
        return pow(p + 2.0, 5.0);
    }
    // d/dx_i (2 + P)^5 = 5 (2 + P)^4 * (-1/2) sin(x_i/2) prod_{j != i} cos(x_j/2)
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&gr)[E]) {
        double cs[E], p = 1.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            cs[e] = valid_dim<G, E>(g, e, m.D) ? cos(x[e] / 2.0) : 1.0;
            p *= cs[e];
        }
        p = g.prod(p);
        const double f = 5.0 * (p + 2.0) * (p + 2.0) * (p + 2.0) * (p + 2.0);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const double others = (cs[e] != 0.0) ? p / cs[e] : 0.0;
            gr[e] = valid_dim<G, E>(g, e, m.D) ? f * (-0.5 * sin(x[e] / 2.0)) * others : 0.0;
        }
    }
};

// ---- Rosenbrock: examples/rosenbrock.py:20-22 ------------------------------------------------
struct Rosenbrock : FlatPrior {
    static constexpr bool kScratch = false;
    static constexpr bool kCheap = true;
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            // x_{d+1}: same slot of the next lane, or slot e+1 of lane 0 when this is the last lane
            double nxt;
            if (G == 1) {
                nxt = (e + 1 < E) ? x[(e + 1 < E) ? e + 1 : e] : 0.0;
            } else {
                double a = __shfl_sync(FULL, x[e], (g.lane + 1) & (G - 1), G);
                double b = (e + 1 < E) ? __shfl_sync(FULL, x[(e + 1 < E) ? e + 1 : e], 0, G) : 0.0;
                nxt = (g.lane == G - 1) ? b : a;
            }
            int d = g.lane + e * G;
            double u = nxt - x[e] * x[e];
            double v = 1.0 - x[e];
            double term = 100.0 * u * u + v * v;
            acc += (d < m.D - 1) ? term : 0.0;
        }
        return -g.sum(acc);
    }
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&gr)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            double nxt, prv;
            if (G == 1) {
                nxt = (e + 1 < E) ? x[(e + 1 < E) ? e + 1 : e] : 0.0;
                prv = (e > 0) ? x[(e > 0) ? e - 1 : 0] : 0.0;
            } else {
                const double a = __shfl_sync(FULL, x[e], (g.lane + 1) & (G - 1), G);
                const double b = (e + 1 < E) ? __shfl_sync(FULL, x[(e + 1 < E) ? e + 1 : e], 0, G) : 0.0;
                nxt = (g.lane == G - 1) ? b : a;
                const double c = __shfl_sync(FULL, x[e], (g.lane + G - 1) & (G - 1), G);
                const double dd = (e > 0) ? __shfl_sync(FULL, x[(e > 0) ? e - 1 : 0], G - 1, G) : 0.0;
                prv = (g.lane == 0) ? dd : c;
            }
            const int d = g.lane + e * G;
            double v = 0.0;
            if (d < m.D - 1) v += -400.0 * x[e] * (nxt - x[e] * x[e]) - 2.0 * (1.0 - x[e]);
            if (d > 0 && d < m.D) v += 200.0 * (x[e] - prv * prv);
            gr[e] = -v;
        }
    }
};

// ---- standard D-dim Ackley, logL = -ackley(x) (examples/ackley.py:21-28 as written
//      has no finite value anywhere, so there is no reference parity for this functor) --------------------
struct Ackley : FlatPrior {
    static constexpr bool kScratch = false;
    static constexpr bool kCheap = true;
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        double s2 = 0.0, sc = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            s2 += masked(m, e, x[e] * x[e]);
            sc += masked(m, e, cos(2.0 * 3.14159265358979323846 * x[e]));
        }
        s2 = g.sum(s2);
        sc = g.sum(sc);
        double a = -20.0 * exp(-0.2 * sqrt(s2 / m.D)) - exp(sc / m.D) + 20.0 + 2.71828182845904523536;
        return -a;
    }
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&gr)[E]) {
        const double twopi = 2.0 * 3.14159265358979323846;
        double s2 = 0.0, sc = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            bool ok = valid_dim<G, E>(g, e, m.D);
            s2 += ok ? x[e] * x[e] : 0.0;
            sc += ok ? cos(twopi * x[e]) : 0.0;
        }
        s2 = g.sum(s2);
        sc = g.sum(sc);
        const double r = sqrt(s2 / m.D);
        const double t1 = (r > 0.0) ? 4.0 * exp(-0.2 * r) / (m.D * r) : 0.0, t2 = twopi / m.D * exp(sc / m.D);
#pragma unroll
        for (int e = 0; e < E; ++e) gr[e] = valid_dim<G, E>(g, e, m.D) ? -(t1 * x[e] + t2 * sin(twopi * x[e])) : 0.0;
    }
};

// ---- ring: tests/test_ring.py:19-23 ; blob = {radius, thickness} ------------------------------
struct Ring : FlatPrior {
    static constexpr bool kScratch = false;
    static constexpr bool kCheap = true;
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        double a = coord<G, E>(g, x, 0), b = coord<G, E>(g, x, 1);
        double r = sqrt(a * a + b * b);
        double dr = m.blob[0] - r;
        return -0.5 * (dr * dr) / (m.blob[1] * m.blob[1]);
    }
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&gr)[E]) {
        const double a = coord<G, E>(g, x, 0), b = coord<G, E>(g, x, 1);
        const double r = sqrt(a * a + b * b);
        const double f = (r > 0.0) ? (m.blob[0] - r) / (m.blob[1] * m.blob[1]) / r : 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) gr[e] = (g.lane + e * G < 2) ? f * x[e] : 0.0;
    }
};

// ---- (mean, sigma) Gaussian with a 1/sigma prior: tests/test_gaussian.py:15-20 ----------------
struct GaussMeanSigma : ModelBase {
    static constexpr bool kFlatPrior = false;
    static constexpr bool kScratch = false;
    static constexpr bool kCheap = true;
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx&) {
        double mean = coord<G, E>(g, x, 0), sigma = coord<G, E>(g, x, 1);
        return -0.5 * (mean * mean) / (sigma * sigma) - log(sigma) - CPN_HALF_LOG_2PI;
    }
    template <int G, int E>
    static __device__ __forceinline__ double log_prior(const Group<G>& g, const double (&x)[E], const ModelCtx&) {
        double sigma = coord<G, E>(g, x, 1);
        return -log(sigma) - 2.30258509299404568402 /*log 10*/ - (-0.05129329438755058) /*log 0.95*/;
    }
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx&, double (&gr)[E]) {
        const double mean = coord<G, E>(g, x, 0), sigma = coord<G, E>(g, x, 1);
        const double v[2] = {-mean / (sigma * sigma), mean * mean / (sigma * sigma * sigma) - 1.0 / sigma};
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            gr[e] = (d == 0) ? v[0] : (d == 1 ? v[1] : 0.0);
        }
    }
    // potential = -log_prior = log(sigma) + const
    template <int G, int E>
    static __device__ __forceinline__ void grad_potential(const Group<G>& g, const double (&x)[E], const ModelCtx&, double (&gv)[E]) {
        const double sigma = coord<G, E>(g, x, 1);
#pragma unroll
        for (int e = 0; e < E; ++e) gv[e] = (g.lane + e * G == 1) ? 1.0 / sigma : 0.0;
    }
};

// exp(x) for x <= 0, straight-line code (no range branches, so that the unrolled copies of a data loop interleave on the
// FP64 pipe): x is clamped at -50 (exp(-50) = 2e-22 vanishes against the 1 it is added to), n = rint(x log2 e) by the
// 1.5 * 2^52 trick, r = x - n ln2 in two pieces, Taylor polynomial of degree 13 on |r| <= ln2 / 2 (truncation 4e-18),
// 2^n applied to the exponent field.  Measured against numpy.exp on 3e6 points of [-50, 0]: <= 2 ulp.
__device__ __forceinline__ double exp_nonpos(double x) {
    x = fmax(x, -50.0);
    const double magic = 6755399441055744.0;
    const double t = fma(x, 1.4426950408889634074, magic);
    const int n = __double2loint(t);
    const double fn = t - magic;
    double r = fma(fn, -6.93147180369123816490e-01, x);
    r = fma(fn, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;              // 1/13!
    p = fma(p, r, 2.08767569878681e-09);            // 1/12!
    p = fma(p, r, 2.505210838544172e-08);           // 1/11!
    p = fma(p, r, 2.755731922398589e-07);           // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);          // 1/9!
    p = fma(p, r, 2.48015873015873e-05);            // 1/8!
    p = fma(p, r, 1.984126984126984e-04);           // 1/7!
    p = fma(p, r, 1.3888888888888889e-03);          // 1/6!
    p = fma(p, r, 8.333333333333333e-03);           // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);          // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);          // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// ---- two-component mixture over a data vector: examples/gaussianmixture.py:22-31 ;
//      blob = {n_data, data[n_data]} ; parameters (mean1, sigma1, mean2, sigma2, weight) ----------
struct GaussMixture : ModelBase {
    static constexpr bool kFlatPrior = false;
    static constexpr bool kScratch = false;
    static constexpr bool kCheap = false;
    static constexpr int kTeam = 4;
    template <int G, int E>
    static __device__ __forceinline__ double log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m) {
        const double m1 = coord<G, E>(g, x, 0), s1 = coord<G, E>(g, x, 1), m2 = coord<G, E>(g, x, 2),
                     s2 = coord<G, E>(g, x, 3), w = coord<G, E>(g, x, 4);
        const int n = (int)m.blob[0];
        const double* data = m.blob + 1;
        const double c1 = log(w) - log(s1), c2 = log(1.0 - w) - log(s2);
        const double i1 = 1.0 / s1, i2 = 1.0 / s2;
        // np.logaddexp(l1, l2) = hi + log1p(exp(lo - hi)), summed over the data.  The log1p terms are collected as a
        // product: every factor 1 + exp(lo - hi) lies in [1, 2], so 512 of them cannot overflow, and one log per 512
        // terms replaces 512 log1p calls (rounding: <= 1 ulp per factor, i.e. < 1e-13 absolute on a sum of O(1e4)).
        // the data are dealt to the lanes of the whole team (m.team warps of G lanes; team == 1 outside team mode)
        const int W = G * m.team;
        double acc = 0.0;
        for (int base = m.team_rank * G + g.lane; base < n; base += 512 * W) {
            const int end = min(n, base + 512 * W);
            double prod = 1.0;
#pragma unroll 8
            for (int i = base; i < end; i += W) {
                const double d = __ldg(data + i);
                const double t1 = (d - m1) * i1, t2 = (d - m2) * i2;
                const double l1 = c1 - 0.5 * t1 * t1, l2 = c2 - 0.5 * t2 * t2;
                const double hi = fmax(l1, l2), lo = fmin(l1, l2);
                acc += hi;
                prod *= 1.0 + exp_nonpos(lo - hi);
            }
            acc += log(prod);
        }
        acc = g.sum(acc);
        if (m.team > 1) {
            // every warp of the team reaches this point together (identical chain state, hence identical control flow);
            // all of them add the partial sums in the same order, so they keep agreeing bit for bit
            if (g.lane == 0) m.team_buf[m.team_rank] = acc;
            __syncthreads();
            acc = 0.0;
            for (int w = 0; w < m.team; ++w) acc += m.team_buf[w];
            __syncthreads();
        }
        return acc;
    }
    template <int G, int E>
    static __device__ __forceinline__ double log_prior(const Group<G>& g, const double (&x)[E], const ModelCtx&) {
        return -log(coord<G, E>(g, x, 1)) - log(coord<G, E>(g, x, 3));
    }
    // responsibilities r_k = exp(l_k - logaddexp(l1, l2)):
    //   d/dmean_k = sum r_k (d - mean_k)/sigma_k^2 ; d/dsigma_k = sum r_k ((d - mean_k)^2/sigma_k^3 - 1/sigma_k) ;
    //   d/dw = sum r_1 / w - r_2 / (1 - w)
    template <int G, int E>
    static __device__ __forceinline__ void grad_log_likelihood(const Group<G>& g, const double (&x)[E], const ModelCtx& m, double (&gr)[E]) {
        const double m1 = coord<G, E>(g, x, 0), s1 = coord<G, E>(g, x, 1), m2 = coord<G, E>(g, x, 2),
                     s2 = coord<G, E>(g, x, 3), w = coord<G, E>(g, x, 4);
        const int n = (int)m.blob[0];
        const double* data = m.blob + 1;
        const double c1 = log(w) - log(s1), c2 = log(1.0 - w) - log(s2);
        const double i1 = 1.0 / s1, i2 = 1.0 / s2;
        double a[5] = {0, 0, 0, 0, 0};
        for (int i = g.lane; i < n; i += G) {
            const double d = __ldg(data + i);
            const double t1 = (d - m1) * i1, t2 = (d - m2) * i2;
            const double l1 = c1 - 0.5 * t1 * t1, l2 = c2 - 0.5 * t2 * t2;
            const double hi = fmax(l1, l2), lae = hi + log1p(exp(fmin(l1, l2) - hi));
            const double r1 = exp(l1 - lae), r2 = exp(l2 - lae);
            a[0] += r1 * t1 * i1;
            a[1] += r1 * (t1 * t1 - 1.0) * i1;
            a[2] += r2 * t2 * i2;
            a[3] += r2 * (t2 * t2 - 1.0) * i2;
            a[4] += r1 / w - r2 / (1.0 - w);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) a[k] = g.sum(a[k]);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 5; ++k) v = (d == k) ? a[k] : v;
            gr[e] = v;
        }
    }
    template <int G, int E>
    static __device__ __forceinline__ void grad_potential(const Group<G>& g, const double (&x)[E], const ModelCtx&, double (&gv)[E]) {
        const double s1 = coord<G, E>(g, x, 1), s2 = coord<G, E>(g, x, 3);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int d = g.lane + e * G;
            gv[e] = (d == 1) ? 1.0 / s1 : (d == 3 ? 1.0 / s2 : 0.0);
        }
    }
};

}  // namespace cpn
