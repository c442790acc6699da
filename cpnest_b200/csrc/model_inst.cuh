// Instantiation + dispatch of the per-model kernels.  A translation unit defines CPN_MODEL and
// CPN_MODEL_ID and includes this file; it exports one launcher pair that api.cu indexes by
// functor id.  (G, E) = lanes per chain, coordinates per lane; G * E >= D.
#pragma once
#include <cuda_runtime.h>
#include "eval_kernel.cuh"
#include "mh_kernel.cuh"
#include "slice_kernel.cuh"
#include "hmc_kernel.cuh"
#include "launch.h"

#ifdef CPN_EXP_G4E16
#define CPN_CONFIGS(X) X(1, 1) X(1, 2) X(1, 4) X(1, 8) X(4, 4) X(4, 16) X(8, 2) X(8, 8) X(16, 4) X(32, 1) X(32, 2) X(32, 4) X(32, 8)
#else
#define CPN_CONFIGS(X) X(1, 1) X(1, 2) X(1, 4) X(1, 8) X(4, 4) X(8, 2) X(8, 8) X(16, 4) X(32, 1) X(32, 2) X(32, 4) X(32, 8)
#endif

namespace cpn {

template <class Model>
static int launch_mh_model(int G, int E, int replay, const MHParams& p, cudaStream_t s) {
    const int nchains = p.j1 - p.j0;
    if (nchains <= 0) return 0;
#define CPN_CASE(GG, EE)                                                                              \
    if (G == GG && E == EE) {                                                                         \
        const int cpb = kMHThreads / GG;                                                              \
        const int nb = (nchains + cpb - 1) / cpb;                                                     \
        const size_t sm = (Model::kScratch ? (size_t)cpb * kScratchRows * p.Dp * sizeof(double) : 0) +               \
                          (!replay && Model::template block_coop<GG> ? (size_t)Model::block_doubles(p.D) * sizeof(double) : 0); \
        const int nbt = (GG == 32 && Model::kTeam > 1) ? nchains : nb;   /* team mode: one chain per block */ \
        if (replay) k_mh_chains<GG, EE, Model, true><<<nb, kMHThreads, sm, s>>>(p);                   \
        else if (p.chain_log != nullptr || p.cycle_len > kCycInline) k_mh_chains<GG, EE, Model, false, true><<<nbt, kMHThreads, sm, s>>>(p); \
        else k_mh_chains<GG, EE, Model, false><<<nbt, kMHThreads, sm, s>>>(p);                        \
        return 0;                                                                                     \
    }
    CPN_CONFIGS(CPN_CASE)
#undef CPN_CASE
    return -1;
}

template <class Model>
static int launch_slice_model(int G, int E, int replay, const SliceParams& p, cudaStream_t s) {
    const int nchains = p.j1 - p.j0;
    if (nchains <= 0) return 0;
#define CPN_CASE(GG, EE)                                                                              \
    if (G == GG && E == EE) {                                                                         \
        const int cpb = kMHThreads / GG;                                                              \
        const int nb = (nchains + cpb - 1) / cpb;                                                     \
        const size_t sm = Model::kScratch ? (size_t)cpb * kScratchRows * p.Dp * sizeof(double) : 0;                  \
        if (replay) k_slice_chains<GG, EE, Model, true><<<nb, kMHThreads, sm, s>>>(p);                \
        else k_slice_chains<GG, EE, Model, false><<<nb, kMHThreads, sm, s>>>(p);                      \
        return 0;                                                                                     \
    }
    CPN_CONFIGS(CPN_CASE)
#undef CPN_CASE
    return -1;
}

template <class Model>
static int launch_hmc_model(int G, int E, int replay, const HMCParams& p, cudaStream_t s) {
    const int nchains = p.j1 - p.j0;
    if (nchains <= 0) return 0;
#define CPN_CASE(GG, EE)                                                                              \
    if (G == GG && E == EE) {                                                                         \
        const int cpb = kMHThreads / GG;                                                              \
        const int nb = (nchains + cpb - 1) / cpb;                                                     \
        const size_t sm = (Model::kScratch ? (size_t)cpb * kScratchRows * p.Dp * sizeof(double) : 0) +  \
                          (size_t)cpb * p.Dp * sizeof(double);   /* + a row per chain: the vector of a mat-vec */ \
        if (replay) k_hmc_chains<GG, EE, Model, true><<<nb, kMHThreads, sm, s>>>(p);                  \
        else k_hmc_chains<GG, EE, Model, false><<<nb, kMHThreads, sm, s>>>(p);                        \
        return 0;                                                                                     \
    }
    CPN_CONFIGS(CPN_CASE)
#undef CPN_CASE
    return -1;
}

template <class Model>
static int launch_eval_model(int G, int E, const EvalParams& p, cudaStream_t s) {
    if (p.n <= 0) return 0;
#define CPN_CASE(GG, EE)                                                                              \
    if (G == GG && E == EE) {                                                                         \
        const int cpb = kEvalThreads / GG;                                                            \
        const int nb = (p.n + cpb - 1) / cpb;                                                         \
        const size_t sm = Model::kScratch ? (size_t)cpb * kScratchRows * p.Dp * sizeof(double) : 0;                  \
        k_eval_points<GG, EE, Model><<<nb, kEvalThreads, sm, s>>>(p);                                 \
        return 0;                                                                                     \
    }
    CPN_CONFIGS(CPN_CASE)
#undef CPN_CASE
    return -1;
}

}  // namespace cpn

#define CPN_DEFINE_MODEL_LAUNCHERS(NAME, MODEL)                                                       \
    namespace cpn {                                                                                   \
    int launch_mh_##NAME(int G, int E, int replay, const MHParams& p, cudaStream_t s) {               \
        return launch_mh_model<MODEL>(G, E, replay, p, s);                                            \
    }                                                                                                 \
    int launch_eval_##NAME(int G, int E, const EvalParams& p, cudaStream_t s) {                       \
        return launch_eval_model<MODEL>(G, E, p, s);                                                  \
    }                                                                                                 \
    int launch_slice_##NAME(int G, int E, int replay, const SliceParams& p, cudaStream_t s) {         \
        return launch_slice_model<MODEL>(G, E, replay, p, s);                                         \
    }                                                                                                 \
    int launch_hmc_##NAME(int G, int E, int replay, const HMCParams& p, cudaStream_t s) {             \
        return launch_hmc_model<MODEL>(G, E, replay, p, s);                                           \
    }                                                                                                 \
    }
