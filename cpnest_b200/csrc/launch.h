// launcher table shared between the per-model translation units and api.cu
#pragma once
#include <cuda_runtime.h>
namespace cpn {
struct MHParams;
struct EvalParams;
struct SliceParams;
struct HMCParams;
#define CPN_DECLARE_MODEL(NAME)                                                          \
    int launch_mh_##NAME(int G, int E, int replay, const MHParams& p, cudaStream_t s);   \
    int launch_eval_##NAME(int G, int E, const EvalParams& p, cudaStream_t s);           \
    int launch_slice_##NAME(int G, int E, int replay, const SliceParams& p, cudaStream_t s); \
    int launch_hmc_##NAME(int G, int E, int replay, const HMCParams& p, cudaStream_t s);
CPN_DECLARE_MODEL(gauss_iid)
CPN_DECLARE_MODEL(gauss_corr)
CPN_DECLARE_MODEL(eggbox)
CPN_DECLARE_MODEL(rosenbrock)
CPN_DECLARE_MODEL(ackley)
CPN_DECLARE_MODEL(ring)
CPN_DECLARE_MODEL(gauss_mean_sigma)
CPN_DECLARE_MODEL(gauss_mixture)
CPN_DECLARE_MODEL(gauss_mix_nd)
}  // namespace cpn
