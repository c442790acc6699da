// Run-time compilation of the chain kernels around a user-written functor (csrc/user_model.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
namespace cpn {
struct RtcModule;
// Compiles { eval, mh, slice, hmc } kernels for the given (G, E) configurations with NVRTC for sm_100a and loads
// them.  Returns nullptr and fills `err` on failure.  `kinds` is a bit mask: 1 eval, 2 mh, 4 slice, 8 hmc.
RtcModule* rtc_compile(const char* user_source, const char* const* include_dirs, int n_dirs, const int (*cfgs)[2], int n_cfgs,
                       int kinds, char* err, size_t err_len);
void rtc_destroy(RtcModule* m);
// Launches kernel `kind` ("eval" | "mh" | "slice" | "hmc") instantiated for (G, E) with one by-value parameter struct.
int rtc_launch(RtcModule* m, const char* kind, int G, int E, const void* params, int nblocks, int nthreads, size_t smem,
               cudaStream_t stream, char* err, size_t err_len);
}  // namespace cpn
