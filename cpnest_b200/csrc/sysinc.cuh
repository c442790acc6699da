// System headers for the device code.  Under NVRTC (user-supplied functors are compiled at run time,
// csrc/rtc.cu) the C library headers are not available: the fixed-width integer types are spelled out and
// the math functions are NVRTC built-ins.
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#else
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#endif
