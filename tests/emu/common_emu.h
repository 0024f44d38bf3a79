// Stand-in for mvster_b200/csrc/common.cuh in the CPU emulation build: the CUDA runtime names the SIMT sources use, as host stubs.
#pragma once
#include "simt_emu.h"
#include <cstdarg>
#include <cstdio>
#include <stdint.h>
#include "../../../../include/mvster_b200.h"

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

namespace mvster {
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
inline int check_launch(const char*) { count_launch(); return MVSTER_OK; }
#define MVSTER_REQUIRE(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            ::mvster::set_error(__VA_ARGS__); \
            return MVSTER_ERR_ARG;           \
        }                                    \
    } while (0)
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
int conv_px2(const float* x, const float* w, const float* bias, const float* skip, float* y,
             int B, int Di, int Hi, int Wi, int Cin, int Cout, int kd, int k, int sd, int s, int relu, cudaStream_t st);
int conv0_bf16(const void* x, const float* w, const float* scale, const float* bias, void* y, int out_bf16, long long NP, int H, int W, int Cin,
               cudaStream_t st);
int conv0_packed_h16(const float* x, const float* w, const float* bias, void* y, unsigned* overflow, long long NP, int H, int W, int Cin,
                     cudaStream_t st);
}  // namespace mvster
