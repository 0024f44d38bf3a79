// Shared helpers for libmvster_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/mvster_b200.h"

namespace mvster {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MVSTER_ERR_CUDA;
    }
    count_launch();
    return MVSTER_OK;
}

#define MVSTER_REQUIRE(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            ::mvster::set_error(__VA_ARGS__); \
            return MVSTER_ERR_ARG;           \
        }                                    \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// conv_simt_px2.cu: two-output-pixels-per-thread CUDA-core convolution; -100 = shape not covered (use the 1-pixel kernel)
int conv_px2(const float* x, const float* w, const float* bias, const float* skip, float* y,
             int B, int Di, int Hi, int Wi, int Cin, int Cout, int kd, int k, int sd, int s, int relu, cudaStream_t st);

// conv_simt_px2.cu: conv0 (G -> 8, 3x3) of the bf16-storage regulariser; -100 = shape not covered
int conv0_bf16(const void* x, const float* w, const float* scale, const float* bias, void* y, int out_bf16, long long NP, int H, int W, int Cin,
               cudaStream_t st);

int conv0_packed_h16(const float* x, const float* w, const float* bias, void* y, unsigned* overflow, long long NP, int H, int W, int Cin,
                     cudaStream_t st);
}  // namespace mvster
