// CUDA-core channels-last convolution, two output pixels per thread.
// Same arithmetic as conv_fwd_kernel (conv_simt.cu) / conv2d_kernel (fpn.cu) - one fp32 FMA chain per output -
// but every broadcast weight read from shared memory now feeds two pixels, halving the shared-memory wavefronts
// per FMA (the single-pixel kernels were bound by the LSU data pipe, not by the FMA pipe).  Generic in the
// kernel extent: (kd, k, k) with kd in {1,3}, k in {1,3,5}; strides 1/2; used for the strided layers of the
// regulariser (mvs4net_utils.py:876-880) and the 5x5 stride-2 / small-channel layers of FPN4 (:424-445).
#include "common.cuh"
#include <stdlib.h>

namespace mvster {

struct ConvPxArgs {
    const float* x; const float* w; const float* bias; const float* skip; float* y;
    int B, Di, Hi, Wi, Do, Ho, Wo, cout, kd, k, sd, s, relu;
};

template <int CIN, int COUT_T>
__global__ void __launch_bounds__(128) conv_px2_kernel(const ConvPxArgs a) {
    extern __shared__ __align__(16) float w_s[];  // [kd*k*k][CIN][COUT_T]
    const int cg = blockIdx.y, taps = a.kd * a.k * a.k, pad = a.k / 2, pz = a.kd / 2;
    for (int i = threadIdx.x; i < taps * CIN * COUT_T; i += blockDim.x) {
        const int o = i % COUT_T, rest = i / COUT_T;
        w_s[i] = __ldg(a.w + (long long)rest * a.cout + cg * COUT_T + o);
    }
    __syncthreads();
    const int wo2 = a.Wo / 2;
    const long long n = (long long)a.B * a.Do * a.Ho * wo2;
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int ox = 2 * (int)(v % wo2), oy = (int)((v / wo2) % a.Ho);
    const int oz = (int)((v / ((long long)wo2 * a.Ho)) % a.Do), b = (int)(v / ((long long)wo2 * a.Ho * a.Do));
    float acc0[COUT_T], acc1[COUT_T];
#pragma unroll
    for (int o = 0; o < COUT_T; ++o) acc0[o] = acc1[o] = a.bias ? __ldg(a.bias + cg * COUT_T + o) : 0.f;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kz = 0; kz < a.kd; ++kz) {
        const int iz = oz * a.sd + kz - pz;
        if ((unsigned)iz >= (unsigned)a.Di) continue;
        for (int ky = 0; ky < a.k; ++ky) {
            const int iy = oy * a.s + ky - pad;
            if ((unsigned)iy >= (unsigned)a.Hi) continue;
            const float* rowp = a.x + (((long long)b * a.Di + iz) * a.Hi + iy) * (long long)a.Wi * CIN;
            for (int kx = 0; kx < a.k; ++kx) {
                const int ix0 = ox * a.s + kx - pad, ix1 = ix0 + a.s;
                const bool v0 = (unsigned)ix0 < (unsigned)a.Wi, v1 = (unsigned)ix1 < (unsigned)a.Wi;
                if (!v0 && !v1) continue;
                const float4* p0 = reinterpret_cast<const float4*>(rowp + (long long)ix0 * CIN);
                const float4* p1 = reinterpret_cast<const float4*>(rowp + (long long)ix1 * CIN);
                const float* wt = w_s + ((kz * a.k + ky) * a.k + kx) * CIN * COUT_T;
#pragma unroll
                for (int c4 = 0; c4 < CIN / 4; ++c4) {
                    const float4 t0 = v0 ? __ldg(p0 + c4) : zero, t1 = v1 ? __ldg(p1 + c4) : zero;
                    const float a0[4] = {t0.x, t0.y, t0.z, t0.w}, a1[4] = {t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float* wr = wt + (c4 * 4 + j) * COUT_T;
#pragma unroll
                        for (int o = 0; o < COUT_T; ++o) {
                            const float wv = wr[o];
                            acc0[o] = fmaf(a0[j], wv, acc0[o]);
                            acc1[o] = fmaf(a1[j], wv, acc1[o]);
                        }
                    }
                }
            }
        }
    }
    const long long vox = ((((long long)b * a.Do + oz) * a.Ho + oy) * a.Wo + ox);
    const long long off = vox * a.cout + cg * COUT_T;
#pragma unroll
    for (int o = 0; o < COUT_T; o += 4) {
        float4 r0 = make_float4(acc0[o], acc0[o + 1], acc0[o + 2], acc0[o + 3]);
        float4 r1 = make_float4(acc1[o], acc1[o + 1], acc1[o + 2], acc1[o + 3]);
        if (a.relu) {
            r0.x = fmaxf(r0.x, 0.f); r0.y = fmaxf(r0.y, 0.f); r0.z = fmaxf(r0.z, 0.f); r0.w = fmaxf(r0.w, 0.f);
            r1.x = fmaxf(r1.x, 0.f); r1.y = fmaxf(r1.y, 0.f); r1.z = fmaxf(r1.z, 0.f); r1.w = fmaxf(r1.w, 0.f);
        }
        if (a.skip) {  // added AFTER the ReLU (mvs4net_utils.py:907-909)
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(a.skip + off + o));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.skip + off + a.cout + o));
            r0.x += s0.x; r0.y += s0.y; r0.z += s0.z; r0.w += s0.w;
            r1.x += s1.x; r1.y += s1.y; r1.z += s1.z; r1.w += s1.w;
        }
        *reinterpret_cast<float4*>(a.y + off + o) = r0;
        *reinterpret_cast<float4*>(a.y + off + a.cout + o) = r1;
    }
}

// Specialisation for the regulariser's first layer (conv0: (1,3,3), stride 1, G -> 8 channels at the stage's full resolution,
// mvs4net_utils.py:875) - the one convolution that stays on the CUDA cores in the tensor-core engine.  conv_px2_kernel is
// instruction-bound there (~350 instructions per voxel at G = 4 for 48 bytes).  Here one thread owns four consecutive voxels
// of a row: each weight vector read from shared memory serves four voxels, the FMAs are packed two output channels per
// instruction (fma.rn.f32x2) with the input value as broadcast operand, the kernel extent is a compile-time 3 x 3.  Same
// accumulation order per output as conv_px2_kernel (ky, kx, c; out-of-image taps contribute fma(0, w, acc) = acc) - same bits.
// Opt-in (MVSTER_CONV0_PX4=1) until it has been timed.
namespace px4 {
#ifdef MVSTER_CPU_EMU
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) { return emu::pack(lo, hi); }
__device__ __forceinline__ float2 unpack2(unsigned long long v) { return emu::unpack(v); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    const float2 x = emu::unpack(a), y = emu::unpack(b), z = emu::unpack(c);
    return emu::pack(fmaf(x.x, y.x, z.x), fmaf(x.y, y.y, z.y));
}
#else
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
#endif
}  // namespace px4

// two fp32 -> the packed fp16 pairs (t1, t2) with v == t1 + 2^-11 t2 to 22 bits: the operand terms of the two-fp16-term tensor-core
// arithmetic (conv_tc3.cu split2h: same instructions, same bits)
#ifdef MVSTER_CPU_EMU
static inline unsigned emu_f2h(float f) {  // fp32 -> fp16 bits, round to nearest even, saturating to +-65504 (cvt.rn.satfinite.f16.f32)
    unsigned u;
    memcpy(&u, &f, 4);
    const unsigned sign = (u >> 16) & 0x8000u;
    u &= 0x7FFFFFFFu;
    if (u > 0x7F800000u) return sign | 0x7FFFu;                 // NaN
    if (u >= 0x477FF000u) return sign | 0x7BFFu;                // >= 65520 rounds past the largest finite value: saturate
    if (u < 0x33000001u) return sign;                           // below half of the smallest subnormal
    const int e = (int)(u >> 23) - 127;
    unsigned m = (u & 0x7FFFFFu) | 0x800000u;
    int shift = e >= -14 ? 13 : 13 + (-14 - e);                 // subnormal results shift further
    const unsigned half = 1u << (shift - 1), rest = m & ((1u << shift) - 1);
    m >>= shift;
    if (rest > half || (rest == half && (m & 1u))) ++m;
    const unsigned base = e >= -14 ? ((unsigned)(e + 15) << 10) - 0x400u : 0u;  // m carries the implicit bit for normals
    return sign | (base + m);
}
static inline float emu_h2f(unsigned h) {
    const int s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023;
    float v = e == 0 ? ldexpf((float)m, -24) : e == 31 ? (m ? NAN : INFINITY) : ldexpf((float)(m | 1024), e - 25);
    return s ? -v : v;
}
__device__ __forceinline__ void split2h_pair(float x, float y, unsigned& t1, unsigned& t2) {
    const unsigned hx = emu_f2h(x), hy = emu_f2h(y);
    t1 = hx | (hy << 16);
    const float rx = (x - emu_h2f(hx)) * 2048.f, ry = (y - emu_h2f(hy)) * 2048.f;
    t2 = emu_f2h(rx) | (emu_f2h(ry) << 16);
}
#else
__device__ __forceinline__ void split2h_pair(float x, float y, unsigned& t1, unsigned& t2) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(t1) : "f"(y), "f"(x));
    float bx, by;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(bx), "=f"(by) : "r"(t1));
    const float rx = (x - bx) * 2048.f, ry = (y - by) * 2048.f;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(t2) : "f"(ry), "f"(rx));
}
#endif

// BF (bf16 storage, mvster_reg2d_bf16): x holds bf16 voxels (the bf16 cost volume), `scale` the per-channel BatchNorm factor that
// stays out of the bf16-valued weights: y = relu(scale * conv(x, w) + bias), accumulators start at zero.
// BF output: out_bf16 = 1 writes y as bf16 NHWC (8 channels = 16 bytes per voxel: the packed-operand layout for C = 8), 0 writes
// fp32 holding the bf16-rounded values (a bf16 build stores this layer's output either way).
template <int CIN, bool BF = false>
__global__ void __launch_bounds__(128) conv0_px4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y,
                                                        long long NP, int H, int W, int relu, const float* __restrict__ scale = nullptr,
                                                        int out_bf16 = 0, unsigned* overflow = nullptr) {   // NP = B * D planes
    // !BF with out_bf16 = 2: y as the packed fp16 pair [NP][1 octet][a1 | a2][H][W][8] - the operand layout of the packed two-term
    // tensor-core layers (mvster_conv_tc3_pb16 with MVSTER_TC3_FP16X2); `overflow` as in mvster_tc3_set_overflow_flag
    using namespace px4;
    __shared__ __align__(16) float w_s[9 * CIN * 8 + 16];
    for (int i = threadIdx.x; i < 9 * CIN * 8; i += blockDim.x) w_s[i] = __ldg(w + i);
    if (threadIdx.x < 8) w_s[9 * CIN * 8 + threadIdx.x] = bias ? __ldg(bias + threadIdx.x) : 0.f;
    if (BF && threadIdx.x >= 8 && threadIdx.x < 16) w_s[9 * CIN * 8 + threadIdx.x] = scale ? __ldg(scale + threadIdx.x - 8) : 1.f;
    __syncthreads();
    const unsigned W4 = (unsigned)W >> 2;
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;                 // quads: < 2^31 (checked on the host)
    if (q >= (unsigned)NP * H * W4) return;
    const unsigned r = q / W4;                                                // row index over all planes
    const int x0 = (int)(q - r * W4) * 4, yy = (int)(r % (unsigned)H);
    const long long pl = r / (unsigned)H;
    unsigned long long acc[4][4];                                             // [voxel][output-channel pair]
    if constexpr (BF) {
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = pack2(0.f, 0.f);
    } else {
        const float4 b0 = *reinterpret_cast<const float4*>(w_s + 9 * CIN * 8), b1 = *reinterpret_cast<const float4*>(w_s + 9 * CIN * 8 + 4);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            acc[p][0] = pack2(b0.x, b0.y); acc[p][1] = pack2(b0.z, b0.w);
            acc[p][2] = pack2(b1.x, b1.y); acc[p][3] = pack2(b1.z, b1.w);
        }
    }
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = yy + ky - 1;
        if ((unsigned)iy >= (unsigned)H) continue;
        float4 t[6][CIN / 4];                                                 // voxels x0-1 .. x0+4
        if constexpr (BF) {  // 4 bf16 channels = 8 bytes; bf16 -> fp32 is a shift / mask
            const uint2* row = reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(x) + ((pl * H + iy) * W + x0) * CIN);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const bool ok = (i > 0 || x0 > 0) && (i < 5 || x0 + 4 < W);
#pragma unroll
                for (int c4 = 0; c4 < CIN / 4; ++c4) {
                    const uint2 u = ok ? __ldg(row + (i - 1) * (CIN / 4) + c4) : make_uint2(0u, 0u);
                    t[i][c4] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                                           __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
                }
            }
        } else {
            const float4* row = reinterpret_cast<const float4*>(x + ((pl * H + iy) * W + x0) * CIN);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const bool ok = (i > 0 || x0 > 0) && (i < 5 || x0 + 4 < W);
#pragma unroll
                for (int c4 = 0; c4 < CIN / 4; ++c4) t[i][c4] = ok ? __ldg(row + (i - 1) * (CIN / 4) + c4) : zero;
            }
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int c4 = 0; c4 < CIN / 4; ++c4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float* wr = w_s + ((ky * 3 + kx) * CIN + c4 * 4 + j) * 8;
                    const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
                    const unsigned long long wp[4] = {pack2(w0.x, w0.y), pack2(w0.z, w0.w), pack2(w1.x, w1.y), pack2(w1.z, w1.w)};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float4 tv = t[p + kx][c4];
                        const float v = j == 0 ? tv.x : j == 1 ? tv.y : j == 2 ? tv.z : tv.w;
                        const unsigned long long vv = pack2(v, v);
#pragma unroll
                        for (int o = 0; o < 4; ++o) acc[p][o] = fma2(vv, wp[o], acc[p][o]);
                    }
                }
            }
        }
    }
    float4* dst = reinterpret_cast<float4*>(y + ((pl * H + yy) * W + x0) * 8);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float2 a0 = unpack2(acc[p][0]), a1 = unpack2(acc[p][1]), a2 = unpack2(acc[p][2]), a3 = unpack2(acc[p][3]);
        float4 r0 = make_float4(a0.x, a0.y, a1.x, a1.y), r1 = make_float4(a2.x, a2.y, a3.x, a3.y);
        if constexpr (BF) {
            const float* bs = w_s + 9 * CIN * 8;  // [8] bias, [8] scale
            r0.x = fmaf(r0.x, bs[8], bs[0]); r0.y = fmaf(r0.y, bs[9], bs[1]); r0.z = fmaf(r0.z, bs[10], bs[2]); r0.w = fmaf(r0.w, bs[11], bs[3]);
            r1.x = fmaf(r1.x, bs[12], bs[4]); r1.y = fmaf(r1.y, bs[13], bs[5]); r1.z = fmaf(r1.z, bs[14], bs[6]); r1.w = fmaf(r1.w, bs[15], bs[7]);
        }
        if (relu) {
            r0.x = fmaxf(r0.x, 0.f); r0.y = fmaxf(r0.y, 0.f); r0.z = fmaxf(r0.z, 0.f); r0.w = fmaxf(r0.w, 0.f);
            r1.x = fmaxf(r1.x, 0.f); r1.y = fmaxf(r1.y, 0.f); r1.z = fmaxf(r1.z, 0.f); r1.w = fmaxf(r1.w, 0.f);
        }
        if constexpr (BF) {
            auto rn = [](float v) -> unsigned {  // fp32 -> bf16 bits, round to nearest even (finite values)
                const unsigned u = __float_as_uint(v);
                return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
            };
            const unsigned b0 = rn(r0.x) | (rn(r0.y) << 16), b1 = rn(r0.z) | (rn(r0.w) << 16);
            const unsigned b2 = rn(r1.x) | (rn(r1.y) << 16), b3 = rn(r1.z) | (rn(r1.w) << 16);
            if (out_bf16) {
                reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(y) + ((pl * H + yy) * W + x0) * 8)[p] = make_uint4(b0, b1, b2, b3);
                continue;
            }
            r0 = make_float4(__uint_as_float(b0 << 16), __uint_as_float(b0 & 0xFFFF0000u), __uint_as_float(b1 << 16), __uint_as_float(b1 & 0xFFFF0000u));
            r1 = make_float4(__uint_as_float(b2 << 16), __uint_as_float(b2 & 0xFFFF0000u), __uint_as_float(b3 << 16), __uint_as_float(b3 & 0xFFFF0000u));
        }
        if (!BF && out_bf16 == 2) {
            uint4 t1, t2;
            split2h_pair(r0.x, r0.y, t1.x, t2.x); split2h_pair(r0.z, r0.w, t1.y, t2.y);
            split2h_pair(r1.x, r1.y, t1.z, t2.z); split2h_pair(r1.z, r1.w, t1.w, t2.w);
            uint4* yp = reinterpret_cast<uint4*>(y) + ((pl * 2) * H + yy) * W + x0 + p;
            yp[0] = t1;
            yp[(long long)H * W] = t2;
            if (overflow) {
                const float m = fmaxf(fmaxf(fmaxf(fabsf(r0.x), fabsf(r0.y)), fmaxf(fabsf(r0.z), fabsf(r0.w))),
                                      fmaxf(fmaxf(fabsf(r1.x), fabsf(r1.y)), fmaxf(fabsf(r1.z), fabsf(r1.w))));
                const bool nan = r0.x != r0.x || r0.y != r0.y || r0.z != r0.z || r0.w != r0.w || r1.x != r1.x || r1.y != r1.y || r1.z != r1.z || r1.w != r1.w;
                if (nan || !(m < 65504.f)) atomicOr(overflow, 1u);
            }
            continue;
        }
        dst[2 * p] = r0;
        dst[2 * p + 1] = r1;
    }
}

template <int CIN, int COUT_T>
static int launch_px2(const ConvPxArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)a.kd * a.k * a.k * CIN * COUT_T * sizeof(float);
    auto k = conv_px2_kernel<CIN, COUT_T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long n = (long long)a.B * a.Do * a.Ho * (a.Wo / 2);
    k<<<dim3(ceil_div(n, 128), a.cout / COUT_T), 128, smem, st>>>(a);
    return check_launch("conv_px2_kernel");
}

template <int CIN>
static int dispatch_px2(const ConvPxArgs& a, cudaStream_t st) {
    if (a.cout % 16 == 0) return launch_px2<CIN, 16>(a, st);
    return launch_px2<CIN, 8>(a, st);
}

// Returns MVSTER_ERR_UNSUPPORTED_SHAPE (-100) when the two-pixel kernel does not cover the layer (the caller then
// uses its one-pixel kernel), otherwise the launch status.
int conv_px2(const float* x, const float* w, const float* bias, const float* skip, float* y,
             int B, int Di, int Hi, int Wi, int Cin, int Cout, int kd, int k, int sd, int s, int relu, cudaStream_t st) {
    ConvPxArgs a;
    a.x = x; a.w = w; a.bias = bias; a.skip = skip; a.y = y;
    a.B = B; a.Di = Di; a.Hi = Hi; a.Wi = Wi;
    a.Do = (Di - 1) / sd + 1; a.Ho = (Hi - 1) / s + 1; a.Wo = (Wi - 1) / s + 1;
    a.cout = Cout; a.kd = kd; a.k = k; a.sd = sd; a.s = s; a.relu = relu;
    const char* sw = getenv("MVSTER_CONV_PX2");  // "0" forces the one-pixel kernels (A/B measurements)
    if (sw && sw[0] == '0') return -100;
    const char* v4 = getenv("MVSTER_CONV0_PX4");  // default on (B200: 45.1 -> 38.9 us at stage 4, same bits; profiles/r02_glue_ab.md); "0" = two-pixel kernel
    if ((!v4 || v4[0] == '1') && kd == 1 && k == 3 && sd == 1 && s == 1 && Cout == 8 && !skip && Wi % 4 == 0 && (Cin == 4 || Cin == 8) &&
        (long long)B * Di * Hi * (Wi / 4) < (1ll << 31)) {
        const long long NP = (long long)B * Di, n = NP * Hi * (Wi / 4);
        if (Cin == 4) conv0_px4_kernel<4><<<ceil_div(n, 128), 128, 0, st>>>(x, w, bias, y, NP, Hi, Wi, relu);
        else conv0_px4_kernel<8><<<ceil_div(n, 128), 128, 0, st>>>(x, w, bias, y, NP, Hi, Wi, relu);
        return check_launch("conv0_px4_kernel");
    }
    if (a.Wo % 2 || Cout % 8 || (size_t)kd * k * k * Cin * (Cout % 16 == 0 ? 16 : 8) * 4 > 200 * 1024) return -100;
    switch (Cin) {
        case 4: return dispatch_px2<4>(a, st);
        case 8: return dispatch_px2<8>(a, st);
        case 16: return dispatch_px2<16>(a, st);
        case 32: return dispatch_px2<32>(a, st);
        case 64: return dispatch_px2<64>(a, st);
    }
    return -100;
}

// conv0 of the bf16-storage regulariser: x = the bf16 cost volume [NP][H][W][Cin], w [9][Cin][8] (bf16-valued fp32); y = the
// rounded output as bf16 NHWC (out_bf16) or as fp32
int conv0_bf16(const void* x, const float* w, const float* scale, const float* bias, void* y, int out_bf16, long long NP, int H, int W, int Cin,
               cudaStream_t st) {
    if (W % 4 || (Cin != 4 && Cin != 8) || NP * H * (W / 4) >= (1ll << 31)) return -100;
    const long long n = NP * H * (W / 4);
    if (Cin == 4) conv0_px4_kernel<4, true><<<ceil_div(n, 128), 128, 0, st>>>((const float*)x, w, bias, (float*)y, NP, H, W, 1, scale, out_bf16);
    else conv0_px4_kernel<8, true><<<ceil_div(n, 128), 128, 0, st>>>((const float*)x, w, bias, (float*)y, NP, H, W, 1, scale, out_bf16);
    return check_launch("conv0_px4_kernel[bf16]");
}

// conv0 of the fp32 regulariser writing its output as the packed fp16 pair (x fp32 cost volume, folded weights): -100 = not covered
int conv0_packed_h16(const float* x, const float* w, const float* bias, void* y, unsigned* overflow, long long NP, int H, int W, int Cin,
                     cudaStream_t st) {
    if (W % 4 || (Cin != 4 && Cin != 8) || NP * H * (W / 4) >= (1ll << 31)) return -100;
    const long long n = NP * H * (W / 4);
    if (Cin == 4) conv0_px4_kernel<4><<<ceil_div(n, 128), 128, 0, st>>>(x, w, bias, (float*)y, NP, H, W, 1, nullptr, 2, overflow);
    else conv0_px4_kernel<8><<<ceil_div(n, 128), 128, 0, st>>>(x, w, bias, (float*)y, NP, H, W, 1, nullptr, 2, overflow);
    return check_launch("conv0_px4_kernel[packed fp16 pair]");
}

}  // namespace mvster
