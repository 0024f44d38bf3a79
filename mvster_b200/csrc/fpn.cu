// Feature-pyramid (FPN4, models/mvs4net_utils.py:419-502) building blocks on channels-last tensors.
// FPN4 is outside the named hot path (SURVEY.md 8f "next" #1) but it feeds it and, through cuDNN fp32,
// was 75 % of the step time; these kernels let the whole forward run inside libmvster_b200:
//   conv_first_kernel   3x3 conv on the 3-channel NCHW image -> NHWC 8 channels (+ folded BN, ReLU)   (:424, conv0.0)
//   conv2d_kernel       k x k (k = 1, 3, 5), stride 1 / 2, CUDA-core fp32, folded BN / bias / ReLU      (:425-445, :452-459)
//   fpn_merge_kernel    top-down merge: bilinear x2 (align_corners) of the coarser map + 1x1 lateral conv (:479-486)
// The 3x3 stride-1 layers with Cin >= 16 (73 % of FPN4's FLOPs) run on the tcgen05 kernel of conv_tc2.cu.
#include "common.cuh"
#include <stdlib.h>

namespace mvster {

struct Conv2dArgs {
    const float* x; const float* w; const float* bias; float* y;
    int N, Hi, Wi, Ho, Wo, cout, k, s, relu;
};

template <int CIN, int COUT_T>
__global__ void __launch_bounds__(128) conv2d_kernel(const Conv2dArgs a) {
    extern __shared__ __align__(16) float w_s[];  // [k*k][CIN][COUT_T]
    const int cg = blockIdx.y, taps = a.k * a.k, pad = a.k / 2;
    for (int i = threadIdx.x; i < taps * CIN * COUT_T; i += blockDim.x) {
        const int o = i % COUT_T, rest = i / COUT_T;
        w_s[i] = __ldg(a.w + (long long)rest * a.cout + cg * COUT_T + o);
    }
    __syncthreads();
    const long long n = (long long)a.N * a.Ho * a.Wo;
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int ox = (int)(v % a.Wo), oy = (int)((v / a.Wo) % a.Ho), b = (int)(v / ((long long)a.Wo * a.Ho));
    float acc[COUT_T];
#pragma unroll
    for (int o = 0; o < COUT_T; ++o) acc[o] = a.bias ? __ldg(a.bias + cg * COUT_T + o) : 0.f;
    for (int ky = 0; ky < a.k; ++ky) {
        const int iy = oy * a.s + ky - pad;
        if ((unsigned)iy >= (unsigned)a.Hi) continue;
        for (int kx = 0; kx < a.k; ++kx) {
            const int ix = ox * a.s + kx - pad;
            if ((unsigned)ix >= (unsigned)a.Wi) continue;
            const float4* px = reinterpret_cast<const float4*>(a.x + (((long long)b * a.Hi + iy) * a.Wi + ix) * CIN);
            const float* wt = w_s + (ky * a.k + kx) * CIN * COUT_T;
#pragma unroll
            for (int c4 = 0; c4 < CIN / 4; ++c4) {
                const float4 t = __ldg(px + c4);
                const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float* wr = wt + (c4 * 4 + j) * COUT_T;
#pragma unroll
                    for (int o = 0; o < COUT_T; ++o) acc[o] = fmaf(tv[j], wr[o], acc[o]);
                }
            }
        }
    }
    const long long off = v * a.cout + cg * COUT_T;
#pragma unroll
    for (int o = 0; o < COUT_T; o += 4) {
        float4 r = make_float4(acc[o], acc[o + 1], acc[o + 2], acc[o + 3]);
        if (a.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
        *reinterpret_cast<float4*>(a.y + off + o) = r;
    }
}

template <int CIN, int COUT_T>
static int launch_conv2d(const Conv2dArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)a.k * a.k * CIN * COUT_T * sizeof(float);
    auto k = conv2d_kernel<CIN, COUT_T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long n = (long long)a.N * a.Ho * a.Wo;
    k<<<dim3(ceil_div(n, 128), a.cout / COUT_T), 128, smem, st>>>(a);
    return check_launch("conv2d_kernel");
}

template <int CIN>
static int dispatch_conv2d(const Conv2dArgs& a, cudaStream_t st) {
    if (a.cout % 16 == 0) return launch_conv2d<CIN, 16>(a, st);
    if (a.cout % 8 == 0) return launch_conv2d<CIN, 8>(a, st);
    set_error("mvster_conv2d_nhwc_f32: Cout=%d must be a multiple of 8", a.cout);
    return MVSTER_ERR_ARG;
}

// 3-channel NCHW image -> 8-channel NHWC feature: 3x3, stride 1, pad 1, bias (folded BN), ReLU.
// w: [9][3][8].  One thread per output pixel; the image is read plane-wise (coalesced along x).
__global__ void __launch_bounds__(128) conv_first_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, int N, int H, int W) {
    __shared__ float w_s[9 * 3 * 8 + 8];
    for (int i = threadIdx.x; i < 216; i += blockDim.x) w_s[i] = __ldg(w + i);
    if (threadIdx.x < 8) w_s[216 + threadIdx.x] = __ldg(bias + threadIdx.x);
    __syncthreads();
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= (long long)N * H * W) return;
    const int x = (int)(v % W), yy = (int)((v / W) % H), b = (int)(v / ((long long)W * H));
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = w_s[216 + o];
    const float* base = img + (long long)b * 3 * H * W;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = yy + ky - 1;
        if ((unsigned)iy >= (unsigned)H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = x + kx - 1;
            if ((unsigned)ix >= (unsigned)W) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float t = __ldg(base + ((long long)c * H + iy) * W + ix);
                const float* wr = w_s + ((ky * 3 + kx) * 3 + c) * 8;
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] = fmaf(t, wr[o], acc[o]);
            }
        }
    }
    float4* dst = reinterpret_cast<float4*>(y + v * 8);
    dst[0] = make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
    dst[1] = make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
}

// conv_first_kernel is instruction-bound (per pixel 216 FMA + 54 LDS.128 of weights + 27 loads ~ 330 instructions for
// 44 bytes of traffic).  Variant: one thread = four consecutive pixels of a row.  A weight vector read from shared memory
// serves four pixels, the three image rows arrive as 128-bit loads plus the two edge columns, and the FMAs are packed
// two output channels per instruction (fma.rn.f32x2): ~135 instructions per pixel.  Same accumulation order per output
// channel as conv_first_kernel (taps in ky, kx, c order; out-of-image taps contribute fma(0, w, acc) = acc), hence the
// same bits.  Requires W % 4 == 0.  Opt-in (MVSTER_CONV_FIRST=2) until it has been timed.
namespace cf4 {
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
}  // namespace cf4
namespace cf4 {
#ifdef MVSTER_CPU_EMU
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    const float2 x = emu::unpack(a), y = emu::unpack(b);
    return emu::pack(x.x * y.x, x.y * y.y);
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    const float2 x = emu::unpack(a), y = emu::unpack(b);
    return emu::pack(x.x + y.x, x.y + y.y);
}
#else
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
#endif
}  // namespace cf4

__global__ void __launch_bounds__(128) conv_first4_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ y, int N, int H, int W) {
    using namespace cf4;
    __shared__ __align__(16) float w_s[9 * 3 * 8 + 8];
    for (int i = threadIdx.x; i < 216; i += blockDim.x) w_s[i] = __ldg(w + i);
    if (threadIdx.x < 8) w_s[216 + threadIdx.x] = __ldg(bias + threadIdx.x);
    __syncthreads();
    const unsigned W4 = (unsigned)W >> 2;
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;                  // quad of pixels: < 2^31 (checked on the host)
    if (q >= (unsigned)N * H * W4) return;
    const unsigned r = q / W4;
    const int x0 = (int)(q - r * W4) * 4, yy = (int)(r % (unsigned)H), b = (int)(r / (unsigned)H);
    unsigned long long acc[4][4];                                             // [pixel][channel pair]
    {
        const float4 b0 = *reinterpret_cast<const float4*>(w_s + 216), b1 = *reinterpret_cast<const float4*>(w_s + 220);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            acc[p][0] = pack2(b0.x, b0.y); acc[p][1] = pack2(b0.z, b0.w);
            acc[p][2] = pack2(b1.x, b1.y); acc[p][3] = pack2(b1.z, b1.w);
        }
    }
    const float* base = img + (long long)b * 3 * H * W;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = yy + ky - 1;
        if ((unsigned)iy >= (unsigned)H) continue;
        float t[3][6];                                                        // [plane][x0-1 .. x0+4]
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* row = base + ((long long)c * H + iy) * W + x0;
            const float4 m = __ldg(reinterpret_cast<const float4*>(row));
            t[c][0] = x0 > 0 ? __ldg(row - 1) : 0.f;
            t[c][1] = m.x; t[c][2] = m.y; t[c][3] = m.z; t[c][4] = m.w;
            t[c][5] = x0 + 4 < W ? __ldg(row + 4) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* wr = w_s + ((ky * 3 + kx) * 3 + c) * 8;
                const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
                const unsigned long long wp[4] = {pack2(w0.x, w0.y), pack2(w0.z, w0.w), pack2(w1.x, w1.y), pack2(w1.z, w1.w)};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float v = t[c][p + kx];
                    const unsigned long long vv = pack2(v, v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[p][j] = fma2(vv, wp[j], acc[p][j]);
                }
            }
        }
    }
    float4* dst = reinterpret_cast<float4*>(y + (((long long)b * H + yy) * W + x0) * 8);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float2 a0 = unpack2(acc[p][0]), a1 = unpack2(acc[p][1]), a2 = unpack2(acc[p][2]), a3 = unpack2(acc[p][3]);
        dst[2 * p] = make_float4(fmaxf(a0.x, 0.f), fmaxf(a0.y, 0.f), fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f));
        dst[2 * p + 1] = make_float4(fmaxf(a2.x, 0.f), fmaxf(a2.y, 0.f), fmaxf(a3.x, 0.f), fmaxf(a3.y, 0.f));
    }
}

// out[n,y,x,:] = bilinear_x2(top)[n,y,x,:] + W_lat . lat[n,y,x,:] + bias     (F.interpolate(scale_factor=2,
// align_corners=True) + 1x1 lateral conv, mvs4net_utils.py:479-486).  top [N][H/2][W/2][64], lat [N][H][W][CL],
// w [CL][64].  One thread = one pixel x 4 output channels (16 lanes per pixel).
template <int CL>
__global__ void __launch_bounds__(128) fpn_merge_kernel(const float* __restrict__ top, const float* __restrict__ lat,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        float* __restrict__ out, int N, int H, int W) {
    __shared__ __align__(16) float w_s[CL * 64 + 64];
    for (int i = threadIdx.x; i < CL * 64; i += blockDim.x) w_s[i] = __ldg(w + i);
    if (threadIdx.x < 64) w_s[CL * 64 + threadIdx.x] = __ldg(bias + threadIdx.x);
    __syncthreads();
    // 16 consecutive lanes own one pixel, lane q its channels 4q..4q+3: every load of `top` and every store of `out`
    // is a fully coalesced 256-byte run per pixel (the 64-channel full-resolution map is the largest tensor of FPN4).
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long v = t >> 4;
    const int q = (int)(t & 15);
    if (v >= (long long)N * H * W) return;
    const int x = (int)(v % W), y = (int)((v / W) % H), b = (int)(v / ((long long)W * H));
    const int Hc = H / 2, Wc = W / 2;
    const float sy = Hc > 1 ? __fdiv_rn((float)(Hc - 1), (float)(H - 1)) : 0.f, sx = Wc > 1 ? __fdiv_rn((float)(Wc - 1), (float)(W - 1)) : 0.f;
    const float fy = __fmul_rn(sy, (float)y), fx = __fmul_rn(sx, (float)x);
    const int y0 = min((int)floorf(fy), Hc - 1), x0 = min((int)floorf(fx), Wc - 1);
    const int y1 = y0 + (y0 < Hc - 1), x1 = x0 + (x0 < Wc - 1);
    const float ly1 = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), lx1 = fminf(fmaxf(fx - (float)x0, 0.f), 1.f);
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float* tb = top + (long long)b * Hc * Wc * 64 + q * 4;
    const float4 a00 = __ldg(reinterpret_cast<const float4*>(tb + ((long long)y0 * Wc + x0) * 64));
    const float4 a01 = __ldg(reinterpret_cast<const float4*>(tb + ((long long)y0 * Wc + x1) * 64));
    const float4 a10 = __ldg(reinterpret_cast<const float4*>(tb + ((long long)y1 * Wc + x0) * 64));
    const float4 a11 = __ldg(reinterpret_cast<const float4*>(tb + ((long long)y1 * Wc + x1) * 64));
    // ATen order: ly0*(lx0*v00 + lx1*v01) + ly1*(lx0*v10 + lx1*v11)
    float4 up;
    up.x = ly0 * (lx0 * a00.x + lx1 * a01.x) + ly1 * (lx0 * a10.x + lx1 * a11.x);
    up.y = ly0 * (lx0 * a00.y + lx1 * a01.y) + ly1 * (lx0 * a10.y + lx1 * a11.y);
    up.z = ly0 * (lx0 * a00.z + lx1 * a01.z) + ly1 * (lx0 * a10.z + lx1 * a11.z);
    up.w = ly0 * (lx0 * a00.w + lx1 * a01.w) + ly1 * (lx0 * a10.w + lx1 * a11.w);
    float4 lv = *reinterpret_cast<const float4*>(w_s + CL * 64 + q * 4);  // lateral 1x1 conv: bias + sum_c lat[c] * w[c][4q..4q+3]
    const float4* pl = reinterpret_cast<const float4*>(lat + v * CL);
#pragma unroll
    for (int c4 = 0; c4 < CL / 4; ++c4) {
        const float4 tt = __ldg(pl + c4);
        const float tv[4] = {tt.x, tt.y, tt.z, tt.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 wr = *reinterpret_cast<const float4*>(w_s + (c4 * 4 + j) * 64 + q * 4);
            lv.x = fmaf(tv[j], wr.x, lv.x); lv.y = fmaf(tv[j], wr.y, lv.y);
            lv.z = fmaf(tv[j], wr.z, lv.z); lv.w = fmaf(tv[j], wr.w, lv.w);
        }
    }
    *reinterpret_cast<float4*>(out + v * 64 + q * 4) = make_float4(up.x + lv.x, up.y + lv.y, up.z + lv.z, up.w + lv.w);
}

// Variant 2 of the merge (opt-in, MVSTER_FPN_MERGE=2; not timed yet): the same 16 lanes per pixel, but every lane keeps its
// 4 output channels for FOUR consecutive pixels of a row, so each weight row is read from shared memory once per 4 pixels and
// 16 independent loads of `top` are in flight per lane.  Per pixel the arithmetic is the v1 expression term for term (same bits).
template <int CL>
__global__ void __launch_bounds__(128) fpn_merge4_kernel(const float* __restrict__ top, const float* __restrict__ lat,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ out, int N, int H, int W) {
    __shared__ __align__(16) float w_s[CL * 64 + 64];
    for (int i = threadIdx.x; i < CL * 64; i += blockDim.x) w_s[i] = __ldg(w + i);
    if (threadIdx.x < 64) w_s[CL * 64 + threadIdx.x] = __ldg(bias + threadIdx.x);
    __syncthreads();
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;   // < 2^31 (checked on the host): 32-bit index arithmetic
    const int q = (int)(t & 15);
    const unsigned g = t >> 4;
    const unsigned gpr = (unsigned)(W + 3) / 4;  // pixel groups per row
    if (g >= (unsigned)N * H * gpr) return;
    const unsigned rr = g / gpr;
    const int xg = (int)(g - rr * gpr), y = (int)(rr % (unsigned)H), b = (int)(rr / (unsigned)H);
    const int Hc = H / 2, Wc = W / 2;
    const float sy = Hc > 1 ? __fdiv_rn((float)(Hc - 1), (float)(H - 1)) : 0.f, sx = Wc > 1 ? __fdiv_rn((float)(Wc - 1), (float)(W - 1)) : 0.f;
    const float fy = __fmul_rn(sy, (float)y);
    const int y0 = min((int)floorf(fy), Hc - 1);
    const int y1 = y0 + (y0 < Hc - 1);
    const float ly1 = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), ly0 = 1.f - ly1;
    const float* tb = top + (long long)b * Hc * Wc * 64 + q * 4;
    float4 up[4], lv[4];
    const float4 b4 = *reinterpret_cast<const float4*>(w_s + CL * 64 + q * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = min(xg * 4 + j, W - 1);  // a ragged last group repeats the last pixel (never stored)
        const float fx = __fmul_rn(sx, (float)x);
        const int x0 = min((int)floorf(fx), Wc - 1);
        const int x1 = x0 + (x0 < Wc - 1);
        const float lx1 = fminf(fmaxf(fx - (float)x0, 0.f), 1.f), lx0 = 1.f - lx1;
        const float4 a00 = __ldg(reinterpret_cast<const float4*>(tb + ((long long)y0 * Wc + x0) * 64));
        const float4 a01 = __ldg(reinterpret_cast<const float4*>(tb + ((long long)y0 * Wc + x1) * 64));
        const float4 a10 = __ldg(reinterpret_cast<const float4*>(tb + ((long long)y1 * Wc + x0) * 64));
        const float4 a11 = __ldg(reinterpret_cast<const float4*>(tb + ((long long)y1 * Wc + x1) * 64));
        up[j].x = ly0 * (lx0 * a00.x + lx1 * a01.x) + ly1 * (lx0 * a10.x + lx1 * a11.x);
        up[j].y = ly0 * (lx0 * a00.y + lx1 * a01.y) + ly1 * (lx0 * a10.y + lx1 * a11.y);
        up[j].z = ly0 * (lx0 * a00.z + lx1 * a01.z) + ly1 * (lx0 * a10.z + lx1 * a11.z);
        up[j].w = ly0 * (lx0 * a00.w + lx1 * a01.w) + ly1 * (lx0 * a10.w + lx1 * a11.w);
        lv[j] = b4;
    }
    const long long row = ((long long)b * H + y) * W;
#pragma unroll
    for (int c4 = 0; c4 < CL / 4; ++c4) {
        float tv[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 tt = __ldg(reinterpret_cast<const float4*>(lat + (row + min(xg * 4 + j, W - 1)) * CL) + c4);
            tv[j][0] = tt.x; tv[j][1] = tt.y; tv[j][2] = tt.z; tv[j][3] = tt.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float4 wr = *reinterpret_cast<const float4*>(w_s + (c4 * 4 + e) * 64 + q * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                lv[j].x = fmaf(tv[j][e], wr.x, lv[j].x); lv[j].y = fmaf(tv[j][e], wr.y, lv[j].y);
                lv[j].z = fmaf(tv[j][e], wr.z, lv[j].z); lv[j].w = fmaf(tv[j][e], wr.w, lv[j].w);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (xg * 4 + j < W)
            *reinterpret_cast<float4*>(out + (row + xg * 4 + j) * 64 + q * 4) =
                make_float4(up[j].x + lv[j].x, up[j].y + lv[j].y, up[j].z + lv[j].z, up[j].w + lv[j].w);
}

// Variant 3 of the merge (opt-in, MVSTER_FPN_MERGE=3; not timed yet), built from the SASS counts: with 16 lanes per pixel every
// lane repeats the pixel's coordinate arithmetic for 4 output channels (v1: ~700 instructions per lane = 11 000 per pixel;
// variant 2: ~3600 per pixel).  Here FOUR lanes own a group of four consecutive pixels and lane q the 16 channels
// {16 i + 4 q + k}: the coordinates are computed once per 16 channels, a weight row read from shared memory (four conflict-free
// 128-bit reads per lane) serves four pixels, and all arithmetic is packed two channels per instruction - ~1100 instructions
// per pixel, which leaves the kernel to its 45.9 / 157 MB of traffic.  The accumulator starts at up + bias (variants 1/2: up +
// (bias + sum)); equal to rounding, checked on the CPU emulation against variant 1 and the float64 definition.
template <int CL>
__global__ void __launch_bounds__(128) fpn_merge5_kernel(const float* __restrict__ top, const float* __restrict__ lat,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ out, int N, int H, int W) {
    using namespace cf4;
    __shared__ __align__(16) float w_s[CL * 64 + 64];
    for (int i = threadIdx.x; i < CL * 64; i += blockDim.x) w_s[i] = __ldg(w + i);
    if (threadIdx.x < 64) w_s[CL * 64 + threadIdx.x] = __ldg(bias + threadIdx.x);
    __syncthreads();
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;   // < 2^31 (checked on the host)
    const int q = (int)(t & 3);
    const unsigned g = t >> 2;
    const unsigned gpr = (unsigned)(W + 3) / 4;                 // pixel groups per row
    if (g >= (unsigned)N * H * gpr) return;
    const unsigned rr = g / gpr;
    const int xg = (int)(g - rr * gpr), y = (int)(rr % (unsigned)H), b = (int)(rr / (unsigned)H);
    const int Hc = H / 2, Wc = W / 2;
    const float sy = Hc > 1 ? __fdiv_rn((float)(Hc - 1), (float)(H - 1)) : 0.f, sx = Wc > 1 ? __fdiv_rn((float)(Wc - 1), (float)(W - 1)) : 0.f;
    const float fy = __fmul_rn(sy, (float)y);
    const int y0 = min((int)floorf(fy), Hc - 1);
    const int y1 = y0 + (y0 < Hc - 1);
    const float ly1 = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), ly0 = 1.f - ly1;
    const unsigned long long ly0v = pack2(ly0, ly0), ly1v = pack2(ly1, ly1);
    const float* tb = top + (long long)b * Hc * Wc * 64 + q * 4;
    unsigned long long acc[4][8];                               // [pixel][chunk i][pair]: channels 16 i + 4 q + {0,1 | 2,3}
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = min(xg * 4 + j, W - 1);                   // a ragged last group repeats the last pixel (never stored)
        const float fx = __fmul_rn(sx, (float)x);
        const int x0 = min((int)floorf(fx), Wc - 1);
        const int x1 = x0 + (x0 < Wc - 1);
        const float lx1 = fminf(fmaxf(fx - (float)x0, 0.f), 1.f), lx0 = 1.f - lx1;
        const unsigned long long lx0v = pack2(lx0, lx0), lx1v = pack2(lx1, lx1);
        const float* p00 = tb + ((long long)y0 * Wc + x0) * 64;
        const float* p01 = tb + ((long long)y0 * Wc + x1) * 64;
        const float* p10 = tb + ((long long)y1 * Wc + x0) * 64;
        const float* p11 = tb + ((long long)y1 * Wc + x1) * 64;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 a00 = __ldg(reinterpret_cast<const float4*>(p00 + 16 * i)), a01 = __ldg(reinterpret_cast<const float4*>(p01 + 16 * i));
            const float4 a10 = __ldg(reinterpret_cast<const float4*>(p10 + 16 * i)), a11 = __ldg(reinterpret_cast<const float4*>(p11 + 16 * i));
            const float4 bb = *reinterpret_cast<const float4*>(w_s + CL * 64 + 16 * i + q * 4);
            // ATen order: ly0*(lx0*v00 + lx1*v01) + ly1*(lx0*v10 + lx1*v11)
            const unsigned long long t_lo = fma2(lx1v, pack2(a01.x, a01.y), mul2(lx0v, pack2(a00.x, a00.y)));
            const unsigned long long b_lo = fma2(lx1v, pack2(a11.x, a11.y), mul2(lx0v, pack2(a10.x, a10.y)));
            const unsigned long long t_hi = fma2(lx1v, pack2(a01.z, a01.w), mul2(lx0v, pack2(a00.z, a00.w)));
            const unsigned long long b_hi = fma2(lx1v, pack2(a11.z, a11.w), mul2(lx0v, pack2(a10.z, a10.w)));
            acc[j][2 * i] = add2(fma2(ly1v, b_lo, mul2(ly0v, t_lo)), pack2(bb.x, bb.y));
            acc[j][2 * i + 1] = add2(fma2(ly1v, b_hi, mul2(ly0v, t_hi)), pack2(bb.z, bb.w));
        }
    }
    const long long row = ((long long)b * H + y) * W;
#pragma unroll
    for (int c4 = 0; c4 < CL / 4; ++c4) {
        float tv[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 tt = __ldg(reinterpret_cast<const float4*>(lat + (row + min(xg * 4 + j, W - 1)) * CL) + c4);
            tv[j][0] = tt.x; tv[j][1] = tt.y; tv[j][2] = tt.z; tv[j][3] = tt.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            unsigned long long wp[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 wr = *reinterpret_cast<const float4*>(w_s + (c4 * 4 + e) * 64 + 16 * i + q * 4);
                wp[2 * i] = pack2(wr.x, wr.y); wp[2 * i + 1] = pack2(wr.z, wr.w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long vv = pack2(tv[j][e], tv[j][e]);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[j][k] = fma2(vv, wp[k], acc[j][k]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (xg * 4 + j >= W) break;
        float* dst = out + (row + xg * 4 + j) * 64 + q * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 lo = unpack2(acc[j][2 * i]), hi = unpack2(acc[j][2 * i + 1]);
            *reinterpret_cast<float4*>(dst + 16 * i) = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_conv2d_nhwc_f32(const float* x, const float* w, const float* bias, float* y,
                                      int N, int H, int W, int Cin, int Cout, int k, int stride, int relu, mvster_stream_t stream) {
    MVSTER_REQUIRE(x && w && y, "mvster_conv2d_nhwc_f32: null pointer");
    MVSTER_REQUIRE(N > 0 && H > 0 && W > 0, "mvster_conv2d_nhwc_f32: bad shape");
    MVSTER_REQUIRE(k == 1 || k == 3 || k == 5, "mvster_conv2d_nhwc_f32: kernel size %d (1, 3 or 5)", k);
    MVSTER_REQUIRE(stride == 1 || stride == 2, "mvster_conv2d_nhwc_f32: stride %d (1 or 2)", stride);
    Conv2dArgs a{x, w, bias, y, N, H, W, (H - 1) / stride + 1, (W - 1) / stride + 1, Cout, k, stride, relu};
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = conv_px2(x, w, bias, nullptr, y, N, 1, H, W, Cin, Cout, 1, k, 1, stride, relu, st);
    if (rc != -100) return rc;  // two-pixels-per-thread kernel covered the layer
    switch (Cin) {
        case 8: return dispatch_conv2d<8>(a, st);
        case 16: return dispatch_conv2d<16>(a, st);
        case 32: return dispatch_conv2d<32>(a, st);
        case 64: return dispatch_conv2d<64>(a, st);
    }
    set_error("mvster_conv2d_nhwc_f32: unsupported Cin=%d (8,16,32,64)", Cin);
    return MVSTER_ERR_ARG;
}

extern "C" int mvster_conv_first_f32(const float* img_nchw, const float* w, const float* bias, float* y,
                                     int N, int H, int W, mvster_stream_t stream) {
    MVSTER_REQUIRE(img_nchw && w && bias && y, "mvster_conv_first_f32: null pointer");
    MVSTER_REQUIRE(N > 0 && H > 0 && W > 0, "mvster_conv_first_f32: bad shape");
    const long long n = (long long)N * H * W;
    // default = the four-pixel kernel (B200, 5 x 512 x 640: 36.9 us vs 51.2 us, same bits - profiles/r02_glue_ab.md); MVSTER_CONV_FIRST=1 = v1
    const char* variant = getenv("MVSTER_CONV_FIRST");
    if ((!variant || atoi(variant) == 2) && W % 4 == 0 && n / 4 < (1ll << 31)) {  // four pixels per thread, packed FMAs
        conv_first4_kernel<<<ceil_div(n / 4, 128), 128, 0, (cudaStream_t)stream>>>(img_nchw, w, bias, y, N, H, W);
        return check_launch("conv_first4_kernel");
    }
    conv_first_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(img_nchw, w, bias, y, N, H, W);
    return check_launch("conv_first_kernel");
}

extern "C" int mvster_fpn_merge_f32(const float* top, const float* lateral, const float* w, const float* bias, float* out,
                                    int N, int H, int W, int Clat, mvster_stream_t stream) {
    MVSTER_REQUIRE(top && lateral && w && bias && out, "mvster_fpn_merge_f32: null pointer");
    MVSTER_REQUIRE(N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "mvster_fpn_merge_f32: H,W must be even");
    cudaStream_t st = (cudaStream_t)stream;
    // measured on B200 (profiles/r02_glue_ab.md): 16 lateral channels 106.5 (v1) / 65.6 (v2) / 69.6 us (v3); 32 lateral channels
    // 51.2 / 32.8 / 30.7 us -> default v2, v3 for the 32-channel level; MVSTER_FPN_MERGE=1|2|3 forces one (A/B)
    const char* env_v = getenv("MVSTER_FPN_MERGE");
    const int mv = env_v ? atoi(env_v) : (Clat == 32 ? 3 : 2);
    if (mv == 2 && (long long)N * H * ((W + 3) / 4) * 16 < (1ll << 31)) {  // four pixels per lane
        const long long n4 = (long long)N * H * ((W + 3) / 4) * 16;
        dim3 grid4(ceil_div(n4, 128));
        if (Clat == 8) fpn_merge4_kernel<8><<<grid4, 128, 0, st>>>(top, lateral, w, bias, out, N, H, W);
        else if (Clat == 16) fpn_merge4_kernel<16><<<grid4, 128, 0, st>>>(top, lateral, w, bias, out, N, H, W);
        else if (Clat == 32) fpn_merge4_kernel<32><<<grid4, 128, 0, st>>>(top, lateral, w, bias, out, N, H, W);
        else MVSTER_REQUIRE(false, "mvster_fpn_merge_f32: unsupported lateral channels %d (8,16,32)", Clat);
        return check_launch("fpn_merge4_kernel");
    }
    if (mv == 3 && (long long)N * H * ((W + 3) / 4) * 4 < (1ll << 31)) {  // four lanes per four pixels, packed FMAs
        const long long n5 = (long long)N * H * ((W + 3) / 4) * 4;
        dim3 grid5(ceil_div(n5, 128));
        if (Clat == 8) fpn_merge5_kernel<8><<<grid5, 128, 0, st>>>(top, lateral, w, bias, out, N, H, W);
        else if (Clat == 16) fpn_merge5_kernel<16><<<grid5, 128, 0, st>>>(top, lateral, w, bias, out, N, H, W);
        else if (Clat == 32) fpn_merge5_kernel<32><<<grid5, 128, 0, st>>>(top, lateral, w, bias, out, N, H, W);
        else MVSTER_REQUIRE(false, "mvster_fpn_merge_f32: unsupported lateral channels %d (8,16,32)", Clat);
        return check_launch("fpn_merge5_kernel");
    }
    const long long n = (long long)N * H * W * 16;  // 16 lanes per pixel
    dim3 grid(ceil_div(n, 128));
    if (Clat == 8) fpn_merge_kernel<8><<<grid, 128, 0, st>>>(top, lateral, w, bias, out, N, H, W);
    else if (Clat == 16) fpn_merge_kernel<16><<<grid, 128, 0, st>>>(top, lateral, w, bias, out, N, H, W);
    else if (Clat == 32) fpn_merge_kernel<32><<<grid, 128, 0, st>>>(top, lateral, w, bias, out, N, H, W);
    else MVSTER_REQUIRE(false, "mvster_fpn_merge_f32: unsupported lateral channels %d (8,16,32)", Clat);
    return check_launch("fpn_merge_kernel");
}

// ---- last pyramid level without materialising the 64-channel full-resolution map ------------------------------
// Reference (mvs4net_utils.py:485-486):  top3 = up2(top2) + inner3(c0);  stage4 = out4(top3)  with out4 a 3x3 conv
// 64 -> 8 without bias.  The conv's channel mixing commutes with the (per-channel, linear) bilinear up-sampling:
//     out4(top3)(p) = sum_{tap, p+tap inside} [ up2(W4[tap] . top2)(p+tap) + (W4[tap] Wi3) . c0(p+tap) + W4[tap] . bi3 ]
// so the 64-channel map at FULL resolution (420 MB for 5 views at 512x640, written and re-read) is never formed:
// U = [W4[tap] . top2]_tap is a 1x1 conv 64 -> 72 at HALF resolution (a GEMM, tensor cores), and this kernel gathers
// 9 bilinear samples of 8 channels plus a 3x3 conv on the 8-channel c0 with composite weights.  4608 MAC per output
// pixel become 1152 + 288 + 576; same result up to fp32 summation order.
namespace mvster {

// U addressing: pixel pitch UC floats, tap t at + t * tap_stride floats (interleaved [..][72]: UC = 72, tap_stride = 8;
// planar [9][N][Hc][Wc][8]: UC = 8, tap_stride = N*Hc*Wc*8 - consecutive half-resolution pixels of one tap are then contiguous,
// so a warp's bilinear samples fall into 4-5 cache lines instead of one line per lane pair)
// eight consecutive floats with ONE 256-bit read-only load (32-byte aligned): half the L1 requests of two 128-bit loads
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ldg_f8(const float* p) {
    F8 r;
#ifdef MVSTER_CPU_EMU  // tests/emu: host build of this source
    for (int i = 0; i < 8; ++i) r.v[i] = p[i];
#else
    unsigned long long q0, q1, q2, q3;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(p));
    r.v[0] = __uint_as_float((unsigned)q0); r.v[1] = __uint_as_float((unsigned)(q0 >> 32));
    r.v[2] = __uint_as_float((unsigned)q1); r.v[3] = __uint_as_float((unsigned)(q1 >> 32));
    r.v[4] = __uint_as_float((unsigned)q2); r.v[5] = __uint_as_float((unsigned)(q2 >> 32));
    r.v[6] = __uint_as_float((unsigned)q3); r.v[7] = __uint_as_float((unsigned)(q3 >> 32));
#endif
    return r;
}

__global__ void __launch_bounds__(128) fpn_out4_gather_kernel(const float* __restrict__ U, int UC, long long tap_stride, const float* __restrict__ c0,
                                                              const float* __restrict__ wc, const float* __restrict__ bt,
                                                              float* __restrict__ out, int N, int H, int W) {
    __shared__ __align__(16) float wc_s[9 * 8 * 8 + 9 * 8];
    for (int i = threadIdx.x; i < 576; i += blockDim.x) wc_s[i] = __ldg(wc + i);
    if (threadIdx.x < 72) wc_s[576 + threadIdx.x] = __ldg(bt + threadIdx.x);
    __syncthreads();
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= (long long)N * H * W) return;
    const int x = (int)(v % W), y = (int)((v / W) % H), b = (int)(v / ((long long)W * H));
    const int Hc = H / 2, Wc = W / 2;
    const float sy = Hc > 1 ? __fdiv_rn((float)(Hc - 1), (float)(H - 1)) : 0.f, sx = Wc > 1 ? __fdiv_rn((float)(Wc - 1), (float)(W - 1)) : 0.f;
    // align_corners=True source rows / columns of the three fine rows y-1..y+1 and columns x-1..x+1
    int ry0[3], ry1[3], rx0[3], rx1[3];
    float wy1[3], wx1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int yy = min(max(y + k - 1, 0), H - 1), xx = min(max(x + k - 1, 0), W - 1);
        const float fy = __fmul_rn(sy, (float)yy), fx = __fmul_rn(sx, (float)xx);
        ry0[k] = min((int)floorf(fy), Hc - 1); rx0[k] = min((int)floorf(fx), Wc - 1);
        ry1[k] = ry0[k] + (ry0[k] < Hc - 1); rx1[k] = rx0[k] + (rx0[k] < Wc - 1);
        wy1[k] = fminf(fmaxf(fy - (float)ry0[k], 0.f), 1.f); wx1[k] = fminf(fmaxf(fx - (float)rx0[k], 0.f), 1.f);
    }
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    const float* Ub = U + (long long)b * Hc * Wc * UC;
    const float* cb = c0 + (long long)b * H * W * 8;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        if ((unsigned)(y + ky - 1) >= (unsigned)H) continue;  // zero padding of the 3x3 conv
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            if ((unsigned)(x + kx - 1) >= (unsigned)W) continue;
            const int tap = ky * 3 + kx;
            // (a) lateral path: composite 3x3 conv on c0 (+ the lateral bias seen through this tap)
            const F8 cpx = ldg_f8(cb + ((long long)(y + ky - 1) * W + (x + kx - 1)) * 8);
            const float* cv = cpx.v;
            const float* wt = wc_s + tap * 64;
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[o] += wc_s[576 + tap * 8 + o];
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] = fmaf(cv[c], wt[c * 8 + o], acc[o]);
            // (b) top-down path: bilinear sample of U_tap (8 channels at half resolution) at the fine position p + tap
            const float ly1 = wy1[ky], ly0 = 1.f - ly1, lx1 = wx1[kx], lx0 = 1.f - lx1;
            const float* ut = Ub + tap * tap_stride;
            const float* u00 = ut + ((long long)ry0[ky] * Wc + rx0[kx]) * UC;
            const float* u01 = ut + ((long long)ry0[ky] * Wc + rx1[kx]) * UC;
            const float* u10 = ut + ((long long)ry1[ky] * Wc + rx0[kx]) * UC;
            const float* u11 = ut + ((long long)ry1[ky] * Wc + rx1[kx]) * UC;
            const F8 a00 = ldg_f8(u00), a01 = ldg_f8(u01), a10 = ldg_f8(u10), a11 = ldg_f8(u11);
#pragma unroll
            for (int o = 0; o < 8; ++o)
                acc[o] += ly0 * (lx0 * a00.v[o] + lx1 * a01.v[o]) + ly1 * (lx0 * a10.v[o] + lx1 * a11.v[o]);
        }
    }
    float4* dst = reinterpret_cast<float4*>(out + v * 8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// ---- variant 2 (opt-in, MVSTER_FPN_GATHER=2; not timed yet): the same sum from shared memory --------------------------
// fpn_out4_gather_kernel pushes 45 256-bit loads per output pixel through the load path (9 taps x (4 bilinear samples of U +
// 1 pixel of c0)) although neighbouring pixels share almost all of them.  Here a CTA owns an 8 x 32 tile of output pixels and
// first stages what the tile needs - the (8+2) x (32+2) halo of c0 and, for each of the 9 tap planes of U, the <= 7 x 20 patch
// of half-resolution pixels its bilinear samples can touch - with coalesced 128-bit loads, split into two 4-channel halves so
// that the per-pixel 128-bit reads of consecutive lanes are conflict-free; the per-pixel arithmetic is the v1 expression, term
// for term, so both kernels return the same bits (tests/test_emu_kernels.py runs them against each other on the CPU).
constexpr int G2_TH = 8, G2_TW = 32, G2_RMAX = 7, G2_CMAX = 20, G2_UPX = G2_RMAX * G2_CMAX, G2_CPX = (G2_TH + 2) * (G2_TW + 2);
constexpr int G2_SMEM_FLOATS = 9 * 2 * G2_UPX * 4 + 2 * G2_CPX * 4 + 576 + 72;

__global__ void __launch_bounds__(256) fpn_out4_gather2_kernel(const float* __restrict__ U, long long tap_stride, const float* __restrict__ c0,
                                                               const float* __restrict__ wc, const float* __restrict__ bt,
                                                               float* __restrict__ out, int N, int H, int W) {
    extern __shared__ __align__(16) float g2_s[];
    float* const u_s = g2_s;                          // [9][2][G2_UPX][4]
    float* const c_s = u_s + 9 * 2 * G2_UPX * 4;      // [2][G2_CPX][4]
    float* const wc_s = c_s + 2 * G2_CPX * 4;         // [9][8][8] then [9][8]
    const int tid = threadIdx.x, b = blockIdx.z;
    const int x0 = blockIdx.x * G2_TW, y0 = blockIdx.y * G2_TH;
    const int Hc = H / 2, Wc = W / 2;
    const float sy = Hc > 1 ? __fdiv_rn((float)(Hc - 1), (float)(H - 1)) : 0.f, sx = Wc > 1 ? __fdiv_rn((float)(Wc - 1), (float)(W - 1)) : 0.f;
    for (int i = tid; i < 576; i += 256) wc_s[i] = __ldg(wc + i);
    if (tid < 72) wc_s[576 + tid] = __ldg(bt + tid);
    // half-resolution patch covered by the tile's bilinear samples (clamped fine rows y0-1 .. y0+TH, columns x0-1 .. x0+TW)
    const int fy_lo = max(y0 - 1, 0), fy_hi = min(y0 + G2_TH, H - 1), fx_lo = max(x0 - 1, 0), fx_hi = min(x0 + G2_TW, W - 1);
    const int r_lo = min((int)floorf(__fmul_rn(sy, (float)fy_lo)), Hc - 1), c_lo = min((int)floorf(__fmul_rn(sx, (float)fx_lo)), Wc - 1);
    int r_hi = min((int)floorf(__fmul_rn(sy, (float)fy_hi)), Hc - 1), c_hi = min((int)floorf(__fmul_rn(sx, (float)fx_hi)), Wc - 1);
    r_hi += (r_hi < Hc - 1); c_hi += (c_hi < Wc - 1);
    const int nr = r_hi - r_lo + 1, nc = c_hi - c_lo + 1;  // <= G2_RMAX, G2_CMAX (scale < 1/2)
    const float* Ub = U + (long long)b * Hc * Wc * 8;
    for (int i = tid; i < 9 * nr * nc * 2; i += 256) {
        const int half = i & 1, px = (i >> 1) % (nr * nc), tap = (i >> 1) / (nr * nc);
        const int r = px / nc, c = px % nc;
        const float4 v = __ldg(reinterpret_cast<const float4*>(Ub + tap * tap_stride + ((long long)(r_lo + r) * Wc + (c_lo + c)) * 8 + half * 4));
        *reinterpret_cast<float4*>(u_s + ((tap * 2 + half) * G2_UPX + r * G2_CMAX + c) * 4) = v;
    }
    const float* cb = c0 + (long long)b * H * W * 8;
    for (int i = tid; i < G2_CPX * 2; i += 256) {
        const int half = i & 1, px = i >> 1, r = px / (G2_TW + 2), c = px % (G2_TW + 2);
        const int yy = y0 - 1 + r, xx = x0 - 1 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W)
            v = __ldg(reinterpret_cast<const float4*>(cb + ((long long)yy * W + xx) * 8 + half * 4));
        *reinterpret_cast<float4*>(c_s + (half * G2_CPX + px) * 4) = v;
    }
    __syncthreads();
    const int tx = tid & 31, ty = tid >> 5, x = x0 + tx, y = y0 + ty;
    if (x >= W || y >= H) return;
    int ry0[3], ry1[3], rx0[3], rx1[3];
    float wy1[3], wx1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int yy = min(max(y + k - 1, 0), H - 1), xx = min(max(x + k - 1, 0), W - 1);
        const float fy = __fmul_rn(sy, (float)yy), fx = __fmul_rn(sx, (float)xx);
        ry0[k] = min((int)floorf(fy), Hc - 1); rx0[k] = min((int)floorf(fx), Wc - 1);
        ry1[k] = ry0[k] + (ry0[k] < Hc - 1); rx1[k] = rx0[k] + (rx0[k] < Wc - 1);
        wy1[k] = fminf(fmaxf(fy - (float)ry0[k], 0.f), 1.f); wx1[k] = fminf(fmaxf(fx - (float)rx0[k], 0.f), 1.f);
        ry0[k] -= r_lo; ry1[k] -= r_lo; rx0[k] -= c_lo; rx1[k] -= c_lo;  // patch coordinates
    }
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        if ((unsigned)(y + ky - 1) >= (unsigned)H) continue;  // zero padding of the 3x3 conv
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            if ((unsigned)(x + kx - 1) >= (unsigned)W) continue;
            const int tap = ky * 3 + kx;
            const int cp = (ty + ky) * (G2_TW + 2) + (tx + kx);
            const float4 cl = *reinterpret_cast<const float4*>(c_s + cp * 4), ch = *reinterpret_cast<const float4*>(c_s + (G2_CPX + cp) * 4);
            const float cv[8] = {cl.x, cl.y, cl.z, cl.w, ch.x, ch.y, ch.z, ch.w};
            const float* wt = wc_s + tap * 64;
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[o] += wc_s[576 + tap * 8 + o];
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] = fmaf(cv[c], wt[c * 8 + o], acc[o]);
            const float ly1 = wy1[ky], ly0 = 1.f - ly1, lx1 = wx1[kx], lx0 = 1.f - lx1;
            const int p00 = ry0[ky] * G2_CMAX + rx0[kx], p01 = ry0[ky] * G2_CMAX + rx1[kx];
            const int p10 = ry1[ky] * G2_CMAX + rx0[kx], p11 = ry1[ky] * G2_CMAX + rx1[kx];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const float* up = u_s + (tap * 2 + half) * G2_UPX * 4;
                const float4 a00 = *reinterpret_cast<const float4*>(up + p00 * 4), a01 = *reinterpret_cast<const float4*>(up + p01 * 4);
                const float4 a10 = *reinterpret_cast<const float4*>(up + p10 * 4), a11 = *reinterpret_cast<const float4*>(up + p11 * 4);
                acc[half * 4 + 0] += ly0 * (lx0 * a00.x + lx1 * a01.x) + ly1 * (lx0 * a10.x + lx1 * a11.x);
                acc[half * 4 + 1] += ly0 * (lx0 * a00.y + lx1 * a01.y) + ly1 * (lx0 * a10.y + lx1 * a11.y);
                acc[half * 4 + 2] += ly0 * (lx0 * a00.z + lx1 * a01.z) + ly1 * (lx0 * a10.z + lx1 * a11.z);
                acc[half * 4 + 3] += ly0 * (lx0 * a00.w + lx1 * a01.w) + ly1 * (lx0 * a10.w + lx1 * a11.w);
            }
        }
    }
    float4* dst = reinterpret_cast<float4*>(out + (((long long)b * H + y) * W + x) * 8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// (A third variant - two pixels per thread, packed FMAs, 1111 instructions per pixel - measured 262 us against 168 us for
// variant 2 and 199 us for variant 1 on B200 and was removed: profiles/r02_glue_ab.md.)

}  // namespace mvster

extern "C" int mvster_fpn_out4_gather_f32(const float* U, int u_channels, const float* c0, const float* w_comp, const float* b_tap,
                                          float* out, int N, int H, int W, mvster_stream_t stream) {
    MVSTER_REQUIRE(U && c0 && w_comp && b_tap && out, "mvster_fpn_out4_gather_f32: null pointer");
    MVSTER_REQUIRE(N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "mvster_fpn_out4_gather_f32: H,W must be even");
    MVSTER_REQUIRE(u_channels == 8 || (u_channels >= 72 && u_channels % 4 == 0),
                   "mvster_fpn_out4_gather_f32: U is [..][>= 72 channels] (9 taps x 8 interleaved) or, with u_channels = 8, planar [9][N][H/2][W/2][8]");
    const long long n = (long long)N * H * W;
    const long long tap_stride = u_channels == 8 ? (long long)N * (H / 2) * (W / 2) * 8 : 8;
    // measured on B200 (5 x 512 x 640, profiles/r02_glue_ab.md): 198.7 (v1) / 168.0 (v2, same bits) / 262.1 us (v3) -> default v2
    const char* variant = getenv("MVSTER_FPN_GATHER");
    if ((!variant || atoi(variant) == 2) && u_channels == 8 && N < 65536) {  // shared-memory tiled variant
        auto k = mvster::fpn_out4_gather2_kernel;
        const size_t smem = (size_t)mvster::G2_SMEM_FLOATS * sizeof(float);
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<dim3(mvster::ceil_div(W, mvster::G2_TW), mvster::ceil_div(H, mvster::G2_TH), N), 256, smem, (cudaStream_t)stream>>>(U, tap_stride, c0, w_comp, b_tap, out, N, H, W);
        return mvster::check_launch("fpn_out4_gather2_kernel");
    }
    mvster::fpn_out4_gather_kernel<<<mvster::ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(U, u_channels, tap_stride, c0, w_comp, b_tap, out, N, H, W);
    return mvster::check_launch("fpn_out4_gather_kernel");
}
