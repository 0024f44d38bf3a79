// Depth-hypothesis schedulers, the depth head and small layout helpers (fp32).
//   init_inverse_range       models/mvs4net_utils.py:71-77
//   schedule_inverse_range   models/mvs4net_utils.py:79-86
//   prob 1x1x1 + softmax + argmax-gather + confidence + inverse range  :900, :1066-1088
//   depth_regression (optional extra output)  models/module.py:935-941
// All of these move a handful of floats per pixel: HBM-bandwidth bound, one thread per pixel.
#include "common.cuh"
#include <math.h>

namespace mvster {

__global__ void hypo_init_inverse_kernel(const float* __restrict__ dv, int n_dv, float* __restrict__ hypo,
                                         int B, int D, long long plane) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n = (long long)B * D * plane;
    if (i >= n) return;
    const int d = (int)((i / plane) % D);
    const int b = (int)(i / (plane * D));
    const float inv_near = __fdiv_rn(1.f, __ldg(dv + (long long)b * n_dv));
    const float inv_far = __fdiv_rn(1.f, __ldg(dv + (long long)b * n_dv + n_dv - 1));
    const float ramp = __fdiv_rn((float)d, (float)(D - 1));
    const float inv = __fadd_rn(inv_far, __fmul_rn(__fsub_rn(inv_near, inv_far), ramp));
    hypo[i] = __fdiv_rn(1.f, inv);
}

// ATen's align_corners=True source index: scale = (in-1)/(out-1); src = scale*dst;
// i0 = min(floor(src), in-1); lambda = clamp(src - i0, 0, 1); i1 = i0 + (i0 < in-1).
__device__ __forceinline__ void ac_index(int dst, int in_size, int out_size, int& i0, int& i1, float& l0, float& l1) {
    const float scale = out_size > 1 ? __fdiv_rn((float)(in_size - 1), (float)(out_size - 1)) : 0.f;
    const float src = __fmul_rn(scale, (float)dst);
    i0 = min((int)floorf(src), in_size - 1);
    l1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.f), 1.f);
    l0 = __fsub_rn(1.f, l1);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
}

template <int D>
__global__ void hypo_schedule_inverse_kernel(const float* __restrict__ inv_min, const float* __restrict__ inv_max,
                                             float* __restrict__ hypo, int B, int H, int W) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long plane = (long long)H * W;
    if (i >= (long long)B * plane) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / plane);
    const int Hc = H / 2, Wc = W / 2;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    ac_index(y, Hc, H, y0, y1, ly0, ly1);
    ac_index(x, Wc, W, x0, x1, lx0, lx1);
    const long long cb = (long long)b * Hc * Wc;
    const long long o00 = cb + (long long)y0 * Wc + x0, o01 = cb + (long long)y0 * Wc + x1;
    const long long o10 = cb + (long long)y1 * Wc + x0, o11 = cb + (long long)y1 * Wc + x1;
    const float mx00 = __ldg(inv_max + o00), mx01 = __ldg(inv_max + o01), mx10 = __ldg(inv_max + o10), mx11 = __ldg(inv_max + o11);
    const float df00 = __fsub_rn(__ldg(inv_min + o00), mx00), df01 = __fsub_rn(__ldg(inv_min + o01), mx01);
    const float df10 = __fsub_rn(__ldg(inv_min + o10), mx10), df11 = __fsub_rn(__ldg(inv_min + o11), mx11);
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float ramp = __fdiv_rn((float)d, (float)(D - 1));
        const float c00 = __fadd_rn(mx00, __fmul_rn(df00, ramp)), c01 = __fadd_rn(mx01, __fmul_rn(df01, ramp));
        const float c10 = __fadd_rn(mx10, __fmul_rn(df10, ramp)), c11 = __fadd_rn(mx11, __fmul_rn(df11, ramp));
        const float top = __fadd_rn(__fmul_rn(lx0, c00), __fmul_rn(lx1, c01));
        const float bot = __fadd_rn(__fmul_rn(lx0, c10), __fmul_rn(lx1, c11));
        const float inv = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
        hypo[((long long)b * D + d) * plane + (long long)y * W + x] = __fdiv_rn(1.f, inv);
    }
}

struct HeadArgs {
    const float* logits; const float* feat8; const float* prob_w; const float* prob_b; const float* hypo;
    float* attn; float* depth; float* conf; float* inv_min; float* inv_max; float* soft;
    int B, H, W;
    float split_itv;
    int round_bf16;  // bf16 storage: the `prob` layer reads feat8 rounded to bf16 (its weights arrive bf16-valued)
};

// fp32 -> nearest-even bf16 value, as fp32 (finite inputs)
__device__ __forceinline__ float bf16_value(float x) {
    const unsigned u = __float_as_uint(x);
    return __uint_as_float((u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u);
}

template <int D>
__global__ void head_kernel(const HeadArgs a) {
    const long long plane = (long long)a.H * a.W;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)a.B * plane) return;
    const int b = (int)(i / plane);
    const long long p = i % plane;
    float lg[D], hy[D];
    if (a.logits) {
#pragma unroll
        for (int d = 0; d < D; ++d) lg[d] = __ldg(a.logits + ((long long)b * D + d) * plane + p);
    } else {
        float w[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] = __ldg(a.prob_w + c);
        const float bias = __ldg(a.prob_b);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float4* f = reinterpret_cast<const float4*>(a.feat8 + (((long long)b * D + d) * plane + p) * 8);
            float4 f0 = __ldg(f), f1 = __ldg(f + 1);
            if (a.round_bf16) {
                f0.x = bf16_value(f0.x); f0.y = bf16_value(f0.y); f0.z = bf16_value(f0.z); f0.w = bf16_value(f0.w);
                f1.x = bf16_value(f1.x); f1.y = bf16_value(f1.y); f1.z = bf16_value(f1.z); f1.w = bf16_value(f1.w);
            }
            float s = f0.x * w[0];
            s = fmaf(f0.y, w[1], s); s = fmaf(f0.z, w[2], s); s = fmaf(f0.w, w[3], s);
            s = fmaf(f1.x, w[4], s); s = fmaf(f1.y, w[5], s); s = fmaf(f1.z, w[6], s); s = fmaf(f1.w, w[7], s);
            lg[d] = s + bias;
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) hy[d] = __ldg(a.hypo + ((long long)b * D + d) * plane + p);
    float m = lg[0];
#pragma unroll
    for (int d = 1; d < D; ++d) m = fmaxf(m, lg[d]);
    float se = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) { lg[d] = expf(lg[d] - m); se += lg[d]; }
    float best = -1.f, soft = 0.f;
    int arg = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float pr = __fdiv_rn(lg[d], se);
        if (a.attn) a.attn[((long long)b * D + d) * plane + p] = pr;
        if (pr > best) { best = pr; arg = d; }  // first maximum wins, like torch.max on CPU
        soft = fmaf(pr, hy[d], soft);
    }
    float dep = hy[0];
#pragma unroll
    for (int d = 1; d < D; ++d) dep = (arg == d) ? hy[d] : dep;
    if (a.depth) a.depth[i] = dep;
    if (a.conf) a.conf[i] = best;
    if (a.soft) a.soft[i] = soft;
    if (a.inv_min || a.inv_max) {
        const float itv = __fsub_rn(__fdiv_rn(1.f, hy[2]), __fdiv_rn(1.f, hy[1]));  // :1084
        const float inv_d = __fdiv_rn(1.f, dep), step = __fmul_rn(a.split_itv, itv);
        if (a.inv_min) a.inv_min[i] = __fadd_rn(inv_d, step);
        if (a.inv_max) a.inv_max[i] = __fsub_rn(inv_d, step);
    }
}

__global__ void upsample_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out,
                                         int B, int H, int W, int f) {
    const int Ho = H * f, Wo = W * f;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)B * Ho * Wo) return;
    const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho), b = (int)(i / ((long long)Wo * Ho));
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    ac_index(y, H, Ho, y0, y1, ly0, ly1);
    ac_index(x, W, Wo, x0, x1, lx0, lx1);
    const float* p = in + (long long)b * H * W;
    const float top = __fadd_rn(__fmul_rn(lx0, __ldg(p + (long long)y0 * W + x0)), __fmul_rn(lx1, __ldg(p + (long long)y0 * W + x1)));
    const float bot = __fadd_rn(__fmul_rn(lx0, __ldg(p + (long long)y1 * W + x0)), __fmul_rn(lx1, __ldg(p + (long long)y1 * W + x1)));
    out[i] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

// [B][C][HW] -> [B][HW][C] through a 32x33 shared tile (coalesced both ways).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, long long HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float* src = in + (long long)b * C * HW;
    float* dst = out + (long long)b * C * HW;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j;
        const long long p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && p < HW) ? src[(long long)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long p = p0 + j;
        const int c = c0 + threadIdx.x;
        if (c < C && p < HW) dst[p * C + c] = tile[threadIdx.x][j];
    }
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread: what a bf16 store of the feature pyramid's outputs applies
__global__ void cast_bf16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, long long n8, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    auto rn = [](float x) -> unsigned {
        const unsigned u = __float_as_uint(x);
        if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (u >> 16) | 0x40u;  // NaN stays NaN
        return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
    };
    if (i < n8) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
        uint4 o;
        o.x = rn(a.x) | (rn(a.y) << 16); o.y = rn(a.z) | (rn(a.w) << 16);
        o.z = rn(b.x) | (rn(b.y) << 16); o.w = rn(b.z) | (rn(b.w) << 16);
        reinterpret_cast<uint4*>(out)[i] = o;
    } else if (i == n8) {  // tail (n % 8 elements)
        for (long long j = 8 * n8; j < n; ++j) out[j] = (uint16_t)rn(in[j]);
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_hypo_init_inverse_f32(const float* depth_values, int n_dv, float* hypo,
                                            int B, int D, int H, int W, mvster_stream_t stream) {
    MVSTER_REQUIRE(depth_values && hypo, "mvster_hypo_init_inverse_f32: null pointer");
    MVSTER_REQUIRE(n_dv >= 1 && B > 0 && D >= 2 && H > 0 && W > 0, "mvster_hypo_init_inverse_f32: bad shape");
    const long long plane = (long long)H * W, n = (long long)B * D * plane;
    hypo_init_inverse_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(depth_values, n_dv, hypo, B, D, plane);
    return check_launch("hypo_init_inverse_kernel");
}

extern "C" int mvster_hypo_schedule_inverse_f32(const float* inv_min, const float* inv_max, float* hypo,
                                                int B, int D, int H, int W, mvster_stream_t stream) {
    MVSTER_REQUIRE(inv_min && inv_max && hypo, "mvster_hypo_schedule_inverse_f32: null pointer");
    MVSTER_REQUIRE(B > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "mvster_hypo_schedule_inverse_f32: H,W must be even");
    const long long n = (long long)B * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 4) hypo_schedule_inverse_kernel<4><<<ceil_div(n, 256), 256, 0, st>>>(inv_min, inv_max, hypo, B, H, W);
    else if (D == 8) hypo_schedule_inverse_kernel<8><<<ceil_div(n, 256), 256, 0, st>>>(inv_min, inv_max, hypo, B, H, W);
    else MVSTER_REQUIRE(false, "mvster_hypo_schedule_inverse_f32: unsupported D=%d (4 or 8)", D);
    return check_launch("hypo_schedule_inverse_kernel");
}

extern "C" int mvster_head_f32(const float* logits, const float* feat8, const float* prob_w, const float* prob_b,
                               const float* hypo, float* attn, float* depth, float* conf,
                               float* inv_min, float* inv_max, float* soft_depth,
                               int B, int D, int H, int W, float split_itv, mvster_stream_t stream) {
    return mvster_head_ex_f32(logits, feat8, prob_w, prob_b, hypo, attn, depth, conf, inv_min, inv_max, soft_depth, B, D, H, W, split_itv, 0, stream);
}

extern "C" int mvster_head_ex_f32(const float* logits, const float* feat8, const float* prob_w, const float* prob_b,
                                  const float* hypo, float* attn, float* depth, float* conf,
                                  float* inv_min, float* inv_max, float* soft_depth,
                                  int B, int D, int H, int W, float split_itv, int flags, mvster_stream_t stream) {
    MVSTER_REQUIRE(hypo, "mvster_head_f32: hypo is null");
    MVSTER_REQUIRE(logits || (feat8 && prob_w && prob_b), "mvster_head_f32: need logits or (feat8, prob_w, prob_b)");
    MVSTER_REQUIRE(B > 0 && H > 0 && W > 0, "mvster_head_f32: bad shape");
    HeadArgs a{logits, feat8, prob_w, prob_b, hypo, attn, depth, conf, inv_min, inv_max, soft_depth, B, H, W, split_itv,
               (flags & MVSTER_HEAD_BF16_INPUT) ? 1 : 0};
    const long long n = (long long)B * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 4) head_kernel<4><<<ceil_div(n, 128), 128, 0, st>>>(a);
    else if (D == 8) head_kernel<8><<<ceil_div(n, 128), 128, 0, st>>>(a);
    else MVSTER_REQUIRE(false, "mvster_head_f32: unsupported D=%d (4 or 8)", D);
    return check_launch("head_kernel");
}

extern "C" int mvster_upsample_bilinear_f32(const float* in, float* out, int B, int H, int W, int factor,
                                            mvster_stream_t stream) {
    MVSTER_REQUIRE(in && out, "mvster_upsample_bilinear_f32: null pointer");
    MVSTER_REQUIRE(B > 0 && H > 0 && W > 0 && factor >= 1, "mvster_upsample_bilinear_f32: bad shape");
    const long long n = (long long)B * H * W * factor * factor;
    upsample_bilinear_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, factor);
    return check_launch("upsample_bilinear_kernel");
}

extern "C" int mvster_nchw_to_nhwc_f32(const float* in, float* out, int B, int C, int H, int W,
                                       mvster_stream_t stream) {
    MVSTER_REQUIRE(in && out, "mvster_nchw_to_nhwc_f32: null pointer");
    MVSTER_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && B < 65536, "mvster_nchw_to_nhwc_f32: bad shape");
    const long long HW = (long long)H * W;
    dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, C, HW);
    return check_launch("nchw_to_nhwc_kernel");
}

// ---- relative pose ----------------------------------------------------------------------------
// models/mvs4net_utils.py:1032-1035 (K @ E[:3,:4] composed in fp32, as the reference's fp32 matmul)
// followed by :24  P = src_proj @ inverse(ref_proj).  The 4x4 inverse and product are evaluated
// in fp64 (Gauss-Jordan, partial pivoting) and rounded once to fp32; the reference's fp32
// LAPACK inverse agrees to ~1e-7 relative.  One thread per (batch, source view).
namespace mvster {

__device__ void compose_full(const float* cam /*[2][4][4]*/, double (&M)[4][4]) {
    const float* E = cam;
    const float* K = cam + 16;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            float s = 0.f;
            for (int k = 0; k < 3; ++k) s = fmaf(K[r * 4 + k], E[k * 4 + c], s);
            M[r][c] = (double)s;
        }
    for (int c = 0; c < 4; ++c) M[3][c] = (double)E[12 + c];
}

__global__ void pose_kernel(const float* __restrict__ proj, float* __restrict__ pose, int B, int Nv, int first_view, int V) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * V) return;
    const int b = i / V, v = i % V;
    double R[4][4], S[4][4], inv[4][4];
    compose_full(proj + ((long long)b * Nv) * 32, R);
    compose_full(proj + ((long long)b * Nv + first_view + v) * 32, S);
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) inv[r][c] = r == c ? 1.0 : 0.0;
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        for (int r = col + 1; r < 4; ++r)
            if (fabs(R[r][col]) > fabs(R[piv][col])) piv = r;
        for (int c = 0; c < 4; ++c) {
            double t = R[col][c]; R[col][c] = R[piv][c]; R[piv][c] = t;
            t = inv[col][c]; inv[col][c] = inv[piv][c]; inv[piv][c] = t;
        }
        const double d = 1.0 / R[col][col];
        for (int c = 0; c < 4; ++c) { R[col][c] *= d; inv[col][c] *= d; }
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = R[r][col];
            for (int c = 0; c < 4; ++c) { R[r][c] -= f * R[col][c]; inv[r][c] -= f * inv[col][c]; }
        }
    }
    float* o = pose + (long long)i * 12;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += S[r][k] * inv[k][c];
            if (c < 3) o[r * 3 + c] = (float)s; else o[9 + r] = (float)s;
        }
    }
}

}  // namespace mvster

extern "C" int mvster_pose_f32(const float* proj, float* pose, int B, int Nv, int first_view, int V,
                               mvster_stream_t stream) {
    MVSTER_REQUIRE(proj && pose, "mvster_pose_f32: null pointer");
    MVSTER_REQUIRE(B > 0 && V > 0 && first_view >= 1 && first_view + V <= Nv, "mvster_pose_f32: views %d..%d outside 1..%d",
                   first_view, first_view + V - 1, Nv - 1);
    mvster::pose_kernel<<<mvster::ceil_div((long long)B * V, 64), 64, 0, (cudaStream_t)stream>>>(proj, pose, B, Nv, first_view, V);
    return mvster::check_launch("pose_kernel");
}

// ---- linear-depth hypotheses (inverse_depth=False) --------------------------------------------
//   init_range      models/mvs4net_utils.py:61-69    d_k = d_min + k * (d_max - d_min)/(D-1)
//   schedule_range  models/mvs4net_utils.py:88-99    per coarse pixel [depth -+ D/2*itv] in D samples,
//                   trilinear (align_corners) x2 up-sampling; itv = ratio_k * (d_max - d_min)/n_dv  (MVS4Net.py:61-63,97)
namespace mvster {

__global__ void hypo_init_linear_kernel(const float* __restrict__ dv, int n_dv, float* __restrict__ hypo,
                                        int B, int D, long long plane) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)B * D * plane) return;
    const int d = (int)((i / plane) % D), b = (int)(i / (plane * D));
    const float lo = __ldg(dv + (long long)b * n_dv), hi = __ldg(dv + (long long)b * n_dv + n_dv - 1);
    const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(D - 1));
    hypo[i] = __fadd_rn(lo, __fmul_rn((float)d, step));
}

template <int D>
__global__ void hypo_schedule_linear_kernel(const float* __restrict__ depth, const float* __restrict__ dv, int n_dv, float ratio,
                                            float* __restrict__ hypo, int B, int H, int W) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long plane = (long long)H * W;
    if (i >= (long long)B * plane) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / plane);
    const int Hc = H / 2, Wc = W / 2;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    ac_index(y, Hc, H, y0, y1, ly0, ly1);
    ac_index(x, Wc, W, x0, x1, lx0, lx1);
    // depth_interval = (d_max - d_min) / n_dv (MVS4Net.py:63), times depth_interals_ratio[k]
    const float itv = __fmul_rn(ratio, __fdiv_rn(__fsub_rn(__ldg(dv + (long long)b * n_dv + n_dv - 1), __ldg(dv + (long long)b * n_dv)), (float)n_dv));
    const float half_span = __fmul_rn((float)D * 0.5f, itv);
    const float* p = depth + (long long)b * Hc * Wc;
    const float c[4] = {__ldg(p + (long long)y0 * Wc + x0), __ldg(p + (long long)y0 * Wc + x1),
                        __ldg(p + (long long)y1 * Wc + x0), __ldg(p + (long long)y1 * Wc + x1)};
    float lo[4], st[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        lo[k] = __fsub_rn(c[k], half_span);
        st[k] = __fdiv_rn(__fsub_rn(__fadd_rn(c[k], half_span), lo[k]), (float)(D - 1));
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = __fadd_rn(lo[k], __fmul_rn((float)d, st[k]));
        const float top = __fadd_rn(__fmul_rn(lx0, s[0]), __fmul_rn(lx1, s[1]));
        const float bot = __fadd_rn(__fmul_rn(lx0, s[2]), __fmul_rn(lx1, s[3]));
        hypo[((long long)b * D + d) * plane + (long long)y * W + x] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    }
}

}  // namespace mvster

extern "C" int mvster_hypo_init_linear_f32(const float* depth_values, int n_dv, float* hypo,
                                           int B, int D, int H, int W, mvster_stream_t stream) {
    MVSTER_REQUIRE(depth_values && hypo, "mvster_hypo_init_linear_f32: null pointer");
    MVSTER_REQUIRE(n_dv >= 1 && B > 0 && D >= 2 && H > 0 && W > 0, "mvster_hypo_init_linear_f32: bad shape");
    const long long plane = (long long)H * W, n = (long long)B * D * plane;
    mvster::hypo_init_linear_kernel<<<mvster::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(depth_values, n_dv, hypo, B, D, plane);
    return mvster::check_launch("hypo_init_linear_kernel");
}

extern "C" int mvster_hypo_schedule_linear_f32(const float* depth, const float* depth_values, int n_dv, float ratio, float* hypo,
                                               int B, int D, int H, int W, mvster_stream_t stream) {
    MVSTER_REQUIRE(depth && depth_values && hypo, "mvster_hypo_schedule_linear_f32: null pointer");
    MVSTER_REQUIRE(B > 0 && n_dv >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "mvster_hypo_schedule_linear_f32: H,W must be even");
    const long long n = (long long)B * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 4) mvster::hypo_schedule_linear_kernel<4><<<mvster::ceil_div(n, 256), 256, 0, st>>>(depth, depth_values, n_dv, ratio, hypo, B, H, W);
    else if (D == 8) mvster::hypo_schedule_linear_kernel<8><<<mvster::ceil_div(n, 256), 256, 0, st>>>(depth, depth_values, n_dv, ratio, hypo, B, H, W);
    else MVSTER_REQUIRE(false, "mvster_hypo_schedule_linear_f32: unsupported D=%d (4 or 8)", D);
    return mvster::check_launch("hypo_schedule_linear_kernel");
}

extern "C" int mvster_cast_bf16(const float* in, void* out, long long n, mvster_stream_t stream) {
    MVSTER_REQUIRE(in && out && n > 0, "mvster_cast_bf16: bad arguments");
    MVSTER_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "mvster_cast_bf16: pointers must be 16-byte aligned");
    const long long n8 = n / 8;
    cast_bf16_kernel<<<ceil_div(n8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<uint16_t*>(out), n8, n);
    return check_launch("cast_bf16_kernel");
}
