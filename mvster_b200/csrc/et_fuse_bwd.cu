// Backward of the fused plane-sweep warp + epipolar-Transformer aggregation (fp32), for training.
//
// Differentiates, for one cascade stage and all source views in ONE launch, what et_fuse.cu computes:
//   homo_warping                          models/mvs4net_utils.py:13-59   (the sampling grid is built under
//                                         torch.no_grad(), :23, and the hypotheses arrive detached, MVS4Net.py:95,
//                                         so gradients flow to the FEATURES only)
//   group correlation                     :1037-1040
//   softmax_D(sum_g cor / temp)/sqrt(C)   :1053
//   weighted sums over views and the final division  :1054-1060
// The reference's autograd keeps, per view, the warped volume [B,C,D,H,W], the replicated reference volume, the
// correlation and the weights alive for the backward pass; here nothing but the forward's cost volume and weight sum is
// saved - the taps are gathered again, the per-view softmax is recomputed in registers, the reference-feature gradient
// is written once and the source-feature gradients are scattered with vector atomics (red.global.add.v2/v4.f32).
//
// With W[d] = 1e-8 + sum_v a_v[d], cost[g,d] = sum_v a_v[d] cor_v[g,d] / W[d], a_v = softmax_d(s_v)/sqrt(C),
// s_v[d] = sum_g cor_v[g,d] / temp and the incoming gradient gc[g,d]:
//   gacc[g,d] = gc[g,d] / W[d]                    gW[d] = -sum_g gc[g,d] cost[g,d] / W[d]
//   ga_v[d]   = sum_g gacc[g,d] cor_v[g,d] + gW[d]
//   gs_v[d]   = a_v[d] (ga_v[d] - sum_d' ga_v[d'] p_v[d'])         (p_v = softmax_d(s_v))
//   gcor_v[g,d] = a_v[d] gacc[g,d] + gs_v[d] / temp
//   gref[c]  += sum_d gcor_v[g(c),d] warped_v[c,d] / (C/G)
//   gsrc_v[c, tap] += bilinear_weight(tap) gcor_v[g(c),d] ref[c] / (C/G)     for the (up to) four valid taps of (d)
//
// Thread mapping = the generic forward kernel's: G consecutive lanes own one reference pixel, lane g the C/G channels of
// group g; sums over groups are warp shuffles.
#include "common.cuh"
#include <math.h>

namespace mvster {

struct EtBwdArgs {
    const float* ref;
    const float* src[MVSTER_MAX_VIEWS];
    float* gsrc[MVSTER_MAX_VIEWS];   // null: that view's features need no gradient
    const float* pose;
    const float* hypo;
    const float* cost;    // [B][D][H][W][G]  normalised forward output
    const float* wsum;    // [B][D][H][W]     sum over views of the weights, WITHOUT the 1e-8 seed (forward with MVSTER_ET_PARTIAL)
    const float* gcost;   // [B][D][H][W][G]
    float* gref;          // [B][H][W][C], overwritten
    int B, V, H, W, Hs, Ws;
    float attn_temp, sqrt_c;
};

namespace bwd {

template <int N>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[N]) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else if constexpr (N == 2) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __ldg(p + i);
    }
}

template <int N>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[N]) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if constexpr (N == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = v[i];
    }
}

// p[0..N) += w * v[0..N), element-wise atomic (the widest reduction the alignment allows: N floats from a N*4-byte aligned address)
template <int N>
__device__ __forceinline__ void scatter_add(float* p, const float (&v)[N], float w) {
#ifdef MVSTER_CPU_EMU
    for (int i = 0; i < N; ++i) atomicAdd(p + i, w * v[i]);
#else
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i)
            atomicAdd(reinterpret_cast<float4*>(p) + i, make_float4(w * v[4 * i], w * v[4 * i + 1], w * v[4 * i + 2], w * v[4 * i + 3]));
    } else if constexpr (N == 2) {
        atomicAdd(reinterpret_cast<float2*>(p), make_float2(w * v[0], w * v[1]));
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) atomicAdd(p + i, w * v[i]);
    }
#endif
}

template <int G>
__device__ __forceinline__ float group_sum(float s) {
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if constexpr (G >= 8) s += __shfl_xor_sync(0xffffffffu, s, 4);
    return s;
}

// the four bilinear taps of one (pixel, view, hypothesis): the forward kernel's arithmetic, operation for operation
// (mvs4net_utils.py:34-45 and grid_sample's align_corners=True un-normalisation)
struct Taps {
    float w_nw, w_ne, w_sw, w_se;
    long long off;   // element offset of the north-west tap's pixel (times C added by the caller)
    bool nw, ne, sw, se;
};

__device__ __forceinline__ Taps make_taps(float rx, float ry, float rz, float tx, float ty, float tz, float dep,
                                          float half_w, float half_h, float lim_x, float lim_y, int Hs, int Ws) {
    const float X = __fadd_rn(__fmul_rn(rx, dep), tx);
    const float Y = __fadd_rn(__fmul_rn(ry, dep), ty);
    float Z = __fadd_rn(__fmul_rn(rz, dep), tz);
    if (Z == 0.f) Z = 1e-9f;
    const float gx = __fsub_rn(__fdiv_rn(__fdiv_rn(X, Z), half_w), 1.f);
    const float gy = __fsub_rn(__fdiv_rn(__fdiv_rn(Y, Z), half_h), 1.f);
    float ix = __fmul_rn(__fadd_rn(gx, 1.f), half_w);
    float iy = __fmul_rn(__fadd_rn(gy, 1.f), half_h);
    ix = fminf(fmaxf(ix, -2.f), lim_x);
    iy = fminf(fmaxf(iy, -2.f), lim_y);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx = ix - x0f, wy = iy - y0f;
    const float ex = 1.f - wx, ey = 1.f - wy;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const bool vx0 = (unsigned)x0 < (unsigned)Ws, vx1 = (unsigned)(x0 + 1) < (unsigned)Ws;
    const bool vy0 = (unsigned)y0 < (unsigned)Hs, vy1 = (unsigned)(y0 + 1) < (unsigned)Hs;
    Taps t;
    t.w_nw = ey * ex; t.w_ne = ey * wx; t.w_sw = wy * ex; t.w_se = wy * wx;
    t.off = (long long)y0 * Ws + x0;
    t.nw = vy0 && vx0; t.ne = vy0 && vx1; t.sw = vy1 && vx0; t.se = vy1 && vx1;
    return t;
}

}  // namespace bwd

template <int CPG, int G, int D>
__global__ void __launch_bounds__(256) et_fuse_bwd_kernel(const EtBwdArgs a) {
    using namespace bwd;
    constexpr int C = CPG * G;
    const int g = threadIdx.x % G;
    const long long npix = (long long)a.B * a.H * a.W;
    long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const bool live = pix < npix;
    if (!live) pix = npix - 1;  // keep the lane for the shuffles, drop its stores and atomics
    const int x = (int)(pix % a.W);
    const int y = (int)((pix / a.W) % a.H);
    const int b = (int)(pix / ((long long)a.W * a.H));
    const long long plane = (long long)a.H * a.W;
    const long long pix_in_b = (long long)y * a.W + x;

    float ref[CPG], gref[CPG];
    load_vec<CPG>(a.ref + pix * C + g * CPG, ref);
#pragma unroll
    for (int c = 0; c < CPG; ++c) gref[c] = 0.f;

    float dep[D], gacc[D], gws[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const long long o = ((long long)b * D + d) * plane + pix_in_b;
        dep[d] = __ldg(a.hypo + o);
        const float Wd = __fadd_rn(1e-8f, __ldg(a.wsum + o));   // mvs4net_utils.py:1022 seed, as et_normalize_kernel adds it
        const float gc = __ldg(a.gcost + o * G + g);
        gacc[d] = gc / Wd;
        gws[d] = -group_sum<G>(gc * __ldg(a.cost + o * G + g)) / Wd;
    }

    const float fx = (float)x, fy = (float)y;
    const float half_w = 0.5f * (float)(a.Ws - 1);
    const float half_h = 0.5f * (float)(a.Hs - 1);
    const float lim_x = (float)a.Ws + 1.f, lim_y = (float)a.Hs + 1.f;
    const long long row = (long long)a.Ws * C;

    for (int v = 0; v < a.V; ++v) {
        const float* P = a.pose + ((long long)b * a.V + v) * 12;
        const float rx = fmaf(__ldg(P + 2), 1.f, fmaf(__ldg(P + 1), fy, __ldg(P + 0) * fx));
        const float ry = fmaf(__ldg(P + 5), 1.f, fmaf(__ldg(P + 4), fy, __ldg(P + 3) * fx));
        const float rz = fmaf(__ldg(P + 8), 1.f, fmaf(__ldg(P + 7), fy, __ldg(P + 6) * fx));
        const float tx = __ldg(P + 9), ty = __ldg(P + 10), tz = __ldg(P + 11);
        const long long view_off = (long long)b * a.Hs * a.Ws * C + g * CPG;
        const float* S = a.src[v] + view_off;

        // forward again: warped features and the lane's group correlation for every hypothesis
        float warped[CPG][D], cor[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const Taps t = make_taps(rx, ry, rz, tx, ty, tz, dep[d], half_w, half_h, lim_x, lim_y, a.Hs, a.Ws);
            const float* base = S + t.off * C;
            float t_nw[CPG], t_ne[CPG], t_sw[CPG], t_se[CPG];
#pragma unroll
            for (int c = 0; c < CPG; ++c) t_nw[c] = t_ne[c] = t_sw[c] = t_se[c] = 0.f;
            if (t.nw) load_vec<CPG>(base, t_nw);
            if (t.ne) load_vec<CPG>(base + C, t_ne);
            if (t.sw) load_vec<CPG>(base + row, t_sw);
            if (t.se) load_vec<CPG>(base + row + C, t_se);
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < CPG; ++c) {
                warped[c][d] = t_nw[c] * t.w_nw + t_ne[c] * t.w_ne + t_sw[c] * t.w_sw + t_se[c] * t.w_se;
                dot += ref[c] * warped[c][d];
            }
            cor[d] = dot * (1.f / CPG);
        }
        // the view's attention over the hypotheses (:1053) and its gradient
        float p[D], m = -INFINITY;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            p[d] = __fdiv_rn(group_sum<G>(cor[d]), a.attn_temp);
            m = fmaxf(m, p[d]);
        }
        float se = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) { p[d] = expf(p[d] - m); se += p[d]; }
        float ga[D], mean_ga = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            p[d] = __fdiv_rn(p[d], se);
            ga[d] = group_sum<G>(gacc[d] * cor[d]) + gws[d];
            mean_ga += ga[d] * p[d];
        }
        float* GS = a.gsrc[v] ? a.gsrc[v] + view_off : nullptr;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float aw = __fdiv_rn(p[d], a.sqrt_c);
            const float gcor = aw * gacc[d] + aw * (ga[d] - mean_ga) / a.attn_temp;
            const float k = gcor * (1.f / CPG);   // through the mean over the group's channels (:1040)
#pragma unroll
            for (int c = 0; c < CPG; ++c) gref[c] += k * warped[c][d];
            if (GS && live) {
                float gw[CPG];
#pragma unroll
                for (int c = 0; c < CPG; ++c) gw[c] = k * ref[c];
                const Taps t = make_taps(rx, ry, rz, tx, ty, tz, dep[d], half_w, half_h, lim_x, lim_y, a.Hs, a.Ws);
                float* base = GS + t.off * C;
                if (t.nw) scatter_add<CPG>(base, gw, t.w_nw);
                if (t.ne) scatter_add<CPG>(base + C, gw, t.w_ne);
                if (t.sw) scatter_add<CPG>(base + row, gw, t.w_sw);
                if (t.se) scatter_add<CPG>(base + row + C, gw, t.w_se);
            }
        }
    }
    if (live) store_vec<CPG>(a.gref + pix * C + g * CPG, gref);
}

template <int CPG, int G, int D>
static int launch_bwd(const EtBwdArgs& a, cudaStream_t st) {
    const long long threads = (long long)a.B * a.H * a.W * G;
    et_fuse_bwd_kernel<CPG, G, D><<<ceil_div(threads, 256), 256, 0, st>>>(a);
    return check_launch("et_fuse_bwd_kernel");
}

template <int CPG, int G>
static int bwd_dispatch_d(const EtBwdArgs& a, int D, cudaStream_t st) {
    if (D == 4) return launch_bwd<CPG, G, 4>(a, st);
    if (D == 8) return launch_bwd<CPG, G, 8>(a, st);
    set_error("mvster_et_fuse_bwd_f32: unsupported D=%d (4 or 8)", D);
    return MVSTER_ERR_ARG;
}

template <int G>
static int bwd_dispatch_cpg(const EtBwdArgs& a, int cpg, int D, cudaStream_t st) {
    switch (cpg) {
        case 1: return bwd_dispatch_d<1, G>(a, D, st);
        case 2: return bwd_dispatch_d<2, G>(a, D, st);
        case 4: return bwd_dispatch_d<4, G>(a, D, st);
        case 8: return bwd_dispatch_d<8, G>(a, D, st);
    }
    set_error("mvster_et_fuse_bwd_f32: unsupported C/G=%d (1,2,4,8)", cpg);
    return MVSTER_ERR_ARG;
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_et_fuse_bwd_f32(const float* ref, const float* const* src_host, int V, const float* pose,
                                      const float* hypo, const float* cost, const float* wsum, const float* grad_cost,
                                      float* grad_ref, float* const* grad_src_host,
                                      int B, int C, int G, int D, int H, int W, int Hs, int Ws,
                                      float attn_temp, mvster_stream_t stream) {
    MVSTER_REQUIRE(ref && src_host && pose && hypo && cost && wsum && grad_cost && grad_ref && grad_src_host,
                   "mvster_et_fuse_bwd_f32: null pointer");
    MVSTER_REQUIRE(V >= 1 && V <= MVSTER_MAX_VIEWS, "mvster_et_fuse_bwd_f32: V=%d outside 1..%d", V, MVSTER_MAX_VIEWS);
    MVSTER_REQUIRE(B > 0 && H > 0 && W > 0 && Hs > 0 && Ws > 0, "mvster_et_fuse_bwd_f32: bad shape");
    MVSTER_REQUIRE(G == 4 || G == 8, "mvster_et_fuse_bwd_f32: unsupported G=%d (4 or 8)", G);
    MVSTER_REQUIRE(C % G == 0, "mvster_et_fuse_bwd_f32: C=%d not divisible by G=%d", C, G);
    MVSTER_REQUIRE(attn_temp != 0.f, "mvster_et_fuse_bwd_f32: attn_temp == 0");
    EtBwdArgs a;
    a.ref = ref;
    for (int v = 0; v < MVSTER_MAX_VIEWS; ++v) {
        a.src[v] = v < V ? src_host[v] : nullptr;
        a.gsrc[v] = v < V ? grad_src_host[v] : nullptr;
    }
    for (int v = 0; v < V; ++v) MVSTER_REQUIRE(a.src[v], "mvster_et_fuse_bwd_f32: src[%d] is null", v);
    a.pose = pose; a.hypo = hypo; a.cost = cost; a.wsum = wsum; a.gcost = grad_cost; a.gref = grad_ref;
    a.B = B; a.V = V; a.H = H; a.W = W; a.Hs = Hs; a.Ws = Ws;
    a.attn_temp = attn_temp;
    a.sqrt_c = (float)sqrt((double)C);
    cudaStream_t st = (cudaStream_t)stream;
    return G == 4 ? bwd_dispatch_cpg<4>(a, C / G, D, st) : bwd_dispatch_cpg<8>(a, C / G, D, st);
}
