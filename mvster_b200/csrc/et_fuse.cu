// Fused plane-sweep homography warp + epipolar-Transformer aggregation (fp32).
//
// Replaces, for one cascade stage and all source views in ONE launch:
//   homo_warping                      models/mvs4net_utils.py:13-59
//   group correlation                 models/mvs4net_utils.py:1037-1040
//   softmax_D(sum_g cor / temp)/sqrt(C)  :1053
//   running weighted sums over views  :1054-1055  and the final division :1060
// The warped volume [B,C,D,H,W], the replicated reference volume and the per-view
// correlation / weight tensors of the reference are never materialised: every
// feature element is read from HBM once (plus L1/L2-served bilinear re-reads) and
// the cost volume is written once.  HBM-bandwidth bound; no tensor cores (per pixel
// it is a C/G-long dot product per (group, hypothesis)).
//
// Thread mapping: G consecutive lanes own one reference pixel, lane g owns the C/G
// channels of group g (channels-last features => the G lanes of a pixel read one
// contiguous C*4-byte run per bilinear tap, 128-bit loads when C/G >= 4).  All D
// hypotheses of the pixel stay in registers; the sum over groups that feeds the
// softmax is a log2(G)-step warp shuffle.
#include "common.cuh"
#include "et_args.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace mvster {

// name of the kernel the last mvster_et_fuse_f32 call of this thread dispatched to (mvster_et_last_kernel)
static thread_local char g_et_kernel[112] = "";
static void note_et_kernel(const char* family, int C, int G, int D, int lpp, int mb) {
    snprintf(g_et_kernel, sizeof g_et_kernel, "%s<C=%d,G=%d,D=%d,LPP=%d,MB=%d>", family, C, G, D, lpp, mb);
}

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

template <int G>
__device__ __forceinline__ float group_sum(float s) {
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if constexpr (G >= 8) s += __shfl_xor_sync(0xffffffffu, s, 4);
    return s;
}

// GROUP = true : cost[g] = mean over the lane's CPG channels of ref*warped (group correlation, :1037-1040)
// GROUP = false: cost[c] = (ref[c] - warped[c])^2 per channel (:1042); the G lanes of a pixel then
//                own CPG output channels each and the cost volume has C = G*CPG channels.
template <int CPG, int G, int D, bool GROUP>
__global__ void __launch_bounds__(256) et_fuse_kernel(const EtArgs a) {
    constexpr int C = CPG * G;
    constexpr int NOUT = GROUP ? 1 : CPG;   // cost channels per lane
    constexpr int CO = GROUP ? G : C;       // cost channels per voxel
    const int g = threadIdx.x % G;
    const long long npix = (long long)a.B * a.H * a.W;
    long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const bool live = pix < npix;
    if (!live) pix = npix - 1;  // keep the lane for the shuffles, drop its stores
    const int x = (int)(pix % a.W);
    const int y = (int)((pix / a.W) % a.H);
    const int b = (int)(pix / ((long long)a.W * a.H));
    const long long plane = (long long)a.H * a.W;
    const long long pix_in_b = (long long)y * a.W + x;

    float ref[CPG];
    load_vec<CPG>(a.ref + pix * C + g * CPG, ref);

    float dep[D], acc[NOUT][D], ws[D];
#pragma unroll
    for (int d = 0; d < D; ++d) dep[d] = __ldg(a.hypo + ((long long)b * D + d) * plane + pix_in_b);
    if (a.flags & MVSTER_ET_ACCUMULATE) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const long long o = ((long long)b * D + d) * plane + pix_in_b;
#pragma unroll
            for (int n = 0; n < NOUT; ++n) acc[n][d] = a.cost[o * CO + g * NOUT + n];
            ws[d] = a.wsum[o];
        }
    } else {
        const float seed = (a.flags & MVSTER_ET_PARTIAL) ? 0.f : 1e-8f;  // mvs4net_utils.py:1022
#pragma unroll
        for (int d = 0; d < D; ++d) {
            ws[d] = seed;
#pragma unroll
            for (int n = 0; n < NOUT; ++n) acc[n][d] = 0.f;
        }
    }
    const bool fuse_d = !(a.flags & MVSTER_ET_NO_FUSE_D);

    const float fx = (float)x, fy = (float)y;
    const float half_w = 0.5f * (float)(a.Ws - 1);  // (Ws-1)/2, exact in fp32
    const float half_h = 0.5f * (float)(a.Hs - 1);
    const float lim_x = (float)a.Ws + 1.f, lim_y = (float)a.Hs + 1.f;

    for (int v = 0; v < a.V; ++v) {
        const float* P = a.pose + ((long long)b * a.V + v) * 12;
        // ray = R @ (x, y, 1): k-ordered FMA chain like a K=3 sgemm (mvs4net_utils.py:33)
        const float rx = fmaf(__ldg(P + 2), 1.f, fmaf(__ldg(P + 1), fy, __ldg(P + 0) * fx));
        const float ry = fmaf(__ldg(P + 5), 1.f, fmaf(__ldg(P + 4), fy, __ldg(P + 3) * fx));
        const float rz = fmaf(__ldg(P + 8), 1.f, fmaf(__ldg(P + 7), fy, __ldg(P + 6) * fx));
        const float tx = __ldg(P + 9), ty = __ldg(P + 10), tz = __ldg(P + 11);
        const float* S = a.src[v] + (long long)b * a.Hs * a.Ws * C + g * CPG;

        float cor[NOUT][D], tot[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            // separately rounded mul / add / div, as the reference's tensor ops (:34-45)
            const float X = __fadd_rn(__fmul_rn(rx, dep[d]), tx);
            const float Y = __fadd_rn(__fmul_rn(ry, dep[d]), ty);
            float Z = __fadd_rn(__fmul_rn(rz, dep[d]), tz);
            if (Z == 0.f) Z = 1e-9f;  // :38-39
            const float gx = __fsub_rn(__fdiv_rn(__fdiv_rn(X, Z), half_w), 1.f);  // :43
            const float gy = __fsub_rn(__fdiv_rn(__fdiv_rn(Y, Z), half_h), 1.f);  // :44
            // grid_sample(align_corners=True) un-normalisation: (g + 1) * (size-1)/2
            float ix = __fmul_rn(__fadd_rn(gx, 1.f), half_w);
            float iy = __fmul_rn(__fadd_rn(gy, 1.f), half_h);
            // outside [-1, size] no tap is valid; clamp so the int conversion is defined
            ix = fminf(fmaxf(ix, -2.f), lim_x);
            iy = fminf(fmaxf(iy, -2.f), lim_y);
            const float x0f = floorf(ix), y0f = floorf(iy);
            const float wx = ix - x0f, wy = iy - y0f;
            const float ex = 1.f - wx, ey = 1.f - wy;
            const int x0 = (int)x0f, y0 = (int)y0f;
            const bool vx0 = (unsigned)x0 < (unsigned)a.Ws, vx1 = (unsigned)(x0 + 1) < (unsigned)a.Ws;
            const bool vy0 = (unsigned)y0 < (unsigned)a.Hs, vy1 = (unsigned)(y0 + 1) < (unsigned)a.Hs;
            const float* base = S + ((long long)y0 * a.Ws + x0) * C;
            float t_nw[CPG], t_ne[CPG], t_sw[CPG], t_se[CPG];
#pragma unroll
            for (int c = 0; c < CPG; ++c) t_nw[c] = t_ne[c] = t_sw[c] = t_se[c] = 0.f;  // zeros padding, per tap
            if (vy0 && vx0) load_vec<CPG>(base, t_nw);
            if (vy0 && vx1) load_vec<CPG>(base + C, t_ne);
            if (vy1 && vx0) load_vec<CPG>(base + (long long)a.Ws * C, t_sw);
            if (vy1 && vx1) load_vec<CPG>(base + (long long)a.Ws * C + C, t_se);
            const float w_nw = ey * ex, w_ne = ey * wx, w_sw = wy * ex, w_se = wy * wx;
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < CPG; ++c) {
                const float warped = t_nw[c] * w_nw + t_ne[c] * w_ne + t_sw[c] * w_sw + t_se[c] * w_se;
                if constexpr (GROUP) {
                    dot += ref[c] * warped;
                } else {
                    const float df = __fsub_rn(ref[c], warped);
                    cor[c][d] = __fmul_rn(df, df);  // :1042
                    dot += cor[c][d];
                }
            }
            if constexpr (GROUP) { cor[0][d] = dot * (1.f / CPG); tot[d] = cor[0][d]; }  // .mean(2), :1040
            else tot[d] = dot;
        }

        // attention over the D hypotheses of this pixel: softmax_d(sum over all cost channels [/ temp]) (:1049-1053)
        float lg[D], m = -INFINITY;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float s = group_sum<G>(tot[d]);
            lg[d] = fuse_d ? __fdiv_rn(s, a.attn_temp) : s;
            m = fmaxf(m, lg[d]);
        }
        float se = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) { lg[d] = expf(lg[d] - m); se += lg[d]; }
        float wmax = 0.f;  // attn_fuse_d=False: one scalar weight per pixel and view = max_d softmax (:1049)
#pragma unroll
        for (int d = 0; d < D; ++d) wmax = fmaxf(wmax, __fdiv_rn(lg[d], se));
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float w = fuse_d ? __fdiv_rn(__fdiv_rn(lg[d], se), a.sqrt_c) : wmax;
            ws[d] = __fadd_rn(ws[d], w);                                  // :1050 / :1054
#pragma unroll
            for (int n = 0; n < NOUT; ++n) acc[n][d] = __fadd_rn(acc[n][d], __fmul_rn(w, cor[n][d]));  // :1051 / :1055
        }
    }

    if (!live) return;
    const bool partial = a.flags & MVSTER_ET_PARTIAL;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const long long o = ((long long)b * D + d) * plane + pix_in_b;
#pragma unroll
        for (int n = 0; n < NOUT; ++n) a.cost[o * CO + g * NOUT + n] = partial ? acc[n][d] : __fdiv_rn(acc[n][d], ws[d]);  // :1058-1060
        if (partial && g == 0) a.wsum[o] = ws[d];
    }
}

__global__ void et_normalize_kernel(float* __restrict__ cost, const float* __restrict__ wsum, long long n, int G) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) cost[i] = __fdiv_rn(cost[i], __fadd_rn(1e-8f, __ldg(wsum + i / G)));
}

}  // namespace mvster
#include "et_fuse_tiled.cuh"
#include "et_fuse_win.cuh"
#include "et_fuse_tma.cuh"
namespace mvster {

template <int CPG, int G, int D>
static int launch_et(const EtArgs& a, cudaStream_t st) {
    const long long threads = (long long)a.B * a.H * a.W * G;
    if (a.flags & MVSTER_ET_SQDIFF) et_fuse_kernel<CPG, G, D, false><<<ceil_div(threads, 256), 256, 0, st>>>(a);
    else et_fuse_kernel<CPG, G, D, true><<<ceil_div(threads, 256), 256, 0, st>>>(a);
    note_et_kernel("et_fuse_kernel", CPG * G, G, D, G, 0);
    return check_launch("et_fuse_kernel");
}

template <int CPG, int G>
static int dispatch_d(const EtArgs& a, int D, cudaStream_t st) {
    if (D == 4) return launch_et<CPG, G, 4>(a, st);
    if (D == 8) return launch_et<CPG, G, 8>(a, st);
    set_error("mvster_et_fuse_f32: unsupported D=%d (4 or 8)", D);
    return MVSTER_ERR_ARG;
}

template <int G>
static int dispatch_cpg(const EtArgs& a, int cpg, int D, cudaStream_t st) {
    switch (cpg) {
        case 1: return dispatch_d<1, G>(a, D, st);
        case 2: return dispatch_d<2, G>(a, D, st);
        case 4: return dispatch_d<4, G>(a, D, st);
        case 8: return dispatch_d<8, G>(a, D, st);
    }
    set_error("mvster_et_fuse_f32: unsupported C/G=%d (1,2,4,8)", cpg);
    return MVSTER_ERR_ARG;
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_et_fuse_f32(const float* ref, const float* const* src_host, int V, const float* pose,
                                  const float* hypo, float* cost, float* wsum,
                                  int B, int C, int G, int D, int H, int W, int Hs, int Ws,
                                  float attn_temp, int flags, mvster_stream_t stream) {
    MVSTER_REQUIRE(ref && src_host && pose && hypo && cost, "mvster_et_fuse_f32: null pointer");
    MVSTER_REQUIRE(V >= 1 && V <= MVSTER_MAX_VIEWS, "mvster_et_fuse_f32: V=%d outside 1..%d", V, MVSTER_MAX_VIEWS);
    MVSTER_REQUIRE(B > 0 && H > 0 && W > 0 && Hs > 0 && Ws > 0, "mvster_et_fuse_f32: bad shape");
    if (flags & MVSTER_ET_SQDIFF) {  // per-channel cost: the cost volume has C channels; pick the lane count here
        MVSTER_REQUIRE(G == C, "mvster_et_fuse_f32: with MVSTER_ET_SQDIFF the cost volume has C channels (pass G == C)");
        MVSTER_REQUIRE(C == 8 || C == 16 || C == 32 || C == 64, "mvster_et_fuse_f32: SQDIFF supports C in {8,16,32,64}, got %d", C);
        G = C >= 32 ? 8 : 4;
    }
    MVSTER_REQUIRE(G == 4 || G == 8, "mvster_et_fuse_f32: unsupported G=%d (4 or 8)", G);
    MVSTER_REQUIRE(C % G == 0, "mvster_et_fuse_f32: C=%d not divisible by G=%d", C, G);
    MVSTER_REQUIRE(!(flags & (MVSTER_ET_PARTIAL | MVSTER_ET_ACCUMULATE)) || wsum,
                   "mvster_et_fuse_f32: wsum required with PARTIAL/ACCUMULATE");
    MVSTER_REQUIRE(attn_temp != 0.f, "mvster_et_fuse_f32: attn_temp == 0");
    EtArgs a;
    a.ref = ref;
    for (int v = 0; v < MVSTER_MAX_VIEWS; ++v) a.src[v] = v < V ? src_host[v] : nullptr;
    for (int v = 0; v < V; ++v) MVSTER_REQUIRE(a.src[v], "mvster_et_fuse_f32: src[%d] is null", v);
    a.pose = pose; a.hypo = hypo; a.cost = cost; a.wsum = wsum;
    a.B = B; a.V = V; a.H = H; a.W = W; a.Hs = Hs; a.Ws = Ws;
    a.attn_temp = attn_temp;
    a.sqrt_c = (float)sqrt((double)C);  // math.sqrt(C) -> fp32 scalar
    a.flags = flags;
    { const char* pf = getenv("MVSTER_ET_PREFETCH"); a.prefetch = pf ? atoi(pf) : 1; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = MVSTER_OK;
    const bool plain = !(flags & (MVSTER_ET_GENERIC | MVSTER_ET_SQDIFF | MVSTER_ET_NO_FUSE_D));
    const bool window = (flags & MVSTER_ET_NO_WINDOW) ? false : (flags & MVSTER_ET_WINDOW) ? true : et_window_default();
    const bool tma = (flags & MVSTER_ET_TMA_OFF) ? false : (flags & MVSTER_ET_TMA_ON) ? true : ettma::et_tma_default();
    if (plain && window && tma && ettma::try_launch_tma(a, C, G, D, st, &rc)) return rc;  // stages 2-4: TMA-staged source boxes
    if (plain && window && try_launch_win(a, C, G, D, st, &rc)) return rc;  // same arithmetic, taps gathered from global memory
    MVSTER_REQUIRE(!(flags & MVSTER_ET_INTERLEAVED) || C / G == 8,
                   "mvster_et_fuse_f32: MVSTER_ET_INTERLEAVED is implemented by the window kernels only ((C,G,D) = (8,4,4), (16,4,4), (32,8,8), "
                   "no GENERIC / NO_WINDOW / SQDIFF / NO_FUSE_D); got C=%d G=%d D=%d flags=%d", C, G, D, flags);
    if (plain && try_launch_tiled(a, C, G, D, st, &rc)) return rc;   // hypotheses unrolled per lane, taps gathered per hypothesis
    return G == 4 ? dispatch_cpg<4>(a, C / G, D, st) : dispatch_cpg<8>(a, C / G, D, st);
}

// bf16 storage (BASELINE configs[2]): features and cost volume are bf16 in HBM, everything in between is the fp32 arithmetic
// of the kernels above (geometry, correlations, softmax, weighted sums); the cost is rounded once, on the way out.
extern "C" int mvster_et_fuse_bf16(const void* ref, const void* const* src_host, int V, const float* pose,
                                   const float* hypo, void* cost, int B, int C, int G, int D, int H, int W, int Hs, int Ws,
                                   float attn_temp, int flags, mvster_stream_t stream) {
    MVSTER_REQUIRE(ref && src_host && pose && hypo && cost, "mvster_et_fuse_bf16: null pointer");
    MVSTER_REQUIRE(V >= 1 && V <= MVSTER_MAX_VIEWS, "mvster_et_fuse_bf16: V=%d outside 1..%d", V, MVSTER_MAX_VIEWS);
    MVSTER_REQUIRE(B > 0 && H > 0 && W > 0 && Hs > 0 && Ws > 0, "mvster_et_fuse_bf16: bad shape");
    MVSTER_REQUIRE(!(flags & ~(MVSTER_ET_WINDOW | MVSTER_ET_NO_WINDOW | MVSTER_ET_INTERLEAVED)),
                   "mvster_et_fuse_bf16: only MVSTER_ET_WINDOW / NO_WINDOW / INTERLEAVED are supported with bf16 storage (flags=%d)", flags);
    MVSTER_REQUIRE(attn_temp != 0.f, "mvster_et_fuse_bf16: attn_temp == 0");
    MVSTER_REQUIRE(((uintptr_t)ref & 15) == 0 && ((uintptr_t)cost & 7) == 0, "mvster_et_fuse_bf16: ref must be 16-byte, cost 8-byte aligned");
    EtArgs a;
    a.ref = reinterpret_cast<const float*>(ref);
    for (int v = 0; v < MVSTER_MAX_VIEWS; ++v) a.src[v] = v < V ? reinterpret_cast<const float*>(src_host[v]) : nullptr;
    for (int v = 0; v < V; ++v) MVSTER_REQUIRE(a.src[v] && ((uintptr_t)a.src[v] & 15) == 0, "mvster_et_fuse_bf16: src[%d] is null or not 16-byte aligned", v);
    a.pose = pose; a.hypo = hypo; a.cost = reinterpret_cast<float*>(cost); a.wsum = nullptr;
    a.B = B; a.V = V; a.H = H; a.W = W; a.Hs = Hs; a.Ws = Ws;
    a.attn_temp = attn_temp;
    a.sqrt_c = (float)sqrt((double)C);
    a.flags = flags;
    { const char* pf = getenv("MVSTER_ET_PREFETCH"); a.prefetch = pf ? atoi(pf) : 1; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = MVSTER_OK;
    const bool window = (flags & MVSTER_ET_NO_WINDOW) ? false : (flags & MVSTER_ET_WINDOW) ? true : et_window_default();
    if (window && try_launch_win_bf16(a, C, G, D, st, &rc)) return rc;
    MVSTER_REQUIRE(!(flags & MVSTER_ET_INTERLEAVED) || C / G == 8,
                   "mvster_et_fuse_bf16: MVSTER_ET_INTERLEAVED needs a window specialisation ((C,G,D) = (8,4,4), (16,4,4), (32,8,8)); got C=%d G=%d D=%d", C, G, D);
    if (try_launch_tiled_bf16(a, C, G, D, st, &rc)) return rc;
    set_error("mvster_et_fuse_bf16: unsupported (C,G,D) = (%d,%d,%d): bf16 storage covers (64,8,8), (32,8,8), (16,4,4), (8,4,4)", C, G, D);
    return MVSTER_ERR_ARG;
}

extern "C" const char* mvster_et_last_kernel(void) { return g_et_kernel; }

extern "C" int mvster_et_normalize_f32(float* cost, const float* wsum, int B, int G, int D, int H, int W,
                                       mvster_stream_t stream) {
    MVSTER_REQUIRE(cost && wsum, "mvster_et_normalize_f32: null pointer");
    const long long n = (long long)B * D * H * W * G;
    MVSTER_REQUIRE(n > 0 && G > 0, "mvster_et_normalize_f32: bad shape");
    et_normalize_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(cost, wsum, n, G);
    return check_launch("et_normalize_kernel");
}
