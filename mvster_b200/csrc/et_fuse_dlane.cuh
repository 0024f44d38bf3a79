// Warp + epipolar-Transformer kernel for the D = 4 stages (cascade stages 3 and 4: 80 % of the path's bytes),
// "depth across lanes" mapping.
//
// et_fuse_tiled_kernel gives one lane a pixel and unrolls its D hypotheses; a warp-wide tap load then touches 32
// different source pixels (14 L1 wavefronts per 256-bit request once the 32-byte pixels straddle 128-byte lines) and the
// LSU data pipe, not HBM, bounds the kernel (profiles/r01_et_fuse_tiled_v3_ncu.md).  At stages 3-4 the D hypotheses of a
// pixel project within ~1.3 source pixels of each other, so here the 4 hypotheses of a pixel sit on 4
// adjacent lanes (x LPP lanes of 8 channels): a warp covers 8 (C = 8) or 4 (C = 16) consecutive pixels, its tap loads fall
// into ~5 cache lines instead of 14 because lanes that hit the same source pixel share the wavefront, the geometry is
// evaluated exactly once per (pixel, hypothesis), nothing is unrolled over D (70 registers -> twice the resident warps),
// and the softmax over D becomes two warp shuffles.
#pragma once

namespace mvster {

template <int C, int G, int LPP>
__global__ void __launch_bounds__(128, 6) et_fuse_dlane_kernel(const EtArgs a) {
    constexpr int D = 4;
    constexpr int GPL = G / LPP;    // groups per lane
    constexpr int CPG = C / G;      // channels per group
    constexpr int LPX = D * LPP;    // lanes per pixel
    constexpr int PXW = 32 / LPX;   // pixels per warp
    static_assert(C / LPP == 8 && GPL >= 1 && CPG % 2 == 0 && CPG <= 8, "a lane owns exactly 8 channels");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPP, d = (lane / LPP) % D;
    int x = blockIdx.x * PXW + lane / LPX;
    int y = blockIdx.y * 4 + warp;
    const int b = blockIdx.z;
    const bool live = x < a.W && y < a.H;
    x = min(x, a.W - 1);
    y = min(y, a.H - 1);
    const int plane = a.H * a.W, pix = y * a.W + x;

    unsigned long long ref[4];
    {
        const Pix8 t = ldg256(a.ref + ((long long)b * plane + pix) * C + sub * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) ref[i] = mul2(t.p[i], pack2(1.f / CPG, 1.f / CPG));  // fold the 1/CPG of .mean(2) (exact)
    }
    const long long o_d = ((long long)b * D + d) * plane + pix;  // this lane's (b, d, y, x)
    const float dep = __ldg(a.hypo + o_d);
    float acc[GPL], ws;
    if (a.flags & MVSTER_ET_ACCUMULATE) {
        ws = a.wsum[o_d];
#pragma unroll
        for (int g = 0; g < GPL; ++g) acc[g] = a.cost[o_d * G + sub * GPL + g];
    } else {
        ws = (a.flags & MVSTER_ET_PARTIAL) ? 0.f : 1e-8f;
#pragma unroll
        for (int g = 0; g < GPL; ++g) acc[g] = 0.f;
    }

    const float fx = (float)x, fy = (float)y;
    const float max_x = (float)(a.Ws - 1), max_y = (float)(a.Hs - 1);
    const float inv_temp_log2e = 1.4426950408889634f / a.attn_temp;
    const int row = a.Ws * C;
    const int lane_base = b * a.Hs * row + sub * 8;  // < 2^31 (checked on the host)

    for (int v = 0; v < a.V; ++v) {
        const float* P = a.pose + ((long long)b * a.V + v) * 12;
        const float rx = fmaf(__ldg(P + 2), 1.f, fmaf(__ldg(P + 1), fy, __ldg(P + 0) * fx));
        const float ry = fmaf(__ldg(P + 5), 1.f, fmaf(__ldg(P + 4), fy, __ldg(P + 3) * fx));
        const float rz = fmaf(__ldg(P + 8), 1.f, fmaf(__ldg(P + 7), fy, __ldg(P + 6) * fx));
        const float* S = a.src[v];
        const float X = __fadd_rn(__fmul_rn(rx, dep), __ldg(P + 9));
        const float Y = __fadd_rn(__fmul_rn(ry, dep), __ldg(P + 10));
        float Z = __fadd_rn(__fmul_rn(rz, dep), __ldg(P + 11));
        if (Z == 0.f) Z = 1e-9f;
        const float rZ = rcp_approx(Z);
        const float ix = div_corrected(X, Z, rZ), iy = div_corrected(Y, Z, rZ);
        int o_nw, o_ne, o_sw, o_se;
        float w_nw, w_ne, w_sw, w_se;
        const bool interior = ix >= 0.f && ix < max_x && iy >= 0.f && iy < max_y;  // false for NaN
        if (__all_sync(0xffffffffu, interior)) {
            const float x0f = floorf(ix), y0f = floorf(iy);
            const float wx = ix - x0f, wy = iy - y0f, ex = 1.f - wx, ey = 1.f - wy;
            o_nw = lane_base + (int)y0f * row + (int)x0f * C;
            o_ne = o_nw + C; o_sw = o_nw + row; o_se = o_sw + C;
            w_nw = ey * ex; w_ne = ey * wx; w_sw = wy * ex; w_se = wy * wx;
        } else {  // zeros padding per tap: clamp the address, zero the weight
            const float cx = fminf(fmaxf(ix, -2.f), max_x + 2.f), cy = fminf(fmaxf(iy, -2.f), max_y + 2.f);
            const float x0f = floorf(cx), y0f = floorf(cy);
            const float wx = cx - x0f, wy = cy - y0f;
            const int x0 = (int)x0f, y0 = (int)y0f;
            const float ex = (unsigned)x0 < (unsigned)a.Ws ? 1.f - wx : 0.f, fxw = (unsigned)(x0 + 1) < (unsigned)a.Ws ? wx : 0.f;
            const float ey = (unsigned)y0 < (unsigned)a.Hs ? 1.f - wy : 0.f, fyw = (unsigned)(y0 + 1) < (unsigned)a.Hs ? wy : 0.f;
            const int xa = min(max(x0, 0), a.Ws - 1) * C, xb = min(max(x0 + 1, 0), a.Ws - 1) * C;
            const int ya = lane_base + min(max(y0, 0), a.Hs - 1) * row, yb = lane_base + min(max(y0 + 1, 0), a.Hs - 1) * row;
            o_nw = ya + xa; o_ne = ya + xb; o_sw = yb + xa; o_se = yb + xb;
            w_nw = ey * ex; w_ne = ey * fxw; w_sw = fyw * ex; w_se = fyw * fxw;
        }
        const Pix8 t_nw = ldg256(S + o_nw), t_ne = ldg256(S + o_ne), t_sw = ldg256(S + o_sw), t_se = ldg256(S + o_se);
        const unsigned long long k_nw = pack2(w_nw, w_nw), k_ne = pack2(w_ne, w_ne), k_sw = pack2(w_sw, w_sw), k_se = pack2(w_se, w_se);
        unsigned long long prod[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned long long wv = mul2(t_nw.p[i], k_nw);
            wv = fma2(t_ne.p[i], k_ne, wv);
            wv = fma2(t_sw.p[i], k_sw, wv);
            wv = fma2(t_se.p[i], k_se, wv);
            prod[i] = mul2(ref[i], wv);
        }
        float cor[GPL];
        if constexpr (CPG == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 q = unpack2(prod[i]); cor[i] = q.x + q.y; }
        } else if constexpr (CPG == 4) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float2 q0 = unpack2(prod[2 * g]), q1 = unpack2(prod[2 * g + 1]);
                cor[g] = (q0.x + q0.y) + (q1.x + q1.y);
            }
        } else {
            const float2 q0 = unpack2(prod[0]), q1 = unpack2(prod[1]), q2 = unpack2(prod[2]), q3 = unpack2(prod[3]);
            cor[0] = ((q0.x + q0.y) + (q1.x + q1.y)) + ((q2.x + q2.y) + (q3.x + q3.y));
        }
        // softmax over the 4 hypotheses of this pixel (lanes d = 0..3) of sum_g cor / temp, then / sqrt(C)  (:1053)
        float s = cor[0];
#pragma unroll
        for (int g = 1; g < GPL; ++g) s += cor[g];
#pragma unroll
        for (int o = 1; o < LPP; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float lg = s * inv_temp_log2e;
        float m = fmaxf(lg, __shfl_xor_sync(0xffffffffu, lg, LPP));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2 * LPP));
        const float e = exp2f(lg - m);
        float se = e + __shfl_xor_sync(0xffffffffu, e, LPP);
        se += __shfl_xor_sync(0xffffffffu, se, 2 * LPP);
        const float rs = rcp_approx(se * a.sqrt_c);
        const float w = e * rs;
        ws += w;
#pragma unroll
        for (int g = 0; g < GPL; ++g) acc[g] = fmaf(w, cor[g], acc[g]);
    }

    if (!live) return;
    const bool partial = a.flags & MVSTER_ET_PARTIAL;
    const float r = partial ? 1.f : __frcp_rn(ws);
    float* dst = a.cost + o_d * G + sub * GPL;
    if constexpr (GPL == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0] * r, acc[1] * r, acc[2] * r, acc[3] * r);
    } else if constexpr (GPL == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(acc[0] * r, acc[1] * r);
    } else {
        dst[0] = acc[0] * r;
    }
    if (partial && sub == 0) a.wsum[o_d] = ws;
}

template <int C, int G, int LPP>
static int launch_et_dlane(const EtArgs& a, cudaStream_t st) {
    dim3 grid(ceil_div(a.W, 32 / (4 * LPP)), ceil_div(a.H, 4), a.B);
    et_fuse_dlane_kernel<C, G, LPP><<<grid, 128, 0, st>>>(a);
    note_et_kernel("et_fuse_dlane_kernel", C, G, 4, LPP, 0);
    return check_launch("et_fuse_dlane_kernel");
}

// MEASURED (B200, cfg2, profiles/r01_et_fuse_dlane_ncu.md): global-load sectors drop 2.5x (18.3 M -> 7.3 M at stage 4) as
// intended, but a 256-bit warp load is executed as 8 four-lane passes and lanes of different passes do not share a
// wavefront, so the request count (x2.3) keeps the LSU data pipe just as busy (66 %): 31.8 / 57.4 us vs 29.3 / 55.3 us for
// et_fuse_tiled_kernel at stages 3 / 4.  Kept as an opt-in experiment: MVSTER_ET_DLANE=1 selects it for the D = 4 stages.
static bool try_launch_dlane(const EtArgs& a, int C, int G, int D, cudaStream_t st, int* rc) {
    if (D != 4 || (long long)a.B * a.Hs * a.Ws * C >= (1ll << 31) || a.B > 65535) return false;
    const char* e = getenv("MVSTER_ET_DLANE");
    if (!e || e[0] != '1') return false;
    if (C == 8 && G == 4) { *rc = launch_et_dlane<8, 4, 1>(a, st); return true; }
    if (C == 16 && G == 4) { *rc = launch_et_dlane<16, 4, 2>(a, st); return true; }
    return false;
}

}  // namespace mvster
