// Window variant of the fused warp + epipolar-Transformer kernel: "correlate, then interpolate".
//
// Bilinear sampling is linear in the taps, so the group correlation of the reference pixel with
// a warped sample is the same bilinear blend applied to the per-tap correlations
//     cor[g][d] = sum_c ref[c] * (sum_t w_t(d) * tap_t[c]) = sum_t w_t(d) * T_t[g],   T_t[g] = sum_{c in g} ref[c] * tap_t[c].
// From cascade stage 2 on the D hypotheses of a pixel span well under one source pixel
// (the inverse-depth range shrinks ~7x per stage), i.e. they all sample inside a 3 x 3 block of source
// pixels.  Per (pixel, view) the kernel therefore
//   1. evaluates the D sampling positions (same arithmetic as et_fuse_tiled_kernel),
//   2. loads the 2 x 2 block at the smallest floor position plus, only in the lanes whose hypotheses
//      straddle a pixel boundary, the third column / row (4 + <=5 predicated 256-bit loads instead of 4*D),
//   3. reduces every tap to its G/LPP per-group correlations T once (C -> G numbers),
//   4. interpolates T per hypothesis with the piecewise-linear form
//        h(u) = T0 + min(u,1) * (T1 - T0) + max(u-1,0) * (T2 - T1),     u = position - window origin in [0,2)
//      along x for the three rows and then along y - exactly the bilinear blend of the cell the
//      hypothesis falls in, without selects or divergence (zero padding never occurs on this path).
// Warps in which some lane does not fit (a tap outside the source image, hypotheses spanning more
// than two cells, NaN) take the per-hypothesis gather of et_fuse_tiled_kernel for that view
// (warp-uniform branch).  L1 data-pipe traffic per stage-4 pixel falls from 64 to ~18 tap loads.
//
// reference: models/mvs4net_utils.py:13-59 (homo_warping), :1037-1060 (group correlation,
// softmax over D, accumulation over views).
#pragma once

namespace mvster {

#ifdef MVSTER_CPU_EMU
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    const float2 x = emu::unpack(a), y = emu::unpack(b);
    return emu::pack(x.x + y.x, x.y + y.y);
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    const float2 x = emu::unpack(a), y = emu::unpack(b);
    return emu::pack(x.x - y.x, x.y - y.y);
}
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
__device__ __forceinline__ void prefetch_l2(const void*) {}
__device__ __forceinline__ void prefetch_l1(const void*) {}
#else
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float min_nan(float a, float b) {  // NaN-propagating (fminf would drop it)
    float d;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float max_nan(float a, float b) {
    float d;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

// Per-group correlations of one 8-channel tap with the lane's (pre-scaled) reference channels,
// packed two groups per 64-bit pair.  CPG = 2: 4 groups (2 pairs); CPG = 4: 2 groups (1 pair).
// IL (MVSTER_ET_INTERLEAVED): the 8 channels of the block are stored group-interleaved - position p holds channel
// perm[p], perm = {0,2,1,3,4,6,5,7} for CPG = 2 and {0,4,1,5,2,6,3,7} for CPG = 4 - so that a 64-bit pair holds the SAME
// channel index of two neighbouring groups and the group sums need no horizontal adds: 4 packed operations per tap
// instead of 4 packed + 4 scalar + the moves that rebuild the pairs.
template <int CPG, int NJ, bool IL = false>
__device__ __forceinline__ void tap_groups(const Pix8& t, const unsigned long long (&ref)[4], unsigned long long (&out)[NJ]) {
    if constexpr (IL && CPG == 2) {
        static_assert(NJ == 2, "8 channels in groups of 2 = 4 groups");
        out[0] = fma2(ref[1], t.p[1], mul2(ref[0], t.p[0]));
        out[1] = fma2(ref[3], t.p[3], mul2(ref[2], t.p[2]));
    } else if constexpr (IL) {
        static_assert(CPG == 4 && NJ == 1, "8 channels in groups of 4 = 2 groups");
        out[0] = fma2(ref[3], t.p[3], fma2(ref[2], t.p[2], fma2(ref[1], t.p[1], mul2(ref[0], t.p[0]))));
    } else if constexpr (CPG == 2) {
        static_assert(NJ == 2, "8 channels in groups of 2 = 4 groups");
        const float2 m0 = unpack2(mul2(ref[0], t.p[0])), m1 = unpack2(mul2(ref[1], t.p[1]));
        const float2 m2 = unpack2(mul2(ref[2], t.p[2])), m3 = unpack2(mul2(ref[3], t.p[3]));
        out[0] = pack2(m0.x + m0.y, m1.x + m1.y);
        out[1] = pack2(m2.x + m2.y, m3.x + m3.y);
    } else {
        static_assert(CPG == 4 && NJ == 1, "8 channels in groups of 4 = 2 groups");
        const float2 q0 = unpack2(fma2(ref[1], t.p[1], mul2(ref[0], t.p[0])));
        const float2 q1 = unpack2(fma2(ref[3], t.p[3], mul2(ref[2], t.p[2])));
        out[0] = pack2(q0.x + q0.y, q1.x + q1.y);
    }
}

// Request the two window rows a lane will sample in source view v (position of the first hypothesis; the others lie
// within a pixel or two) without holding registers: level 2 = CCTL.PF2 (towards L2), level 1 = CCTL.PF1 (towards L1).
template <int C, int LEVEL, typename FT>
__device__ __forceinline__ void win_prefetch(const float4* pose_s, int v, const FT* src, float fx, float fy, float d0,
                                             float max_x, float max_y, int lane_base, int row) {
    const float4 q0 = pose_s[v * 3], q1 = pose_s[v * 3 + 1], q2 = pose_s[v * 3 + 2];
    const float X = fmaf(fmaf(q0.z, 1.f, fmaf(q0.y, fy, q0.x * fx)), d0, q2.y);
    const float Y = fmaf(fmaf(q1.y, 1.f, fmaf(q1.x, fy, q0.w * fx)), d0, q2.z);
    const float Z = fmaf(fmaf(q2.x, 1.f, fmaf(q1.w, fy, q1.z * fx)), d0, q2.w);
    const float r = rcp_approx(Z);
    const float px = X * r, py = Y * r;
    if (px >= 0.f && px < max_x && py >= 0.f && py < max_y) {
        const FT* p = src + (lane_base + (int)py * row + (int)px * C);
        if constexpr (LEVEL == 2) {
            prefetch_l2(p);
            prefetch_l2(p + row);
        } else {
            prefetch_l1(p);
            prefetch_l1(p + row);
        }
    }
}

template <int C, int G, int D, int LPP, int MB, bool IL = false, bool BF = false>
__global__ void __launch_bounds__(128, MB) et_fuse_win_kernel(const EtArgs a) {
    typedef typename EtTypes<BF>::feat FT;  // BF: bf16 features in, bf16 cost volume out (no PARTIAL / ACCUMULATE: checked on the host)
    constexpr int CPL = C / LPP;   // channels per lane
    constexpr int GPL = G / LPP;   // groups per lane
    constexpr int CPG = C / G;     // channels per group
    constexpr int PXW = 32 / LPP;  // pixels per warp (tile width)
    constexpr int NJ = GPL / 2;    // packed group pairs per lane
    static_assert(CPL == 8 && (GPL == 2 || GPL == 4) && (CPG == 2 || CPG == 4) && D % 2 == 0,
                  "a lane owns 8 channels = 2 or 4 whole groups");
    __shared__ float4 pose_s[MVSTER_MAX_VIEWS * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPP;
    int x = blockIdx.x * PXW + lane / LPP;
    int y = blockIdx.y * 4 + warp;
    const int b = blockIdx.z;
    if (threadIdx.x < a.V * 3)
        pose_s[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(a.pose + (long long)b * a.V * 12) + threadIdx.x);
    const bool live = x < a.W && y < a.H;
    x = min(x, a.W - 1);
    y = min(y, a.H - 1);
    const int plane = a.H * a.W, pix = y * a.W + x;

    unsigned long long ref[4];
    {
        const Pix8 t = ldg256(reinterpret_cast<const FT*>(a.ref) + ((long long)b * plane + pix) * C + sub * CPL);
        // fold the 1/CPG of .mean(2) into the reference features (power of two: exact)
#pragma unroll
        for (int i = 0; i < 4; ++i) ref[i] = mul2(t.p[i], pack2(1.f / CPG, 1.f / CPG));
    }
    unsigned long long dep2[D / 2];  // hypotheses, two per packed pair
    float ws[D];
    unsigned long long acc2[NJ][D];
    {
        const float* hp = a.hypo + (long long)b * D * plane + pix;
#pragma unroll
        for (int k = 0; k < D / 2; ++k)
            dep2[k] = pack2(__ldg(hp + (long long)(2 * k) * plane), __ldg(hp + (long long)(2 * k + 1) * plane));
    }
    if (!BF && (a.flags & MVSTER_ET_ACCUMULATE)) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const long long o = ((long long)b * D + d) * plane + pix;
            ws[d] = a.wsum[o];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float2 t = *reinterpret_cast<const float2*>(a.cost + o * G + sub * GPL + 2 * j);
                acc2[j][d] = pack2(t.x, t.y);
            }
        }
    } else {
        const float seed = (a.flags & MVSTER_ET_PARTIAL) ? 0.f : 1e-8f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            ws[d] = seed;
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc2[j][d] = 0ull;
        }
    }
    __syncthreads();  // pose_s

    const float fx = (float)x, fy = (float)y;
    const float max_x = (float)(a.Ws - 1), max_y = (float)(a.Hs - 1);
    const float inv_temp_log2e = 1.4426950408889634f / a.attn_temp;
    const int row = a.Ws * C;
    const int lane_base = b * a.Hs * row + sub * CPL;  // < 2^31 (checked on the host)
    const float dfirst = unpack2(dep2[0]).x;
    if (a.prefetch & 1)  // every view's window rows towards L2 before the loop
        for (int v = 0; v < a.V; ++v) win_prefetch<C, 2>(pose_s, v, reinterpret_cast<const FT*>(a.src[v]), fx, fy, dfirst, max_x, max_y, lane_base, row);

    for (int v = 0; v < a.V; ++v) {
        if ((a.prefetch & 2) && v + 1 < a.V)  // the next view's rows towards L1 while this view is processed
            win_prefetch<C, 1>(pose_s, v + 1, reinterpret_cast<const FT*>(a.src[v + 1]), fx, fy, dfirst, max_x, max_y, lane_base, row);
        const float4 q0 = pose_s[v * 3], q1 = pose_s[v * 3 + 1], q2 = pose_s[v * 3 + 2];  // R (row-major 3x3), t
        const float rx = fmaf(q0.z, 1.f, fmaf(q0.y, fy, q0.x * fx));
        const float ry = fmaf(q1.y, 1.f, fmaf(q1.x, fy, q0.w * fx));
        const float nrz = -fmaf(q2.x, 1.f, fmaf(q1.w, fy, q1.z * fx));
        const float tx = q2.y, ty = q2.z, ntz = -q2.w;
        const FT* S = reinterpret_cast<const FT*>(a.src[v]);  // warp-uniform base; batch/lane offsets live in the 32-bit tap offsets

        // sampling positions of the D hypotheses, two per packed operation; bit-identical to the scalar
        // sequence X = rx*d + tx (separate mul, add), Z == 0 -> 1e-9, ix = X / Z (reciprocal + FMA residual)
        float ix[D], iy[D];
#pragma unroll
        for (int k = 0; k < D / 2; ++k) {
            // product and sum rounded separately like the reference (rot_xyz * depth, then + trans).  ptxas contracts
            // mul.rn.f32x2 + add.rn.f32x2 (even fma.f32x2 with a zero addend) into one FFMA2, so the products are scalar
            const float2 dd = unpack2(dep2[k]);
            const unsigned long long X2 = add2(pack2(__fmul_rn(rx, dd.x), __fmul_rn(rx, dd.y)), pack2(tx, tx));
            const unsigned long long Y2 = add2(pack2(__fmul_rn(ry, dd.x), __fmul_rn(ry, dd.y)), pack2(ty, ty));
            float2 zn = unpack2(add2(pack2(__fmul_rn(nrz, dd.x), __fmul_rn(nrz, dd.y)), pack2(ntz, ntz)));  // -Z
            if (zn.x == 0.f) zn.x = -1e-9f;
            if (zn.y == 0.f) zn.y = -1e-9f;
            const float r0 = rcp_approx(-zn.x), r1 = rcp_approx(-zn.y);
            const unsigned long long Zn2 = pack2(zn.x, zn.y), r2 = pack2(r0, r1);
            const unsigned long long qx = mul2(X2, r2), qy = mul2(Y2, r2);
            const float2 px = unpack2(fma2(fma2(qx, Zn2, X2), r2, qx));
            const float2 py = unpack2(fma2(fma2(qy, Zn2, Y2), r2, qy));
            ix[2 * k] = px.x; ix[2 * k + 1] = px.y;
            iy[2 * k] = py.x; iy[2 * k + 1] = py.y;
        }
        // window origin = smallest floor over the hypotheses; the window fits when all floors are within one cell of it
        float bxf = floorf(ix[0]), byf = floorf(iy[0]), mxf = bxf, myf = byf;
#pragma unroll
        for (int d = 1; d < D; ++d) {
            const float fxd = floorf(ix[d]), fyd = floorf(iy[d]);
            bxf = min_nan(bxf, fxd); mxf = max_nan(mxf, fxd);
            byf = min_nan(byf, fyd); myf = max_nan(myf, fyd);
        }
        // every tap of the used window inside the source image (mxf + 1 <= Ws - 1), at most two cells per axis; false for NaN
        const bool fits = bxf >= 0.f && mxf < max_x && byf >= 0.f && myf < max_y && (mxf - bxf) <= 1.f && (myf - byf) <= 1.f;

        unsigned long long cor2[NJ][D];
        if (__all_sync(0xffffffffu, fits)) {
            const bool needx = mxf > bxf, needy = myf > byf;
            const FT* p0 = S + (lane_base + (int)byf * row + (int)bxf * C);
            const FT* p1 = p0 + row;
            const Pix8 t00 = ldg256(p0), t01 = ldg256(p0 + C), t10 = ldg256(p1), t11 = ldg256(p1 + C);
            unsigned long long T[3][3][NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) T[0][2][j] = T[1][2][j] = T[2][0][j] = T[2][1][j] = T[2][2][j] = 0ull;
            if (needx) {
                const Pix8 t02 = ldg256(p0 + 2 * C), t12 = ldg256(p1 + 2 * C);
                tap_groups<CPG, NJ, IL>(t02, ref, T[0][2]);
                tap_groups<CPG, NJ, IL>(t12, ref, T[1][2]);
            }
            tap_groups<CPG, NJ, IL>(t00, ref, T[0][0]);
            tap_groups<CPG, NJ, IL>(t01, ref, T[0][1]);
            tap_groups<CPG, NJ, IL>(t10, ref, T[1][0]);
            tap_groups<CPG, NJ, IL>(t11, ref, T[1][1]);
            if (needy) {
                const FT* p2 = p1 + row;
                const Pix8 t20 = ldg256(p2), t21 = ldg256(p2 + C);
                tap_groups<CPG, NJ, IL>(t20, ref, T[2][0]);
                tap_groups<CPG, NJ, IL>(t21, ref, T[2][1]);
                if (needx) {
                    const Pix8 t22 = ldg256(p2 + 2 * C);
                    tap_groups<CPG, NJ, IL>(t22, ref, T[2][2]);
                }
            }
            unsigned long long dx0[3][NJ], dx1[3][NJ];  // column differences
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    dx0[r][j] = sub2(T[r][1][j], T[r][0][j]);
                    dx1[r][j] = sub2(T[r][2][j], T[r][1][j]);
                }
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float ux = ix[d] - bxf, uy = iy[d] - byf;  // exact, in [0,2)
                const float ax = fminf(ux, 1.f), bx = fmaxf(ux - 1.f, 0.f);
                const float ay = fminf(uy, 1.f), by = fmaxf(uy - 1.f, 0.f);
                const unsigned long long ax2 = pack2(ax, ax), bx2 = pack2(bx, bx), ay2 = pack2(ay, ay), by2 = pack2(by, by);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const unsigned long long h0 = fma2(bx2, dx1[0][j], fma2(ax2, dx0[0][j], T[0][0][j]));
                    const unsigned long long h1 = fma2(bx2, dx1[1][j], fma2(ax2, dx0[1][j], T[1][0][j]));
                    const unsigned long long h2 = fma2(bx2, dx1[2][j], fma2(ax2, dx0[2][j], T[2][0][j]));
                    cor2[j][d] = fma2(by2, sub2(h2, h1), fma2(ay2, sub2(h1, h0), h0));
                }
            }
        } else {
            // per-hypothesis gather with zeros padding per tap (et_fuse_tiled_kernel's arithmetic)
#pragma unroll
            for (int d = 0; d < D; ++d) {
                int o_nw, o_ne, o_sw, o_se;
                float w_nw, w_ne, w_sw, w_se;
                const float px = ix[d], py = iy[d];
                const bool interior = px >= 0.f && px < max_x && py >= 0.f && py < max_y;  // false for NaN
                if (__all_sync(0xffffffffu, interior)) {
                    const float x0f = floorf(px), y0f = floorf(py);
                    const float wx = px - x0f, wy = py - y0f, ex = 1.f - wx, ey = 1.f - wy;
                    o_nw = lane_base + (int)y0f * row + (int)x0f * C;
                    o_ne = o_nw + C; o_sw = o_nw + row; o_se = o_sw + C;
                    w_nw = ey * ex; w_ne = ey * wx; w_sw = wy * ex; w_se = wy * wx;
                } else {
                    const float cx = fminf(fmaxf(px, -2.f), max_x + 2.f), cy = fminf(fmaxf(py, -2.f), max_y + 2.f);
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
                Pix8 wv;  // bilinear(warped) per channel pair
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    unsigned long long s = mul2(t_nw.p[i], pack2(w_nw, w_nw));
                    s = fma2(t_ne.p[i], pack2(w_ne, w_ne), s);
                    s = fma2(t_sw.p[i], pack2(w_sw, w_sw), s);
                    wv.p[i] = fma2(t_se.p[i], pack2(w_se, w_se), s);
                }
                unsigned long long g2[NJ];
                tap_groups<CPG, NJ, IL>(wv, ref, g2);
#pragma unroll
                for (int j = 0; j < NJ; ++j) cor2[j][d] = g2[j];
            }
        }

        // softmax over D of (sum over all G groups) / temp, then / sqrt(C)   (mvs4net_utils.py:1053)
        float lg[D], m = -INFINITY;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            unsigned long long t2 = cor2[0][d];
            if constexpr (NJ == 2) t2 = add2(t2, cor2[1][d]);
            const float2 t = unpack2(t2);
            float s = t.x + t.y;
#pragma unroll
            for (int o = 1; o < LPP; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            lg[d] = s * inv_temp_log2e;
            m = fmaxf(m, lg[d]);
        }
        float se = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {  // arguments <= 0: the bare MUFU.EX2 (results below 2^-126 flush to 0)
            lg[d] = ex2_approx(lg[d] - m);
            se += lg[d];
        }
        const float rs = rcp_approx(se * a.sqrt_c);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float w = lg[d] * rs;
            ws[d] += w;
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc2[j][d] = fma2(pack2(w, w), cor2[j][d], acc2[j][d]);
        }
    }

    if (!live) return;
    const bool partial = !BF && (a.flags & MVSTER_ET_PARTIAL);
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const long long o = ((long long)b * D + d) * plane + pix;
        const float r = partial ? 1.f : __frcp_rn(ws[d]);
        if constexpr (BF) {  // bf16 cost volume: the lane's GPL = 2 NJ groups as 16-bit values
            bf16_t* dstb = reinterpret_cast<bf16_t*>(a.cost) + o * G + sub * GPL;
            const float2 u0 = unpack2(acc2[0][d]);
            if constexpr (NJ == 2) {
                const float2 u1 = unpack2(acc2[1][d]);
                *reinterpret_cast<uint2*>(dstb) = make_uint2(f2bf2(u0.x * r, u0.y * r), f2bf2(u1.x * r, u1.y * r));
            } else {
                *reinterpret_cast<uint32_t*>(dstb) = f2bf2(u0.x * r, u0.y * r);
            }
            continue;
        }
        float* dst = a.cost + o * G + sub * GPL;
        if constexpr (NJ == 2) {
            const float2 u0 = unpack2(acc2[0][d]), u1 = unpack2(acc2[1][d]);
            *reinterpret_cast<float4*>(dst) = partial ? make_float4(u0.x, u0.y, u1.x, u1.y)
                                                      : make_float4(u0.x * r, u0.y * r, u1.x * r, u1.y * r);
        } else {
            const float2 u0 = unpack2(acc2[0][d]);
            *reinterpret_cast<float2*>(dst) = partial ? make_float2(u0.x, u0.y) : make_float2(u0.x * r, u0.y * r);
        }
        if (partial && sub == 0) a.wsum[o] = ws[d];
    }
}

// A software-pipelined variant (next view's positions evaluated and its four unconditional taps requested while the current
// view is interpolated) was measured and dropped: it needs 168-196 registers (8-12 warps per SM) and was slower at every
// stage (profiles/r01_et_fuse_win_ncu.md: stage 4 50.9 us vs 47.0 us).

template <int C, int G, int D, int LPP, int MB>
static int launch_et_win(const EtArgs& a, cudaStream_t st) {
    dim3 grid(ceil_div(a.W, 32 / LPP), ceil_div(a.H, 4), a.B);
    if (a.flags & MVSTER_ET_INTERLEAVED) et_fuse_win_kernel<C, G, D, LPP, MB, true><<<grid, 128, 0, st>>>(a);
    else et_fuse_win_kernel<C, G, D, LPP, MB, false><<<grid, 128, 0, st>>>(a);
    note_et_kernel(a.flags & MVSTER_ET_INTERLEAVED ? "et_fuse_win_kernel[interleaved]" : "et_fuse_win_kernel", C, G, D, LPP, MB);
    return check_launch("et_fuse_win_kernel");
}

template <int C, int G, int D, int LPP, int MB>
static int launch_et_win_bf16(const EtArgs& a, cudaStream_t st) {
    dim3 grid(ceil_div(a.W, 32 / LPP), ceil_div(a.H, 4), a.B);
    if (a.flags & MVSTER_ET_INTERLEAVED) et_fuse_win_kernel<C, G, D, LPP, MB, true, true><<<grid, 128, 0, st>>>(a);
    else et_fuse_win_kernel<C, G, D, LPP, MB, false, true><<<grid, 128, 0, st>>>(a);
    note_et_kernel(a.flags & MVSTER_ET_INTERLEAVED ? "et_fuse_win_kernel[bf16, interleaved]" : "et_fuse_win_kernel[bf16]", C, G, D, LPP, MB);
    return check_launch("et_fuse_win_kernel[bf16]");
}

// MVSTER_ET_WIN=0/1 overrides the built-in default (A/B measurements); the MVSTER_ET_WINDOW /
// MVSTER_ET_NO_WINDOW flags of the call win over both.
static bool et_window_default() {
    const char* e = getenv("MVSTER_ET_WIN");
    return e ? atoi(e) != 0 : true;  // measured on B200 (cfg2, tools/et_ab.py): stages 2-4 33.5 / 29.1 / 54.9 -> 24.5 / 24.6 / 45.1 us
}

// Returns true if a window specialisation exists for (C,G,D) and was launched into *rc.
static bool try_launch_win(const EtArgs& a, int C, int G, int D, cudaStream_t st, int* rc) {
    if ((long long)a.B * a.Hs * a.Ws * C >= (1ll << 31) || a.B > 65535) return false;  // 32-bit tap offsets
    if (a.V * 3 > 128 || ((uintptr_t)a.pose & 15)) return false;      // pose staged with 128-bit loads
    const char* e = getenv("MVSTER_ET_WIN_MB");  // resident CTAs per SM the kernel is compiled for (A/B measurements)
    const int mb = e ? atoi(e) : 0;
    if (C == 8 && G == 4 && D == 4) {
        *rc = mb == 3 ? launch_et_win<8, 4, 4, 1, 3>(a, st) : mb == 5 ? launch_et_win<8, 4, 4, 1, 5>(a, st) : launch_et_win<8, 4, 4, 1, 4>(a, st);
        return true;
    }
    if (C == 16 && G == 4 && D == 4) {
        *rc = mb == 3 ? launch_et_win<16, 4, 4, 2, 3>(a, st) : mb == 4 ? launch_et_win<16, 4, 4, 2, 4>(a, st) : launch_et_win<16, 4, 4, 2, 5>(a, st);
        return true;
    }
    if (C == 32 && G == 8 && D == 8) {
        *rc = mb == 2 ? launch_et_win<32, 8, 8, 4, 2>(a, st) : mb == 4 ? launch_et_win<32, 8, 8, 4, 4>(a, st) : launch_et_win<32, 8, 8, 4, 3>(a, st);
        return true;
    }
    return false;
}

// bf16 storage: the same three window specialisations (pose staged the same way)
static bool try_launch_win_bf16(const EtArgs& a, int C, int G, int D, cudaStream_t st, int* rc) {
    if ((long long)a.B * a.Hs * a.Ws * C >= (1ll << 31) || a.B > 65535) return false;
    if (a.V * 3 > 128 || ((uintptr_t)a.pose & 15)) return false;
    if (C == 8 && G == 4 && D == 4) { *rc = launch_et_win_bf16<8, 4, 4, 1, 4>(a, st); return true; }
    if (C == 16 && G == 4 && D == 4) { *rc = launch_et_win_bf16<16, 4, 4, 2, 5>(a, st); return true; }
    if (C == 32 && G == 8 && D == 8) { *rc = launch_et_win_bf16<32, 8, 8, 4, 3>(a, st); return true; }
    return false;
}

}  // namespace mvster
