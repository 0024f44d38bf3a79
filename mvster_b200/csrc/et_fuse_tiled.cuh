// Tiled, instruction-lean variant of the fused warp + epipolar-Transformer kernel for the shipped
// (C, G, D) combinations.  Same arithmetic as et_fuse_kernel (et_fuse.cu) with these changes:
//   * LPP lanes per reference pixel instead of G: the plane-sweep geometry (the dominant
//     instruction cost at C = 8..16) is evaluated once per pixel (LPP = 1) or twice/4x for the wide
//     stages, and every lane moves C/LPP contiguous channels per tap with 128-bit loads.
//   * (32/LPP) x 4-pixel CTA tiles (one warp per row): the lower taps of row y are the upper taps of
//     row y+1, so the vertical bilinear overlap is served by L1 instead of L2.
//   * packed fp32 math (FFMA2, fma.rn.f32x2 - new on sm_100) for the 4-tap blend and the
//     group dot products; divisions by a shared denominator use one reciprocal + an FMA residual
//     correction (<= 1 ulp from IEEE); softmax uses ex2.approx.
// The kernel is HBM-bandwidth bound by design: per pixel it must move (1+V)*C*4 + D*4 + G*D*4
// bytes (240 B at stage 4) and now issues ~1.3 k instructions for them (5.5 instr/B is the
// B200 issue/HBM balance point).
#pragma once

namespace mvster {

#ifdef MVSTER_CPU_EMU  // tests/emu: the same helpers in portable C++ so that the kernel sources below run on the host
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) { return emu::pack(lo, hi); }
__device__ __forceinline__ float2 unpack2(unsigned long long v) { return emu::unpack(v); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    const float2 x = emu::unpack(a), y = emu::unpack(b), z = emu::unpack(c);
    return emu::pack(fmaf(x.x, y.x, z.x), fmaf(x.y, y.y, z.y));
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    const float2 x = emu::unpack(a), y = emu::unpack(b);
    return emu::pack(x.x * y.x, x.y * y.y);
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
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
#endif

struct Pix8 { unsigned long long p[4]; };  // 8 fp32 channels as 4 packed pairs
// One 256-bit read-only load (LDG.E.256, new on sm_100): a lane fetches its whole 8-channel tap, a
// warp request is 1 KB of full 128-byte lines for the L1 data pipe (two 128-bit loads cost twice
// the wavefronts because each quarter-warp then straddles two lines).
__device__ __forceinline__ Pix8 ldg256(const float* p) {
    Pix8 r;
#ifdef MVSTER_CPU_EMU
    for (int i = 0; i < 4; ++i) r.p[i] = emu::pack(p[2 * i], p[2 * i + 1]);
#else
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.p[0]), "=l"(r.p[1]), "=l"(r.p[2]), "=l"(r.p[3]) : "l"(p));
#endif
    return r;
}

// bf16 storage (mvster_et_fuse_bf16): a lane's 8-channel tap is 16 bytes = one 128-bit load; bf16 -> fp32 is a shift / mask of
// the packed words, so the arithmetic below is unchanged (fp32 geometry, correlations and softmax).
typedef uint16_t bf16_t;
__device__ __forceinline__ Pix8 ldg256(const bf16_t* p) {
    Pix8 r;
#ifdef MVSTER_CPU_EMU
    for (int i = 0; i < 4; ++i) {
        const uint32_t lo = (uint32_t)p[2 * i] << 16, hi = (uint32_t)p[2 * i + 1] << 16;
        float fl, fh;
        memcpy(&fl, &lo, 4); memcpy(&fh, &hi, 4);
        r.p[i] = emu::pack(fl, fh);
    }
#else
    uint32_t w[4];
    asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
#pragma unroll
    for (int i = 0; i < 4; ++i) r.p[i] = pack2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
#endif
    return r;
}
// fp32 -> bf16, round to nearest even (NaN stays NaN): the rounding of torch's .to(torch.bfloat16)
__device__ __forceinline__ uint32_t f2bf(float x) {
#ifdef MVSTER_CPU_EMU
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (u >> 16) | 0x40u;
    return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
#else
    uint16_t r;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ uint32_t f2bf2(float lo, float hi) { return f2bf(lo) | (f2bf(hi) << 16); }
// feature / cost element types of a kernel instantiation
template <bool BF> struct EtTypes { typedef float feat; typedef float cost; };
template <> struct EtTypes<true> { typedef bf16_t feat; typedef bf16_t cost; };

__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP
#ifdef MVSTER_CPU_EMU
    return 1.f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2 (results below 2^-126 flush to 0)
#ifdef MVSTER_CPU_EMU
    return exp2f(x);
#else
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// a / b given r ~= 1/b: one Newton-style residual correction (exact residual through FMA).
__device__ __forceinline__ float div_corrected(float a, float b, float r) {
    const float q = a * r;
    return fmaf(fmaf(-q, b, a), r, q);
}

template <int C, int G, int D, int LPP, int MB, bool BF = false>
__global__ void __launch_bounds__(128, MB) et_fuse_tiled_kernel(const EtArgs a) {
    typedef typename EtTypes<BF>::feat FT;  // BF: bf16 features in, bf16 cost volume out (no PARTIAL / ACCUMULATE: checked on the host)
    constexpr int CPL = C / LPP;   // channels per lane
    constexpr int GPL = G / LPP;   // groups per lane
    constexpr int CPG = C / G;     // channels per group
    constexpr int PXW = 32 / LPP;  // pixels per warp (tile width)
    constexpr int NP = CPL / 2;    // packed pairs per lane
    static_assert(CPL == 8 && GPL >= 1 && CPG % 2 == 0 && CPG <= 8, "a lane owns exactly 8 channels (one 256-bit load per tap)");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPP;
    int x = blockIdx.x * PXW + lane / LPP;
    int y = blockIdx.y * 4 + warp;
    const int b = blockIdx.z;
    const bool live = x < a.W && y < a.H;
    x = min(x, a.W - 1);
    y = min(y, a.H - 1);
    const int plane = a.H * a.W, pix = y * a.W + x;

    unsigned long long ref[NP];
    {
        const Pix8 t = ldg256(reinterpret_cast<const FT*>(a.ref) + ((long long)b * plane + pix) * C + sub * CPL);
#pragma unroll
        for (int i = 0; i < NP; ++i) ref[i] = t.p[i];
    }
    float dep[D], ws[D], acc[GPL][D];
    const float* hp = a.hypo + (long long)b * D * plane + pix;
#pragma unroll
    for (int d = 0; d < D; ++d) dep[d] = __ldg(hp + (long long)d * plane);
    if (!BF && (a.flags & MVSTER_ET_ACCUMULATE)) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const long long o = ((long long)b * D + d) * plane + pix;
            ws[d] = a.wsum[o];
#pragma unroll
            for (int g = 0; g < GPL; ++g) acc[g][d] = a.cost[o * G + sub * GPL + g];
        }
    } else {
        const float seed = (a.flags & MVSTER_ET_PARTIAL) ? 0.f : 1e-8f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            ws[d] = seed;
#pragma unroll
            for (int g = 0; g < GPL; ++g) acc[g][d] = 0.f;
        }
    }

    // fold the 1/CPG of .mean(2) into the reference features (power of two: exact)
#pragma unroll
    for (int i = 0; i < NP; ++i) ref[i] = mul2(ref[i], pack2(1.f / CPG, 1.f / CPG));

    const float fx = (float)x, fy = (float)y;
    const float max_x = (float)(a.Ws - 1), max_y = (float)(a.Hs - 1);
    const float inv_temp_log2e = 1.4426950408889634f / a.attn_temp;
    const int row = a.Ws * C;
    const int lane_base = b * a.Hs * row + sub * CPL;  // < 2^31 (checked on the host)

    for (int v = 0; v < a.V; ++v) {
        const float* P = a.pose + ((long long)b * a.V + v) * 12;
        const float rx = fmaf(__ldg(P + 2), 1.f, fmaf(__ldg(P + 1), fy, __ldg(P + 0) * fx));
        const float ry = fmaf(__ldg(P + 5), 1.f, fmaf(__ldg(P + 4), fy, __ldg(P + 3) * fx));
        const float rz = fmaf(__ldg(P + 8), 1.f, fmaf(__ldg(P + 7), fy, __ldg(P + 6) * fx));
        const float tx = __ldg(P + 9), ty = __ldg(P + 10), tz = __ldg(P + 11);
        const FT* S = reinterpret_cast<const FT*>(a.src[v]);  // warp-uniform base; batch/lane offsets live in the 32-bit tap offsets

        float cor[GPL][D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float X = __fadd_rn(__fmul_rn(rx, dep[d]), tx);
            const float Y = __fadd_rn(__fmul_rn(ry, dep[d]), ty);
            float Z = __fadd_rn(__fmul_rn(rz, dep[d]), tz);
            if (Z == 0.f) Z = 1e-9f;
            const float rZ = rcp_approx(Z);
            // Sampling position in source pixels.  The reference normalises to [-1,1] and grid_sample
            // maps back (mvs4net_utils.py:43-44 + align_corners=True): an identity up to <= 4e-5 px of
            // fp32 rounding, skipped here.
            const float ix = div_corrected(X, Z, rZ), iy = div_corrected(Y, Z, rZ);
            int o_nw, o_ne, o_sw, o_se;
            float w_nw, w_ne, w_sw, w_se;
            const bool interior = ix >= 0.f && ix < max_x && iy >= 0.f && iy < max_y;  // false for NaN
            if (__all_sync(0xffffffffu, interior)) {  // warp-uniform fast path: all 4 taps of every lane in bounds
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
            const unsigned long long k_nw = pack2(w_nw, w_nw), k_ne = pack2(w_ne, w_ne);
            const unsigned long long k_sw = pack2(w_sw, w_sw), k_se = pack2(w_se, w_se);
            unsigned long long prod[NP];  // ref * bilinear(warped) per channel pair
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                unsigned long long wv = mul2(t_nw.p[i], k_nw);
                wv = fma2(t_ne.p[i], k_ne, wv);
                wv = fma2(t_sw.p[i], k_sw, wv);
                wv = fma2(t_se.p[i], k_se, wv);
                prod[i] = wv;
            }
            float gsum[GPL];
            if constexpr (CPG == 2) {         // 4 groups of one pair
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float2 q = unpack2(mul2(ref[i], prod[i])); gsum[i] = q.x + q.y; }
            } else if constexpr (CPG == 4) {  // 2 groups of two pairs
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const float2 q = unpack2(fma2(ref[2 * g + 1], prod[2 * g + 1], mul2(ref[2 * g], prod[2 * g])));
                    gsum[g] = q.x + q.y;
                }
            } else {                          // one group of four pairs
                unsigned long long q2 = mul2(ref[0], prod[0]);
                q2 = fma2(ref[1], prod[1], q2); q2 = fma2(ref[2], prod[2], q2); q2 = fma2(ref[3], prod[3], q2);
                const float2 q = unpack2(q2);
                gsum[0] = q.x + q.y;
            }
#pragma unroll
            for (int g = 0; g < GPL; ++g) cor[g][d] = gsum[g];
        }

        // softmax over D of (sum over all G groups) / temp, then / sqrt(C)   (mvs4net_utils.py:1053)
        float lg[D], m = -INFINITY;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float s = cor[0][d];
#pragma unroll
            for (int g = 1; g < GPL; ++g) s += cor[g][d];
#pragma unroll
            for (int o = 1; o < LPP; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            lg[d] = s * inv_temp_log2e;
            m = fmaxf(m, lg[d]);
        }
        float se = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) { lg[d] = exp2f(lg[d] - m); se += lg[d]; }
        const float rs = rcp_approx(se * a.sqrt_c);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float w = lg[d] * rs;
            ws[d] += w;
#pragma unroll
            for (int g = 0; g < GPL; ++g) acc[g][d] = fmaf(w, cor[g][d], acc[g][d]);
        }
    }

    if (!live) return;
    const bool partial = !BF && (a.flags & MVSTER_ET_PARTIAL);
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const long long o = ((long long)b * D + d) * plane + pix;
        const float r = partial ? 1.f : __frcp_rn(ws[d]);
        float out[GPL];
#pragma unroll
        for (int g = 0; g < GPL; ++g) out[g] = partial ? acc[g][d] : acc[g][d] * r;
        if constexpr (BF) {  // bf16 cost volume: the lane's GPL groups as 16-bit values (GPL even, or 1 with LPP = 8)
            bf16_t* dstb = reinterpret_cast<bf16_t*>(a.cost) + o * G + sub * GPL;
            if constexpr (GPL % 4 == 0) {
#pragma unroll
                for (int g = 0; g < GPL; g += 4) *reinterpret_cast<uint2*>(dstb + g) = make_uint2(f2bf2(out[g], out[g + 1]), f2bf2(out[g + 2], out[g + 3]));
            } else if constexpr (GPL == 2) {
                *reinterpret_cast<uint32_t*>(dstb) = f2bf2(out[0], out[1]);
            } else {
#pragma unroll
                for (int g = 0; g < GPL; ++g) dstb[g] = (bf16_t)f2bf(out[g]);
            }
            continue;
        }
        float* dst = a.cost + o * G + sub * GPL;
        if constexpr (GPL % 4 == 0) {
#pragma unroll
            for (int g = 0; g < GPL; g += 4) *reinterpret_cast<float4*>(dst + g) = make_float4(out[g], out[g + 1], out[g + 2], out[g + 3]);
        } else if constexpr (GPL == 2) {
            *reinterpret_cast<float2*>(dst) = make_float2(out[0], out[1]);
        } else {
#pragma unroll
            for (int g = 0; g < GPL; ++g) dst[g] = out[g];
        }
        if (partial && sub == 0) a.wsum[o] = ws[d];
    }
}

template <int C, int G, int D, int LPP, int MB>
static int launch_et_tiled(const EtArgs& a, cudaStream_t st) {
    dim3 grid(ceil_div(a.W, 32 / LPP), ceil_div(a.H, 4), a.B);
    et_fuse_tiled_kernel<C, G, D, LPP, MB><<<grid, 128, 0, st>>>(a);
    note_et_kernel("et_fuse_tiled_kernel", C, G, D, LPP, MB);
    return check_launch("et_fuse_tiled_kernel");
}

template <int C, int G, int D, int LPP, int MB>
static int launch_et_tiled_bf16(const EtArgs& a, cudaStream_t st) {
    dim3 grid(ceil_div(a.W, 32 / LPP), ceil_div(a.H, 4), a.B);
    et_fuse_tiled_kernel<C, G, D, LPP, MB, true><<<grid, 128, 0, st>>>(a);
    note_et_kernel("et_fuse_tiled_kernel[bf16]", C, G, D, LPP, MB);
    return check_launch("et_fuse_tiled_kernel[bf16]");
}

// Resident CTAs per SM the D = 4 kernels are compiled for (register cap 65536 / (128 * MB)): 4 -> 128 regs,
// 5 -> 96 (12 B spilled), 6 -> 80 (68 B spilled).  MVSTER_ET_MB overrides the default for A/B measurements.
static int et_min_blocks() {
    const char* e = getenv("MVSTER_ET_MB");
    const int v = e ? atoi(e) : 5;  // measured on B200 (cfg2): stage 3 32.7 -> 29.2 us, stage 4 56.8 -> 55.2 us vs MB = 4
    return (v == 4 || v == 6) ? v : 5;
}

// Returns 1 if a tiled specialisation exists for (C,G,D) and was launched into *rc.
static bool try_launch_tiled(const EtArgs& a, int C, int G, int D, cudaStream_t st, int* rc) {
    if ((long long)a.B * a.Hs * a.Ws * C >= (1ll << 31) || a.B > 65535) return false;  // 32-bit tap offsets
    const int mb = et_min_blocks();
    if (C == 8 && G == 4 && D == 4) {
        *rc = mb == 6 ? launch_et_tiled<8, 4, 4, 1, 6>(a, st) : mb == 5 ? launch_et_tiled<8, 4, 4, 1, 5>(a, st) : launch_et_tiled<8, 4, 4, 1, 4>(a, st);
        return true;
    }
    if (C == 16 && G == 4 && D == 4) {
        *rc = mb == 6 ? launch_et_tiled<16, 4, 4, 2, 6>(a, st) : mb == 5 ? launch_et_tiled<16, 4, 4, 2, 5>(a, st) : launch_et_tiled<16, 4, 4, 2, 4>(a, st);
        return true;
    }
    if (C == 32 && G == 8 && D == 8) { *rc = launch_et_tiled<32, 8, 8, 4, 4>(a, st); return true; }
    if (C == 64 && G == 8 && D == 8) { *rc = launch_et_tiled<64, 8, 8, 8, 4>(a, st); return true; }
    return false;
}

// bf16 storage: same specialisations (natural channel order only: the tiled kernel has no interleaved form)
static bool try_launch_tiled_bf16(const EtArgs& a, int C, int G, int D, cudaStream_t st, int* rc) {
    if ((long long)a.B * a.Hs * a.Ws * C >= (1ll << 31) || a.B > 65535) return false;
    if (C == 8 && G == 4 && D == 4) { *rc = launch_et_tiled_bf16<8, 4, 4, 1, 5>(a, st); return true; }
    if (C == 16 && G == 4 && D == 4) { *rc = launch_et_tiled_bf16<16, 4, 4, 2, 5>(a, st); return true; }
    if (C == 32 && G == 8 && D == 8) { *rc = launch_et_tiled_bf16<32, 8, 8, 4, 4>(a, st); return true; }
    if (C == 64 && G == 8 && D == 8) { *rc = launch_et_tiled_bf16<64, 8, 8, 8, 4>(a, st); return true; }
    return false;
}

}  // namespace mvster
