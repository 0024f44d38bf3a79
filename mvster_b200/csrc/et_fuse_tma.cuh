// TMA-staged variant of the fused warp + epipolar-Transformer kernel (cascade stages 2-4: C = 8 / 16 / 32).
//
// What bounds et_fuse_win_kernel (profiles/r01_et_fuse_win_ncu.md): every (pixel, view) waits one DRAM round trip for its
// window taps (3.8 long-scoreboard stall cycles per issued instruction at 14 warps / SM) and 15 % of the launch is the tail of
// a 4.3-wave grid.  Here the source features a reference tile can touch are brought into shared memory AHEAD of their use:
//
//   * persistent CTAs (grid = 2 x SMs or fewer), each walking reference tiles of 8 rows x 32 pixels;
//   * one PRODUCER warp per CTA runs ahead of the consumers: for the next (tile, view) it reads the tile's hypotheses,
//     projects the tile corners at the nearest and farthest hypothesis (a homography is monotone along each axis, so the corners
//     bound the footprint), and issues the footprint as TMA tile loads (cp.async.bulk.tensor, 2 source rows x BW pixels x C
//     channels per op, only the rows the footprint covers) into a ring of NSLOT shared-memory boxes, completion on an mbarrier
//     per slot.  TMA's out-of-bounds zero fill IS grid_sample's zeros padding, so border tiles need no special case, and the
//     tensor map's 32/64/128-byte swizzle makes the consumers' 16-byte reads at a 32-byte-per-lane stride bank-conflict free;
//   * CONSUMER warps (one row segment of the tile each, one lane per (pixel, 8 channels)) wait on the slot's mbarrier and run
//     the "correlate, then interpolate" arithmetic of et_fuse_win.cuh with the 3 x 3 tap window read from shared memory
//     (LDS with immediate offsets: no per-tap address arithmetic, no bounds tests), then release the slot;
//   * a warp whose window is not inside the box (footprint larger than the box, hypotheses spanning more than two source
//     cells, degenerate geometry, NaN) takes the per-hypothesis global gather of et_fuse_win.cuh for that view: the kernel is
//     valid for any input;
//   * the next tile's reference features and hypotheses are requested (plain coalesced loads) while the current tile is
//     computed.
//
// reference: models/mvs4net_utils.py:13-59 (homo_warping), :1037-1060 (group correlation, softmax over D, accumulation).
#pragma once
#ifndef MVSTER_CPU_EMU
#include "tc_ptx.cuh"
#endif

namespace mvster {
namespace ettma {

constexpr int TW = 32;             // reference tile width (pixels); its height TH depends on the channel count (Cfg)
constexpr int BW = 56;             // source box per (tile, view): BH rows x BW pixels (x C channels)
constexpr int RB = 4;              // source rows per TMA op
constexpr int MAX_SLOT = 4;        // boxes in flight per CTA (Cfg::NSLOT <= MAX_SLOT)
constexpr int MAXV = 12;           // tensor maps per launch (more views: chained launches through the partial sums)

#ifdef MVSTER_CPU_EMU
// ---- host twins of the mbarrier / TMA primitives (tests/emu): same protocol, synchronous copy with zero fill + swizzle
struct Mbar {
    std::atomic<int> pending{0};
    std::atomic<long long> tx{0};
    std::atomic<int> phase{0};
    int count = 0;
};
struct SrcMap { const float* base; int C, Ws, Hs, B; };
inline void mbar_try_complete(Mbar* m) {
    // single completer: the thread whose arrive / complete_tx brings (pending, tx) to (0, 0)
    int exp_p = 0;
    if (m->tx.load() == 0 && m->pending.compare_exchange_strong(exp_p, m->count)) m->phase.fetch_xor(1);
}
inline void mbar_init(Mbar* m, int count) { m->count = count; m->pending = count; m->tx = 0; m->phase = 0; }
inline void mbar_arrive(Mbar* m) { m->pending.fetch_sub(1); mbar_try_complete(m); }
inline void mbar_arrive_expect_tx(Mbar* m, int bytes) { m->tx.fetch_add(bytes); m->pending.fetch_sub(1); mbar_try_complete(m); }
inline void mbar_wait(Mbar* m, int parity) { while (m->phase.load() == parity) std::this_thread::yield(); }
inline unsigned swizzle_off(unsigned off, unsigned mask) { return off ^ (((off >> 7) & mask) << 4); }
// RB rows x BW pixels x C channels from (x0, y0) of image b, zero filled outside, written with the tensor map's swizzle
inline void tma_rows(unsigned char* slot, unsigned dst_off, const SrcMap* m, Mbar* bar, int x0, int y0, int b) {
    const unsigned mask = m->C / 4 - 1;
    for (int r = 0; r < RB; ++r)
        for (int x = 0; x < BW; ++x)
            for (int c = 0; c < m->C; ++c) {
                const int yy = y0 + r, xx = x0 + x;
                const bool in = yy >= 0 && yy < m->Hs && xx >= 0 && xx < m->Ws && b >= 0 && b < m->B;
                const float v = in ? m->base[(((long long)b * m->Hs + yy) * m->Ws + xx) * m->C + c] : 0.f;
                const unsigned off = dst_off + ((r * BW + x) * m->C + c) * 4;
                std::memcpy(slot + swizzle_off(off, mask), &v, 4);
            }
    bar->tx.fetch_sub((long long)RB * BW * m->C * 4);
    mbar_try_complete(bar);
}
inline void lds128(const unsigned char* smem, unsigned off, unsigned long long& a, unsigned long long& b) {
    std::memcpy(&a, smem + off, 8);
    std::memcpy(&b, smem + off + 8, 8);
}
#else
typedef unsigned long long Mbar;  // the 64-bit mbarrier object
struct SrcMap { CUtensorMap m; };
__device__ __forceinline__ void mbar_init(Mbar* m, int count) { ptx::mbar_init(ptx::smem_u32(m), count); }
__device__ __forceinline__ void mbar_arrive(Mbar* m) { ptx::mbar_arrive(ptx::smem_u32(m)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(Mbar* m, int bytes) { ptx::mbar_expect_tx(ptx::smem_u32(m), bytes); }
__device__ __forceinline__ void mbar_wait(Mbar* m, int parity) { ptx::mbar_wait_parked(ptx::smem_u32(m), parity, 20000u); }
__device__ __forceinline__ void tma_rows(unsigned char* slot, unsigned dst_off, const SrcMap* m, Mbar* bar, int x0, int y0, int b) {
    ptx::tma_load_4d(ptx::smem_u32(slot) + dst_off, &m->m, ptx::smem_u32(bar), 0, x0, y0, b);
}
__device__ __forceinline__ void lds128(const unsigned char* smem, unsigned off, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(ptx::smem_u32(smem) + off));
}
#endif

struct Maps { SrcMap m[MAXV]; };

// One 8-channel tap (two swizzled 16-byte chunks): chunk h of the 32-byte run at logical offset `off` lives at off ^ sw ^ 16 h.
template <int ROW_OFF>
__device__ __forceinline__ Pix8 lds_tap(const unsigned char* smem, unsigned a_lo, unsigned a_hi) {
    Pix8 t;
    lds128(smem, a_lo + ROW_OFF, t.p[0], t.p[1]);
    lds128(smem, a_hi + ROW_OFF, t.p[2], t.p[3]);
    return t;
}

// TH reference tile rows.  Warps per CTA = TH * LPP consumers + 1 producer; the register file is split over 4 scheduler
// partitions, so 8 warps (x 2 CTAs per SM) or 13..16 warps (x 1) leave 128 registers per thread, 9 or 17 warps only 96.
template <int C, int LPP, int TH_>
struct Cfg {
    static constexpr int TH = TH_;
    static constexpr int BH = TH >= 12 ? TH + 13 : TH >= 6 ? TH + 13 : TH + 9;  // box rows (footprint of TH rows + the depth-dependent shift), a multiple of RB
    static constexpr int NSLOT = LPP == 4 ? 2 : (C == 8 && TH >= 12) ? 4 : 3;  // boxes in flight per CTA
    static constexpr int NCW = TH * LPP;                  // consumer warps (a warp covers 32 / LPP pixels of one tile row)
    static constexpr int THREADS = (NCW + 1) * 32;        // + the producer warp
    static constexpr int PITCH = BW * C * 4;              // bytes per box row
    static constexpr int SLOT_BYTES = BH * PITCH;
    static constexpr int SLOT_STRIDE = (SLOT_BYTES + 1023) / 1024 * 1024;
    static constexpr int CTRL_BYTES = 512;                // mbarriers + box descriptors
    static constexpr int SMEM = NSLOT * SLOT_STRIDE + CTRL_BYTES + 1024;  // + alignment slack
    static constexpr unsigned SWZ_MASK = C / 4 - 1;       // 32 / 64 / 128-byte swizzle for C = 8 / 16 / 32
    static constexpr int MIN_CTAS = THREADS <= 256 ? 2 : 1;
};

struct Ctrl {
    Mbar full[MAX_SLOT];
    Mbar empty[MAX_SLOT];
    int4 box[MAX_SLOT];  // x0, y0, rows loaded (0 = no box: consumers gather from global memory), unused
};

// Producer: footprint of the tile in source view v.  Lane = tile column; for each of its TH pixels the sampling position at
// the pixel's smallest and largest hypothesis (the position is monotone in the depth), reduced over the warp.  The arithmetic
// need not match the consumers' bit for bit - they test their own window against the box - so the rotated pixel is advanced
// row by row and a 1/64-pixel margin absorbs the difference.  rows = 0: no box (footprint larger than the box, depth <= 0,
// a point behind the source camera, NaN); the consumers then gather from global memory.
template <int TH>
__device__ __forceinline__ void tile_footprint(const float* __restrict__ pose, float x, float y0, const float (&dlo)[TH], const float (&dhi)[TH],
                                               bool depth_ok, int BH, int& x0, int& y0_out, int& rows) {
    float rx = fmaf(pose[0], x, fmaf(pose[1], y0, pose[2]));
    float ry = fmaf(pose[3], x, fmaf(pose[4], y0, pose[5]));
    float rz = fmaf(pose[6], x, fmaf(pose[7], y0, pose[8]));
    const unsigned long long tx2 = pack2(pose[9], pose[9]), ty2 = pack2(pose[10], pose[10]), tz2 = pack2(pose[11], pose[11]);
    float umin = INFINITY, umax = -INFINITY, wmin = INFINITY, wmax = -INFINITY, zmin = INFINITY;
#pragma unroll
    for (int r = 0; r < TH; ++r) {  // both depths of the pixel per packed operation
        const unsigned long long d2 = pack2(dlo[r], dhi[r]);
        const float2 z = unpack2(fma2(pack2(rz, rz), d2, tz2));
        const unsigned long long i2 = pack2(rcp_approx(z.x), rcp_approx(z.y));
        const float2 u = unpack2(mul2(fma2(pack2(rx, rx), d2, tx2), i2)), w = unpack2(mul2(fma2(pack2(ry, ry), d2, ty2), i2));
        umin = fminf(umin, fminf(u.x, u.y)); umax = fmaxf(umax, fmaxf(u.x, u.y));
        wmin = fminf(wmin, fminf(w.x, w.y)); wmax = fmaxf(wmax, fmaxf(w.x, w.y));
        zmin = fminf(zmin, fminf(z.x, z.y));
        rx += pose[1]; ry += pose[4]; rz += pose[7];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o));
        umax = fmaxf(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        wmin = fminf(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
        wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        zmin = fminf(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
    }
    // taps floor(u) .. floor(u) + 1 of every sample (the consumers' third column / row is the +1 of a larger floor)
    const float m = 1.f / 64.f;
    const float ulo = floorf(umin - m), uhi = floorf(umax + m), wlo = floorf(wmin - m), whi = floorf(wmax + m);
    const float wcols = uhi - ulo + 2.f, wrows = whi - wlo + 2.f;
    const bool ok = depth_ok && zmin > 0.f && fabsf(ulo) < 1e6f && fabsf(uhi) < 1e6f && fabsf(wlo) < 1e6f && fabsf(whi) < 1e6f &&
                    wcols <= (float)BW && wrows <= (float)BH;  // every comparison is false for NaN
    x0 = ok ? (int)ulo : 0;
    y0_out = ok ? (int)wlo : 0;
    rows = ok ? (((int)wrows + RB - 1) / RB) * RB : 0;
}

template <int C, int G, int D, int LPP, int TH_, bool IL>
__global__ void __launch_bounds__(Cfg<C, LPP, TH_>::THREADS, Cfg<C, LPP, TH_>::MIN_CTAS)
et_fuse_tma_kernel(const EtArgs a, const __grid_constant__ Maps maps, int tiles_x, int tiles_y, int ntiles) {
    using K = Cfg<C, LPP, TH_>;
    constexpr int TH = K::TH, NSLOT = K::NSLOT;
    constexpr int GPL = G / LPP;   // groups per lane
    constexpr int CPG = C / G;     // channels per group
    constexpr int PXW = 32 / LPP;  // pixels per warp
    constexpr int NJ = GPL / 2;    // packed group pairs per lane
    static_assert(C / LPP == 8 && (GPL == 2 || GPL == 4) && (CPG == 2 || CPG == 4) && D % 2 == 0, "a lane owns 8 channels = 2 or 4 whole groups");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
#ifdef MVSTER_CPU_EMU
    unsigned char* smem = smem_raw;
#else
    unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
#endif
    Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem + NSLOT * K::SLOT_STRIDE);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int plane = a.H * a.W;
    if (threadIdx.x == 0) {
#ifdef MVSTER_CPU_EMU
        for (int s = 0; s < NSLOT; ++s) { new (&ctrl->full[s]) Mbar(); new (&ctrl->empty[s]) Mbar(); }
#endif
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(&ctrl->full[s], 1);
            mbar_init(&ctrl->empty[s], K::NCW);
        }
#ifndef MVSTER_CPU_EMU
        ptx::fence_barrier_init();
#endif
    }
    __syncthreads();

    if (warp == K::NCW) {
        // ------------------------------------------------------------------ producer warp
        // Lane = tile column.  Per pixel the farthest / nearest hypothesis (first and last plane: the schedules are monotone in
        // d; with an unsorted volume the consumers' own in-box test sends the affected warps to the global-gather path).  The
        // next tile's planes are requested before this tile's boxes are issued, so the producer never waits for DRAM itself.
        int slot = 0, phase = 0;
        float dlo[TH], dhi[TH], dan[TH], dbn[TH];
        auto request_range = [&](int tile, float (&da)[TH], float (&db)[TH]) {
            const int b = tile / (tiles_x * tiles_y), t2 = tile % (tiles_x * tiles_y);
            const int x_lo = (t2 % tiles_x) * TW, y_lo = (t2 / tiles_x) * TH;
            const int xx = min(x_lo + lane, a.W - 1);
            const float* hp = a.hypo + (long long)b * D * plane;
#pragma unroll
            for (int r = 0; r < TH; ++r) {
                const int yy = min(y_lo + r, a.H - 1);
                da[r] = __ldg(hp + yy * a.W + xx);
                db[r] = __ldg(hp + (long long)(D - 1) * plane + yy * a.W + xx);
            }
        };
        if ((int)blockIdx.x < ntiles) request_range(blockIdx.x, dan, dbn);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b = tile / (tiles_x * tiles_y), t2 = tile % (tiles_x * tiles_y);
            const int ty = t2 / tiles_x, tx = t2 % tiles_x;
            const int x_lo = tx * TW, y_lo = ty * TH;
            const int xx = min(x_lo + lane, a.W - 1);
            bool bad = false;
#pragma unroll
            for (int r = 0; r < TH; ++r) {
                dlo[r] = fminf(dan[r], dbn[r]);
                dhi[r] = fmaxf(dan[r], dbn[r]);
                bad = bad || !(dlo[r] > 0.f) || !(dhi[r] < 1e30f);
            }
            const bool depth_ok = __all_sync(0xffffffffu, !bad);
            if (tile + (int)gridDim.x < ntiles) request_range(tile + gridDim.x, dan, dbn);  // in flight while this tile's boxes go out
            for (int v = 0; v < a.V; ++v) {
                int x0, y0, rows;
                tile_footprint<TH>(a.pose + ((long long)b * a.V + v) * 12, (float)xx, (float)y_lo, dlo, dhi, depth_ok, K::BH, x0, y0, rows);
                mbar_wait(&ctrl->empty[slot], phase ^ 1);  // all consumer warps have released the slot's previous box
                if (lane == 0) {
                    ctrl->box[slot] = make_int4(x0, y0, rows, 0);
                    if (rows > 0) {
                        mbar_arrive_expect_tx(&ctrl->full[slot], rows * K::PITCH);
                        for (int r = 0; r < rows; r += RB)
                            tma_rows(smem + slot * K::SLOT_STRIDE, r * K::PITCH, &maps.m[v], &ctrl->full[slot], x0, y0 + r, b);
                    } else {
                        mbar_arrive(&ctrl->full[slot]);
                    }
                }
                __syncwarp();
                if (++slot == NSLOT) { slot = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int sub = lane % LPP;
    const int wrow = warp / LPP, wcol = (warp % LPP) * PXW + lane / LPP;
    const float max_x = (float)(a.Ws - 1), max_y = (float)(a.Hs - 1);
    const float inv_temp_log2e = 1.4426950408889634f / a.attn_temp;
    const int srow = a.Ws * C;
    const bool partial = a.flags & MVSTER_ET_PARTIAL;

    // inputs of a tile: the lane's 8 reference channels (pre-scaled by 1 / CPG: a power of two, exact) and D hypotheses
    unsigned long long ref_n[4], dep_n[D / 2];
    auto request_tile = [&](int tile) {
        const int b = tile / (tiles_x * tiles_y), t2 = tile % (tiles_x * tiles_y);
        const int x = min((t2 % tiles_x) * TW + wcol, a.W - 1), y = min((t2 / tiles_x) * TH + wrow, a.H - 1);
        const int pix = y * a.W + x;
        const Pix8 t = ldg256(a.ref + ((long long)b * plane + pix) * C + sub * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) ref_n[i] = t.p[i];
        const float* hp = a.hypo + (long long)b * D * plane + pix;
#pragma unroll
        for (int k = 0; k < D / 2; ++k) dep_n[k] = pack2(__ldg(hp + (long long)(2 * k) * plane), __ldg(hp + (long long)(2 * k + 1) * plane));
    };
    if ((int)blockIdx.x < ntiles) request_tile(blockIdx.x);

    int slot = 0, phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / (tiles_x * tiles_y), t2 = tile % (tiles_x * tiles_y);
        int x = (t2 % tiles_x) * TW + wcol, y = (t2 / tiles_x) * TH + wrow;
        const bool live = x < a.W && y < a.H;
        x = min(x, a.W - 1);
        y = min(y, a.H - 1);
        const int pix = y * a.W + x;
        unsigned long long ref[4], dep2[D / 2];
#pragma unroll
        for (int i = 0; i < 4; ++i) ref[i] = mul2(ref_n[i], pack2(1.f / CPG, 1.f / CPG));
#pragma unroll
        for (int k = 0; k < D / 2; ++k) dep2[k] = dep_n[k];
        if (tile + (int)gridDim.x < ntiles) request_tile(tile + gridDim.x);  // in flight while this tile is computed

        float ws[D];
        unsigned long long acc2[NJ][D];
        if (a.flags & MVSTER_ET_ACCUMULATE) {
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
            const float seed = partial ? 0.f : 1e-8f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                ws[d] = seed;
#pragma unroll
                for (int j = 0; j < NJ; ++j) acc2[j][d] = 0ull;
            }
        }
        const float fx = (float)x, fy = (float)y;
        const int lane_base = b * a.Hs * srow + sub * 8;  // < 2^31 (checked on the host)

        for (int v = 0; v < a.V; ++v) {
            const float4* pp = reinterpret_cast<const float4*>(a.pose + ((long long)b * a.V + v) * 12);
            const float4 q0 = __ldg(pp), q1 = __ldg(pp + 1), q2 = __ldg(pp + 2);  // R (row-major 3x3), t
            const float rx = fmaf(q0.z, 1.f, fmaf(q0.y, fy, q0.x * fx));
            const float ry = fmaf(q1.y, 1.f, fmaf(q1.x, fy, q0.w * fx));
            const float nrz = -fmaf(q2.x, 1.f, fmaf(q1.w, fy, q1.z * fx));
            const float tx = q2.y, ty = q2.z, ntz = -q2.w;
            // sampling positions of the D hypotheses, two per packed operation; bit-identical to et_fuse_win_kernel's (and to the
            // scalar sequence X = rx*d + tx with separately rounded product and sum, Z == 0 -> 1e-9, u = X / Z through one
            // reciprocal + FMA residual): a fused multiply-add here moves the positions by up to 1e-4 pixel against the
            // reference and, on white-noise-like features, the cost volume by 3e-4 of its maximum (measured)
            unsigned long long ix2[D / 2], iy2[D / 2];
            float ix[D], iy[D];
#pragma unroll
            for (int k = 0; k < D / 2; ++k) {
                const float2 dd = unpack2(dep2[k]);
                const unsigned long long X2 = add2(pack2(__fmul_rn(rx, dd.x), __fmul_rn(rx, dd.y)), pack2(tx, tx));
                const unsigned long long Y2 = add2(pack2(__fmul_rn(ry, dd.x), __fmul_rn(ry, dd.y)), pack2(ty, ty));
                float2 zn = unpack2(add2(pack2(__fmul_rn(nrz, dd.x), __fmul_rn(nrz, dd.y)), pack2(ntz, ntz)));  // -Z
                if (zn.x == 0.f) zn.x = -1e-9f;
                if (zn.y == 0.f) zn.y = -1e-9f;
                const float r0 = rcp_approx(-zn.x), r1 = rcp_approx(-zn.y);
                const unsigned long long Zn2 = pack2(zn.x, zn.y), r2 = pack2(r0, r1);
                const unsigned long long qx = mul2(X2, r2), qy = mul2(Y2, r2);
                ix2[k] = fma2(fma2(qx, Zn2, X2), r2, qx);
                iy2[k] = fma2(fma2(qy, Zn2, Y2), r2, qy);
                const float2 px = unpack2(ix2[k]), py = unpack2(iy2[k]);
                ix[2 * k] = px.x; ix[2 * k + 1] = px.y;
                iy[2 * k] = py.x; iy[2 * k + 1] = py.y;
            }
            float bxf = floorf(ix[0]), byf = floorf(iy[0]), mxf = bxf, myf = byf;
#pragma unroll
            for (int d = 1; d < D; ++d) {
                const float fxd = floorf(ix[d]), fyd = floorf(iy[d]);
                bxf = min_nan(bxf, fxd); mxf = max_nan(mxf, fxd);
                byf = min_nan(byf, fyd); myf = max_nan(myf, fyd);
            }
            const bool needx = mxf > bxf, needy = myf > byf;
            // at most two cells per axis (false for NaN), window origin representable as an int
            const bool span_ok = (mxf - bxf) <= 1.f && (myf - byf) <= 1.f && fabsf(bxf) < 1e6f && fabsf(byf) < 1e6f;

            mbar_wait(&ctrl->full[slot], phase);
            const int4 box = ctrl->box[slot];
            const int rx0 = (int)bxf - box.x, ry0 = (int)byf - box.y;
            // every tap the lane will read (2 + needx columns, 2 + needy rows) inside the loaded part of the box
            const bool in_box = span_ok && rx0 >= 0 && rx0 + 1 + (needx ? 1 : 0) < BW && ry0 >= 0 && ry0 + 1 + (needy ? 1 : 0) < box.z;

            unsigned long long cor2[NJ][D];
            if (__all_sync(0xffffffffu, in_box)) {
                const unsigned char* sb = smem + slot * K::SLOT_STRIDE;
                // logical byte offsets of the three window columns in row ry0, then the swizzle (address bits 4.. ^= bits 7..)
                unsigned alo[3], ahi[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const unsigned col = (unsigned)((rx0 + c) * (C * 4) + sub * 32);
                    const unsigned lo = col ^ (((col >> 7) & K::SWZ_MASK) << 4);
                    alo[c] = lo + (unsigned)ry0 * K::PITCH;
                    ahi[c] = alo[c] ^ 16u;
                }
                unsigned long long T[3][3][NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) T[0][2][j] = T[1][2][j] = T[2][0][j] = T[2][1][j] = T[2][2][j] = 0ull;
                const Pix8 t00 = lds_tap<0>(sb, alo[0], ahi[0]), t01 = lds_tap<0>(sb, alo[1], ahi[1]);
                const Pix8 t10 = lds_tap<K::PITCH>(sb, alo[0], ahi[0]), t11 = lds_tap<K::PITCH>(sb, alo[1], ahi[1]);
                if (needx) {
                    const Pix8 t02 = lds_tap<0>(sb, alo[2], ahi[2]), t12 = lds_tap<K::PITCH>(sb, alo[2], ahi[2]);
                    tap_groups<CPG, NJ, IL>(t02, ref, T[0][2]);
                    tap_groups<CPG, NJ, IL>(t12, ref, T[1][2]);
                }
                tap_groups<CPG, NJ, IL>(t00, ref, T[0][0]);
                tap_groups<CPG, NJ, IL>(t01, ref, T[0][1]);
                tap_groups<CPG, NJ, IL>(t10, ref, T[1][0]);
                tap_groups<CPG, NJ, IL>(t11, ref, T[1][1]);
                if (needy) {
                    const Pix8 t20 = lds_tap<2 * K::PITCH>(sb, alo[0], ahi[0]), t21 = lds_tap<2 * K::PITCH>(sb, alo[1], ahi[1]);
                    tap_groups<CPG, NJ, IL>(t20, ref, T[2][0]);
                    tap_groups<CPG, NJ, IL>(t21, ref, T[2][1]);
                    if (needx) {
                        const Pix8 t22 = lds_tap<2 * K::PITCH>(sb, alo[2], ahi[2]);
                        tap_groups<CPG, NJ, IL>(t22, ref, T[2][2]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctrl->empty[slot]);  // the warp holds everything it needs from the box
                unsigned long long dx0[3][NJ], dx1[3][NJ];  // column differences
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        dx0[r][j] = sub2(T[r][1][j], T[r][0][j]);
                        dx1[r][j] = sub2(T[r][2][j], T[r][1][j]);
                    }
#pragma unroll
                for (int k = 0; k < D / 2; ++k) {
                    // offsets from the window origin (exact, in [0,2)) and their split u = a + b, a = min(u, 1), b = u - a
                    const unsigned long long ux2 = sub2(ix2[k], pack2(bxf, bxf)), uy2 = sub2(iy2[k], pack2(byf, byf));
                    const float2 ux = unpack2(ux2), uy = unpack2(uy2);
                    const float axs[2] = {fminf(ux.x, 1.f), fminf(ux.y, 1.f)}, ays[2] = {fminf(uy.x, 1.f), fminf(uy.y, 1.f)};
                    const float2 bxs = unpack2(sub2(ux2, pack2(axs[0], axs[1]))), bys = unpack2(sub2(uy2, pack2(ays[0], ays[1])));
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int d = 2 * k + e;
                        const float bx = e ? bxs.y : bxs.x, by = e ? bys.y : bys.x;
                        const unsigned long long ax2 = pack2(axs[e], axs[e]), bx2 = pack2(bx, bx), ay2 = pack2(ays[e], ays[e]), by2 = pack2(by, by);
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const unsigned long long h0 = fma2(bx2, dx1[0][j], fma2(ax2, dx0[0][j], T[0][0][j]));
                            const unsigned long long h1 = fma2(bx2, dx1[1][j], fma2(ax2, dx0[1][j], T[1][0][j]));
                            const unsigned long long h2 = fma2(bx2, dx1[2][j], fma2(ax2, dx0[2][j], T[2][0][j]));
                            cor2[j][d] = fma2(by2, sub2(h2, h1), fma2(ay2, sub2(h1, h0), h0));
                        }
                    }
                }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctrl->empty[slot]);  // nothing is read from the box on this path
                // per-hypothesis gather from global memory with zeros padding per tap (et_fuse_win.cuh's fallback)
                const float* S = a.src[v];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    int o_nw, o_ne, o_sw, o_se;
                    float w_nw, w_ne, w_sw, w_se;
                    const float px = ix[d], py = iy[d];
                    {
                        const float cx = fminf(fmaxf(px, -2.f), max_x + 2.f), cy = fminf(fmaxf(py, -2.f), max_y + 2.f);
                        const float x0f = floorf(cx), y0f = floorf(cy);
                        const float wx = cx - x0f, wy = cy - y0f;
                        const int x0 = (int)x0f, y0 = (int)y0f;
                        const float ex = (unsigned)x0 < (unsigned)a.Ws ? 1.f - wx : 0.f, fxw = (unsigned)(x0 + 1) < (unsigned)a.Ws ? wx : 0.f;
                        const float ey = (unsigned)y0 < (unsigned)a.Hs ? 1.f - wy : 0.f, fyw = (unsigned)(y0 + 1) < (unsigned)a.Hs ? wy : 0.f;
                        const int xa = min(max(x0, 0), a.Ws - 1) * C, xb = min(max(x0 + 1, 0), a.Ws - 1) * C;
                        const int ya = lane_base + min(max(y0, 0), a.Hs - 1) * srow, yb = lane_base + min(max(y0 + 1, 0), a.Hs - 1) * srow;
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
            if (++slot == NSLOT) { slot = 0; phase ^= 1; }

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
            for (int d = 0; d < D; ++d) {
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

        if (live) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const long long o = ((long long)b * D + d) * plane + pix;
                const float r = partial ? 1.f : __frcp_rn(ws[d]);
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
    }
}

#ifndef MVSTER_CPU_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// [B][Hs][Ws][C] fp32 source features; box = C channels x BW pixels x RB rows of one image, swizzled by the pixel's byte width
static bool make_map(SrcMap* m, const float* src, int B, int Hs, int Ws, int C) {
    EncodeTiledFn enc = encode_fn();
    if (!enc || ((uintptr_t)src & 15)) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)Ws * C * 4, (cuuint64_t)Hs * Ws * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)C, BW, RB, 1}, es[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle swz = C == 8 ? CU_TENSOR_MAP_SWIZZLE_32B : C == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    return enc(&m->m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static int sm_count_here() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cache[dev]) cudaDeviceGetAttribute(&cache[dev], cudaDevAttrMultiProcessorCount, dev);
    return cache[dev];
}
#else
static bool make_map(SrcMap* m, const float* src, int B, int Hs, int Ws, int C) {
    m->base = src; m->C = C; m->Ws = Ws; m->Hs = Hs; m->B = B;
    return true;
}
static int sm_count_here() { return 2; }
#endif

template <int C, int G, int D, int LPP, int TH_>
static int launch_et_tma(const EtArgs& a, const Maps& maps, cudaStream_t st) {
    using K = Cfg<C, LPP, TH_>;
    const int tiles_x = ceil_div(a.W, TW), tiles_y = ceil_div(a.H, K::TH);
    const long long ntiles = (long long)tiles_x * tiles_y * a.B;
    const bool il = a.flags & MVSTER_ET_INTERLEAVED;
    if (cudaFuncSetAttribute(et_fuse_tma_kernel<C, G, D, LPP, TH_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(et_fuse_tma_kernel<C, G, D, LPP, TH_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM) != cudaSuccess) {
        cudaGetLastError();
        return -100;
    }
    const int slots = sm_count_here() * K::MIN_CTAS;
    const int grid = (int)(ntiles < slots ? ntiles : slots);
    if (il) et_fuse_tma_kernel<C, G, D, LPP, TH_, true><<<grid, K::THREADS, K::SMEM, st>>>(a, maps, tiles_x, tiles_y, (int)ntiles);
    else et_fuse_tma_kernel<C, G, D, LPP, TH_, false><<<grid, K::THREADS, K::SMEM, st>>>(a, maps, tiles_x, tiles_y, (int)ntiles);
    note_et_kernel(il ? (TH_ == 15 ? "et_fuse_tma_kernel[15x32 tiles, interleaved]" : TH_ == 7 ? "et_fuse_tma_kernel[7x32 tiles, interleaved]" : "et_fuse_tma_kernel[3x32 tiles, interleaved]")
                      : (TH_ == 15 ? "et_fuse_tma_kernel[15x32 tiles]" : TH_ == 7 ? "et_fuse_tma_kernel[7x32 tiles]" : "et_fuse_tma_kernel[3x32 tiles]"),
                   C, G, D, LPP, K::MIN_CTAS);
    return check_launch("et_fuse_tma_kernel");
}

// MVSTER_ET_TMA=0/1 overrides the built-in default; the MVSTER_ET_TMA_ON / MVSTER_ET_TMA_OFF flags of the call win over both.
static bool et_tma_default() {
    const char* e = getenv("MVSTER_ET_TMA");
    return e ? atoi(e) != 0 : true;
}

// Returns true if a TMA-staged specialisation exists for (C,G,D), the inputs qualify, and it was launched into *rc.
static bool try_launch_tma(const EtArgs& a, int C, int G, int D, cudaStream_t st, int* rc) {
    if ((long long)a.B * a.Hs * a.Ws * C >= (1ll << 31) || (long long)a.B * a.H * a.W * D >= (1ll << 31)) return false;  // 32-bit offsets
    if (a.V > MAXV || ((uintptr_t)a.pose & 15) || ((uintptr_t)a.ref & 31)) return false;
    const bool known = (C == 8 && G == 4 && D == 4) || (C == 16 && G == 4 && D == 4) || (C == 32 && G == 8 && D == 8);
    if (!known) return false;
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    for (int v = 0; v < a.V; ++v)
        if (!make_map(&maps.m[v], a.src[v], a.B, a.Hs, a.Ws, C)) return false;
    int r = -100;
    // measured on B200 (cfg2 stage 4, profiles/r02_et_fuse_tma_ncu.md): 7 x 32 tiles, two CTAs per SM 45.0 us; 15 x 32 tiles, one CTA 46.6 us
    const char* the = getenv("MVSTER_ET_TMA_TH");
    const int th8 = the ? atoi(the) : 7;  // A/B: 15 = 15 x 32 tiles, one CTA per SM
    if (C == 8) r = th8 == 7 ? launch_et_tma<8, 4, 4, 1, 7>(a, maps, st) : launch_et_tma<8, 4, 4, 1, 15>(a, maps, st);
    else if (C == 16) r = launch_et_tma<16, 4, 4, 2, 7>(a, maps, st);
    else r = launch_et_tma<32, 8, 8, 4, 3>(a, maps, st);
    if (r == -100) return false;
    *rc = r;
    return true;
}

}  // namespace ettma
}  // namespace mvster
