// Implicit-GEMM convolution on tcgen05, generation 3: persistent, fully pipelined, error-compensated BF16.
//
// What changed against conv_tc2.cu, and the measurement behind each change (tools/mma_microbench.cu on B200: one
// M128 x K8 kind::tf32 MMA costs 39 / 40 / 48 / 64 / 128 cycles at N = 16 / 32 / 64 / 128 / 256 - below N = 64 the
// instruction is bound by the 4 KB A-operand fetch from shared memory, not by the multipliers, whatever the layout):
//   * Arithmetic: the fp32 operands are split into three bf16 terms (a = a1 + a2 + a3 exactly to 24 bits, same for w) and
//     the six significant products a1w1, a1w2, a1w3, a2w1, a2w2, a3w1 (dropped terms <= 2^-24 relative) are computed by
//     THREE kind::f16 MMAs per 16 channels: a1 x [w1|w2|w3] (N = 3*NC), a2 x [w1|w2] (N = 2*NC), a3 x [w1] (N = NC) - the
//     weight splits sit side by side along N, which is free while the MMA is A-fetch bound.  3xTF32 needed six MMAs (K = 8)
//     for the same 16 channels.  The three column blocks of the accumulator are summed in the epilogue.
//   * Persistent CTAs (one per SM) walk groups of T tiles; two accumulator sets in TMEM (2 x 192 columns) let the epilogue
//     of group g overlap the MMAs of group g+1; TMA staging, operand conversion, MMA issue and epilogue are separate warps
//     connected by mbarrier rings, so no stage waits for a prologue.
//   * A small per-layer plan (stages x taps) generalises the addressing: a stage is one TMA-staged halo tile (18 x 10 pixels
//     x <= 16 channels of one input plane, optionally of one stride-2 parity class via the tensor map's element strides), a
//     tap is a 16-byte-granular shift of the A descriptor inside it.  Stride-1 3x3(x3) and 1x1, and stride-2 3x3 / 5x5
//     convolutions (4 parity classes, no wasted taps) run on the same kernel.
//   * tcgen05 / TMA issue is guarded by elect.sync (see tc_ptx.cuh: a lane test makes the compiler serialise every issue).
//
//   * Second arithmetic (NS = 2, MVSTER_TC3_FP16X2): two FP16 terms per operand, a = a1 + 2^-11 a2 with a1 = fp16(a) and
//     a2 = fp16(2^11 (a - a1)) (22 significand bits; the scaling keeps the residual out of fp16's subnormal range), same for w.
//     TWO MMAs per 16 channels: a1 x [w1|w2] (N = 2*NC) and a2 x [w1] accumulated onto the SECOND column block, so block 0
//     holds a1w1 and block 1 holds a1w2 + a2w1; the epilogue returns block0 + 2^-11 block1 (dropped: 2^-22 a2w2).  Same
//     fp32-class accuracy with 2/3 of the MMAs, operand conversions and operand bytes; valid for |x| < 65504 (fp16 range; larger
//     inputs saturate), which is why the 3 x bf16 arithmetic stays available.
//
//   * Third arithmetic (NS = 1, MVSTER_TC3_BF16X1): ONE bf16 term per operand - the bf16-storage configuration (BASELINE
//     configs[2]): the converters round the fp32 activations to bf16 (round to nearest even, the rounding a bf16 store applies),
//     the weights are bf16 values, one MMA per tap and 16 channels, products exact, fp32 accumulation; an optional per-channel
//     fp32 scale (the BatchNorm factor, kept out of the bf16 weights) is applied to the accumulator before the bias.
//
//   * Packed operands (PB, mvster_conv_tc3_pb16 / mvster_deconv_tc3_pb16): the activations live in HBM as 16-bit operand terms
//     in the operand's own order - octet-planar [plane][C/8][NS terms][H][W][8 channels]; NS = 1: bf16 (bf16 storage), NS = 2:
//     the fp16 pair (a1, a2) of the two-term arithmetic, 32 bits per element like fp32, split ONCE by the producing layer's
//     epilogue instead of by every consumer - so a TMA box lands the halo tile of a stage directly in the operand ring in the
//     un-swizzled K-major layout the MMA descriptors address: no fp32 staging ring, no converter warps (224 threads); the
//     epilogue writes the same layout (8 channels of a pixel = one 16-byte store per term, 8 pixels of a tile row = 128
//     contiguous bytes) and reads the skip tensor from it.
//     tools/tma_microbench.cu: TMA stages such boxes as fast as the fp32 ones (0.40 us per box and SM, 6 in flight).
//
// warp 0 = activation producer (TMA), warp 1 = TMEM owner + MMA issuer, warp 2 = weight producer, warps 3-10 = fp32 -> 3 x bf16
// (or 2 x fp16) converters, warps 11-14 = epilogue.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "conv_tc3_plan.h"
#include <stdlib.h>
#include <type_traits>

namespace mvster {
namespace tc3 {
using namespace ptx;

constexpr int F_BYTES = 4 * QBYTES;          // fp32 staging of one tile-stage: [180 pixels][<= 16 channels], ONE TMA box
constexpr int A_SPLIT = 2 * PLANE;           // one 16-bit term of one tile-stage: 2 channel octets
constexpr int NCONV = 256;                 // converter threads (8 warps: one warp per SM sub-partition was latency-bound)
constexpr int CTEAM = NCONV / 2;           // ... in two teams that take alternate tile-stages
constexpr int THREADS = 96 + NCONV + 128;
constexpr int NB = 10;
#ifndef MVSTER_TC3_SKIP_DUMMY
#define MVSTER_TC3_SKIP_DUMMY 0  // 1 (A/B builds, not yet run on a GPU): skip the MMA slots of tiles a shrunk group does not have
#endif
#ifndef MVSTER_TC3_NF_SMALL
#define MVSTER_TC3_NF_SMALL 4   // fp32 staging boxes in flight for the N <= 32 layers (A/B builds: tools/tma_microbench.cu, profiles/r02_tma_microbench.md)
#endif

struct Args {
    const uint8_t* w;
    const float* bias; const float* skip; float* y;
    const float* scale;  // per output channel, applied to the accumulator before the bias (nullptr = 1)
    int round_out;       // one-term arithmetic with fp32 activations: round the stored output to bf16 values (MVSTER_TC3_ROUND_OUT)
    int out_pb16;        // packed-operand kernels: y is packed like the input, [plane][Cout/8][NS][Hout][Wout][8] (else fp32 NDHWC); skip always is
    int D, Ho, Wo, cout, relu, sx, nstage, T, tiles_x, tiles_per_plane, groups_per_plane, total_groups, zero_a;
    // epilogue addressing (see there): ncls column blocks of Cout channels; up = 2 for the depth-to-space scatter of a transposed conv
    int ncls, py0, up, lg_cout, cls_a, cls_b;
    // MVSTER_TC3_DEBUG (stage ablation for profiling; results are garbage): 1 = epilogue without global loads/stores,
    // 2 = converters only hand the slots over, 8 = no activation loads
    int debug;
    // resident != 0: the layer's nslab weight slabs fit in shared memory next to the rings - they are loaded once per CTA and
    // stay; otherwise they stream through a ring of NB slabs per group of tiles (large Cin * taps * Cout)
    int resident, nslab;
    int cin_merged;  // > 0: x_map is the 3-D map with (channels x pixels) merged; value = channels per pixel
    // two-fp16-term arithmetic only: if non-null, the converters track max |x| of the layer's input and set *overflow when a value
    // does not fit the fp16 terms (|x| >= 65504 or non-finite) - mvster_tc3_set_overflow_flag; null = no tracking, no cost
    unsigned* overflow;
};

// plain / planar-block / depth-to-space output addressing of a launch; false if an offset would not fit 32 bits
static bool set_output_mode(Args& a, int ncls, int py0, bool d2s, long long block_stride) {
    a.ncls = ncls; a.py0 = py0; a.up = d2s ? 2 : 1;
    a.lg_cout = 0;
    while ((1 << a.lg_cout) < a.cout) ++a.lg_cout;
    long long ca = 0, cb = 0;
    if (d2s) { ca = 2ll * a.Wo * a.cout; cb = a.cout; }
    else if (ncls > 1) { ca = 2 * block_stride; cb = block_stride; }
    if ((1 << a.lg_cout) != a.cout || ca * (ncls / 2 + 1) >= (1ll << 31)) return false;
    a.cls_a = (int)ca; a.cls_b = (int)cb;
    return true;
}

template <int NC, int NS, bool PB = false>
struct Cfg {
    static_assert(!PB || NS <= 2, "packed operands: one bf16 term or the fp16 pair");
    static constexpr int TMAX = NC > 64 ? 1 : 64 / NC;      // tiles accumulated side by side: 3*NC*TMAX <= 240 TMEM columns per set
    static constexpr int NF = PB ? 0 : (NC <= 32 ? MVSTER_TC3_NF_SMALL : 4);  // fp32 staging ring (TMA boxes in flight); none with packed operands
    // 16-bit operand ring (tile-stages).  Packed operands: the ring is what TMA has in flight, and a box takes ~2.3 us to land
    // (tools/tma_microbench.cu) while a group's MMAs hold up to TMAX slots - MVSTER_TC3_DEEP_RING=1 (A/B builds) deepens it where
    // the layer's weights still fit next to it
#ifndef MVSTER_TC3_DEEP_RING
#define MVSTER_TC3_DEEP_RING 0
#endif
    static constexpr int NA_PB = NS == 1 ? (MVSTER_TC3_DEEP_RING && NC <= 32 ? 24 : 16)
                                         : (MVSTER_TC3_DEEP_RING ? (NC == 16 ? 16 : NC == 32 ? 12 : 10) : 10);
    static constexpr int NA = PB ? NA_PB : (NC >= 64 ? 6 : 8);
    static constexpr int A_BYTES = NS * A_SPLIT;            // a1 | a2 | a3   (NS = 2: a1 | a2)
    static constexpr int B_BYTES = 96 * NC;                 // one (stage, tap) weight slab: [2 K-halves][3*NC rows][8 x 16 bit]
    static constexpr int BAR_BYTES = 1024;                  // mbarriers + TMEM slot
    static constexpr int NTHREADS = PB ? 96 + 128 : THREADS;
    static constexpr int SMEM_FIXED = 1024 + NF * F_BYTES + NA * A_BYTES + BAR_BYTES;  // alignment slack, rings, barriers; + weight slabs
    static constexpr int SMEM_MAX = 232448;                                         // 227 KB opt-in limit per CTA on sm_100
};

__device__ __forceinline__ uint32_t bf16x2_rn(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// two fp32 -> three packed bf16 pairs with v == t1 + t2 + t3 to 24 bits (the subtractions are exact in fp32)
__device__ __forceinline__ void split3(float x, float y, uint32_t& t1, uint32_t& t2, uint32_t& t3) {
    t1 = bf16x2_rn(x, y);
    float rx = x - __uint_as_float(t1 << 16), ry = y - __uint_as_float(t1 & 0xFFFF0000u);
    t2 = bf16x2_rn(rx, ry);
    rx -= __uint_as_float(t2 << 16);
    ry -= __uint_as_float(t2 & 0xFFFF0000u);
    t3 = bf16x2_rn(rx, ry);
}
// two fp32 -> two packed fp16 pairs with v == t1 + 2^-11 t2 to 22 bits (saturating: |v| >= 65504 is outside this arithmetic)
__device__ __forceinline__ void split2h(float x, float y, uint32_t& t1, uint32_t& t2) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(t1) : "f"(y), "f"(x));
    float bx, by;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(bx), "=f"(by) : "r"(t1));
    const float rx = (x - bx) * 2048.f, ry = (y - by) * 2048.f;  // exact: the residual has <= 13 significant bits
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(t2) : "f"(ry), "f"(rx));
}

// All MMAs of one tile and one tap in ONE asm block: the three (two) A descriptors differ by a constant, the accumulator
// blocks by a constant column offset, and every other operand is an immediate - so the issuing thread spends one add and the
// register -> uniform-register moves of (lo, d) per tile instead of re-materialising every operand per MMA (the issue thread,
// not the tensor pipe, bounded the N <= 32 layers: ncu source page, profiles/r01_conv_tc3_h16_ncu.md).
template <int NC, int NS, bool PB = false>
__device__ __forceinline__ void mma_tile(uint32_t lo, uint64_t bd, uint32_t d, uint32_t accumulate) {
    // distance of the second operand term from the first: converter layout [a1: octets][a2: octets]; packed layout [octet][a1 | a2]
    constexpr int TERM2 = (PB ? PLANE : A_SPLIT) >> 4;
    constexpr uint32_t A_HI32 = (uint32_t)((HW_ * 16) >> 4) | (1u << 14);
    if constexpr (NS == 1) {
        asm volatile(
            "{\n\t.reg .pred pa;\n\t.reg .b64 a1;\n\t"
            "setp.ne.b32 pa, %3, 0;\n\t"
            "mov.b64 a1, {%1, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, %2, %5, pa;\n\t}"  // bf16(a) x bf16(w)
            ::"r"(d), "r"(lo), "l"(bd), "r"(accumulate), "n"(A_HI32), "n"(idesc_bf16_m128(NC)) : "memory");
    } else if constexpr (NS == 3) {
        asm volatile(
            "{\n\t.reg .pred pa, pt;\n\t.reg .b32 l2, l3;\n\t.reg .b64 a1, a2, a3;\n\t"
            "setp.ne.b32 pa, %3, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
            "add.u32 l2, %1, %4;\n\tadd.u32 l3, %1, %5;\n\t"
            "mov.b64 a1, {%1, %6};\n\tmov.b64 a2, {l2, %6};\n\tmov.b64 a3, {l3, %6};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, %2, %7, pa;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, %2, %8, pt;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, %2, %9, pt;\n\t}"
            ::"r"(d), "r"(lo), "l"(bd), "r"(accumulate), "n"(A_SPLIT >> 4), "n"(2 * A_SPLIT >> 4), "n"(A_HI32),
              "n"(idesc_bf16_m128(3 * NC)), "n"(idesc_bf16_m128(2 * NC)), "n"(idesc_bf16_m128(NC)) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred pa, pt;\n\t.reg .b32 l2, d2;\n\t.reg .b64 a1, a2;\n\t"
            "setp.ne.b32 pa, %3, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
            "add.u32 l2, %1, %4;\n\tadd.u32 d2, %0, %5;\n\t"
            "mov.b64 a1, {%1, %6};\n\tmov.b64 a2, {l2, %6};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, %2, %7, pa;\n\t"   // a1 x [w1|w2]: block 0 = a1 w1, block 1 = a1 w2
            "tcgen05.mma.cta_group::1.kind::f16 [d2], a2, %2, %8, pt;\n\t}"  // a2 x [w1] onto block 1
            ::"r"(d), "r"(lo), "l"(bd), "r"(accumulate), "n"(TERM2), "n"(NC), "n"(A_HI32),
              "n"(idesc_f16_m128(2 * NC)), "n"(idesc_f16_m128(NC)) : "memory");
    }
}

// MG (opt-in, MVSTER_TC3_MERGE=1; not timed yet): the activation tensor map has channels and pixels MERGED into one dimension
// (stride-1 layers whose stage covers all channels, Cin <= 16), so a halo-tile row is one contiguous run for the TMA unit instead
// of HW_ pixel-sized pieces; the bytes land in shared memory in the same order.
template <int NC, int NS, bool MG = false, bool PB = false>
__global__ void __launch_bounds__(THREADS, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ Plan plan, const Args a) {
    using C = Cfg<NC, NS, PB>;
    constexpr int A_BYTES = C::A_BYTES;
    constexpr int NF = C::NF;
    constexpr int EPI0 = PB ? 3 : 3 + NCONV / 32;  // first epilogue warp (packed operands: no converter warps)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t f_base = base, a_base = f_base + NF * F_BYTES, bar_base = a_base + C::NA * A_BYTES, b_base = bar_base + C::BAR_BYTES;
    auto F_FULL = [&](uint32_t s) { return bar_base + 8u * s; };
    auto F_EMPTY = [&](uint32_t s) { return bar_base + 8u * (NF + s); };
    auto A_FULL = [&](uint32_t s) { return bar_base + 8u * (2 * NF + s); };
    auto A_EMPTY = [&](uint32_t s) { return bar_base + 8u * (2 * NF + C::NA + s); };
    auto B_FULL = [&](uint32_t s) { return bar_base + 8u * (2 * NF + 2 * C::NA + s); };
    auto B_EMPTY = [&](uint32_t s) { return bar_base + 8u * (2 * NF + 2 * C::NA + NB + s); };
    auto ACC_FULL = [&](uint32_t s) { return bar_base + 8u * (2 * NF + 2 * C::NA + 2 * NB + s); };
    auto ACC_EMPTY = [&](uint32_t s) { return bar_base + 8u * (2 * NF + 2 * C::NA + 2 * NB + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * NF + 2 * C::NA + 2 * NB + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NF; ++s) { mbar_init(F_FULL(s), 1); mbar_init(F_EMPTY(s), CTEAM); }
        for (int s = 0; s < C::NA; ++s) { mbar_init(A_FULL(s), PB ? 1 : CTEAM); mbar_init(A_EMPTY(s), 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(B_FULL(s), 1); mbar_init(B_EMPTY(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(ACC_FULL(s), 1); mbar_init(ACC_EMPTY(s), 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    if (a.zero_a) {  // Cin < 16: the converters never write the upper channel planes, the MMAs read them as zeros
        uint4* p = reinterpret_cast<uint4*>(smem_raw + (a_base - raw));
        for (int i = threadIdx.x; i < C::NA * A_BYTES / 16; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

    // every role walks the same sequence: groups g (T tiles of one output plane) x stages x tiles / taps
#define MVSTER_TC3_GROUP_HEAD                                                                         \
    const int plane = g / a.groups_per_plane, tile0 = (g % a.groups_per_plane) * a.T, z = plane % a.D; \
    const int Tg = min(a.T, a.tiles_per_plane - tile0);
#define MVSTER_TC3_STAGE_SKIP(s) ((unsigned)(z + plan.st[s].dz) >= (unsigned)a.D)

    if (warp == 0) {
        // ------------------------------------------------------------------ activation producer (TMA halo tiles)
        // Its own warp: sharing one issue thread with the weight stream (a ring of ~1 stage of taps) tied the activation
        // prefetch distance to the MMA progress and left the converters waiting for data half of the time.
        if constexpr (PB) {
            // packed operands: the box IS the operand tile - [<= 2 channel octets][18 rows][10 pixels][8 x bf16] - straight into the ring
            if (elect_one()) {
                uint32_t au = 0;
                for (int g = blockIdx.x; g < a.total_groups; g += gridDim.x) {
                    MVSTER_TC3_GROUP_HEAD
                    for (int s = 0; s < a.nstage; ++s) {
                        if (MVSTER_TC3_STAGE_SKIP(s)) continue;
                        const Stage st = plan.st[s];
                        const int noct = (st.nq >> 1) * NS, oct0 = (st.c0 >> 3) * NS;  // nq = 2 / 4 channel quads = 1 / 2 octets of NS term planes
                        for (int t = 0; t < Tg; ++t, ++au) {
                            const uint32_t as = au % C::NA;
                            mbar_wait(A_EMPTY(as), ((au / C::NA) & 1) ^ 1);
                            mbar_expect_tx(A_FULL(as), noct * PLANE);
                            const int ti = tile0 + t, y0 = (ti / a.tiles_x) * TH, x0 = (ti % a.tiles_x) * TW;
                            const uint32_t dst = a_base + as * A_BYTES;
                            if (a.sx == 1)  // 4-D map, pixels x channels-of-an-octet merged: a halo row is one 160-byte run
                                tma_load_4d(dst, &x_map, A_FULL(as), (x0 + st.ox) * 8, y0 + st.oy, oct0, plane + st.dz);
                            else            // 5-D map with element strides 2 along x and y: one parity class of the input
                                tma_load_5d(dst, &x_map, A_FULL(as), 0, a.sx * x0 + st.ox, a.sx * y0 + st.oy, oct0, plane + st.dz);
                        }
                    }
                }
            }
        } else if (elect_one()) {
            uint32_t fu = 0;
            for (int g = blockIdx.x; g < a.total_groups; g += gridDim.x) {
                MVSTER_TC3_GROUP_HEAD
                for (int s = 0; s < a.nstage; ++s) {
                    if (MVSTER_TC3_STAGE_SKIP(s)) continue;  // depth padding: the plane contributes nothing
                    const Stage st = plan.st[s];
                    for (int t = 0; t < Tg; ++t, ++fu) {
                        const uint32_t fs = fu % NF;
                        mbar_wait(F_EMPTY(fs), ((fu / NF) & 1) ^ 1);
                        if (a.debug & 8) { mbar_arrive(F_FULL(fs)); continue; }
                        mbar_expect_tx(F_FULL(fs), st.nq * QBYTES);
                        const int ti = tile0 + t, y0 = (ti / a.tiles_x) * TH, x0 = (ti % a.tiles_x) * TW;
                        // one box = [18][10] pixels x min(Cin,16) channels (64-byte rows); the converters re-lay it out for the MMA
                        if constexpr (MG)
                            tma_load_3d(f_base + fs * F_BYTES, &x_map, F_FULL(fs), (x0 + st.ox) * a.cin_merged, y0 + st.oy, plane + st.dz);
                        else
                            tma_load_4d(f_base + fs * F_BYTES, &x_map, F_FULL(fs), st.c0, a.sx * x0 + st.ox, a.sx * y0 + st.oy, plane + st.dz);
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ weight producer (one bulk copy per (stage, tap))
        if (elect_one()) {
            if (a.resident) {  // the whole layer once: B_FULL(0) completes when every slab has landed
                mbar_expect_tx(B_FULL(0), (uint32_t)a.nslab * C::B_BYTES);
                for (int i = 0; i < a.nslab; ++i)
                    bulk_load(b_base + i * C::B_BYTES, a.w + (size_t)i * C::B_BYTES, C::B_BYTES, B_FULL(0));
            } else {
                uint32_t bu = 0;
                for (int g = blockIdx.x; g < a.total_groups; g += gridDim.x) {
                    MVSTER_TC3_GROUP_HEAD
                    (void)tile0; (void)Tg;
                    for (int s = 0; s < a.nstage; ++s) {
                        if (MVSTER_TC3_STAGE_SKIP(s)) continue;
                        const int ntap = plan.st[s].ntap, slab0 = plan.st[s].slab0;
                        for (int tap = 0; tap < ntap; ++tap, ++bu) {
                            const uint32_t sb = bu % NB;
                            mbar_wait(B_EMPTY(sb), ((bu / NB) & 1) ^ 1);
                            mbar_expect_tx(B_FULL(sb), C::B_BYTES);
                            bulk_load(b_base + sb * C::B_BYTES, a.w + (size_t)(slab0 + tap) * C::B_BYTES, C::B_BYTES, B_FULL(sb));
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // ONE thread runs the whole role (waits included) and keeps the per-tap scalar work to a few adds: with a per-tap
        // elect / descriptor rebuild / warp sync the issue thread needed ~1000 cycles per tap whatever the MMA sizes and the
        // tensor pipe idled a third of the time (ncu source page, profiles/r01_conv_tc3_ncu.md).
        if (elect_one()) {
            // descriptor = (hi << 32) | lo; lo = start address >> 4 | LBO >> 4 << 16 (taps / splits / slots only move the address)
            constexpr uint64_t B_HI = (uint64_t)(128 >> 4) | (1ull << 14);
            constexpr uint32_t B_LBO = (uint32_t)((3 * NC * 16) >> 4) << 16;  // the A operand's LBO comes with each tap (plan.a_desc)
            // the role is instantiated twice (weights resident / streamed) so that the tap loop carries no mode test: with one
            // tile per group (N >= 64) the issue thread's ~40 instructions per tap cost as much as the tap's MMAs
            auto run_role = [&](auto resident_tag) {
            constexpr bool RES = decltype(resident_tag)::value;
            uint32_t a_slot = 0, a_par = 0, b_slot = 0, b_par = 0, gc = 0;
            if constexpr (RES) {
                mbar_wait(B_FULL(0), 0);
                tc_fence_after();
            }
            for (int g = blockIdx.x; g < a.total_groups; g += gridDim.x, ++gc) {
                MVSTER_TC3_GROUP_HEAD
                const uint32_t set = gc & 1;
                mbar_wait(ACC_EMPTY(set), ((gc >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator set
                tc_fence_after();
                const uint32_t d0 = tmem_base + set * 256u;
                uint32_t accumulate = 0;
                for (int s = 0; s < a.nstage; ++s) {
                    if (MVSTER_TC3_STAGE_SKIP(s)) continue;
                    const int ntap = plan.st[s].ntap;
                    uint32_t alo[C::TMAX], aslot[C::TMAX];
#pragma unroll
                    for (int t = 0; t < C::TMAX; ++t)
                        if (t < Tg) {
                            mbar_wait(A_FULL(a_slot), a_par);
                            aslot[t] = a_slot;
                            alo[t] = ((a_base + a_slot * A_BYTES) & 0x3FFFFu) >> 4;
                            if (++a_slot == C::NA) { a_slot = 0; a_par ^= 1; }
                        } else {
                            alo[t] = alo[0];
                        }
                    tc_fence_after();
                    uint32_t b_res = b_base + (uint32_t)plan.st[s].slab0 * C::B_BYTES;  // resident: this stage's first slab
#pragma unroll 1
                    for (int tap = 0; tap < ntap; ++tap) {
                        uint32_t b_addr;
                        if constexpr (RES) {
                            b_addr = b_res;
                            b_res += C::B_BYTES;
                        } else {
                            mbar_wait(B_FULL(b_slot), b_par);
                            tc_fence_after();
                            b_addr = b_base + b_slot * C::B_BYTES;
                        }
                        const uint64_t bd = (B_HI << 32) | (((b_addr & 0x3FFFFu) >> 4) | B_LBO);
                        const uint32_t shift = plan.a_desc[s][tap];  // start shift (bits 0-13) + LBO (bits 16-29): one add per tile
                        // every tile slot of the group is issued, also past Tg (a ragged last group of a plane): alo[] then repeats
                        // tile 0's operands and the MMAs land in accumulator columns the epilogue never reads - straight-line code
#pragma unroll
                        for (int t = 0; t < C::TMAX; ++t) {
#if MVSTER_TC3_SKIP_DUMMY  // A/B build: do not issue the slots of tiles the group does not have (profiles/r02_packed_operands.md)
                            if (t >= Tg) continue;
#endif
                            mma_tile<NC, NS, PB>(alo[t] + shift, bd, d0 + (uint32_t)(t * 3 * NC), accumulate);
                        }
                        if constexpr (!RES) {
                            umma_commit(B_EMPTY(b_slot));
                            if (++b_slot == NB) { b_slot = 0; b_par ^= 1; }
                        }
                        accumulate = 1;
                    }
#pragma unroll
                    for (int t = 0; t < C::TMAX; ++t)
                        if (t < Tg) umma_commit(A_EMPTY(aslot[t]));
                }
                umma_commit(ACC_FULL(set));
            }
            };
            if (a.resident) run_role(std::true_type{});
            else run_role(std::false_type{});
        }
    } else if (warp < EPI0) {
        // ------------------------------------------------------------------ converters: fp32 halo tile -> a1 | a2 | a3 (bf16)
        // Two teams of CTEAM threads take alternate tile-stages, so one team's barrier waits and proxy fence overlap the other
        // team's conversion (with all 8 warps on one tile-stage the role cost ~1100 cycles per tile-stage whatever the tap count).
        const int team = (threadIdx.x - 96) / CTEAM, tid = (threadIdx.x - 96) % CTEAM;
        uint32_t u = 0;
        float amax = 0.f;  // NaN-propagating running max of |input| (fp16 arithmetic with a.overflow only)
        for (int g = blockIdx.x; g < a.total_groups; g += gridDim.x) {
            MVSTER_TC3_GROUP_HEAD
            for (int s = 0; s < a.nstage; ++s) {
                if (MVSTER_TC3_STAGE_SKIP(s)) continue;
                const int nq = plan.st[s].nq, items = HPIX * nq, qsh = nq >> 1;  // nq in {1,2,4}: pixel = item >> qsh
                for (int t = 0; t < Tg; ++t, ++u) {
                    if ((int)(u & 1) != team) continue;
                    constexpr int NFD = NF ? NF : 1;  // (this role does not exist with packed operands: NF = 0)
                    const uint32_t fs = u % NFD, as = u % C::NA;
                    mbar_wait(F_FULL(fs), (u / NFD) & 1);
                    mbar_wait(A_EMPTY(as), ((u / C::NA) & 1) ^ 1);
                    const uint8_t* F = smem_raw + (f_base + fs * F_BYTES - raw);
                    uint8_t* A = smem_raw + (a_base + as * A_BYTES - raw);
                    // item = one channel quad of one halo pixel = 16 contiguous bytes of F; <= 6 items per thread, all loads first
                    if (a.debug & 2) {
                        mbar_arrive(A_FULL(as));
                        mbar_arrive(F_EMPTY(fs));
                        continue;
                    }
                    constexpr int IT = (HPIX * 4 + CTEAM - 1) / CTEAM;
                    float4 v[IT];
#pragma unroll
                    for (int k = 0; k < IT; ++k) {
                        const int i = tid + k * CTEAM;
                        if (i < items) v[k] = *reinterpret_cast<const float4*>(F + i * 16);
                    }
#pragma unroll
                    for (int k = 0; k < IT; ++k) {
                        const int i = tid + k * CTEAM;
                        if (i < items) {
                            const int p = i >> qsh, q = i & (nq - 1);
                            uint8_t* dst = A + (q >> 1) * PLANE + p * 16 + (q & 1) * 8;
                            if constexpr (NS == 1) {
                                *reinterpret_cast<uint2*>(dst) = make_uint2(bf16x2_rn(v[k].x, v[k].y), bf16x2_rn(v[k].z, v[k].w));
                            } else if constexpr (NS == 3) {
                                uint2 t1, t2, t3;
                                split3(v[k].x, v[k].y, t1.x, t2.x, t3.x);
                                split3(v[k].z, v[k].w, t1.y, t2.y, t3.y);
                                *reinterpret_cast<uint2*>(dst) = t1;
                                *reinterpret_cast<uint2*>(dst + A_SPLIT) = t2;
                                *reinterpret_cast<uint2*>(dst + 2 * A_SPLIT) = t3;
                            } else {
                                if (a.overflow) {
                                    float m4;
                                    asm("{\n\t.reg .f32 t, u;\n\tmax.NaN.f32 t, %1, %2;\n\tmax.NaN.f32 u, %3, %4;\n\tmax.NaN.f32 %0, t, u;\n\t}"
                                        : "=f"(m4) : "f"(fabsf(v[k].x)), "f"(fabsf(v[k].y)), "f"(fabsf(v[k].z)), "f"(fabsf(v[k].w)));
                                    asm("max.NaN.f32 %0, %0, %1;" : "+f"(amax) : "f"(m4));
                                }
                                uint2 t1, t2;
                                split2h(v[k].x, v[k].y, t1.x, t2.x);
                                split2h(v[k].z, v[k].w, t1.y, t2.y);
                                *reinterpret_cast<uint2*>(dst) = t1;
                                *reinterpret_cast<uint2*>(dst + A_SPLIT) = t2;
                            }
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive(A_FULL(as));
                    mbar_arrive(F_EMPTY(fs));
                }
            }
        }
        if constexpr (NS == 2) {
            if (a.overflow && !(amax < 65504.f)) atomicOr(a.overflow, 1u);  // also true for NaN
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3, r = q * 32 + lane;  // TMEM lane quarter, accumulator row = pixel in the tile
        uint32_t gc = 0;
        [[maybe_unused]] bool bad = false;  // packed fp16-pair output: a stored value outside the fp16 terms' range
        for (int g = blockIdx.x; g < a.total_groups; g += gridDim.x, ++gc) {
            MVSTER_TC3_GROUP_HEAD
            (void)z;
            const uint32_t set = gc & 1;
            mbar_wait(ACC_FULL(set), (gc >> 1) & 1);
            tc_fence_after();
            for (int t = 0; t < Tg; ++t) {
                const int ti = tile0 + t, yy = (ti / a.tiles_x) * TH + r / TW, xx = (ti % a.tiles_x) * TW + r % TW;
                const bool ok = yy < a.Ho && xx < a.Wo && !(a.debug & 1);
                const uint32_t col = tmem_base + ((uint32_t)(q * 32) << 16) + set * 256u + (uint32_t)(t * 3 * NC);
                const int ncol = a.ncls * a.cout;
                // Column c = class * Cout + channel (Cout a power of two).  Plain conv: one class.  Transposed conv (depth-to-space,
                // up = 2): class (py, px) of input pixel (yy, xx) is output pixel (2 yy + py, 2 xx + px): cls_a = one output row,
                // cls_b = one output pixel.  Planar blocks: class j is its own tensor, cls_b = its distance, cls_a = 2 cls_b.
                // One 64-bit base per tile and pixel; everything per column is 32-bit.
                const long long base = (((long long)plane * (a.up * a.Ho) + a.up * yy + a.py0) * (a.up * a.Wo) + a.up * xx) * a.cout;
                float* const yb = a.y + base;
                const float* const sb = a.skip + base;
                auto out_offset = [&](int c, int& ch) -> int {
                    const int cls = c >> a.lg_cout;
                    ch = c & (a.cout - 1);
                    return (cls >> 1) * a.cls_a + (cls & 1) * a.cls_b + ch;
                };
                if constexpr (PB) {
                    // Packed operands: the skip tensor and (out_pb16) the output are octet-planar 16-bit terms,
                    // [plane][Cout/8][NS][Hout][Wout][8]: 8 accumulator columns = one channel octet of one output pixel = one 16-byte
                    // access per term.  Column c = class * Cout + channel as above; class (py, px) of a transposed layer is output
                    // pixel (2 yy + py0 + py, 2 xx + px).
                    const int Hout = a.up * a.Ho, Wout = a.up * a.Wo, noct_out = a.cout >> 3;
                    const long long term_stride = (long long)Hout * Wout;  // 16-byte units between the a1 and the a2 plane of an octet
                    const uint4* const skip_pb = reinterpret_cast<const uint4*>(a.skip);
                    uint4* const y_pb = reinterpret_cast<uint4*>(a.y);
#pragma unroll
                    for (int c0 = 0; c0 < NC; c0 += 16) {
                        uint32_t v1[16], v2[16];
                        tmem_ld16(col + c0, v1);
                        if constexpr (NS == 2) tmem_ld16(col + NC + c0, v2);
                        tmem_ld_wait();
                        if (ok && c0 < ncol) {
#pragma unroll
                            for (int h8 = 0; h8 < 16; h8 += 8) {
                                const int c = c0 + h8;
                                if (c >= ncol) break;
                                const int cls = c >> a.lg_cout, ch = c & (a.cout - 1);
                                const int Y = a.up * yy + a.py0 + (a.up == 2 ? (cls >> 1) : 0), X = a.up * xx + (a.up == 2 ? (cls & 1) : 0);
                                const long long pix_o = ((long long)(plane * noct_out + (ch >> 3)) * NS * Hout + Y) * Wout + X;  // 16-byte units
                                float o[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    if constexpr (NS == 2) o[e] = fmaf(__uint_as_float(v2[h8 + e]), 1.f / 2048.f, __uint_as_float(v1[h8 + e]));
                                    else o[e] = __uint_as_float(v1[h8 + e]);
                                    if (a.scale) o[e] *= __ldg(a.scale + ch + e);
                                    if (a.bias) o[e] += __ldg(a.bias + ch + e);
                                    if (a.relu) o[e] = fmaxf(o[e], 0.f);
                                }
                                if (a.skip) {
                                    const uint4 s4 = __ldg(skip_pb + pix_o);
                                    if constexpr (NS == 2) {  // skip = a1 + 2^-11 a2 (exact in fp32)
                                        const uint4 r4 = __ldg(skip_pb + pix_o + term_stride);
                                        const uint32_t w1[4] = {s4.x, s4.y, s4.z, s4.w}, w2[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            float a1x, a1y, a2x, a2y;
                                            asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(a1x), "=f"(a1y) : "r"(w1[e]));
                                            asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(a2x), "=f"(a2y) : "r"(w2[e]));
                                            o[2 * e] += fmaf(a2x, 1.f / 2048.f, a1x);
                                            o[2 * e + 1] += fmaf(a2y, 1.f / 2048.f, a1y);
                                        }
                                    } else {
                                        o[0] += __uint_as_float(s4.x << 16); o[1] += __uint_as_float(s4.x & 0xFFFF0000u);
                                        o[2] += __uint_as_float(s4.y << 16); o[3] += __uint_as_float(s4.y & 0xFFFF0000u);
                                        o[4] += __uint_as_float(s4.z << 16); o[5] += __uint_as_float(s4.z & 0xFFFF0000u);
                                        o[6] += __uint_as_float(s4.w << 16); o[7] += __uint_as_float(s4.w & 0xFFFF0000u);
                                    }
                                }
                                if (a.out_pb16) {
                                    if constexpr (NS == 2) {  // the next layer's operand terms, split here once: a = a1 + 2^-11 a2
                                        if (a.overflow) {     // range check of the fp16 terms (mvster_tc3_set_overflow_flag), on what this layer stores
#pragma unroll
                                            for (int e = 0; e < 8; ++e) bad |= !(fabsf(o[e]) < 65504.f);  // also true for NaN
                                        }
                                        uint4 t1, t2;
                                        split2h(o[0], o[1], t1.x, t2.x); split2h(o[2], o[3], t1.y, t2.y);
                                        split2h(o[4], o[5], t1.z, t2.z); split2h(o[6], o[7], t1.w, t2.w);
                                        y_pb[pix_o] = t1;
                                        y_pb[pix_o + term_stride] = t2;
                                    } else {
                                        y_pb[pix_o] = make_uint4(bf16x2_rn(o[0], o[1]), bf16x2_rn(o[2], o[3]), bf16x2_rn(o[4], o[5]), bf16x2_rn(o[6], o[7]));
                                    }
                                } else {  // fp32 NDHWC (the regulariser's last layer, read by the head)
                                    float* dstf = a.y + (((long long)plane * Hout + Y) * Wout + X) * a.cout + ch;
                                    *reinterpret_cast<float4*>(dstf) = make_float4(o[0], o[1], o[2], o[3]);
                                    *reinterpret_cast<float4*>(dstf + 4) = make_float4(o[4], o[5], o[6], o[7]);
                                }
                            }
                        }
                    }
                    continue;
                }
                constexpr int RC = NC % 32 ? 16 : 32;  // columns per round: their skip values are fetched up front, all in flight
#pragma unroll
                for (int cb = 0; cb < NC; cb += RC) {
                    float4 sk[RC / 4];
                    if (a.skip && ok) {
#pragma unroll
                        for (int jg = 0; jg < RC / 4; ++jg) {
                            int ch;
                            if (cb + 4 * jg < ncol) sk[jg] = __ldg(reinterpret_cast<const float4*>(sb + out_offset(cb + 4 * jg, ch)));
                        }
                    }
#pragma unroll
                    for (int c0 = cb; c0 < cb + RC; c0 += 16) {
                        uint32_t v1[16], v2[16], v3[16];
                        tmem_ld16(col + c0, v1);
                        if constexpr (NS >= 2) tmem_ld16(col + NC + c0, v2);
                        if constexpr (NS == 3) tmem_ld16(col + 2 * NC + c0, v3);
                        tmem_ld_wait();
                        if (ok && c0 < ncol) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                if (c0 + j >= ncol) break;
                                int ch;
                                const int o_off = out_offset(c0 + j, ch);
                                float o[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    if constexpr (NS == 3) o[e] = (__uint_as_float(v3[j + e]) + __uint_as_float(v2[j + e])) + __uint_as_float(v1[j + e]);
                                    else if constexpr (NS == 2) o[e] = fmaf(__uint_as_float(v2[j + e]), 1.f / 2048.f, __uint_as_float(v1[j + e]));
                                    else o[e] = __uint_as_float(v1[j + e]);
                                    if (a.scale) o[e] *= __ldg(a.scale + ch + e);
                                    if (a.bias) o[e] += __ldg(a.bias + ch + e);
                                    if (a.relu) o[e] = fmaxf(o[e], 0.f);
                                }
                                if (a.skip) {
                                    const float4 s4 = sk[(c0 - cb + j) / 4];
                                    o[0] += s4.x; o[1] += s4.y; o[2] += s4.z; o[3] += s4.w;
                                }
                                if constexpr (NS == 1) {
                                    if (a.round_out) {  // what a bf16 store of this layer's output keeps
                                        const uint32_t p0 = bf16x2_rn(o[0], o[1]), p1 = bf16x2_rn(o[2], o[3]);
                                        o[0] = __uint_as_float(p0 << 16); o[1] = __uint_as_float(p0 & 0xFFFF0000u);
                                        o[2] = __uint_as_float(p1 << 16); o[3] = __uint_as_float(p1 & 0xFFFF0000u);
                                    }
                                }
                                *reinterpret_cast<float4*>(yb + o_off) = make_float4(o[0], o[1], o[2], o[3]);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(ACC_EMPTY(set));
        }
        if constexpr (PB && NS == 2) {
            if (a.overflow && bad) atomicOr(a.overflow, 1u);
        }
    }
#undef MVSTER_TC3_GROUP_HEAD
#undef MVSTER_TC3_STAGE_SKIP
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

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

// Persistent CTAs keep their SM for the whole launch, so the grid size is also how much of the GPU a launch claims.
// mvster_set_sm_budget(n) caps the grid of the launches that follow (0 = every SM): the host uses it to run the small
// early cascade stages on a second stream next to the feature pyramid's large layers (engine.py).
// Thread-local: under nn.DataParallel (test_mvs4.py:196) every replica drives its own device from its own host thread, and a
// cap set by one replica must not leak into another replica's launches (or into a CUDA-graph capture running beside it).
static thread_local int g_sm_budget = 0;
// Device word that the two-fp16-term launches of the calling thread OR a 1 into when a layer input does not fit the fp16 terms
// (mvster_tc3_set_overflow_flag; nullptr = off).  Thread-local for the same reason as the SM cap.
static thread_local unsigned* g_overflow_flag = nullptr;

// SM count of the CURRENT device (replicas may sit on different devices; cached per device id).
static int current_sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int n = cache[dev];
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = n;
    }
    return n;
}

template <int NC, int NS, bool MG = false, bool PB = false>
static int launch_ns(const CUtensorMap& xm, const Plan& plan, Args& a, long long total_tiles, int sms, cudaStream_t st) {
    using C = Cfg<NC, NS, PB>;
    if (g_sm_budget > 0 && g_sm_budget < sms) sms = g_sm_budget;
    int T = C::TMAX;
    while (T > 1 && total_tiles < (long long)T * 2 * sms) T >>= 1;  // keep every SM busy before widening the groups
    if (T > a.tiles_per_plane) T = a.tiles_per_plane;
    a.T = T;
    a.groups_per_plane = ceil_div(a.tiles_per_plane, T);
    a.total_groups = (int)(total_tiles / a.tiles_per_plane) * a.groups_per_plane;
    auto k = conv_tc3_kernel<NC, NS, MG, PB>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_MAX) != cudaSuccess) {
        set_error("conv_tc3_kernel: cannot reserve %d bytes of shared memory", C::SMEM_MAX);
        cudaGetLastError();
        return MVSTER_ERR_CUDA;
    }
    // weights resident in shared memory when the whole layer fits next to the rings (MVSTER_TC3_STREAM=1 forces the ring)
    static const bool force_stream = getenv("MVSTER_TC3_STREAM") && atoi(getenv("MVSTER_TC3_STREAM")) != 0;
    a.resident = !force_stream && a.nslab > 0 && C::SMEM_FIXED + a.nslab * C::B_BYTES <= C::SMEM_MAX;
    const int grid = a.total_groups < sms ? a.total_groups : sms;
    // MVSTER_TC3_STREAM_SMALL=1 (A/B): a launch whose CTAs each process ONE group reuses no slab, and the resident form makes its
    // first MMA wait for the whole layer's weights (up to 110 KB per CTA); streamed, the first MMA starts after the first slab
    static const bool stream_small = getenv("MVSTER_TC3_STREAM_SMALL") && atoi(getenv("MVSTER_TC3_STREAM_SMALL")) != 0;
    if (stream_small && a.total_groups <= sms) a.resident = 0;
    const int smem = C::SMEM_FIXED + (a.resident ? a.nslab : NB) * C::B_BYTES;
    // (Programmatic dependent launch between consecutive convolutions - griddepcontrol.launch_dependents at the top, .wait before
    // the first activation load and the epilogue - was measured and removed: -60 us (3.4 %) per step with one stream, but with
    // the two-stream forward the replayed graph stalled for 7-100 ms every few steps; profiles/r02_pdl.md.)
    k<<<grid, C::NTHREADS, smem, st>>>(xm, plan, a);
    return check_launch("conv_tc3_kernel");
}

// terms = operand terms of the arithmetic: 3 (bf16 x 3, the default), 2 (MVSTER_TC3_FP16X2), 1 (MVSTER_TC3_BF16X1)
static int arith_terms(int relu_arg) { return (relu_arg & MVSTER_TC3_BF16X1) ? 1 : (relu_arg & MVSTER_TC3_FP16X2) ? 2 : 3; }

template <int NC>
static int launch(const CUtensorMap& xm, const Plan& plan, Args& a, long long total_tiles, int sms, cudaStream_t st, int terms) {
    if (terms == 1) return launch_ns<NC, 1>(xm, plan, a, total_tiles, sms, st);  // (the merged-dimension map is not instantiated for it)
    const bool fp16x2 = terms == 2;
    if constexpr (NC <= 32) {  // merged-dimension activation map: only small-Cin layers qualify (conv_tc3_run decides)
        if (a.cin_merged > 0)
            return fp16x2 ? launch_ns<NC, 2, true>(xm, plan, a, total_tiles, sms, st) : launch_ns<NC, 3, true>(xm, plan, a, total_tiles, sms, st);
    }
    return fp16x2 ? launch_ns<NC, 2>(xm, plan, a, total_tiles, sms, st) : launch_ns<NC, 3>(xm, plan, a, total_tiles, sms, st);
}

}  // namespace tc3
}  // namespace mvster

using namespace mvster;

extern "C" void mvster_set_sm_budget(int n) { tc3::g_sm_budget = n > 0 ? n : 0; }
extern "C" void mvster_tc3_set_overflow_flag(unsigned* device_flag) { tc3::g_overflow_flag = device_flag; }
extern "C" unsigned* mvster_tc3_overflow_flag(void) { return tc3::g_overflow_flag; }

extern "C" int mvster_conv_tc3_supported(int Cin, int Cout, int kd, int k, int stride_hw) {
    return tc3::supported(Cin, Cout, kd, k, stride_hw);
}

extern "C" int mvster_conv_tc3_plan(int Cin, int kd, int k, int stride_hw, int* slabs, int max_slabs) {
    if (!tc3::supported(Cin, 16, kd, k, stride_hw)) return -1;
    const int n = tc3::build_plan(Cin, kd, k, stride_hw, nullptr, nullptr);
    if (slabs) {
        if (n > max_slabs) return -1;
        tc3::build_plan(Cin, kd, k, stride_hw, nullptr, reinterpret_cast<int(*)[6]>(slabs));
    }
    return n;
}

extern "C" size_t mvster_conv_tc3_packed_bytes(int Cin, int Cout, int kd, int k, int stride_hw) {
    if (!tc3::supported(Cin, Cout, kd, k, stride_hw)) return 0;
    const int NC = Cout < 16 ? 16 : (Cout > 64 ? 80 : Cout);
    return (size_t)tc3::build_plan(Cin, kd, k, stride_hw, nullptr, nullptr) * 96 * NC;
}

// block == 0: y [B][D][Ho][Wo][Cout].  block > 0 (divides Cout): output channels [j*block, (j+1)*block) go to a separate
// tensor [B][D][Ho][Wo][block] at y + j*block_stride (floats).
static int conv_tc3_run(const float* x, const void* w_packed, const float* scale, const float* bias, const float* skip, float* y,
                        int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int stride_hw, int relu,
                        int block, long long block_stride, mvster_stream_t stream) {
    using namespace mvster::tc3;
    MVSTER_REQUIRE(x && w_packed && y, "mvster_conv_tc3_f32: null pointer");
    MVSTER_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "mvster_conv_tc3_f32: bad shape");
    MVSTER_REQUIRE(supported(Cin, Cout, kd, k, stride_hw), "mvster_conv_tc3_f32: unsupported layer Cin=%d Cout=%d kd=%d k=%d stride=%d",
                   Cin, Cout, kd, k, stride_hw);
    MVSTER_REQUIRE(((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)x & 15) == 0, "mvster_conv_tc3_f32: x and w_packed must be 16-byte aligned");
    EncodeTiledFn enc = encode_fn();
    MVSTER_REQUIRE(enc, "mvster_conv_tc3_f32: cuTensorMapEncodeTiled is unavailable in this driver");
    const int sms = tc3::current_sm_count();
    const int s = stride_hw;
    CUtensorMap xm;
    static const bool want_merge = getenv("MVSTER_TC3_MERGE") && atoi(getenv("MVSTER_TC3_MERGE")) != 0;
    const int terms = arith_terms(relu);
    const bool merged = want_merge && terms != 1 && s == 1 && Cin <= 16 && Cout <= 32 && (long long)W * Cin < (1ll << 31);
    if (merged) {  // [B*D][H][W*C]: a box row = the halo row's HW_ pixels x Cin channels as ONE contiguous run
        cuuint64_t dims[3] = {(cuuint64_t)W * Cin, (cuuint64_t)H, (cuuint64_t)B * D};
        cuuint64_t strides[2] = {(cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4};
        cuuint32_t box[3] = {(cuuint32_t)(HW_ * Cin), (cuuint32_t)HH_, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVSTER_REQUIRE(r == CUDA_SUCCESS, "mvster_conv_tc3_f32: merged activation tensor map rejected (CUresult %d)", (int)r);
    } else {   // activations [B*D][H][W][C]; a box is <= 16 channels of an 18 x 10 pixel halo patch, every s-th pixel
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * D};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4};
        cuuint32_t box[4] = {(cuuint32_t)(Cin < 16 ? Cin : 16), (cuuint32_t)(HW_ * s), (cuuint32_t)(HH_ * s), 1}, es[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVSTER_REQUIRE(r == CUDA_SUCCESS, "mvster_conv_tc3_f32: activation tensor map rejected (CUresult %d)", (int)r);
    }
    Plan plan;
    memset(&plan, 0, sizeof(plan));
    const int nslab = build_plan(Cin, kd, k, s, &plan, nullptr);
    Args a;
    a.nslab = nslab;
    a.cin_merged = merged ? Cin : 0;
    a.w = (const uint8_t*)w_packed; a.bias = bias; a.skip = skip; a.y = y; a.scale = scale;
    a.round_out = (terms == 1 && (relu & MVSTER_TC3_ROUND_OUT)) ? 1 : 0; a.out_pb16 = 0;
    a.D = D; a.Ho = (H - 1) / s + 1; a.Wo = (W - 1) / s + 1; a.cout = Cout; a.relu = relu & 1; a.sx = s;
    a.nstage = kd * (s == 2 ? 4 : 1) * ((Cin + 15) / 16);
    a.tiles_x = ceil_div(a.Wo, TW);
    a.tiles_per_plane = a.tiles_x * ceil_div(a.Ho, TH);
    { const char* dbg = getenv("MVSTER_TC3_DEBUG"); a.debug = dbg ? atoi(dbg) : 0; }
    a.overflow = tc3::g_overflow_flag;
    a.zero_a = Cin < 8;  // Cin = 4: each 16-byte row holds 4 real channels, the other 4 must read as zero (Cin = 8 fills the one plane it reads)
    const bool blocks = block > 0 && block < Cout;
    if (blocks) a.cout = block;
    MVSTER_REQUIRE(set_output_mode(a, blocks ? Cout / block : 1, 0, false, block_stride),
                   "mvster_conv_tc3_f32: output addressing does not fit (Cout block %d must be a power of two, blocks < 2^31 floats apart)", a.cout);
    const long long total_tiles = (long long)a.tiles_per_plane * B * D;
    cudaStream_t st = (cudaStream_t)stream;
    MVSTER_REQUIRE(Cout <= 64 || blocks, "mvster_conv_tc3_f32: Cout = %d only as separate channel blocks", Cout);
    const int NC = Cout < 16 ? 16 : (Cout > 64 ? 80 : Cout);
    if (NC == 16) return launch<16>(xm, plan, a, total_tiles, sms, st, terms);
    if (NC == 32) return launch<32>(xm, plan, a, total_tiles, sms, st, terms);
    if (NC == 80) return launch<80>(xm, plan, a, total_tiles, sms, st, terms);
    return launch<64>(xm, plan, a, total_tiles, sms, st, terms);
}

extern "C" int mvster_conv_tc3_f32(const float* x, const void* w_packed, const float* bias, const float* skip, float* y,
                                   int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int stride_hw, int relu,
                                   mvster_stream_t stream) {
    return conv_tc3_run(x, w_packed, nullptr, bias, skip, y, B, D, H, W, Cin, Cout, kd, k, stride_hw, relu, 0, 0, stream);
}

extern "C" int mvster_conv_tc3_scaled_f32(const float* x, const void* w_packed, const float* scale, const float* bias, const float* skip,
                                          float* y, int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int stride_hw, int relu,
                                          mvster_stream_t stream) {
    return conv_tc3_run(x, w_packed, scale, bias, skip, y, B, D, H, W, Cin, Cout, kd, k, stride_hw, relu, 0, 0, stream);
}

extern "C" int mvster_pointwise_tc3_blocks_ex_f32(const float* x, const void* w_packed, float* y, int N, int H, int W, int Cin, int Cout,
                                                  int block, long long block_stride_floats, int flags, mvster_stream_t stream) {
    MVSTER_REQUIRE(block >= 4 && block % 4 == 0 && Cout % block == 0 && (Cout == block || block_stride_floats % 4 == 0),
                   "mvster_pointwise_tc3_blocks_f32: bad block %d for Cout %d", block, Cout);
    return conv_tc3_run(x, w_packed, nullptr, nullptr, nullptr, y, N, 1, H, W, Cin, Cout, 1, 1, 1, flags & MVSTER_TC3_FP16X2, block,
                        block_stride_floats, stream);
}

extern "C" int mvster_pointwise_tc3_blocks_f32(const float* x, const void* w_packed, float* y, int N, int H, int W, int Cin, int Cout,
                                               int block, long long block_stride_floats, mvster_stream_t stream) {
    return mvster_pointwise_tc3_blocks_ex_f32(x, w_packed, y, N, H, W, Cin, Cout, block, block_stride_floats, 0, stream);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Transposed convolution, kernel (1,3,3), stride (1,2,2), padding 1, output padding 1 (Deconv3d of reg2d, mvs4net_utils.py:
// 893-897): out[2i - 1 + k] += in[i] w[k].  Output pixel (2y + py, 2x + px) depends on input rows y + dy with
//     py = 0: (dy = 0, ky = 1)            py = 1: (dy = 0, ky = 2), (dy = 1, ky = 0)        (columns alike)
// so the layer is a 2 x 2 stride-1 convolution on the INPUT grid (taps (dy,dx) in {0,1}^2 = halo offsets 1..2) producing the
// output parity classes side by side along N - [class (py,px)][Cout] columns, zero weights where a tap does not feed a class -
// followed by a depth-to-space scatter in the epilogue.  rows = -1: all four classes in one launch (4*Cout <= 64 columns);
// rows = 0 / 1: only the output rows of parity py (2*Cout columns; two launches cover the layer when 4*Cout > 64).
// Slab order: [16-channel chunk][tap (dy,dx) row-major over the taps the launch needs]; rows = 0 needs only dy = 0.
using mvster::tc3::deconv_ncls;
using mvster::tc3::deconv_ntap;

extern "C" int mvster_deconv_tc3_supported(int Cin, int Cout, int rows) {
    const int n = deconv_ncls(rows) * Cout;
    return (Cin == 16 || Cin == 32 || Cin == 64) && (Cout == 8 || Cout == 16 || Cout == 32) && rows >= -1 && rows <= 1 &&
           (n == 16 || n == 32 || n == 64);
}

extern "C" size_t mvster_deconv_tc3_packed_bytes(int Cin, int Cout, int rows) {
    if (!mvster_deconv_tc3_supported(Cin, Cout, rows)) return 0;
    return (size_t)(Cin / 16) * deconv_ntap(rows) * 96 * deconv_ncls(rows) * Cout;
}

extern "C" int mvster_deconv_tc3_f32(const float* x, const void* w_packed, const float* bias, const float* skip, float* y,
                                     int B, int D, int H, int W, int Cin, int Cout, int rows, int relu, mvster_stream_t stream) {
    return mvster_deconv_tc3_scaled_f32(x, w_packed, nullptr, bias, skip, y, B, D, H, W, Cin, Cout, rows, relu, stream);
}

extern "C" int mvster_deconv_tc3_scaled_f32(const float* x, const void* w_packed, const float* scale, const float* bias, const float* skip,
                                            float* y, int B, int D, int H, int W, int Cin, int Cout, int rows, int relu,
                                            mvster_stream_t stream) {
    using namespace mvster::tc3;
    MVSTER_REQUIRE(x && w_packed && y, "mvster_deconv_tc3_f32: null pointer");
    MVSTER_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "mvster_deconv_tc3_f32: bad shape");
    MVSTER_REQUIRE(mvster_deconv_tc3_supported(Cin, Cout, rows), "mvster_deconv_tc3_f32: unsupported layer Cin=%d Cout=%d rows=%d", Cin, Cout, rows);
    MVSTER_REQUIRE(((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)x & 15) == 0, "mvster_deconv_tc3_f32: x and w_packed must be 16-byte aligned");
    EncodeTiledFn enc = encode_fn();
    MVSTER_REQUIRE(enc, "mvster_deconv_tc3_f32: cuTensorMapEncodeTiled is unavailable in this driver");
    const int sms = tc3::current_sm_count();
    CUtensorMap xm;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * D};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4};
        cuuint32_t box[4] = {16, HW_, HH_, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVSTER_REQUIRE(r == CUDA_SUCCESS, "mvster_deconv_tc3_f32: activation tensor map rejected (CUresult %d)", (int)r);
    }
    Plan plan;
    memset(&plan, 0, sizeof(plan));
    const int kch = Cin / 16, ntap = deconv_ntap(rows);
    for (int kc = 0; kc < kch; ++kc) {
        Stage S;
        S.c0 = (short)(kc * 16); S.nq = 4; S.ox = -1; S.oy = -1; S.dz = 0; S.ntap = (short)ntap; S.slab0 = (short)(kc * ntap); S.pad = 0;
        plan.st[kc] = S;
        for (int t = 0; t < ntap; ++t) plan.a_desc[kc][t] = tap_desc((t / 2 + 1) * HW_ + (t % 2 + 1), PLANE >> 4);  // halo (1 + dy, 1 + dx)
    }
    Args a;
    a.w = (const uint8_t*)w_packed; a.bias = bias; a.skip = skip; a.y = y; a.scale = scale;
    const int terms = arith_terms(relu);
    a.round_out = (terms == 1 && (relu & MVSTER_TC3_ROUND_OUT)) ? 1 : 0; a.out_pb16 = 0;
    a.D = D; a.Ho = H; a.Wo = W; a.cout = Cout; a.relu = relu & 1; a.sx = 1; a.nstage = kch;
    a.nslab = kch * ntap;
    a.cin_merged = 0;
    a.tiles_x = ceil_div(W, TW);
    a.tiles_per_plane = a.tiles_x * ceil_div(H, TH);
    a.zero_a = 0;
    { const char* dbg = getenv("MVSTER_TC3_DEBUG"); a.debug = dbg ? atoi(dbg) : 0; }
    a.overflow = tc3::g_overflow_flag;
    MVSTER_REQUIRE(set_output_mode(a, deconv_ncls(rows), rows == 1 ? 1 : 0, true, 0), "mvster_deconv_tc3_f32: output row pitch does not fit 32 bits");
    const long long total_tiles = (long long)a.tiles_per_plane * B * D;
    cudaStream_t st = (cudaStream_t)stream;
    const int NC = a.ncls * Cout;
    if (NC == 16) return launch<16>(xm, plan, a, total_tiles, sms, st, terms);
    if (NC == 32) return launch<32>(xm, plan, a, total_tiles, sms, st, terms);
    return launch<64>(xm, plan, a, total_tiles, sms, st, terms);
}


// ---------------------------------------------------------------------------------------------------------------------------
// Packed operands: x, skip and (MVSTER_TC3_OUT_PB16) y are octet-planar 16-bit terms, [B*D][C/8][NT][H][W][8 channels]; NT = 1:
// bf16, NT = 2 (MVSTER_TC3_FP16X2): the fp16 pair.  The (octet, term) planes of a stage are consecutive: one box dimension.
static int encode_pb16_map(CUtensorMap* xm, const void* x, int P, int H, int W, int Cin, int s, int NT, const char* who) {
    using namespace mvster::tc3;
    EncodeTiledFn enc = encode_fn();
    MVSTER_REQUIRE(enc, "%s: cuTensorMapEncodeTiled is unavailable in this driver", who);
    const cuuint32_t noct = (Cin >= 16 ? 2 : 1) * NT;  // (octet, term) planes per stage
    const CUtensorMapDataType dt = NT == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r;
    if (s == 1) {  // {W*8 elements, H, C/8, P}: pixels and the 8 channels of an octet are one contiguous run
        cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)(Cin / 8 * NT), (cuuint64_t)P};
        cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * 2 * Cin * NT};
        cuuint32_t box[4] = {(cuuint32_t)(HW_ * 8), (cuuint32_t)HH_, noct, 1}, es[4] = {1, 1, 1, 1};
        r = enc(xm, dt, 4, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {       // {8, W, H, C/8, P} with element strides s along x and y: one parity class of a stride-2 layer per box
        cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(Cin / 8 * NT), (cuuint64_t)P};
        cuuint64_t strides[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * 2 * Cin * NT};
        cuuint32_t box[5] = {8, (cuuint32_t)(HW_ * s), (cuuint32_t)(HH_ * s), noct, 1}, es[5] = {1, (cuuint32_t)s, (cuuint32_t)s, 1, 1};
        r = enc(xm, dt, 5, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    MVSTER_REQUIRE(r == CUDA_SUCCESS, "%s: packed activation tensor map rejected (CUresult %d)", who, (int)r);
    return MVSTER_OK;
}

template <int NC>
static int launch_pb(const CUtensorMap& xm, const mvster::tc3::Plan& plan, mvster::tc3::Args& a, long long total_tiles, int sms, cudaStream_t st, int NT) {
    if (NT == 2) return mvster::tc3::launch_ns<NC, 2, false, true>(xm, plan, a, total_tiles, sms, st);
    return mvster::tc3::launch_ns<NC, 1, false, true>(xm, plan, a, total_tiles, sms, st);
}

// fp16-pair operands: the two channel octets of an MMA's K halves are 2 planes apart ([octet][a1 | a2]); the paired-tap layers
// (Cin <= 8: second K half = a second tap of the same plane) keep their tap distance
static void widen_octet_lbo(mvster::tc3::Plan& plan, int nstage) {
    using namespace mvster::tc3;
    for (int s = 0; s < nstage; ++s)
        for (int t = 0; t < plan.st[s].ntap; ++t)
            if ((plan.a_desc[s][t] >> 16) == (uint32_t)(PLANE >> 4)) plan.a_desc[s][t] = tap_desc(plan.a_desc[s][t] & 0xFFFFu, 2 * PLANE >> 4);
}

extern "C" int mvster_conv_tc3_pb16(const void* x, const void* w_packed, const float* scale, const float* bias, const void* skip, void* y,
                                    int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int stride_hw, int flags,
                                    mvster_stream_t stream) {
    using namespace mvster::tc3;
    MVSTER_REQUIRE(x && w_packed && y, "mvster_conv_tc3_pb16: null pointer");
    MVSTER_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "mvster_conv_tc3_pb16: bad shape");
    MVSTER_REQUIRE(supported(Cin, Cout, kd, k, stride_hw) && Cin >= 8 && Cout <= 64, "mvster_conv_tc3_pb16: unsupported layer Cin=%d Cout=%d kd=%d k=%d stride=%d",
                   Cin, Cout, kd, k, stride_hw);
    MVSTER_REQUIRE(((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)skip & 15) == 0,
                   "mvster_conv_tc3_pb16: pointers must be 16-byte aligned");
    const int s = stride_hw, NT = (flags & MVSTER_TC3_FP16X2) ? 2 : 1;
    CUtensorMap xm;
    int rc = encode_pb16_map(&xm, x, B * D, H, W, Cin, s, NT, "mvster_conv_tc3_pb16");
    if (rc != MVSTER_OK) return rc;
    Plan plan;
    memset(&plan, 0, sizeof(plan));
    Args a;
    memset(&a, 0, sizeof(a));
    a.nslab = build_plan(Cin, kd, k, s, &plan, nullptr);
    if (NT == 2 && Cin >= 16) widen_octet_lbo(plan, kd * (s == 2 ? 4 : 1) * ((Cin + 15) / 16));
    a.overflow = NT == 2 ? tc3::g_overflow_flag : nullptr;
    a.w = (const uint8_t*)w_packed; a.bias = bias; a.skip = (const float*)skip; a.y = (float*)y; a.scale = scale;
    a.out_pb16 = (flags & MVSTER_TC3_OUT_PB16) ? 1 : 0;
    a.D = D; a.Ho = (H - 1) / s + 1; a.Wo = (W - 1) / s + 1; a.cout = Cout; a.relu = flags & 1; a.sx = s;
    a.nstage = kd * (s == 2 ? 4 : 1) * ((Cin + 15) / 16);
    a.tiles_x = ceil_div(a.Wo, TW);
    a.tiles_per_plane = a.tiles_x * ceil_div(a.Ho, TH);
    MVSTER_REQUIRE(set_output_mode(a, 1, 0, false, 0), "mvster_conv_tc3_pb16: Cout must be a power of two");
    const long long total_tiles = (long long)a.tiles_per_plane * B * D;
    const int sms = current_sm_count();
    cudaStream_t st = (cudaStream_t)stream;
    const int NC = Cout < 16 ? 16 : Cout;
    if (NC == 16) return launch_pb<16>(xm, plan, a, total_tiles, sms, st, NT);
    if (NC == 32) return launch_pb<32>(xm, plan, a, total_tiles, sms, st, NT);
    return launch_pb<64>(xm, plan, a, total_tiles, sms, st, NT);
}

extern "C" int mvster_deconv_tc3_pb16(const void* x, const void* w_packed, const float* scale, const float* bias, const void* skip, void* y,
                                      int B, int D, int H, int W, int Cin, int Cout, int rows, int flags, mvster_stream_t stream) {
    using namespace mvster::tc3;
    MVSTER_REQUIRE(x && w_packed && y, "mvster_deconv_tc3_pb16: null pointer");
    MVSTER_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "mvster_deconv_tc3_pb16: bad shape");
    MVSTER_REQUIRE(mvster_deconv_tc3_supported(Cin, Cout, rows), "mvster_deconv_tc3_pb16: unsupported layer Cin=%d Cout=%d rows=%d", Cin, Cout, rows);
    MVSTER_REQUIRE(((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)skip & 15) == 0,
                   "mvster_deconv_tc3_pb16: pointers must be 16-byte aligned");
    const int NT = (flags & MVSTER_TC3_FP16X2) ? 2 : 1;
    CUtensorMap xm;
    int rc = encode_pb16_map(&xm, x, B * D, H, W, Cin, 1, NT, "mvster_deconv_tc3_pb16");
    if (rc != MVSTER_OK) return rc;
    Plan plan;
    memset(&plan, 0, sizeof(plan));
    const int kch = Cin / 16, ntap = deconv_ntap(rows);
    for (int kc = 0; kc < kch; ++kc) {
        Stage S;
        S.c0 = (short)(kc * 16); S.nq = 4; S.ox = -1; S.oy = -1; S.dz = 0; S.ntap = (short)ntap; S.slab0 = (short)(kc * ntap); S.pad = 0;
        plan.st[kc] = S;
        for (int t = 0; t < ntap; ++t) plan.a_desc[kc][t] = tap_desc((t / 2 + 1) * HW_ + (t % 2 + 1), NT * PLANE >> 4);  // halo (1 + dy, 1 + dx)
    }
    Args a;
    memset(&a, 0, sizeof(a));
    a.overflow = NT == 2 ? tc3::g_overflow_flag : nullptr;
    a.w = (const uint8_t*)w_packed; a.bias = bias; a.skip = (const float*)skip; a.y = (float*)y; a.scale = scale;
    a.out_pb16 = (flags & MVSTER_TC3_OUT_PB16) ? 1 : 0;
    a.D = D; a.Ho = H; a.Wo = W; a.cout = Cout; a.relu = flags & 1; a.sx = 1; a.nstage = kch;
    a.nslab = kch * ntap;
    a.tiles_x = ceil_div(W, TW);
    a.tiles_per_plane = a.tiles_x * ceil_div(H, TH);
    MVSTER_REQUIRE(set_output_mode(a, deconv_ncls(rows), rows == 1 ? 1 : 0, true, 0), "mvster_deconv_tc3_pb16: output row pitch does not fit 32 bits");
    const long long total_tiles = (long long)a.tiles_per_plane * B * D;
    const int sms = current_sm_count();
    cudaStream_t st = (cudaStream_t)stream;
    const int NC = a.ncls * Cout;
    if (NC == 16) return launch_pb<16>(xm, plan, a, total_tiles, sms, st, NT);
    if (NC == 32) return launch_pb<32>(xm, plan, a, total_tiles, sms, st, NT);
    return launch_pb<64>(xm, plan, a, total_tiles, sms, st, NT);
}
