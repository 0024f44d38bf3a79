// Implicit-GEMM 3-D convolution on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   y[m, n] = relu( sum_tap sum_c  x[pixel(m) + tap, c] * w[tap][c][n] + bias[n] ) (+ skip[m, n])
//
// for the stride-1 (kd,3,3) layers of the regularisation U-Nets (reg2d conv2/conv4/conv6:
// 69 % of reg2d's FLOPs, mvs4net_utils.py:877-881; ConvBnReLU3D :116-123 with BN folded).
//
// Mapping (one CTA, 192 threads, no cluster):
//   * M tile  = 128 output voxels = an 8 x 16 pixel patch of one (batch, depth) plane.  The A operand
//     of tap (kz,ky,kx) is ONE TMA box load [8][16][KC] of the channels-last activation tensor at the
//     shifted origin; out-of-bounds rows/columns/planes are zero-filled by TMA, which is exactly the
//     convolution's zero padding (incl. the depth axis).  The box lands K-major with 128-byte (KC=32)
//     or 64-byte (KC=16) hardware swizzle - the layout tcgen05.mma consumes directly.
//   * N = Cout (16/32/64), K per tap = Cin, walked in chunks of KC <= 32 channels, 8 (tf32) per MMA.
//   * B operand = the tap's [Cout][KC] weight slab, K-major, streamed by TMA once per (tap, chunk) and
//     reused by up to T = 4 M-tiles whose fp32 accumulators live side by side in TMEM (T*Cout columns).
//   * fp32 fidelity: "3xTF32" - A and W are split hi + lo (hi = top 19 bits, lo = remainder), and
//     D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (fp32 accumulate in TMEM); the dropped A_lo*W_lo term is
//     ~2^-22 relative.  W is pre-split on the host; A is split in shared memory by the 4 epilogue
//     warps between the TMA arrival and the MMA (generic-proxy writes + fence.proxy.async).
//     NPASS = 1 skips the split (plain TF32) for reduced-precision configurations.
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//     warps 2-5 = A splitter, then epilogue (tcgen05.ld -> bias/ReLU/skip -> 128-bit global stores).
//   * pipelines: a_full (TMA->split), a_ready (split->MMA), a_empty (tcgen05.commit->TMA),
//     b_full / b_empty, acc_full (last commit -> epilogue).
#include "common.cuh"
#include <cuda.h>

namespace mvster {

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a converged warp; see tc_ptx.cuh for why single-thread tcgen05/TMA issue is guarded by this
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && (spin & 63u) == 63u) {  // watchdog: a protocol bug must fail loudly within ~2 s, never hang the GPU
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// K-major, hardware-swizzled shared-memory matrix descriptor (sm_100 "version 1"):
// start address, LBO (unused for swizzled K-major, canonical value 1), SBO = 8 rows * row bytes, layout type.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout_type << 61);
}

struct TcArgs {
    const float* bias; const float* skip; float* y;
    int B, D, H, W, kd, relu, tiles_x, tiles_y, n_tiles, T;
};

constexpr int TC_THREADS = 192;
constexpr int TILE_W = 16, TILE_H = 8;

template <int NC, int KC, int KCHUNKS, int NPASS>
struct TcCfg {
    static constexpr int SA = 3, SB = 2, TMAX = 4;
    static constexpr int A_BYTES = 128 * KC * 4, B_BYTES = NC * KC * 4;
    static constexpr int A_STAGE = A_BYTES * (NPASS == 3 ? 2 : 1), B_STAGE = B_BYTES * (NPASS == 3 ? 2 : 1);
    static constexpr int SMEM = 1024 + SA * A_STAGE + SB * B_STAGE + 256;
    static constexpr int TCOLS = (TMAX * NC) < 32 ? 32 : TMAX * NC;  // 64 / 128 / 256: powers of two
    static constexpr uint32_t LAYOUT = KC == 32 ? 2u : 4u;           // SWIZZLE_128B : SWIZZLE_64B
    static constexpr uint32_t SBO = 8 * KC * 4;
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NC >> 3) << 17) | ((128u >> 4) << 24);
};

template <int NC, int KC, int KCHUNKS, int NPASS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap w_map, const TcArgs a) {
    using Cfg = TcCfg<NC, KC, KCHUNKS, NPASS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = base, b_base = base + Cfg::SA * Cfg::A_STAGE;
    const uint32_t bar_base = b_base + Cfg::SB * Cfg::B_STAGE;  // 8-byte barriers
    auto A_FULL = [&](int s) { return bar_base + 8u * s; };
    auto A_READY = [&](int s) { return bar_base + 8u * (Cfg::SA + s); };
    auto A_EMPTY = [&](int s) { return bar_base + 8u * (2 * Cfg::SA + s); };
    auto B_FULL = [&](int s) { return bar_base + 8u * (3 * Cfg::SA + s); };
    auto B_EMPTY = [&](int s) { return bar_base + 8u * (3 * Cfg::SA + Cfg::SB + s); };
    const uint32_t ACC_FULL = bar_base + 8u * (3 * Cfg::SA + 2 * Cfg::SB);
    const uint32_t tmem_slot = ACC_FULL + 8u;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile0 = blockIdx.x * a.T;
    const int T = min(a.T, a.n_tiles - tile0);
    const int taps = a.kd * 9;
    const int n_iter = taps * KCHUNKS * T;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::SA; ++s) { mbar_init(A_FULL(s), 1); mbar_init(A_READY(s), 128); mbar_init(A_EMPTY(s), 1); }
        for (int s = 0; s < Cfg::SB; ++s) { mbar_init(B_FULL(s), 1); mbar_init(B_EMPTY(s), 1); }
        mbar_init(ACC_FULL, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    auto tile_coords = [&](int t, int& b, int& z, int& y0, int& x0) {
        int ti = tile0 + t;
        const int tx = ti % a.tiles_x; ti /= a.tiles_x;
        const int ty = ti % a.tiles_y; ti /= a.tiles_y;
        z = ti % a.D; b = ti / a.D;
        y0 = ty * TILE_H; x0 = tx * TILE_W;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int it = 0, bi = 0;
            for (int tap = 0; tap < taps; ++tap) {
                const int kz = tap / 9 - a.kd / 2, ky = (tap % 9) / 3 - 1, kx = tap % 3 - 1;
                for (int kc = 0; kc < KCHUNKS; ++kc, ++bi) {
                    const int sb = bi % Cfg::SB;
                    mbar_wait(B_EMPTY(sb), ((bi / Cfg::SB) & 1) ^ 1);
                    mbar_expect_tx(B_FULL(sb), Cfg::B_STAGE);
                    const uint32_t bdst = b_base + sb * Cfg::B_STAGE;
                    const int wrow = (tap * KCHUNKS + kc) * NC;
                    tma_load_2d(bdst, &w_map, B_FULL(sb), 0, wrow);
                    if (NPASS == 3) tma_load_2d(bdst + Cfg::B_BYTES, &w_map, B_FULL(sb), 0, taps * KCHUNKS * NC + wrow);
                    for (int t = 0; t < T; ++t, ++it) {
                        const int s = it % Cfg::SA;
                        mbar_wait(A_EMPTY(s), ((it / Cfg::SA) & 1) ^ 1);
                        mbar_expect_tx(A_FULL(s), Cfg::A_BYTES);
                        int b, z, y0, x0;
                        tile_coords(t, b, z, y0, x0);
                        tma_load_5d(a_base + s * Cfg::A_STAGE, &x_map, A_FULL(s), kc * KC, x0 + kx, y0 + ky, z + kz, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        int it = 0, bi = 0;
        for (int tap = 0; tap < taps; ++tap) {
            for (int kc = 0; kc < KCHUNKS; ++kc, ++bi) {
                const int sb = bi % Cfg::SB;
                mbar_wait(B_FULL(sb), (bi / Cfg::SB) & 1);
                const uint32_t b_hi = b_base + sb * Cfg::B_STAGE, b_lo = b_hi + Cfg::B_BYTES;
                for (int t = 0; t < T; ++t, ++it) {
                    const int s = it % Cfg::SA;
                    mbar_wait(NPASS == 3 ? A_READY(s) : A_FULL(s), (it / Cfg::SA) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t a_hi = a_base + s * Cfg::A_STAGE, a_lo = a_hi + Cfg::A_BYTES;
                        const uint32_t d = tmem_base + (uint32_t)(t * NC);
                        const bool first = (tap == 0 && kc == 0);
#pragma unroll
                        for (int pass = 0; pass < NPASS; ++pass) {
                            const uint32_t aa = (pass == 1) ? a_lo : a_hi, bb = (pass == 2) ? b_lo : b_hi;
#pragma unroll
                            for (int k = 0; k < KC / 8; ++k) {
                                umma_tf32(d, smem_desc(aa + k * 32, Cfg::SBO, Cfg::LAYOUT), smem_desc(bb + k * 32, Cfg::SBO, Cfg::LAYOUT),
                                          Cfg::IDESC, (first && pass == 0 && k == 0) ? 0u : 1u);
                            }
                        }
                        umma_commit(A_EMPTY(s));
                        if (t == T - 1) umma_commit(B_EMPTY(sb));
                    }
                    __syncwarp();
                }
            }
        }
        if (elect_one()) umma_commit(ACC_FULL);
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ A splitter (3xTF32 only)
        const int tid = threadIdx.x - 64;  // 0..127
        if (NPASS == 3) {
            for (int it = 0; it < n_iter; ++it) {
                const int s = it % Cfg::SA;
                mbar_wait(A_FULL(s), (it / Cfg::SA) & 1);
                uint8_t* hi_p = smem_raw + (a_base + s * Cfg::A_STAGE - smem_u32(smem_raw));
                uint4* hi4 = reinterpret_cast<uint4*>(hi_p);
                uint4* lo4 = reinterpret_cast<uint4*>(hi_p + Cfg::A_BYTES);
#pragma unroll
                for (int i = 0; i < Cfg::A_BYTES / 16 / 128; ++i) {
                    uint4 v = hi4[i * 128 + tid];
                    uint4 h, l;
                    h.x = v.x & 0xFFFFE000u; h.y = v.y & 0xFFFFE000u; h.z = v.z & 0xFFFFE000u; h.w = v.w & 0xFFFFE000u;
                    l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)) & 0xFFFFE000u;
                    l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)) & 0xFFFFE000u;
                    l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)) & 0xFFFFE000u;
                    l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)) & 0xFFFFE000u;
                    hi4[i * 128 + tid] = h;
                    lo4[i * 128 + tid] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the MMA (async proxy)
                mbar_arrive(A_READY(s));
            }
        }
        // ------------------------------------------------------------------ epilogue
        mbar_wait(ACC_FULL, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;          // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;     // accumulator row = pixel within the tile
        for (int t = 0; t < T; ++t) {
            int b, z, y0, x0;
            tile_coords(t, b, z, y0, x0);
            const int yy = y0 + r / TILE_W, xx = x0 + r % TILE_W;
            const bool ok = yy < a.H && xx < a.W;
            const long long off = ((((long long)b * a.D + z) * a.H + yy) * a.W + xx) * NC;
#pragma unroll
            for (int c0 = 0; c0 < NC; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NC + c0), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 o;
                        o.x = __uint_as_float(v[j]) + (a.bias ? __ldg(a.bias + c0 + j) : 0.f);
                        o.y = __uint_as_float(v[j + 1]) + (a.bias ? __ldg(a.bias + c0 + j + 1) : 0.f);
                        o.z = __uint_as_float(v[j + 2]) + (a.bias ? __ldg(a.bias + c0 + j + 2) : 0.f);
                        o.w = __uint_as_float(v[j + 3]) + (a.bias ? __ldg(a.bias + c0 + j + 3) : 0.f);
                        if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        if (a.skip) {
                            const float4 sk = __ldg(reinterpret_cast<const float4*>(a.skip + off + c0 + j));
                            o.x += sk.x; o.y += sk.y; o.z += sk.z; o.w += sk.w;
                        }
                        *reinterpret_cast<float4*>(a.y + off + c0 + j) = o;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TCOLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------ host side
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

template <int NC, int KC, int KCHUNKS, int NPASS>
static int launch_tc(const CUtensorMap& xm, const CUtensorMap& wm, const TcArgs& a, cudaStream_t st) {
    using Cfg = TcCfg<NC, KC, KCHUNKS, NPASS>;
    auto k = conv_tc_kernel<NC, KC, KCHUNKS, NPASS>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) {
        set_error("conv_tc_kernel: cannot reserve %d bytes of shared memory", Cfg::SMEM);
        cudaGetLastError();
        return MVSTER_ERR_CUDA;
    }
    k<<<ceil_div(a.n_tiles, a.T), TC_THREADS, Cfg::SMEM, st>>>(xm, wm, a);
    return check_launch("conv_tc_kernel");
}

template <int NC, int KC, int KCHUNKS>
static int dispatch_pass(const CUtensorMap& xm, const CUtensorMap& wm, const TcArgs& a, int npass, cudaStream_t st) {
    if (npass == 3) return launch_tc<NC, KC, KCHUNKS, 3>(xm, wm, a, st);
    return launch_tc<NC, KC, KCHUNKS, 1>(xm, wm, a, st);
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_conv3d_tc_supported(int Cin, int Cout, int kd, int stride_hw, int transposed) {
    const bool cin_ok = Cin == 16 || Cin == 32 || Cin == 64, cout_ok = Cout == 16 || Cout == 32 || Cout == 64;
    return cin_ok && cout_ok && (kd == 1 || kd == 3) && stride_hw == 1 && !transposed;
}

extern "C" int mvster_conv3d_tc_f32(const float* x, const float* w_packed, const float* bias, const float* skip, float* y,
                                    int B, int D, int H, int W, int Cin, int Cout, int kd, int relu, int npass,
                                    mvster_stream_t stream) {
    MVSTER_REQUIRE(x && w_packed && y, "mvster_conv3d_tc_f32: null pointer");
    MVSTER_REQUIRE(mvster_conv3d_tc_supported(Cin, Cout, kd, 1, 0), "mvster_conv3d_tc_f32: unsupported layer Cin=%d Cout=%d kd=%d", Cin, Cout, kd);
    MVSTER_REQUIRE(npass == 1 || npass == 3, "mvster_conv3d_tc_f32: npass must be 1 (tf32) or 3 (3xtf32)");
    MVSTER_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "mvster_conv3d_tc_f32: bad shape");
    EncodeTiledFn enc = encode_fn();
    MVSTER_REQUIRE(enc, "mvster_conv3d_tc_f32: cuTensorMapEncodeTiled is unavailable in this driver");
    const int KC = Cin >= 32 ? 32 : 16, KCH = Cin / KC, taps = kd * 9;
    const CUtensorMapSwizzle swz = KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

    CUtensorMap xm, wm;
    {
        cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
        cuuint64_t strides[4] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4, (cuuint64_t)D * H * W * Cin * 4};
        cuuint32_t box[5] = {(cuuint32_t)KC, TILE_W, TILE_H, 1, 1}, es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVSTER_REQUIRE(r == CUDA_SUCCESS, "mvster_conv3d_tc_f32: activation tensor map rejected (CUresult %d)", (int)r);
    }
    {
        const int rows = taps * KCH * Cout * (npass == 3 ? 2 : 1);
        cuuint64_t dims[2] = {(cuuint64_t)KC, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)KC * 4};
        cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)Cout}, es[2] = {1, 1};
        CUresult r = enc(&wm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w_packed, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVSTER_REQUIRE(r == CUDA_SUCCESS, "mvster_conv3d_tc_f32: weight tensor map rejected (CUresult %d)", (int)r);
    }
    TcArgs a;
    a.bias = bias; a.skip = skip; a.y = y;
    a.B = B; a.D = D; a.H = H; a.W = W; a.kd = kd; a.relu = relu;
    a.tiles_x = ceil_div(W, TILE_W); a.tiles_y = ceil_div(H, TILE_H);
    a.n_tiles = a.tiles_x * a.tiles_y * D * B;
    a.T = a.n_tiles >= 4 * 296 ? 4 : (a.n_tiles >= 2 * 296 ? 2 : 1);  // keep >= 2 CTAs per SM worth of groups when possible
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 16) {
        if (Cout == 16) return dispatch_pass<16, 16, 1>(xm, wm, a, npass, st);
        if (Cout == 32) return dispatch_pass<32, 16, 1>(xm, wm, a, npass, st);
        return dispatch_pass<64, 16, 1>(xm, wm, a, npass, st);
    }
    if (Cin == 32) {
        if (Cout == 16) return dispatch_pass<16, 32, 1>(xm, wm, a, npass, st);
        if (Cout == 32) return dispatch_pass<32, 32, 1>(xm, wm, a, npass, st);
        return dispatch_pass<64, 32, 1>(xm, wm, a, npass, st);
    }
    if (Cout == 16) return dispatch_pass<16, 32, 2>(xm, wm, a, npass, st);
    if (Cout == 32) return dispatch_pass<32, 32, 2>(xm, wm, a, npass, st);
    return dispatch_pass<64, 32, 2>(xm, wm, a, npass, st);
}
