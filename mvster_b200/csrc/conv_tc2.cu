// Implicit-GEMM convolution on tcgen05, generation 2: each activation tile is staged ONCE per depth plane
// (halo included) and the 9 in-plane taps are addressed by shifting the UMMA descriptor, instead of
// re-loading a shifted copy of the tile per tap (conv_tc.cu, 27 loads per tile, L2->SMEM fill bound).
//
//   * Output tile = 16 x 8 pixels (M = 128) of one (batch, depth) plane.  Its input halo (18 x 10 pixels)
//     for 16 channels is four 4-D TMA box loads (one per channel quad, box = [18][10][4 floats]): it lands in
//     shared memory as [4 channel quads][18][10][4 floats], i.e. pixels 16 bytes apart along x - exactly the
//     un-swizzled K-major "core matrix" layout of tcgen05 (8 rows x 16 B contiguous).  Out-of-image halo
//     pixels are zero-filled by TMA = the convolution's zero padding; out-of-range depth planes are skipped.
//   * A operand of tap (ky,kx), K-step k: descriptor start = tile + ((ky*10 + kx)*16 B) + 2k quad planes,
//     LBO = one quad plane (2880 B), SBO = one halo row (160 B).  No data movement per tap.
//   * B operand: the tap's [Cout][16] weight slab (64-byte swizzle, K-major), streamed by TMA through a 4-deep
//     ring and shared by the T <= 4 tiles a CTA accumulates side by side in TMEM (T*N fp32 columns).
//   * 3xTF32 for fp32 fidelity: A split hi/lo in shared memory by the epilogue warps once per staged tile
//     (not once per tap), W pre-split on the host: D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo.
//   * Stage = (depth tap kz, 16-channel chunk): T staged tiles, 9 taps x 2 K-steps x NPASS MMAs per tile.
//   * warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = splitter then epilogue.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace mvster {
namespace tc2 {
using namespace ptx;

constexpr int TW = 8, TH = 16, HW_ = TW + 2, HH_ = TH + 2;
constexpr int QBYTES = HH_ * HW_ * 16;       // bytes TMA writes per channel quad of the halo tile (2880)
constexpr int QPLANE = 2944;                 // quad-plane pitch in shared memory: 2880 rounded up to 128 B (TMA destination alignment)
constexpr int A_HI = 4 * QPLANE;             // 16 channels of one tile (11776 B incl. padding)
constexpr int SA = 2;
constexpr int THREADS = 192;

struct Args {
    const float* bias; const float* skip; float* y;
    int B, D, H, W, cin, cout, kd, ntap, relu;  // ntap = in-plane taps: 9 (3x3) or 1 (1x1, the centre tap)
    int tiles_x, tiles_per_plane, groups_per_plane, T;
};

template <int NC, int NPASS>
struct Cfg {
    // tiles accumulated side by side in TMEM, and depth of the per-tap weight ring (> 9 = more than one full
    // stage of taps in flight: the weight stream never waits for the MMAs of the current stage)
    static constexpr int TMAX = NC >= 64 ? 2 : 4;
    static constexpr int SB = NC == 80 ? 8 : (NC == 64 ? 12 : 10);
    static constexpr int A_TILE = A_HI * (NPASS == 3 ? 2 : 1);
    static constexpr int A_STAGE = TMAX * A_TILE;
    static constexpr int B_HALF = NC * 64;                    // [NC][16] fp32
    static constexpr int B_STAGE = B_HALF * (NPASS == 3 ? 2 : 1);
    static constexpr int SMEM = 1024 + SA * A_STAGE + SB * B_STAGE + 256;
    static constexpr int TCOLS = TMAX * NC <= 64 ? 64 : (TMAX * NC <= 128 ? 128 : 256);  // power of two >= TMAX*NC
};

template <int NC, int NPASS>
__global__ void __launch_bounds__(THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap w_map, const Args a) {
    using C = Cfg<NC, NPASS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t a_base = base, b_base = base + SA * C::A_STAGE, bar_base = b_base + C::SB * C::B_STAGE;
    auto A_FULL = [&](int s) { return bar_base + 8u * s; };
    auto A_READY = [&](int s) { return bar_base + 8u * (SA + s); };
    auto A_EMPTY = [&](int s) { return bar_base + 8u * (2 * SA + s); };
    auto B_FULL = [&](int s) { return bar_base + 8u * (3 * SA + s); };
    auto B_EMPTY = [&](int s) { return bar_base + 8u * (3 * SA + C::SB + s); };
    const uint32_t ACC_FULL = bar_base + 8u * (3 * SA + 2 * C::SB);
    const uint32_t tmem_slot = ACC_FULL + 8u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int plane = blockIdx.x / a.groups_per_plane, grp = blockIdx.x % a.groups_per_plane;  // plane = b*D + z
    const int z = plane % a.D;
    const int tile0 = grp * a.T;
    const int T = min(a.T, a.tiles_per_plane - tile0);
    const int kch = a.cin / 16, pz = a.kd / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(A_FULL(s), 1); mbar_init(A_READY(s), 128); mbar_init(A_EMPTY(s), 1); }
        for (int s = 0; s < C::SB; ++s) { mbar_init(B_FULL(s), 1); mbar_init(B_EMPTY(s), 1); }
        mbar_init(ACC_FULL, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

    auto tile_xy = [&](int t, int& y0, int& x0) {
        const int ti = tile0 + t;
        y0 = (ti / a.tiles_x) * TH;
        x0 = (ti % a.tiles_x) * TW;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int sa_it = 0, sb_it = 0;
            for (int kz = 0; kz < a.kd; ++kz) {
                const int zz = z + kz - pz;
                if (zz < 0 || zz >= a.D) continue;  // depth padding: the plane contributes nothing
                for (int kc = 0; kc < kch; ++kc, ++sa_it) {
                    const int s = sa_it % SA;
                    mbar_wait(A_EMPTY(s), ((sa_it / SA) & 1) ^ 1);
                    mbar_expect_tx(A_FULL(s), T * 4 * QBYTES);
                    for (int t = 0; t < T; ++t) {
                        int y0, x0;
                        tile_xy(t, y0, x0);
                        const uint32_t dst = a_base + s * C::A_STAGE + t * C::A_TILE;
#pragma unroll
                        for (int cq = 0; cq < 4; ++cq)  // one 4-channel box per quad plane: [18][10][4 floats]
                            tma_load_4d(dst + cq * QPLANE, &x_map, A_FULL(s), kc * 16 + cq * 4, x0 - 1, y0 - 1, plane + kz - pz);
                    }
                    for (int tap = 0; tap < a.ntap; ++tap, ++sb_it) {
                        const int sb = sb_it % C::SB;
                        mbar_wait(B_EMPTY(sb), ((sb_it / C::SB) & 1) ^ 1);
                        mbar_expect_tx(B_FULL(sb), C::B_STAGE);
                        const int wrow = ((kz * a.ntap + tap) * kch + kc) * NC;
                        tma_load_2d(b_base + sb * C::B_STAGE, &w_map, B_FULL(sb), 0, wrow);
                        if (NPASS == 3) tma_load_2d(b_base + sb * C::B_STAGE + C::B_HALF, &w_map, B_FULL(sb), 0, a.kd * a.ntap * kch * NC + wrow);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t IDESC = idesc_tf32_m128(NC);
        int sa_it = 0, sb_it = 0;
        bool fresh = true;
        for (int kz = 0; kz < a.kd; ++kz) {
            const int zz = z + kz - pz;
            if (zz < 0 || zz >= a.D) continue;
            for (int kc = 0; kc < kch; ++kc, ++sa_it) {
                const int s = sa_it % SA;
                mbar_wait(NPASS == 3 ? A_READY(s) : A_FULL(s), (sa_it / SA) & 1);
                for (int tap = 0; tap < a.ntap; ++tap, ++sb_it) {
                    const int sb = sb_it % C::SB;
                    mbar_wait(B_FULL(sb), (sb_it / C::SB) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        // Descriptors differ only in their 16-byte-granular start-address field (low 14 bits, no carry
                        // out: shared memory < 256 KB), so every operand is the stage descriptor plus a small constant:
                        // the single issuing thread spends ~2 integer adds per MMA instead of rebuilding descriptors.
                        const uint64_t a0 = smem_desc(a_base + s * C::A_STAGE, QPLANE, HW_ * 16, 0) +
                                            (uint64_t)(a.ntap == 1 ? HW_ + 1 : (tap / 3) * HW_ + tap % 3);  // (ky*10 + kx) * 16 B
                        const uint64_t b0 = smem_desc(b_base + sb * C::B_STAGE, 16, 512, 4);
                        const uint32_t first = (fresh && tap == 0) ? 0u : 1u;
                        for (int t = 0; t < T; ++t) {
                            const uint64_t at = a0 + (uint64_t)(t * (C::A_TILE >> 4));
                            const uint32_t d = tmem_base + (uint32_t)(t * NC);
                            umma_tf32(d, at, b0, IDESC, first);                                   // A_hi * W_hi, channels 0-7
                            umma_tf32(d, at + (2 * QPLANE >> 4), b0 + 2, IDESC, 1u);              //              channels 8-15
                            if (NPASS == 3) {
                                umma_tf32(d, at + (A_HI >> 4), b0, IDESC, 1u);                    // A_lo * W_hi
                                umma_tf32(d, at + ((A_HI + 2 * QPLANE) >> 4), b0 + 2, IDESC, 1u);
                                umma_tf32(d, at, b0 + (C::B_HALF >> 4), IDESC, 1u);               // A_hi * W_lo
                                umma_tf32(d, at + (2 * QPLANE >> 4), b0 + (C::B_HALF >> 4) + 2, IDESC, 1u);
                            }
                        }
                        umma_commit(B_EMPTY(sb));
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(A_EMPTY(s));
                __syncwarp();
                fresh = false;
            }
        }
        if (elect_one()) umma_commit(ACC_FULL);
        __syncwarp();
    } else {
        const int tid = threadIdx.x - 64;  // 0..127
        // ------------------------------------------------------------------ A splitter (3xTF32): once per staged tile
        if (NPASS == 3) {
            int sa_it = 0;
            for (int kz = 0; kz < a.kd; ++kz) {
                const int zz = z + kz - pz;
                if (zz < 0 || zz >= a.D) continue;
                for (int kc = 0; kc < kch; ++kc, ++sa_it) {
                    const int s = sa_it % SA;
                    mbar_wait(A_FULL(s), (sa_it / SA) & 1);
                    for (int t = 0; t < T; ++t) {
                        uint8_t* p = smem_raw + (a_base + s * C::A_STAGE + t * C::A_TILE - raw);
                        uint4* hi4 = reinterpret_cast<uint4*>(p);
                        uint4* lo4 = reinterpret_cast<uint4*>(p + A_HI);
                        for (int i = tid; i < A_HI / 16; i += 128) {
                            const uint4 v = hi4[i];
                            uint4 h, l;
                            h.x = v.x & 0xFFFFE000u; h.y = v.y & 0xFFFFE000u; h.z = v.z & 0xFFFFE000u; h.w = v.w & 0xFFFFE000u;
                            l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x)) & 0xFFFFE000u;
                            l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y)) & 0xFFFFE000u;
                            l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z)) & 0xFFFFE000u;
                            l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w)) & 0xFFFFE000u;
                            hi4[i] = h;
                            lo4[i] = l;
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive(A_READY(s));
                }
            }
        }
        // ------------------------------------------------------------------ epilogue
        mbar_wait(ACC_FULL, 0);
        tc_fence_after();
        const int q = warp & 3, r = q * 32 + lane;  // TMEM lane quarter, accumulator row = pixel in the tile
        for (int t = 0; t < T; ++t) {
            int y0, x0;
            tile_xy(t, y0, x0);
            const int yy = y0 + r / TW, xx = x0 + r % TW;
            const bool ok = yy < a.H && xx < a.W;
            const long long off = (((long long)plane * a.H + yy) * a.W + xx) * a.cout;
#pragma unroll
            for (int c0 = 0; c0 < NC; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NC + c0), v);
                tmem_ld_wait();
                if (ok && c0 < a.cout) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        if (c0 + j >= a.cout) break;
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
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TCOLS);
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

template <int NC, int NPASS>
static int launch(const CUtensorMap& xm, const CUtensorMap& wm, const Args& a, int grid, cudaStream_t st) {
    using C = Cfg<NC, NPASS>;
    auto k = conv_tc2_kernel<NC, NPASS>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) {
        set_error("conv_tc2_kernel: cannot reserve %d bytes of shared memory", C::SMEM);
        cudaGetLastError();
        return MVSTER_ERR_CUDA;
    }
    k<<<grid, THREADS, C::SMEM, st>>>(xm, wm, a);
    return check_launch("conv_tc2_kernel");
}

}  // namespace tc2
}  // namespace mvster

using namespace mvster;

extern "C" int mvster_conv3d_tc2_supported(int Cin, int Cout, int kd, int stride_hw, int transposed) {
    const bool cin_ok = Cin == 16 || Cin == 32 || Cin == 64, cout_ok = Cout == 8 || Cout == 16 || Cout == 32 || Cout == 64;
    return cin_ok && cout_ok && (kd == 1 || kd == 3) && stride_hw == 1 && !transposed;
}

static int run_tc2(const float* x, const float* w_packed, const float* bias, const float* skip, float* y,
                   int B, int D, int H, int W, int Cin, int Cout, int kd, int ntap, int relu, int npass, mvster_stream_t stream);

extern "C" int mvster_conv3d_tc2_f32(const float* x, const float* w_packed, const float* bias, const float* skip, float* y,
                                     int B, int D, int H, int W, int Cin, int Cout, int kd, int relu, int npass,
                                     mvster_stream_t stream) {
    MVSTER_REQUIRE(mvster_conv3d_tc2_supported(Cin, Cout, kd, 1, 0), "mvster_conv3d_tc2_f32: unsupported layer Cin=%d Cout=%d kd=%d", Cin, Cout, kd);
    return run_tc2(x, w_packed, bias, skip, y, B, D, H, W, Cin, Cout, kd, 9, relu, npass, stream);
}

/* Point-wise (1x1) convolution = plain GEMM [pixels x Cin] x [Cin x Cout] on the same kernel (centre tap only).
 * Cout in {8,16,32,64} or 65..80 (padded to N = 80).  w_packed: packing.pack_tc2_weights of a [1][Cin][Cout] tensor. */
extern "C" int mvster_pointwise_tc2_f32(const float* x, const float* w_packed, const float* bias, float* y,
                                        int N, int H, int W, int Cin, int Cout, int relu, int npass, mvster_stream_t stream) {
    MVSTER_REQUIRE((Cin == 16 || Cin == 32 || Cin == 64) && (Cout == 8 || Cout == 16 || Cout == 32 || Cout == 64 || (Cout > 64 && Cout <= 80 && Cout % 4 == 0)),
                   "mvster_pointwise_tc2_f32: unsupported Cin=%d Cout=%d", Cin, Cout);
    return run_tc2(x, w_packed, bias, nullptr, y, N, 1, H, W, Cin, Cout, 1, 1, relu, npass, stream);
}

static int run_tc2(const float* x, const float* w_packed, const float* bias, const float* skip, float* y,
                   int B, int D, int H, int W, int Cin, int Cout, int kd, int ntap, int relu, int npass, mvster_stream_t stream) {
    using namespace mvster::tc2;
    MVSTER_REQUIRE(x && w_packed && y, "mvster_conv3d_tc2_f32: null pointer");
    MVSTER_REQUIRE(npass == 1 || npass == 3, "mvster_conv3d_tc2_f32: npass must be 1 (tf32) or 3 (3xtf32)");
    MVSTER_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "mvster_conv3d_tc2_f32: bad shape");
    EncodeTiledFn enc = encode_fn();
    MVSTER_REQUIRE(enc, "mvster_conv3d_tc2_f32: cuTensorMapEncodeTiled is unavailable in this driver");
    const int NC = Cout < 16 ? 16 : (Cout > 64 ? 80 : Cout), kch = Cin / 16, taps = kd * ntap;

    CUtensorMap xm, wm;
    {   // activations [B*D][H][W][C]; a box is one channel QUAD (16 B) of an 18 x 10 pixel halo patch
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * D};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4};
        cuuint32_t box[4] = {4, HW_, HH_, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVSTER_REQUIRE(r == CUDA_SUCCESS, "mvster_conv3d_tc2_f32: activation tensor map rejected (CUresult %d)", (int)r);
    }
    {
        const int rows = taps * kch * NC * (npass == 3 ? 2 : 1);
        cuuint64_t dims[2] = {16, (cuuint64_t)rows};
        cuuint64_t strides[1] = {64};
        cuuint32_t box[2] = {16, (cuuint32_t)NC}, es[2] = {1, 1};
        CUresult r = enc(&wm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w_packed, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MVSTER_REQUIRE(r == CUDA_SUCCESS, "mvster_conv3d_tc2_f32: weight tensor map rejected (CUresult %d)", (int)r);
    }
    Args a;
    a.bias = bias; a.skip = skip; a.y = y;
    a.B = B; a.D = D; a.H = H; a.W = W; a.cin = Cin; a.cout = Cout; a.kd = kd; a.ntap = ntap; a.relu = relu;
    a.tiles_x = ceil_div(W, TW);
    a.tiles_per_plane = a.tiles_x * ceil_div(H, TH);
    const long long total_tiles = (long long)a.tiles_per_plane * B * D;
    const int tmax = NC >= 64 ? 2 : 4;  // Cfg<NC, *>::TMAX
    a.T = total_tiles >= 4 * 296 ? 4 : (total_tiles >= 2 * 296 ? 2 : 1);
    a.T = a.T > tmax ? tmax : a.T;
    a.T = a.T > a.tiles_per_plane ? a.tiles_per_plane : a.T;
    a.groups_per_plane = ceil_div(a.tiles_per_plane, a.T);
    const int grid = a.groups_per_plane * B * D;
    cudaStream_t st = (cudaStream_t)stream;
    if (NC == 16) return npass == 3 ? launch<16, 3>(xm, wm, a, grid, st) : launch<16, 1>(xm, wm, a, grid, st);
    if (NC == 32) return npass == 3 ? launch<32, 3>(xm, wm, a, grid, st) : launch<32, 1>(xm, wm, a, grid, st);
    if (NC == 80) return npass == 3 ? launch<80, 3>(xm, wm, a, grid, st) : launch<80, 1>(xm, wm, a, grid, st);
    return npass == 3 ? launch<64, 3>(xm, wm, a, grid, st) : launch<64, 1>(xm, wm, a, grid, st);
}
