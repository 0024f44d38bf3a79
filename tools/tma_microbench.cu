// Micro-benchmark: how fast does TMA stage the halo tiles of conv_tc3_kernel, as a function of the BOX SHAPE and of the number of
// boxes a CTA keeps in flight?  One persistent CTA per SM, one producer thread (wait slot empty -> expect_tx -> cp.async.bulk.tensor),
// one consumer warp that only waits for the bytes and hands the slot back - the "loads + hand-shakes" skeleton of the convolution
// kernel with everything else removed.  The activation tensor is P planes of H x W pixels with 16 channels (64 bytes per pixel as
// fp32 or as an fp16 pair); a tile is the 18 x 10-pixel halo patch of a 16 x 8 output tile.  Layouts:
//   0  fp32 NHWC, 4-D map {16 ch, W, H, P}, box {16, 10, 18, 1}: 180 pieces of 64 B          (what conv_tc3 does today)
//   1  fp32 NHWC, 3-D map {W*16, H, P}, box {160, 18, 1}: 18 runs of 640 B                   (MVSTER_TC3_MERGE)
//   2  16-bit octet-planar [P][4 planes][H][W][8], 4-D map {W*8, H, 4, P}, box {80, 18, 4, 1}: 72 runs of 160 B, lands in the
//      UMMA operand layout directly (no converter pass)                                      (the "pre-split operands" design)
//   3  like 2 with 2 planes (one bf16 term): 36 runs of 160 B, half the bytes
//   4  16-bit pixel-interleaved [P][H][W][4][8], 5-D map {8, W, H, 4, P}, box {8, 10, 18, 4, 1}: 720 pieces of 16 B
//   5  1-D bulk copies of 11520 contiguous bytes (upper bound: no tensor addressing at all)
//   6  layout 0 with a box of TWO horizontally adjacent tiles {16, 18, 18, 1} = 20.7 KB (is the cost per box or per byte?)
//   7  layout 0 with FOUR tiles per box {16, 34, 18, 1} = 39.2 KB
//   8  layout 0 + an L2 prefetch (cp.async.bulk.prefetch.tensor) of the box 8 tiles ahead: more bytes in flight than the ring holds
// Output: one line per (layout, depth): us for all tiles, GB/s, us per tile and SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mvster_b200/csrc -o tools/_build/tma_microbench tools/tma_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "tc_ptx.cuh"

using namespace mvster::ptx;

constexpr int MAXD = 12;
constexpr int SLOT = 11520;  // bytes of one staged tile (layouts 0, 1, 2, 4, 5, 8); layout 3: 5760; layouts 6 / 7: 2 / 4 tiles per box
constexpr int SMEM_RING = 12 * SLOT;

struct Params {
    int layout, depth, tiles_x, tiles_per_plane, total_tiles, bytes, T, slot, tiles_per_box;
    const uint8_t* flat;
};

__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d_(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    tma_load_5d(dst, map, bar, c0, c1, c2, c3, c4);
}

__global__ void __launch_bounds__(64, 1) bench(const __grid_constant__ CUtensorMap map, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    const uint32_t bar = base + SMEM_RING;
    auto FULL = [&](uint32_t s) { return bar + 8u * s; };
    auto EMPTY = [&](uint32_t s) { return bar + 8u * (MAXD + s); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.depth; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    // groups of T consecutive tiles, strided over the grid, like the convolution kernel
    const int groups = (p.total_tiles + p.T - 1) / p.T;
    if (warp == 0) {
        if (elect_one()) {
            uint32_t u = 0;
            for (int g = blockIdx.x; g < groups; g += gridDim.x)
                for (int t = 0; t < p.T; ++t, ++u) {
                    const int ti = g * p.T + t;
                    if (ti >= p.total_tiles) break;
                    const uint32_t s = u % p.depth;
                    mbar_wait(EMPTY(s), ((u / p.depth) & 1) ^ 1);
                    mbar_expect_tx(FULL(s), p.bytes);
                    const int plane = ti / p.tiles_per_plane, r = ti % p.tiles_per_plane;
                    const int y0 = (r / p.tiles_x) * 16 - 1, x0 = (r % p.tiles_x) * 8 * p.tiles_per_box - 1;
                    const uint32_t dst = base + s * p.slot;
                    if (p.layout == 8) {
                        const int tj = ti + 8 * (int)gridDim.x;  // two of this CTA's own groups (8 tiles) ahead: groups of T = 4 tiles are strided over the grid
                        if (tj < p.total_tiles) {
                            const int pl2 = tj / p.tiles_per_plane, r2 = tj % p.tiles_per_plane;
                            tma_prefetch_4d(&map, 0, (r2 % p.tiles_x) * 8 - 1, (r2 / p.tiles_x) * 16 - 1, pl2);
                        }
                    }
                    if (p.layout == 0 || p.layout >= 6) tma_load_4d(dst, &map, FULL(s), 0, x0, y0, plane);
                    else if (p.layout == 1) tma_load_3d(dst, &map, FULL(s), x0 * 16, y0, plane);
                    else if (p.layout == 2 || p.layout == 3) tma_load_4d(dst, &map, FULL(s), x0 * 8, y0, 0, plane);
                    else if (p.layout == 4) tma_load_5d_(dst, &map, FULL(s), 0, x0, y0, 0, plane);
                    else if (p.layout == 5) bulk_load(dst, p.flat + (size_t)ti * SLOT, SLOT, FULL(s));
                }
        }
    } else {
        uint32_t u = 0;
        for (int g = blockIdx.x; g < groups; g += gridDim.x)
            for (int t = 0; t < p.T; ++t, ++u) {
                if (g * p.T + t >= p.total_tiles) break;
                const uint32_t s = u % p.depth;
                mbar_wait(FULL(s), (u / p.depth) & 1);
                __syncwarp();
                if (threadIdx.x == 32) mbar_arrive(EMPTY(s));
            }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 8, H = argc > 2 ? atoi(argv[2]) : 512, W = argc > 3 ? atoi(argv[3]) : 640;
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const size_t bytes = (size_t)P * H * W * 64;
    uint8_t* x;
    cudaMalloc(&x, bytes + SLOT);
    cudaMemset(x, 1, bytes + SLOT);
    uint8_t* flush;
    const size_t flush_bytes = 512u << 20;
    cudaMalloc(&flush, flush_bytes);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int tiles_x = (W + 7) / 8, tiles_per_plane = tiles_x * ((H + 15) / 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_RING + 2048);
    printf("tensor: %d planes x %d x %d pixels x 64 B = %.1f MB; %d tiles of 18 x 10 pixels; %d SMs\n", P, H, W, bytes / 1e6, P * tiles_per_plane, sms);
    for (int layout = 0; layout <= 8; ++layout) {
        CUtensorMap map;
        memset(&map, 0, sizeof map);
        CUresult r = CUDA_SUCCESS;
        int box_bytes = SLOT, tiles_per_box = 1;
        if (layout == 6 || layout == 7) {
            tiles_per_box = layout == 6 ? 2 : 4;
            const cuuint32_t bw = 8 * tiles_per_box + 2;
            cuuint64_t dims[4] = {16, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P}, st[3] = {64, (cuuint64_t)W * 64, (cuuint64_t)H * W * 64};
            cuuint32_t box[4] = {16, bw, 18, 1}, es[4] = {1, 1, 1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            box_bytes = 64 * bw * 18;
        } else if (layout == 0 || layout == 8) {
            cuuint64_t dims[4] = {16, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P}, st[3] = {64, (cuuint64_t)W * 64, (cuuint64_t)H * W * 64};
            cuuint32_t box[4] = {16, 10, 18, 1}, es[4] = {1, 1, 1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else if (layout == 1) {
            cuuint64_t dims[3] = {(cuuint64_t)W * 16, (cuuint64_t)H, (cuuint64_t)P}, st[2] = {(cuuint64_t)W * 64, (cuuint64_t)H * W * 64};
            cuuint32_t box[3] = {160, 18, 1}, es[3] = {1, 1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else if (layout == 2 || layout == 3) {
            const int J = layout == 2 ? 4 : 2;
            cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)J, (cuuint64_t)P};
            cuuint64_t st[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * 16 * J};
            cuuint32_t box[4] = {80, 18, (cuuint32_t)J, 1}, es[4] = {1, 1, 1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, x, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            box_bytes = 80 * 2 * 18 * J;
        } else if (layout == 4) {
            cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, 4, (cuuint64_t)P};
            cuuint64_t st[4] = {64, (cuuint64_t)W * 64, 16, (cuuint64_t)H * W * 64};
            cuuint32_t box[5] = {8, 10, 18, 4, 1}, es[5] = {1, 1, 1, 1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, x, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("layout %d: tensor map rejected (CUresult %d)\n", layout, (int)r); continue; }
        for (int depth : {2, 4, 6, 8, 12}) {
            Params p;
            p.layout = layout; p.depth = depth; p.tiles_x = tiles_x; p.tiles_per_plane = tiles_per_plane;
            const int txb = (tiles_x + tiles_per_box - 1) / tiles_per_box;  // boxes per tile row
            p.tiles_x = txb; p.tiles_per_plane = txb * ((H + 15) / 16); p.tiles_per_box = tiles_per_box;
            p.slot = (box_bytes + 127) / 128 * 128;
            if (depth * p.slot > SMEM_RING) continue;
            p.total_tiles = P * p.tiles_per_plane; p.bytes = box_bytes; p.T = tiles_per_box == 1 ? 4 : 1; p.flat = x;
            if (layout == 5 && (size_t)p.total_tiles * SLOT > bytes) p.total_tiles = (int)(bytes / SLOT);
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaMemsetAsync(flush, rep, flush_bytes);
                cudaEvent_t a, b;
                cudaEventCreate(&a); cudaEventCreate(&b);
                cudaEventRecord(a);
                bench<<<sms, 64, SMEM_RING + 2048>>>(map, p);
                cudaEventRecord(b);
                cudaError_t e = cudaEventSynchronize(b);
                if (e != cudaSuccess) { printf("layout %d depth %d: %s\n", layout, depth, cudaGetErrorString(e)); return 1; }
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (rep && ms < best) best = ms;
                cudaEventDestroy(a); cudaEventDestroy(b);
            }
            const double moved = (double)p.total_tiles * box_bytes;
            printf("layout %d depth %2d: %8.1f us  %7.1f GB/s (box bytes)  %.3f us per box and SM (%d tile(s) per box)\n", layout, depth, best * 1e3,
                   moved / (best * 1e-3) / 1e9, best * 1e3 / ((double)p.total_tiles / sms), tiles_per_box);
        }
    }
    return 0;
}
