// Micro-benchmark: cycles per tcgen05.mma (M = 128, one K step: K = 8 for kind::tf32, K = 16 for kind::f16 / bf16) as a
// function of N and of the shared-memory layout of the A operand (the conv kernels address their taps by shifting an
// un-swizzled K-major descriptor inside a staged halo tile).  Issue loop: one elected thread, unrolled, no waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mvster_b200/csrc -o tools/_build/mma_microbench tools/mma_microbench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_ptx.cuh"

using namespace mvster::ptx;

struct Case {
    int kind, n, layout, lbo, sbo, a_step, nacc, b_layout, b_lbo, b_sbo, kstep_a, kstep_b, reps, n2, n3;
};

template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0) umma_tf32(d, a, b, idesc, acc); else umma_bf16(d, a, b, idesc, acc);
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) bench_kernel(Case c, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
    const uint32_t bar_a = smem_u32(&bar), slot_a = smem_u32(&slot);
    if (threadIdx.x == 0) { mbar_init(bar_a, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(slot_a, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {
        const uint32_t id1 = KIND == 0 ? idesc_tf32_m128(c.n) : idesc_bf16_m128(c.n);
        const uint32_t id2 = KIND == 0 ? idesc_tf32_m128(c.n2) : idesc_bf16_m128(c.n2);
        const uint32_t id3 = KIND == 0 ? idesc_tf32_m128(c.n3) : idesc_bf16_m128(c.n3);
        const uint64_t a0 = smem_desc(base, c.lbo, c.sbo, c.layout);
        const uint64_t b0 = smem_desc(base + 128 * 1024, c.b_lbo, c.b_sbo, c.b_layout);
        const uint64_t a1 = a0 + (uint64_t)c.kstep_a, b1 = b0 + (uint64_t)c.kstep_b;
        const uint32_t d1 = tmem + (uint32_t)((c.nacc > 1 ? 1 : 0) * c.n);
        long long best = 1ll << 62;
        for (int rep = 0; rep < 5; ++rep) {
            const long long t0 = clock64();
            if (elect_one()) {
                for (int i = 0; i < c.reps; i += 6) {
                    const uint64_t sh = (uint64_t)(((i / 6) % 3) * c.a_step);
                    mma<KIND>(tmem, a0 + sh, b0, id1, i ? 1u : 0u);
                    mma<KIND>(tmem, a1 + sh, b1, id2, 1u);
                    mma<KIND>(tmem, a0 + sh + 1, b0, id3, 1u);
                    mma<KIND>(d1, a1 + sh + 1, b1, id1, i ? 1u : 0u);
                    mma<KIND>(d1, a0 + sh + 10, b0, id2, 1u);
                    mma<KIND>(d1, a1 + sh + 10, b1, id3, 1u);
                }
                umma_commit(bar_a);
            }
            __syncwarp();
            mbar_wait(bar_a, rep & 1);
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (threadIdx.x == 0) out[0] = best;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(bench_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    std::vector<std::pair<const char*, Case>> cases;
    const int R = 1152;  // multiple of 6
    for (int kind = 0; kind < 2; ++kind)
        for (int n : {16, 32, 48, 64, 96, 128, 192, 256}) {
            const int nacc = n >= 128 ? 1 : 2;
            // B un-swizzled K-major: [2 K-halves][n rows][16 B]
            auto mk = [&](int layout, int lbo, int sbo, int step, int ks) {
                return Case{kind, n, layout, lbo, sbo, step, nacc, 0, n * 16, 128, ks, 0, R, n, n};
            };
            cases.push_back({"halo tile: noswz lbo2880 sbo160, tap shifts", mk(0, 2880, 160, 1, 0)});
            cases.push_back({"dense:     noswz lbo2048 sbo128            ", mk(0, 2048, 128, 0, 0)});
            cases.push_back({"sw128 rows of 128 B (sbo1024), k +32B      ", mk(2, 16, 1024, 0, 2)});
        }
    // the per-tap MMA triples of the conv kernels: N = 3*NC, 2*NC, NC
    for (int kind = 0; kind < 2; ++kind)
        for (int nc : {16, 32, 64}) {
            Case c{kind, 3 * nc, 0, 2880, 160, 1, 1, 0, 3 * nc * 16, 128, 0, 0, R, 2 * nc, nc};
            cases.push_back({"halo tile: triple N = 3NC, 2NC, NC (avg)   ", c});
        }
    printf("%-46s %5s %5s %10s\n", "A layout", "kind", "N", "cyc/MMA");
    for (auto& pc : cases) {
        if (pc.second.kind == 0) bench_kernel<0><<<1, 128, 200 * 1024>>>(pc.second, d_out);
        else bench_kernel<1><<<1, 128, 200 * 1024>>>(pc.second, d_out);
        long long h = 0;
        cudaError_t e = cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", pc.first, cudaGetErrorString(e)); return 1; }
        printf("%-46s %5s %5d %10.1f\n", pc.first, pc.second.kind ? "bf16" : "tf32", pc.second.n, (double)h / pc.second.reps);
    }
    return 0;
}
