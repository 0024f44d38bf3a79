// CPU build of SIMT kernel sources of libmvster_b200 (see simt_emu.h): the window warp + ET kernel and the geometric-consistency
// kernel, exported with plain host pointers for tests/test_emu_kernels.py.
#include "simt_emu.h"

#define MVSTER_MAX_VIEWS 16
#define MVSTER_ET_PARTIAL 1
#define MVSTER_ET_ACCUMULATE 2
typedef void* mvster_stream_t;
#define MVSTER_OK 0
#define MVSTER_ERR_ARG (-1)
#define MVSTER_REQUIRE(cond, ...) do { if (!(cond)) return MVSTER_ERR_ARG; } while (0)
namespace mvster {
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int check_launch(const char*) { return 0; }
}

#include "../../mvster_b200/csrc/et_args.cuh"
#include "../../mvster_b200/csrc/et_fuse_tiled.cuh"   // packed-fp32 helpers only (the tiled kernel is excluded under MVSTER_CPU_EMU)
#include "../../mvster_b200/csrc/et_fuse_win.cuh"

using namespace mvster;

template <int C, int G, int D, int LPP>
static void run_win(const EtArgs& a) {
    dim3 grid(ceil_div(a.W, 32 / LPP), ceil_div(a.H, 4), a.B);
    emu::launch(grid, dim3(128), [&] { et_fuse_win_kernel<C, G, D, LPP, 4>(a); });
}

extern "C" int emu_et_fuse_win(const float* ref, const float* const* src, int V, const float* pose, const float* hypo, float* cost,
                               float* wsum, int B, int C, int G, int D, int H, int W, int Hs, int Ws, float attn_temp, int flags) {
    EtArgs a;
    a.ref = ref;
    for (int v = 0; v < MVSTER_MAX_VIEWS; ++v) a.src[v] = v < V ? src[v] : nullptr;
    a.pose = pose; a.hypo = hypo; a.cost = cost; a.wsum = wsum;
    a.B = B; a.V = V; a.H = H; a.W = W; a.Hs = Hs; a.Ws = Ws;
    a.attn_temp = attn_temp; a.sqrt_c = (float)std::sqrt((double)C); a.flags = flags; a.prefetch = 1;
    if (C == 8 && G == 4 && D == 4) run_win<8, 4, 4, 1>(a);
    else if (C == 16 && G == 4 && D == 4) run_win<16, 4, 4, 2>(a);
    else if (C == 32 && G == 8 && D == 8) run_win<32, 8, 8, 4>(a);
    else return -1;
    return 0;
}

// ---- geometric-consistency kernel (csrc/fusion.cu) ----
#include "../../mvster_b200/csrc/fusion.cu"

extern "C" int emu_geo_consistency(const float* depth_ref, const float* depth_src, const double* mats, float* depth_reproj, float* x_src,
                                   float* y_src, unsigned char* mask, int* mask_sum, float* depth_sum, int H, int W, int Hs, int Ws,
                                   float dist_thres, float rel_thres) {
    GeoArgs a;
    a.depth_ref = depth_ref; a.depth_src = depth_src;
    const double* m = mats;
    for (int i = 0; i < 9; ++i) a.kr_inv[i] = *m++;
    for (int i = 0; i < 12; ++i) a.t_rs[i] = *m++;
    for (int i = 0; i < 9; ++i) a.ks[i] = *m++;
    for (int i = 0; i < 9; ++i) a.ks_inv[i] = *m++;
    for (int i = 0; i < 12; ++i) a.t_sr[i] = *m++;
    for (int i = 0; i < 9; ++i) a.kr[i] = *m++;
    a.depth_reproj = depth_reproj; a.x_src = x_src; a.y_src = y_src; a.mask = mask; a.mask_sum = mask_sum; a.depth_sum = depth_sum;
    a.H = H; a.W = W; a.Hs = Hs; a.Ws = Ws; a.dist_thres = (double)dist_thres; a.rel_thres = rel_thres;
    emu::launch(dim3(ceil_div(W, 32), ceil_div(H, 8)), dim3(256), [&] { geo_consistency_kernel(a); });
    return 0;
}
