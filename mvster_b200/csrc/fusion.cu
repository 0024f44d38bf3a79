// Geometric-consistency check between a reference depth map and one source depth map - the step right after the forward in
// the reference's test_mvs4.py (reproject_with_depth :271-309, check_geometric_consistency :312-328, accumulation of
// filter_depth :362-378), there numpy + cv2.remap in a 4-process pool.  SURVEY.md 8f "next" #3.
//
// One thread per reference pixel, whole chain in registers (HBM-bound: reads 4 B of reference depth and a 2x2 patch of
// the source depth, writes <= 17 B).  Arithmetic follows the reference: float64 geometry on camera matrices that the host
// prepared with the reference's own numpy calls (float32 inverses / products, passed as doubles), float32 casts where the
// reference casts, and cv2.remap(INTER_LINEAR)'s fixed-point sampling for the source-depth lookup: coordinates rounded to
// 1/32 pixel, float32 weights, zeros outside the image, products and sums rounded separately.
#include "common.cuh"

namespace mvster {

struct GeoArgs {
    const float* depth_ref;  // [H][W]
    const float* depth_src;  // [Hs][Ws]
    double kr_inv[9], t_rs[12], ks[9], ks_inv[9], t_sr[12], kr[9];
    float* depth_reproj;     // [H][W]  zero where the check fails
    float* x_src; float* y_src;  // [H][W] or null
    unsigned char* mask;     // [H][W] or null
    int* mask_sum;           // [H][W] += mask, or null
    float* depth_sum;        // [H][W] += depth_reproj, or null
    int H, W, Hs, Ws;
    double dist_thres;
    float rel_thres;
};

__device__ __forceinline__ float remap_tap(const float* __restrict__ s, int y, int x, int Hs, int Ws) {
    return ((unsigned)x < (unsigned)Ws && (unsigned)y < (unsigned)Hs) ? __ldg(s + (long long)y * Ws + x) : 0.f;
}

// cv2.remap(src, x, y, INTER_LINEAR), BORDER_CONSTANT 0, for one float32 coordinate pair
__device__ __forceinline__ float remap_bilinear(const float* __restrict__ s, float x, float y, int Hs, int Ws) {
    // OpenCV saturates the fixed-point coordinates to int; keep far-away / non-finite coordinates outside the image instead
    if (!(x > -4.f && x < (float)Ws + 4.f && y > -4.f && y < (float)Hs + 4.f)) return 0.f;
    const int sx = __float2int_rn(__fmul_rn(x, 32.f)), sy = __float2int_rn(__fmul_rn(y, 32.f));
    const int ix = sx >> 5, iy = sy >> 5;
    const float fx = (float)(sx & 31) * 0.03125f, fy = (float)(sy & 31) * 0.03125f;
    const float ex = __fsub_rn(1.f, fx), ey = __fsub_rn(1.f, fy);
    const float w00 = __fmul_rn(ey, ex), w01 = __fmul_rn(ey, fx), w10 = __fmul_rn(fy, ex), w11 = __fmul_rn(fy, fx);
    float r = __fmul_rn(remap_tap(s, iy, ix, Hs, Ws), w00);
    r = __fadd_rn(r, __fmul_rn(remap_tap(s, iy, ix + 1, Hs, Ws), w01));
    r = __fadd_rn(r, __fmul_rn(remap_tap(s, iy + 1, ix, Hs, Ws), w10));
    r = __fadd_rn(r, __fmul_rn(remap_tap(s, iy + 1, ix + 1, Hs, Ws), w11));
    return r;
}

__device__ __forceinline__ void mat3(const double* m, double a, double b, double c, double& x, double& y, double& z) {
    x = m[0] * a + m[1] * b + m[2] * c;
    y = m[3] * a + m[4] * b + m[5] * c;
    z = m[6] * a + m[7] * b + m[8] * c;
}
__device__ __forceinline__ void rigid(const double* m, double a, double b, double c, double& x, double& y, double& z) {
    x = m[0] * a + m[1] * b + m[2] * c + m[3];   // [R|t] . (a, b, c, 1)
    y = m[4] * a + m[5] * b + m[6] * c + m[7];
    z = m[8] * a + m[9] * b + m[10] * c + m[11];
}

__global__ void __launch_bounds__(256) geo_consistency_kernel(const GeoArgs a) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= a.W || y >= a.H) return;
    const long long p = (long long)y * a.W + x;
    const float d_ref = __ldg(a.depth_ref + p);
    const double d = (double)d_ref;
    double rx, ry, rz, sx, sy, sz, kx, ky, kz;
    mat3(a.kr_inv, (double)x * d, (double)y * d, d, rx, ry, rz);       // reference camera space
    rigid(a.t_rs, rx, ry, rz, sx, sy, sz);                             // source camera space
    mat3(a.ks, sx, sy, sz, kx, ky, kz);
    const double u = kx / kz, v = ky / kz;                             // source pixel (float64)
    const float xs = (float)u, ys = (float)v;
    const float d_src = remap_bilinear(a.depth_src, xs, ys, a.Hs, a.Ws);
    const double ds = (double)d_src;
    mat3(a.ks_inv, u * ds, v * ds, ds, sx, sy, sz);                    // back-projected with the SOURCE depth
    rigid(a.t_sr, sx, sy, sz, rx, ry, rz);                             // reference camera space again
    float d_rep = (float)rz;
    mat3(a.kr, rx, ry, rz, kx, ky, kz);
    const float xr = (float)(kx / kz), yr = (float)(ky / kz);
    const double ex = (double)xr - (double)x, ey = (double)yr - (double)y;
    const double dist = sqrt(ex * ex + ey * ey);
    const float rel = __fdiv_rn(fabsf(__fsub_rn(d_rep, d_ref)), d_ref);
    const bool ok = dist < a.dist_thres && rel < a.rel_thres;          // NaN compares false, as in numpy
    if (!ok) d_rep = 0.f;
    a.depth_reproj[p] = d_rep;
    if (a.x_src) a.x_src[p] = xs;
    if (a.y_src) a.y_src[p] = ys;
    if (a.mask) a.mask[p] = ok ? 1 : 0;
    if (a.mask_sum) a.mask_sum[p] += ok ? 1 : 0;
    if (a.depth_sum) a.depth_sum[p] = __fadd_rn(a.depth_sum[p], d_rep);
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_geo_consistency_f32(const float* depth_ref, const float* depth_src, const double* mats, float* depth_reproj,
                                          float* x_src, float* y_src, unsigned char* mask, int* mask_sum, float* depth_sum,
                                          int H, int W, int Hs, int Ws, float dist_thres, float rel_thres, mvster_stream_t stream) {
    MVSTER_REQUIRE(depth_ref && depth_src && mats && depth_reproj, "mvster_geo_consistency_f32: null pointer");
    MVSTER_REQUIRE(H > 0 && W > 0 && Hs > 0 && Ws > 0, "mvster_geo_consistency_f32: bad shape");
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
    a.H = H; a.W = W; a.Hs = Hs; a.Ws = Ws;
    a.dist_thres = (double)dist_thres; a.rel_thres = rel_thres;
    dim3 grid(ceil_div(W, 32), ceil_div(H, 8));
    geo_consistency_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("geo_consistency_kernel");
}
