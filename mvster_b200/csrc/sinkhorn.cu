// Fused entropy-regularised optimal-transport loss of the training step (models/mvs4net_utils.py:1096-1142 `sinkhorn`,
// called per stage from MVS4net_loss, MVS4Net.py:149), forward AND backward in one launch.
//
// The reference builds, per pixel, a D x D (or D x (D+1), `continuous`) ground-cost matrix, runs `iters` log-domain Sinkhorn
// iterations as ~6 ATen kernels each on [B, HW, D, D] tensors, and lets autograd keep every intermediate of every iteration for
// the backward pass.  Here one thread owns one pixel: the matrix (<= 72 entries), both scaling vectors and the whole iteration
// history (<= 33 x 8 + 32 x 9 floats, thread-local) never leave the SM; the thread writes the pixel's transport cost and - by
// reverse-mode differentiation of its own history - d cost / d attn[d].  The host reduces (mean over the masked pixels) and
// scales the gradient with the upstream one (mvster_b200/train_ops.py).
//
//   log_mu_j = log(mu_j + 1e-12), log_nu_i = log(attn_i + 1e-12), M = cost / eps, u = v = 0
//   repeat iters:  v_j = log_mu_j - LSE_i(M_ij + u_i);   u_i = log_nu_i - LSE_j(M_ij + v_j)
//   T_ij = exp(M_ij + u_i + v_j);   loss = sum_ij T_ij cost_ij
// discrete:   mu = one-hot of argmin_d |hypo_d - gt| (first minimum), cost_ij = |i - j|
// continuous: one extra column j = D with mu = e_D, cost_iD = |(1/gt - 1/hypo_0) / (1/hypo_2 - 1/hypo_1) - i| (10 where !mask)
#include "common.cuh"
#include <math.h>

namespace mvster {

constexpr int SK_MAX_ITERS = 32;

template <int D, bool CONT>
__global__ void __launch_bounds__(128) sinkhorn_kernel(const float* __restrict__ gt, const float* __restrict__ hypo, const float* __restrict__ attn,
                                                       const unsigned char* __restrict__ mask, int iters, float eps, float* __restrict__ loss_px,
                                                       float* __restrict__ grad_attn, long long npix, long long plane) {
    constexpr int NJ = CONT ? D + 1 : D;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i0 >= npix) return;
    const long long b = i0 / plane, p = i0 % plane;
    const float* hp = hypo + b * D * plane + p;
    const float* ap = attn + b * D * plane + p;
    const bool m = mask[i0] != 0;
    const float g = gt[i0];
    float nu[D], log_nu[D], log_mu[NJ], cost[D][NJ];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        nu[i] = ap[(long long)i * plane];
        log_nu[i] = logf(nu[i] + 1e-12f);
    }
    if constexpr (CONT) {
        const float itv = 1.f / hp[2 * plane] - 1.f / hp[plane];
        float dist = (1.f / g - 1.f / hp[0]) / itv;
        if (!m) dist = 10.f;
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
            for (int j = 0; j < D; ++j) cost[i][j] = fabsf((float)(i - j));
            cost[i][D] = fabsf(dist - (float)i);
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) log_mu[j] = logf((j == D ? 1.f : 0.f) + 1e-12f);
    } else {
        int arg = 0;
        float best = fabsf(hp[0] - g);
#pragma unroll
        for (int d = 1; d < D; ++d) {
            const float e = fabsf(hp[(long long)d * plane] - g);
            if (e < best) { best = e; arg = d; }  // first minimum, like torch.min(dim)
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) cost[i][j] = fabsf((float)(i - j));
#pragma unroll
        for (int j = 0; j < NJ; ++j) log_mu[j] = logf((j == arg ? 1.f : 0.f) + 1e-12f);
    }
    float M[D][NJ];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) M[i][j] = cost[i][j] / eps;

    float uh[SK_MAX_ITERS + 1][D], vh[SK_MAX_ITERS][NJ];  // iteration history (thread-local)
#pragma unroll
    for (int i = 0; i < D; ++i) uh[0][i] = 0.f;
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {  // v_j = log_mu_j - logsumexp_i(M_ij + u_i)
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < D; ++i) mx = fmaxf(mx, M[i][j] + uh[k][i]);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) s += expf(M[i][j] + uh[k][i] - mx);
            vh[k][j] = log_mu[j] - (mx + logf(s));
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {  // u_i = log_nu_i - logsumexp_j(M_ij + v_j)
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NJ; ++j) mx = fmaxf(mx, M[i][j] + vh[k][j]);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) s += expf(M[i][j] + vh[k][j] - mx);
            uh[k + 1][i] = log_nu[i] - (mx + logf(s));
        }
    }
    // transport cost and its adjoints w.r.t. the final u, v
    float ub[D], vb[NJ], loss = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) ub[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) vb[j] = 0.f;
    if (iters > 0) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float t = expf(M[i][j] + uh[iters][i] + vh[iters - 1][j]) * cost[i][j];
                loss += t;
                ub[i] += t;
                vb[j] += t;
            }
    } else {  // no iteration: u = v = 0
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) loss += expf(M[i][j]) * cost[i][j];
    }
    loss_px[i0] = m ? loss : 0.f;
    if (!grad_attn) return;
    float ab[D];  // adjoint of log_nu
#pragma unroll
    for (int i = 0; i < D; ++i) ab[i] = 0.f;
    for (int k = iters - 1; k >= 0; --k) {
        // u_{k+1,i} = log_nu_i - LSE_j(M_ij + v_{k,j}):  ab_i += ub_i;  vb_j -= sum_i ub_i softmax_j(M_ij + v_{k,j})
        float vnew[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) vnew[j] = vb[j];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            ab[i] += ub[i];
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NJ; ++j) mx = fmaxf(mx, M[i][j] + vh[k][j]);
            float e[NJ], s = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) { e[j] = expf(M[i][j] + vh[k][j] - mx); s += e[j]; }
            const float r = ub[i] / s;
#pragma unroll
            for (int j = 0; j < NJ; ++j) vnew[j] -= r * e[j];
        }
        // v_{k,j} = log_mu_j - LSE_i(M_ij + u_{k,i}):  ub_i(prev) = - sum_j vb_j softmax_i(M_ij + u_{k,i})   (v_{k-1} is not read again)
        float unew[D];
#pragma unroll
        for (int i = 0; i < D; ++i) unew[i] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < D; ++i) mx = fmaxf(mx, M[i][j] + uh[k][i]);
            float e[D], s = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) { e[i] = expf(M[i][j] + uh[k][i] - mx); s += e[i]; }
            const float r = vnew[j] / s;
#pragma unroll
            for (int i = 0; i < D; ++i) unew[i] -= r * e[i];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) ub[i] = unew[i];
#pragma unroll
        for (int j = 0; j < NJ; ++j) vb[j] = 0.f;
    }
    float* gp = grad_attn + b * D * plane + p;
#pragma unroll
    for (int i = 0; i < D; ++i) gp[(long long)i * plane] = m ? ab[i] / (nu[i] + 1e-12f) : 0.f;  // d log(attn + 1e-12) / d attn
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_sinkhorn_f32(const float* gt_depth, const float* hypo, const float* attn, const unsigned char* mask,
                                   float* loss_px, float* grad_attn, int B, int D, int H, int W, int iters, float eps, int continuous,
                                   mvster_stream_t stream) {
    MVSTER_REQUIRE(gt_depth && hypo && attn && mask && loss_px, "mvster_sinkhorn_f32: null pointer");
    MVSTER_REQUIRE(B > 0 && H > 0 && W > 0, "mvster_sinkhorn_f32: bad shape");
    MVSTER_REQUIRE(iters >= 0 && iters <= SK_MAX_ITERS, "mvster_sinkhorn_f32: iters=%d outside 0..%d", iters, SK_MAX_ITERS);
    MVSTER_REQUIRE(eps != 0.f, "mvster_sinkhorn_f32: eps == 0");
    const long long plane = (long long)H * W, n = (long long)B * plane;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div(n, 128);
    if (D == 4 && !continuous) sinkhorn_kernel<4, false><<<grid, 128, 0, st>>>(gt_depth, hypo, attn, mask, iters, eps, loss_px, grad_attn, n, plane);
    else if (D == 4) sinkhorn_kernel<4, true><<<grid, 128, 0, st>>>(gt_depth, hypo, attn, mask, iters, eps, loss_px, grad_attn, n, plane);
    else if (D == 8 && !continuous) sinkhorn_kernel<8, false><<<grid, 128, 0, st>>>(gt_depth, hypo, attn, mask, iters, eps, loss_px, grad_attn, n, plane);
    else if (D == 8) sinkhorn_kernel<8, true><<<grid, 128, 0, st>>>(gt_depth, hypo, attn, mask, iters, eps, loss_px, grad_attn, n, plane);
    else MVSTER_REQUIRE(false, "mvster_sinkhorn_f32: unsupported D=%d (4 or 8)", D);
    return check_launch("sinkhorn_kernel");
}
