// Argument block of the fused warp + epipolar-Transformer kernels (et_fuse.cu and its .cuh variants).
#pragma once

namespace mvster {

struct EtArgs {
    const float* ref;
    const float* src[MVSTER_MAX_VIEWS];
    const float* pose;  // [B][V][12]
    const float* hypo;  // [B][D][H][W]
    float* cost;        // [B][D][H][W][G]
    float* wsum;        // [B][D][H][W] or nullptr
    int B, V, H, W, Hs, Ws;
    float attn_temp, sqrt_c;
    int flags;
    int prefetch;  // window kernels: L2 prefetch of every view's window rows before the view loop
};

}  // namespace mvster
