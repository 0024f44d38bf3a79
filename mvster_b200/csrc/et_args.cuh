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
    const int4* boxes;  // TMA kernel: [tiles][V] source boxes (x0, y0, rows, 0) from et_tile_boxes_kernel, or nullptr
};

}  // namespace mvster
