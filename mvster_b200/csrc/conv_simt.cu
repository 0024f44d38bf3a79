// Channels-last (NDHWC) 3-D convolution layers with folded BatchNorm on the CUDA cores (exact
// fp32 FMA), and the reg2d U-Net driver built from them.
//   ConvBnReLU3D                      models/mvs4net_utils.py:116-123
//   ConvTranspose3d + BN + ReLU       models/mvs4net_utils.py:885-898
//   reg2d.forward                     models/mvs4net_utils.py:902-912
// This is the exact-fp32 path (bit-faithful to an fp32 FMA chain; it is what the parity tests
// pin) and the path for the 4..8-channel full-resolution layers, which are HBM/L1-bandwidth
// bound (12-24 FLOP/B).  One thread = one output voxel x COUT_T output channels; the layer's
// weights for that channel slice sit in shared memory ([tap][cin][cout], warp-uniform
// broadcast reads), activations are read straight from global/L1 with 128-bit loads (each
// voxel's Cin channels are contiguous).
#include "common.cuh"

namespace mvster {

struct ConvArgs {
    const float* x; const float* w; const float* bias; const float* skip; float* y;
    int B, Di, Hi, Wi, Do, Ho, Wo;
    int cout, kd, sd, s, relu;
};

template <int CIN, int COUT_T>
__global__ void __launch_bounds__(128) conv_fwd_kernel(const ConvArgs a) {
    extern __shared__ __align__(16) float w_s[];  // [kd*9][CIN][COUT_T]
    const int cg = blockIdx.y;                    // output-channel slice
    const int taps = a.kd * 9;
    for (int i = threadIdx.x; i < taps * CIN * COUT_T; i += blockDim.x) {
        const int o = i % COUT_T, rest = i / COUT_T;  // rest = tap*CIN + cin
        w_s[i] = __ldg(a.w + (long long)rest * a.cout + cg * COUT_T + o);
    }
    __syncthreads();
    const long long nvox = (long long)a.B * a.Do * a.Ho * a.Wo;
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    const int ox = (int)(v % a.Wo), oy = (int)((v / a.Wo) % a.Ho);
    const int oz = (int)((v / ((long long)a.Wo * a.Ho)) % a.Do), b = (int)(v / ((long long)a.Wo * a.Ho * a.Do));

    float acc[COUT_T];
#pragma unroll
    for (int o = 0; o < COUT_T; ++o) acc[o] = a.bias ? __ldg(a.bias + cg * COUT_T + o) : 0.f;

    const int pz = a.kd / 2;
    for (int kz = 0; kz < a.kd; ++kz) {
        const int iz = oz * a.sd + kz - pz;
        if ((unsigned)iz >= (unsigned)a.Di) continue;
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * a.s + ky - 1;
            if ((unsigned)iy >= (unsigned)a.Hi) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = ox * a.s + kx - 1;
                if ((unsigned)ix >= (unsigned)a.Wi) continue;
                const float4* px = reinterpret_cast<const float4*>(
                    a.x + ((((long long)b * a.Di + iz) * a.Hi + iy) * a.Wi + ix) * CIN);
                const float* wt = w_s + ((kz * 3 + ky) * 3 + kx) * CIN * COUT_T;
#pragma unroll
                for (int c4 = 0; c4 < CIN / 4; ++c4) {
                    const float4 t = __ldg(px + c4);
                    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float* wr = wt + (c4 * 4 + j) * COUT_T;
#pragma unroll
                        for (int o = 0; o < COUT_T; ++o) acc[o] = fmaf(tv[j], wr[o], acc[o]);
                    }
                }
            }
        }
    }
    const long long off = v * a.cout + cg * COUT_T;
#pragma unroll
    for (int o = 0; o < COUT_T; ++o) {
        float r = acc[o];
        if (a.relu) r = fmaxf(r, 0.f);
        if (a.skip) r += __ldg(a.skip + off + o);  // skip is added AFTER the ReLU (:907-909)
        acc[o] = r;
    }
    if constexpr (COUT_T % 4 == 0) {
#pragma unroll
        for (int o = 0; o < COUT_T; o += 4)
            *reinterpret_cast<float4*>(a.y + off + o) = make_float4(acc[o], acc[o + 1], acc[o + 2], acc[o + 3]);
    } else {
#pragma unroll
        for (int o = 0; o < COUT_T; ++o) a.y[off + o] = acc[o];
    }
}

// ConvTranspose3d kernel (1,3,3), stride (1,2,2), padding (0,1,1), output_padding (0,1,1):
// out[2i-1+ky][2j-1+kx] += in[i][j] * w[ky][kx].  One thread owns the 2x2 output quad whose
// top-left input is (i,j): exactly the 9 (Cin x Cout) tap products, no parity branches.
template <int CIN, int COUT_T>
__global__ void __launch_bounds__(128) deconv_fwd_kernel(const ConvArgs a) {
    extern __shared__ __align__(16) float w_s[];  // [9][CIN][COUT_T]
    const int cg = blockIdx.y;
    for (int i = threadIdx.x; i < 9 * CIN * COUT_T; i += blockDim.x) {
        const int o = i % COUT_T, rest = i / COUT_T;
        w_s[i] = __ldg(a.w + (long long)rest * a.cout + cg * COUT_T + o);
    }
    __syncthreads();
    const long long nq = (long long)a.B * a.Di * a.Hi * a.Wi;
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int j = (int)(q % a.Wi), i = (int)((q / a.Wi) % a.Hi);
    const long long bz = q / ((long long)a.Wi * a.Hi);  // b*Di + z
    const bool right = j + 1 < a.Wi, down = i + 1 < a.Hi;

    float acc[4][COUT_T];
#pragma unroll
    for (int o = 0; o < COUT_T; ++o) {
        const float bv = a.bias ? __ldg(a.bias + cg * COUT_T + o) : 0.f;
        acc[0][o] = acc[1][o] = acc[2][o] = acc[3][o] = bv;
    }
    const float4* p00 = reinterpret_cast<const float4*>(a.x + ((bz * a.Hi + i) * a.Wi + j) * CIN);
    const float4* p01 = p00 + CIN / 4;
    const float4* p10 = p00 + (long long)a.Wi * (CIN / 4);
    const float4* p11 = p10 + CIN / 4;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int c4 = 0; c4 < CIN / 4; ++c4) {
        const float4 t00 = __ldg(p00 + c4);
        const float4 t01 = right ? __ldg(p01 + c4) : zero;
        const float4 t10 = down ? __ldg(p10 + c4) : zero;
        const float4 t11 = (right && down) ? __ldg(p11 + c4) : zero;
        const float v00[4] = {t00.x, t00.y, t00.z, t00.w}, v01[4] = {t01.x, t01.y, t01.z, t01.w};
        const float v10[4] = {t10.x, t10.y, t10.z, t10.w}, v11[4] = {t11.x, t11.y, t11.z, t11.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int c = c4 * 4 + jj;
            auto W = [&](int ky, int kx) { return w_s + ((ky * 3 + kx) * CIN + c) * COUT_T; };
#pragma unroll
            for (int o = 0; o < COUT_T; ++o) {
                acc[0][o] = fmaf(v00[jj], W(1, 1)[o], acc[0][o]);
                acc[1][o] = fmaf(v00[jj], W(1, 2)[o], fmaf(v01[jj], W(1, 0)[o], acc[1][o]));
                acc[2][o] = fmaf(v00[jj], W(2, 1)[o], fmaf(v10[jj], W(0, 1)[o], acc[2][o]));
                acc[3][o] = fmaf(v00[jj], W(2, 2)[o], fmaf(v01[jj], W(2, 0)[o],
                            fmaf(v10[jj], W(0, 2)[o], fmaf(v11[jj], W(0, 0)[o], acc[3][o]))));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int oy = 2 * i + (k >> 1), ox = 2 * j + (k & 1);
        const long long off = ((bz * a.Ho + oy) * a.Wo + ox) * a.cout + cg * COUT_T;
#pragma unroll
        for (int o = 0; o < COUT_T; ++o) {
            float r = acc[k][o];
            if (a.relu) r = fmaxf(r, 0.f);
            if (a.skip) r += __ldg(a.skip + off + o);
            acc[k][o] = r;
        }
        if constexpr (COUT_T % 4 == 0) {
#pragma unroll
            for (int o = 0; o < COUT_T; o += 4)
                *reinterpret_cast<float4*>(a.y + off + o) = make_float4(acc[k][o], acc[k][o + 1], acc[k][o + 2], acc[k][o + 3]);
        } else {
#pragma unroll
            for (int o = 0; o < COUT_T; ++o) a.y[off + o] = acc[k][o];
        }
    }
}

template <int CIN, int COUT_T>
static int launch_conv(const ConvArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)a.kd * 9 * CIN * COUT_T * sizeof(float);
    auto k = conv_fwd_kernel<CIN, COUT_T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long n = (long long)a.B * a.Do * a.Ho * a.Wo;
    k<<<dim3(ceil_div(n, 128), a.cout / COUT_T), 128, smem, st>>>(a);
    return check_launch("conv_fwd_kernel");
}

template <int CIN, int COUT_T>
static int launch_deconv(const ConvArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)9 * CIN * COUT_T * sizeof(float);
    auto k = deconv_fwd_kernel<CIN, COUT_T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long n = (long long)a.B * a.Di * a.Hi * a.Wi;
    k<<<dim3(ceil_div(n, 128), a.cout / COUT_T), 128, smem, st>>>(a);
    return check_launch("deconv_fwd_kernel");
}

template <int CIN>
static int dispatch_cout(const ConvArgs& a, bool transposed, cudaStream_t st) {
    if (transposed) {  // 4 accumulator sets per thread: keep the channel slice at 8
        if (a.cout % 8 == 0) return launch_deconv<CIN, 8>(a, st);
        set_error("mvster_conv3d_ndhwc_f32: transposed conv needs Cout %% 8 == 0 (got %d)", a.cout);
        return MVSTER_ERR_ARG;
    }
    if (a.cout % 16 == 0) return launch_conv<CIN, 16>(a, st);
    if (a.cout % 8 == 0) return launch_conv<CIN, 8>(a, st);
    if (a.cout == 1) return launch_conv<CIN, 1>(a, st);
    set_error("mvster_conv3d_ndhwc_f32: unsupported Cout=%d (1 or a multiple of 8)", a.cout);
    return MVSTER_ERR_ARG;
}

static int run_conv(const float* x, const float* w, const float* bias, const float* skip, float* y,
                    int B, int Di, int Hi, int Wi, int Cin, int Cout, int kd, int sd, int s, int transposed,
                    int relu, cudaStream_t st) {
    MVSTER_REQUIRE(x && w && y, "mvster_conv3d_ndhwc_f32: null pointer");
    MVSTER_REQUIRE(B > 0 && Di > 0 && Hi > 0 && Wi > 0, "mvster_conv3d_ndhwc_f32: bad shape");
    MVSTER_REQUIRE(kd == 1 || kd == 3, "mvster_conv3d_ndhwc_f32: kd=%d (1 or 3)", kd);
    MVSTER_REQUIRE((sd == 1 || sd == 2) && (s == 1 || s == 2), "mvster_conv3d_ndhwc_f32: strides must be 1 or 2");
    ConvArgs a;
    a.x = x; a.w = w; a.bias = bias; a.skip = skip; a.y = y;
    a.B = B; a.Di = Di; a.Hi = Hi; a.Wi = Wi;
    a.cout = Cout; a.kd = kd; a.sd = sd; a.s = s; a.relu = relu;
    if (transposed) {
        MVSTER_REQUIRE(kd == 1 && sd == 1 && s == 2, "mvster_conv3d_ndhwc_f32: transposed supports kernel (1,3,3) stride (1,2,2) only");
        a.Do = Di; a.Ho = 2 * Hi; a.Wo = 2 * Wi;
    } else {
        a.Do = (Di - 1) / sd + 1; a.Ho = (Hi - 1) / s + 1; a.Wo = (Wi - 1) / s + 1;  // pad = k/2, k = 3 (or 1 along D)
        const int rc = conv_px2(x, w, bias, skip, y, B, Di, Hi, Wi, Cin, Cout, kd, 3, sd, s, relu, st);
        if (rc != -100) return rc;  // two-pixels-per-thread kernel covered the layer
    }
    switch (Cin) {
        case 4: return dispatch_cout<4>(a, transposed, st);
        case 8: return dispatch_cout<8>(a, transposed, st);
        case 16: return dispatch_cout<16>(a, transposed, st);
        case 32: return dispatch_cout<32>(a, transposed, st);
        case 64: return dispatch_cout<64>(a, transposed, st);
    }
    set_error("mvster_conv3d_ndhwc_f32: unsupported Cin=%d (4,8,16,32,64)", Cin);
    return MVSTER_ERR_ARG;
}

// ---- reg2d layer table ----------------------------------------------------------------------
struct Layer { int cin, cout, kd, s, transposed; };
static void reg2d_layers(int G, Layer (&L)[MVSTER_REG2D_LAYERS]) {
    const Layer t[MVSTER_REG2D_LAYERS] = {
        {G, 8, 1, 1, 0},    // conv0  (1,3,3)
        {8, 16, 1, 2, 0},   // conv1  (1,3,3) s2
        {16, 16, 3, 1, 0},  // conv2  3x3x3
        {16, 32, 1, 2, 0},  // conv3
        {32, 32, 3, 1, 0},  // conv4
        {32, 64, 1, 2, 0},  // conv5
        {64, 64, 3, 1, 0},  // conv6
        {64, 32, 1, 2, 1},  // conv7  transposed, + conv4
        {32, 16, 1, 2, 1},  // conv9  transposed, + conv2
        {16, 8, 1, 2, 1},   // conv11 transposed, + conv0
    };
    for (int i = 0; i < MVSTER_REG2D_LAYERS; ++i) L[i] = t[i];
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_conv3d_ndhwc_f32(const float* x, const float* w, const float* bias, const float* skip, float* y,
                                       int B, int Di, int Hi, int Wi, int Cin, int Cout,
                                       int kd, int stride_d, int stride_hw, int transposed, int relu,
                                       mvster_stream_t stream) {
    return run_conv(x, w, bias, skip, y, B, Di, Hi, Wi, Cin, Cout, kd, stride_d, stride_hw, transposed, relu,
                    (cudaStream_t)stream);
}

extern "C" int mvster_reg2d_layer_info(int G, int layer, int64_t* info) {
    MVSTER_REQUIRE(info && layer >= 0 && layer < MVSTER_REG2D_LAYERS, "mvster_reg2d_layer_info: bad layer %d", layer);
    Layer L[MVSTER_REG2D_LAYERS];
    reg2d_layers(G, L);
    int64_t off = 0;
    for (int i = 0; i <= layer; ++i) {
        const int64_t taps = L[i].kd * 9, nw = taps * L[i].cin * L[i].cout;
        if (i == layer) {
            info[0] = L[i].cin; info[1] = L[i].cout; info[2] = L[i].kd; info[3] = L[i].s; info[4] = L[i].transposed;
            info[5] = off; info[6] = off + nw; info[7] = taps;
        }
        off += nw + L[i].cout;
    }
    return MVSTER_OK;
}

extern "C" size_t mvster_reg2d_blob_floats(int G) {
    int64_t info[8];
    if (mvster_reg2d_layer_info(G, MVSTER_REG2D_LAYERS - 1, info) != MVSTER_OK) return 0;
    return (size_t)(info[6] + info[1]);
}

// activations (floats, N = B*D*H*W): c0 8N | c1 4N | c2 4N | c3 2N | c4 2N | c5 N | c6 N | u7 2N | u9 4N  = 28N
extern "C" size_t mvster_reg2d_workspace_floats(int B, int D, int H, int W) {
    return (size_t)28 * B * D * H * W;
}

static int reg2d_run(const float* blob, const float* tc_blob, int npass, int gen, const float* cost, float* feat8, float* ws,
                     int B, int G, int D, int H, int W, mvster_stream_t stream);

extern "C" int mvster_reg2d_f32(const float* blob, const float* cost, float* feat8, float* ws,
                                int B, int G, int D, int H, int W, mvster_stream_t stream) {
    return reg2d_run(blob, nullptr, 0, 0, cost, feat8, ws, B, G, D, H, W, stream);
}

extern "C" size_t mvster_reg2d_tc_blob_floats(void) { return (size_t)2 * 27 * (16 * 16 + 32 * 32 + 64 * 64); }

extern "C" int mvster_reg2d_tc_f32(const float* blob, const float* tc_blob, const float* cost, float* feat8, float* ws,
                                   int B, int G, int D, int H, int W, int npass, int kernel_gen, mvster_stream_t stream) {
    MVSTER_REQUIRE(tc_blob, "mvster_reg2d_tc_f32: tc_blob is null");
    MVSTER_REQUIRE(npass == 1 || npass == 3, "mvster_reg2d_tc_f32: npass must be 1 or 3");
    MVSTER_REQUIRE(kernel_gen == 2, "mvster_reg2d_tc_f32: kernel_gen must be 2 (the first-generation TF32 kernel was removed)");
    return reg2d_run(blob, tc_blob, npass, kernel_gen, cost, feat8, ws, B, G, D, H, W, stream);
}

// Generation-3 blob: slab streams of conv0..conv6 back to back, then the transposed layers conv7 (output rows of parity 0,
// then parity 1: 4*32 columns do not fit one launch), conv9, conv11 (all four parity classes in one launch).
static int tc3_deconv_rows(const Layer& l) { return 4 * l.cout <= 64 ? -1 : 0; }  // -1: one launch, 0: two launches (rows 0, 1)
static size_t tc3_layer_bytes(const Layer& l) {
    if (!l.transposed) return mvster_conv_tc3_packed_bytes(l.cin, l.cout, l.kd, 3, l.s);
    if (tc3_deconv_rows(l) < 0) return mvster_deconv_tc3_packed_bytes(l.cin, l.cout, -1);
    return mvster_deconv_tc3_packed_bytes(l.cin, l.cout, 0) + mvster_deconv_tc3_packed_bytes(l.cin, l.cout, 1);
}
static size_t tc3_layer_offset(const Layer (&L)[MVSTER_REG2D_LAYERS], int layer) {
    size_t off = 0;
    for (int i = 0; i < layer; ++i) off += tc3_layer_bytes(L[i]);
    return off;
}

extern "C" size_t mvster_reg2d_tc3_blob_bytes(int G) {
    Layer L[MVSTER_REG2D_LAYERS];
    reg2d_layers(G, L);
    return tc3_layer_offset(L, MVSTER_REG2D_LAYERS);
}

// Two-fp16-term regulariser with PACKED activations between the layers: conv0 (CUDA cores, fp32 cost volume in) and every
// tensor-core epilogue store the fp16 pair (a1, a2) of their output in the operand layout of the next layer; the last layer
// (conv11 + conv0) writes fp32 for the head.  Same blobs as the unpacked form (folded weights, 'tc3h_blob').
static int reg2d_packed_h16(const float* blob, const void* tc3_blob, const float* cost, float* feat8, float* ws,
                            int B, int G, int D, int H, int W, mvster_stream_t stream) {
    MVSTER_REQUIRE(blob && cost && feat8 && ws, "mvster_reg2d_tc3_f32: null pointer");
    MVSTER_REQUIRE(H % 8 == 0 && W % 8 == 0, "mvster_reg2d_f32: H,W must be multiples of 8 (got %dx%d)", H, W);
    Layer L[MVSTER_REG2D_LAYERS];
    reg2d_layers(G, L);
    const size_t N = (size_t)B * D * H * W;
    float* c0 = ws;           float* c1 = c0 + 8 * N;  float* c2 = c1 + 4 * N;  float* c3 = c2 + 4 * N;   // 32 bits per element, as fp32
    float* c4 = c3 + 2 * N;   float* c5 = c4 + 2 * N;  float* c6 = c5 + N;      float* u7 = c6 + N;
    float* u9 = u7 + 2 * N;
    const float* in[MVSTER_REG2D_LAYERS] = {cost, c0, c1, c2, c3, c4, c5, c6, u7, u9};
    float* out[MVSTER_REG2D_LAYERS] = {c0, c1, c2, c3, c4, c5, c6, u7, u9, feat8};
    const float* skip[MVSTER_REG2D_LAYERS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, c4, c2, c0};
    const int div[MVSTER_REG2D_LAYERS] = {1, 1, 2, 2, 4, 4, 8, 8, 4, 2};
    for (int i = 0; i < MVSTER_REG2D_LAYERS; ++i) {
        int64_t info[8];
        mvster_reg2d_layer_info(G, i, info);
        const uint8_t* wl = (const uint8_t*)tc3_blob + tc3_layer_offset(L, i);
        const float* bias = blob + info[6];
        const int h = H / div[i], w = W / div[i];
        const int pf = 1 | MVSTER_TC3_FP16X2 | (i == MVSTER_REG2D_LAYERS - 1 ? 0 : MVSTER_TC3_OUT_PB16);
        int rc;
        if (i == 0) {
            rc = conv0_packed_h16(cost, blob + info[5], bias, c0, mvster_tc3_overflow_flag(), (long long)B * D, H, W, G, (cudaStream_t)stream);
            MVSTER_REQUIRE(rc != -100, "mvster_reg2d_tc3_f32: conv0 shape not covered by the packed form");
        } else if (!L[i].transposed) {
            rc = mvster_conv_tc3_pb16(in[i], wl, nullptr, bias, skip[i], out[i], B, D, h, w, L[i].cin, L[i].cout, L[i].kd, 3, L[i].s, pf, stream);
        } else {
            const int rows = tc3_deconv_rows(L[i]);
            rc = mvster_deconv_tc3_pb16(in[i], wl, nullptr, bias, skip[i], out[i], B, D, h, w, L[i].cin, L[i].cout, rows, pf, stream);
            if (rc == MVSTER_OK && rows == 0)
                rc = mvster_deconv_tc3_pb16(in[i], wl + mvster_deconv_tc3_packed_bytes(L[i].cin, L[i].cout, 0), nullptr, bias, skip[i], out[i],
                                            B, D, h, w, L[i].cin, L[i].cout, 1, pf, stream);
        }
        if (rc != MVSTER_OK) return rc;
    }
    return MVSTER_OK;
}

extern "C" int mvster_reg2d_tc3_ex_f32(const float* blob, const void* tc3_blob, const float* cost, float* feat8, float* ws,
                                       int B, int G, int D, int H, int W, int flags, mvster_stream_t stream) {
    MVSTER_REQUIRE(tc3_blob, "mvster_reg2d_tc3_f32: tc3_blob is null");
    // packed form: needs the four-voxel conv0 kernel (G in {4, 8}, 32-bit quad index); anything else takes the unpacked chain
    if ((flags & MVSTER_TC3_FP16X2) && (flags & MVSTER_REG2D_PACKED) && (G == 4 || G == 8) && W % 4 == 0 &&
        (long long)B * D * H * (W / 4) < (1ll << 31))
        return reg2d_packed_h16(blob, tc3_blob, cost, feat8, ws, B, G, D, H, W, stream);
    // npass carries the arithmetic of the generation-3 layers: 3 = three bf16 terms, 2 = two fp16 terms
    return reg2d_run(blob, (const float*)tc3_blob, (flags & MVSTER_TC3_FP16X2) ? 2 : 3, 3, cost, feat8, ws, B, G, D, H, W, stream);
}

extern "C" int mvster_reg2d_tc3_f32(const float* blob, const void* tc3_blob, const float* cost, float* feat8, float* ws,
                                    int B, int G, int D, int H, int W, mvster_stream_t stream) {
    return mvster_reg2d_tc3_ex_f32(blob, tc3_blob, cost, feat8, ws, B, G, D, H, W, 0, stream);
}

extern "C" int mvster_reg2d_bf16(const float* blob_q, const void* tc3_blob, const float* scales, const void* cost_bf16, float* feat8,
                                 float* ws, int B, int G, int D, int H, int W, int flags, mvster_stream_t stream) {
    MVSTER_REQUIRE(blob_q && tc3_blob && scales && cost_bf16 && feat8 && ws, "mvster_reg2d_bf16: null pointer");
    MVSTER_REQUIRE(G == 4 || G == 8, "mvster_reg2d_bf16: unsupported G=%d (4 or 8)", G);
    MVSTER_REQUIRE(H % 8 == 0 && W % 8 == 0, "mvster_reg2d_bf16: H,W must be multiples of 8 (got %dx%d)", H, W);
    cudaStream_t st = (cudaStream_t)stream;
    Layer L[MVSTER_REG2D_LAYERS];
    reg2d_layers(G, L);
    const size_t N = (size_t)B * D * H * W;
    // same workspace map as the fp32 network (offsets in floats); packed bf16 activations use the first half of each slot
    float* c0 = ws;           float* c1 = c0 + 8 * N;  float* c2 = c1 + 4 * N;  float* c3 = c2 + 4 * N;
    float* c4 = c3 + 2 * N;   float* c5 = c4 + 2 * N;  float* c6 = c5 + N;      float* u7 = c6 + N;
    float* u9 = u7 + 2 * N;
    const float* in[MVSTER_REG2D_LAYERS] = {nullptr, c0, c1, c2, c3, c4, c5, c6, u7, u9};
    float* out[MVSTER_REG2D_LAYERS] = {c0, c1, c2, c3, c4, c5, c6, u7, u9, feat8};
    const float* skip[MVSTER_REG2D_LAYERS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, c4, c2, c0};
    const int div[MVSTER_REG2D_LAYERS] = {1, 1, 2, 2, 4, 4, 8, 8, 4, 2};  // input resolution divisor
    const bool packed = flags & MVSTER_REG2D_BF16_PACKED;
    const float* sc = scales;
    for (int i = 0; i < MVSTER_REG2D_LAYERS; sc += L[i].cout, ++i) {
        int64_t info[8];
        mvster_reg2d_layer_info(G, i, info);
        const uint8_t* wl = (const uint8_t*)tc3_blob + tc3_layer_offset(L, i);
        const float* bias = blob_q + info[6];
        const bool last = i == MVSTER_REG2D_LAYERS - 1;  // conv11: its sum with conv0 stays fp32 (the head's `prob` rounds its own operand)
        const int h = H / div[i], w = W / div[i];
        int rc;
        if (i == 0) {  // G -> 8 at full resolution: CUDA cores (K and N would be mostly padding on the tensor core)
            rc = conv0_bf16(cost_bf16, blob_q + info[5], sc, bias, c0, packed ? 1 : 0, (long long)B * D, H, W, G, st);
            MVSTER_REQUIRE(rc != -100, "mvster_reg2d_bf16: conv0 shape not covered (W %% 4, G in {4,8})");
        } else if (packed) {
            const int pf = 1 | (last ? 0 : MVSTER_TC3_OUT_PB16);
            if (!L[i].transposed) {
                rc = mvster_conv_tc3_pb16(in[i], wl, sc, bias, skip[i], out[i], B, D, h, w, L[i].cin, L[i].cout, L[i].kd, 3, L[i].s, pf, stream);
            } else {
                const int rows = tc3_deconv_rows(L[i]);
                rc = mvster_deconv_tc3_pb16(in[i], wl, sc, bias, skip[i], out[i], B, D, h, w, L[i].cin, L[i].cout, rows, pf, stream);
                if (rc == MVSTER_OK && rows == 0)
                    rc = mvster_deconv_tc3_pb16(in[i], wl + mvster_deconv_tc3_packed_bytes(L[i].cin, L[i].cout, 0), sc, bias, skip[i], out[i],
                                                B, D, h, w, L[i].cin, L[i].cout, 1, pf, stream);
            }
        } else {
            const int rf = 1 | MVSTER_TC3_BF16X1 | (last ? 0 : MVSTER_TC3_ROUND_OUT);
            if (!L[i].transposed) {
                rc = mvster_conv_tc3_scaled_f32(in[i], wl, sc, bias, skip[i], out[i], B, D, h, w, L[i].cin, L[i].cout, L[i].kd, 3, L[i].s, rf, stream);
            } else {
                const int rows = tc3_deconv_rows(L[i]);
                rc = mvster_deconv_tc3_scaled_f32(in[i], wl, sc, bias, skip[i], out[i], B, D, h, w, L[i].cin, L[i].cout, rows, rf, stream);
                if (rc == MVSTER_OK && rows == 0)
                    rc = mvster_deconv_tc3_scaled_f32(in[i], wl + mvster_deconv_tc3_packed_bytes(L[i].cin, L[i].cout, 0), sc, bias, skip[i], out[i],
                                                      B, D, h, w, L[i].cin, L[i].cout, 1, rf, stream);
            }
        }
        if (rc != MVSTER_OK) return rc;
    }
    return MVSTER_OK;
}

static int reg2d_run(const float* blob, const float* tc_blob, int npass, int gen, const float* cost, float* feat8, float* ws,
                     int B, int G, int D, int H, int W, mvster_stream_t stream) {
    MVSTER_REQUIRE(blob && cost && feat8 && ws, "mvster_reg2d_f32: null pointer");
    MVSTER_REQUIRE(G == 4 || G == 8 || G == 16 || G == 32 || G == 64, "mvster_reg2d_f32: unsupported G=%d", G);
    MVSTER_REQUIRE(H % 8 == 0 && W % 8 == 0, "mvster_reg2d_f32: H,W must be multiples of 8 (got %dx%d)", H, W);
    cudaStream_t st = (cudaStream_t)stream;
    Layer L[MVSTER_REG2D_LAYERS];
    reg2d_layers(G, L);
    const size_t N = (size_t)B * D * H * W;
    float* c0 = ws;           float* c1 = c0 + 8 * N;  float* c2 = c1 + 4 * N;  float* c3 = c2 + 4 * N;
    float* c4 = c3 + 2 * N;   float* c5 = c4 + 2 * N;  float* c6 = c5 + N;      float* u7 = c6 + N;
    float* u9 = u7 + 2 * N;
    const float* in[MVSTER_REG2D_LAYERS] = {cost, c0, c1, c2, c3, c4, c5, c6, u7, u9};
    float* out[MVSTER_REG2D_LAYERS] = {c0, c1, c2, c3, c4, c5, c6, u7, u9, feat8};
    const float* skip[MVSTER_REG2D_LAYERS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, c4, c2, c0};
    const int div[MVSTER_REG2D_LAYERS] = {1, 1, 2, 2, 4, 4, 8, 8, 4, 2};  // input resolution divisor
    const int relu3 = 1 | ((gen == 3 && npass == 2) ? MVSTER_TC3_FP16X2 : 0);  // ReLU + arithmetic of the generation-3 layers
    for (int i = 0; i < MVSTER_REG2D_LAYERS; ++i) {
        int64_t info[8];
        mvster_reg2d_layer_info(G, i, info);
        int rc;
        if (tc_blob && gen == 3 && !L[i].transposed && i > 0) {
            // conv1..conv6 (stride 1 and stride (1,2,2)) on the persistent 3 x bf16 tcgen05 kernel.  conv0 (G -> 8 channels at
            // full resolution) stays on the CUDA cores: with K and N mostly padding the tensor-core tile costs as much as a
            // 16 -> 16 layer (measured 101 us vs 42 us at cfg2 stage 4).
            rc = mvster_conv_tc3_f32(in[i], (const uint8_t*)tc_blob + tc3_layer_offset(L, i), blob + info[6], skip[i], out[i], B, D,
                                     H / div[i], W / div[i], L[i].cin, L[i].cout, L[i].kd, 3, L[i].s, relu3, stream);
        } else if (tc_blob && gen == 3 && L[i].transposed) {
            // conv7 / conv9 / conv11: 2x2 convolution on the input grid + depth-to-space epilogue (mvster_deconv_tc3_f32)
            const uint8_t* wl = (const uint8_t*)tc_blob + tc3_layer_offset(L, i);
            const int rows = tc3_deconv_rows(L[i]);
            rc = mvster_deconv_tc3_f32(in[i], wl, blob + info[6], skip[i], out[i], B, D, H / div[i], W / div[i], L[i].cin, L[i].cout,
                                       rows, relu3, stream);
            if (rc == MVSTER_OK && rows == 0)
                rc = mvster_deconv_tc3_f32(in[i], wl + mvster_deconv_tc3_packed_bytes(L[i].cin, L[i].cout, 0), blob + info[6], skip[i],
                                           out[i], B, D, H / div[i], W / div[i], L[i].cin, L[i].cout, 1, relu3, stream);
        } else if (tc_blob && gen != 3 && L[i].kd == 3) {
            // conv2 / conv4 / conv6 (3x3x3, 69 % of the FLOPs) on the tcgen05 tensor cores; their [hi|lo]
            // K-major slabs sit back to back in tc_blob (2*27*Cin*Cout floats each).
            const size_t off = i == 2 ? 0 : (i == 4 ? (size_t)2 * 27 * 16 * 16 : (size_t)2 * 27 * (16 * 16 + 32 * 32));
            rc = mvster_conv3d_tc2_f32(in[i], tc_blob + off, blob + info[6], skip[i], out[i], B, D, H / div[i], W / div[i], L[i].cin, L[i].cout,
                                       3, 1, npass, stream);
        } else {
            rc = run_conv(in[i], blob + info[5], blob + info[6], skip[i], out[i], B, D, H / div[i], W / div[i],
                          L[i].cin, L[i].cout, L[i].kd, 1, L[i].s, L[i].transposed, 1, st);
        }
        if (rc != MVSTER_OK) return rc;
    }
    return MVSTER_OK;
}

// ================================================================================================
// reg3d (mvs4net_utils.py:914-965): full 3-D U-Net - 3x3x3 kernels, stride 2 along D as well,
// depth `down_size` in {1,2,3} (MVS4Net.py:48: 3,3,2,2 per stage), prob = 3x3x3 conv 8 -> 1 without bias.
// ================================================================================================
namespace mvster {

// ConvTranspose3d(3, stride=2, padding=1, output_padding=1): out[o] += in[i] * w[k] with o = 2i - 1 + k.
// Gather form: one thread per output voxel x COUT_T channels; only taps with matching parity contribute
// (27/8 on average).  reg3d is an optional path of the reference, so simplicity wins over speed here.
template <int CIN, int COUT_T>
__global__ void __launch_bounds__(128) deconv3d_gather_kernel(const ConvArgs a) {
    extern __shared__ __align__(16) float w_s[];  // [27][CIN][COUT_T]
    const int cg = blockIdx.y;
    for (int i = threadIdx.x; i < 27 * CIN * COUT_T; i += blockDim.x) {
        const int o = i % COUT_T, rest = i / COUT_T;
        w_s[i] = __ldg(a.w + (long long)rest * a.cout + cg * COUT_T + o);
    }
    __syncthreads();
    const long long nvox = (long long)a.B * a.Do * a.Ho * a.Wo;
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    const int ox = (int)(v % a.Wo), oy = (int)((v / a.Wo) % a.Ho);
    const int oz = (int)((v / ((long long)a.Wo * a.Ho)) % a.Do), b = (int)(v / ((long long)a.Wo * a.Ho * a.Do));
    float acc[COUT_T];
#pragma unroll
    for (int o = 0; o < COUT_T; ++o) acc[o] = a.bias ? __ldg(a.bias + cg * COUT_T + o) : 0.f;
    for (int kz = 0; kz < 3; ++kz) {
        const int tz = oz + 1 - kz;
        if (tz < 0 || (tz & 1) || (tz >> 1) >= a.Di) continue;
        for (int ky = 0; ky < 3; ++ky) {
            const int ty = oy + 1 - ky;
            if (ty < 0 || (ty & 1) || (ty >> 1) >= a.Hi) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int tx = ox + 1 - kx;
                if (tx < 0 || (tx & 1) || (tx >> 1) >= a.Wi) continue;
                const float4* px = reinterpret_cast<const float4*>(
                    a.x + ((((long long)b * a.Di + (tz >> 1)) * a.Hi + (ty >> 1)) * a.Wi + (tx >> 1)) * CIN);
                const float* wt = w_s + ((kz * 3 + ky) * 3 + kx) * CIN * COUT_T;
#pragma unroll
                for (int c4 = 0; c4 < CIN / 4; ++c4) {
                    const float4 t = __ldg(px + c4);
                    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float* wr = wt + (c4 * 4 + j) * COUT_T;
#pragma unroll
                        for (int o = 0; o < COUT_T; ++o) acc[o] = fmaf(tv[j], wr[o], acc[o]);
                    }
                }
            }
        }
    }
    const long long off = v * a.cout + cg * COUT_T;
#pragma unroll
    for (int o = 0; o < COUT_T; ++o) {
        float r = acc[o];
        if (a.relu) r = fmaxf(r, 0.f);
        if (a.skip) r += __ldg(a.skip + off + o);
        a.y[off + o] = r;
    }
}

template <int CIN>
static int launch_deconv3d(const ConvArgs& a, cudaStream_t st) {
    constexpr int CT = 8;
    const size_t smem = (size_t)27 * CIN * CT * sizeof(float);
    auto k = deconv3d_gather_kernel<CIN, CT>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long n = (long long)a.B * a.Do * a.Ho * a.Wo;
    k<<<dim3(ceil_div(n, 128), a.cout / CT), 128, smem, st>>>(a);
    return check_launch("deconv3d_gather_kernel");
}

static int run_deconv3d(const float* x, const float* w, const float* bias, const float* skip, float* y,
                        int B, int Di, int Hi, int Wi, int Cin, int Cout, cudaStream_t st) {
    MVSTER_REQUIRE(Cout % 8 == 0, "reg3d transposed conv needs Cout %% 8 == 0 (got %d)", Cout);
    ConvArgs a;
    a.x = x; a.w = w; a.bias = bias; a.skip = skip; a.y = y;
    a.B = B; a.Di = Di; a.Hi = Hi; a.Wi = Wi; a.Do = 2 * Di; a.Ho = 2 * Hi; a.Wo = 2 * Wi;
    a.cout = Cout; a.kd = 3; a.sd = 2; a.s = 2; a.relu = 1;
    switch (Cin) {
        case 16: return launch_deconv3d<16>(a, st);
        case 32: return launch_deconv3d<32>(a, st);
        case 64: return launch_deconv3d<64>(a, st);
    }
    set_error("reg3d transposed conv: unsupported Cin=%d", Cin);
    return MVSTER_ERR_ARG;
}

// layer order in the reg3d blob: only the layers that exist for `down_size`, in this order
struct Layer3 { const char* name; int cin, cout, stride, transposed, level_in; };
static int reg3d_layers(int G, int down, Layer3 (&L)[12]) {
    int n = 0;
    L[n++] = {"conv0", G, 8, 1, 0, 0};
    L[n++] = {"conv1", 8, 16, 2, 0, 0};
    L[n++] = {"conv2", 16, 16, 1, 0, 1};
    if (down >= 2) { L[n++] = {"conv3", 16, 32, 2, 0, 1}; L[n++] = {"conv4", 32, 32, 1, 0, 2}; }
    if (down >= 3) { L[n++] = {"conv5", 32, 64, 2, 0, 2}; L[n++] = {"conv6", 64, 64, 1, 0, 3}; L[n++] = {"conv7", 64, 32, 2, 1, 3}; }
    if (down >= 2) L[n++] = {"conv9", 32, 16, 2, 1, 2};
    L[n++] = {"conv11", 16, 8, 2, 1, 1};
    L[n++] = {"prob", 8, 1, 1, 0, 0};
    return n;
}

}  // namespace mvster

extern "C" int mvster_reg3d_num_layers(int down_size) {
    mvster::Layer3 L[12];
    return (down_size >= 1 && down_size <= 3) ? mvster::reg3d_layers(8, down_size, L) : 0;
}

/* info[8] = {Cin, Cout, stride, transposed, w_offset, bias_offset (== -1 for prob: no bias), n_taps (27), name id
 * (0,1,2,3,4,5,6,7,9,11 for conv*, 100 for prob)} */
extern "C" int mvster_reg3d_layer_info(int G, int down_size, int layer, int64_t* info) {
    mvster::Layer3 L[12];
    MVSTER_REQUIRE(info && down_size >= 1 && down_size <= 3, "mvster_reg3d_layer_info: bad arguments");
    const int n = mvster::reg3d_layers(G, down_size, L);
    MVSTER_REQUIRE(layer >= 0 && layer < n, "mvster_reg3d_layer_info: bad layer %d", layer);
    int64_t off = 0;
    for (int i = 0; i <= layer; ++i) {
        const int64_t nw = 27ll * L[i].cin * L[i].cout;
        const bool has_bias = i != n - 1;
        if (i == layer) {
            info[0] = L[i].cin; info[1] = L[i].cout; info[2] = L[i].stride; info[3] = L[i].transposed;
            info[4] = off; info[5] = has_bias ? off + nw : -1; info[6] = 27;
            info[7] = i == n - 1 ? 100 : atoi(L[i].name + 4);
        }
        off += nw + (has_bias ? L[i].cout : 0);
    }
    return MVSTER_OK;
}

extern "C" size_t mvster_reg3d_blob_floats(int G, int down_size) {
    int64_t info[8];
    const int n = mvster_reg3d_num_layers(down_size);
    if (n == 0 || mvster_reg3d_layer_info(G, down_size, n - 1, info) != MVSTER_OK) return 0;
    return (size_t)(info[4] + 27 * info[0] * info[1]);
}

extern "C" size_t mvster_reg3d_workspace_floats(int B, int D, int H, int W) { return (size_t)26 * B * D * H * W; }

extern "C" int mvster_reg3d_f32(const float* blob, const float* cost, float* logits, float* ws,
                                int B, int G, int D, int H, int W, int down_size, mvster_stream_t stream) {
    using namespace mvster;
    MVSTER_REQUIRE(blob && cost && logits && ws, "mvster_reg3d_f32: null pointer");
    MVSTER_REQUIRE(down_size >= 1 && down_size <= 3, "mvster_reg3d_f32: down_size=%d (1..3)", down_size);
    const int m = 1 << down_size;
    MVSTER_REQUIRE(D % m == 0 && H % m == 0 && W % m == 0, "mvster_reg3d_f32: D,H,W must be multiples of %d", m);
    MVSTER_REQUIRE(G == 4 || G == 8 || G == 16 || G == 32 || G == 64, "mvster_reg3d_f32: unsupported G=%d", G);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = (size_t)B * D * H * W;
    // activations: c0 8N | c1 2N | c2 2N | c3 .5N | c4 .5N | c5 .125N | c6 .125N | u7 .5N | u9 2N | u11 8N  (< 26N)
    float* c0 = ws;            float* c1 = c0 + 8 * N;       float* c2 = c1 + 2 * N;      float* c3 = c2 + 2 * N;
    float* c4 = c3 + N / 2;    float* c5 = c4 + N / 2;       float* c6 = c5 + N / 8;      float* u7 = c6 + N / 8;
    float* u9 = u7 + N / 2;    float* u11 = u9 + 2 * N;
    Layer3 L[12];
    const int n = reg3d_layers(G, down_size, L);
    auto W_ = [&](int i, int64_t (&info)[8]) { mvster_reg3d_layer_info(G, down_size, i, info); };
    int rc = MVSTER_OK, li = 0;
    int64_t info[8];
    auto conv = [&](const float* x, float* y, int lvl, int cin, int cout, int stride, int relu) {
        W_(li++, info);
        const int d = D >> lvl, h = H >> lvl, w = W >> lvl;
        return run_conv(x, blob + info[4], info[5] >= 0 ? blob + info[5] : nullptr, nullptr, y, B, d, h, w, cin, cout, 3, stride, stride, 0, relu, st);
    };
    auto up = [&](const float* x, const float* skip, float* y, int lvl, int cin, int cout) {
        W_(li++, info);
        return run_deconv3d(x, blob + info[4], blob + info[5], skip, y, B, D >> lvl, H >> lvl, W >> lvl, cin, cout, st);
    };
    if ((rc = conv(cost, c0, 0, G, 8, 1, 1))) return rc;
    if ((rc = conv(c0, c1, 0, 8, 16, 2, 1))) return rc;
    if ((rc = conv(c1, c2, 1, 16, 16, 1, 1))) return rc;
    const float* x = c2;
    if (down_size >= 2) {
        if ((rc = conv(c2, c3, 1, 16, 32, 2, 1))) return rc;
        if ((rc = conv(c3, c4, 2, 32, 32, 1, 1))) return rc;
        x = c4;
        if (down_size >= 3) {
            if ((rc = conv(c4, c5, 2, 32, 64, 2, 1))) return rc;
            if ((rc = conv(c5, c6, 3, 64, 64, 1, 1))) return rc;
            if ((rc = up(c6, c4, u7, 3, 64, 32))) return rc;   // x = conv4 + conv7(x)
            x = u7;
        }
        if ((rc = up(x, c2, u9, 2, 32, 16))) return rc;        // x = conv2 + conv9(x)
        x = u9;
    }
    if ((rc = up(x, c0, u11, 1, 16, 8))) return rc;            // x = conv0 + conv11(x)
    if ((rc = conv(u11, logits, 0, 8, 1, 1, 0))) return rc;    // prob: 3x3x3, no bias, no ReLU
    (void)n;
    return MVSTER_OK;
}
