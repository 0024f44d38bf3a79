// Host stand-ins for the tensor-core entry points that the SIMT sources and the Python drivers reference.
//
// The tcgen05 / TMA kernels themselves cannot run on the host.  What CAN be checked without a GPU is everything around them: the
// weight packers (slab order, operand terms, two-taps-per-MMA pairing, parity classes of the transposed form), the blob offsets and
// the layer chaining of the reg2d / feature-pyramid drivers, the per-channel scale of the bf16-storage arithmetic.  So the
// generation-3 entry points are implemented here as: decode the PACKED weight slabs with the library's own layer plan
// (conv_tc3_plan.h, the header conv_tc3.cu is built from) and evaluate the layer by a direct convolution on host pointers, with
// the operand rounding of the requested arithmetic.  Test infrastructure only (tests/emu); the generation-2 entry points stay
// unavailable.
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include "conv_tc3_plan.h"

typedef void* mvster_stream_t;
#define MVSTER_TC3_FP16X2 256
#define MVSTER_TC3_BF16X1 512
#define MVSTER_TC3_OUT_PB16 1024
#define MVSTER_TC3_ROUND_OUT 2048

namespace {

float bf16_to_float(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
float half_to_float(uint16_t h) {
    const int s = h >> 15, e = (h >> 10) & 31, m = h & 1023;
    float v;
    if (e == 0) v = std::ldexp((float)m, -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = std::ldexp((float)(m | 1024), e - 25);
    return s ? -v : v;
}
float round_bf16(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return x;
    u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
    std::memcpy(&x, &u, 4);
    return x;
}
int terms_of(int relu) { return (relu & MVSTER_TC3_BF16X1) ? 1 : (relu & MVSTER_TC3_FP16X2) ? 2 : 3; }

// slab i -> [16 K rows][ncol] fp32: the value the MMAs of the given arithmetic multiply with (sum of the operand terms)
std::vector<float> decode_slabs(const void* w_packed, int nslab, int ncol, int terms) {
    const uint16_t* raw = static_cast<const uint16_t*>(w_packed);  // [slab][2 K-halves][3 * ncol rows][8]
    std::vector<float> out((size_t)nslab * 16 * ncol);
    for (int i = 0; i < nslab; ++i)
        for (int k = 0; k < 16; ++k)
            for (int n = 0; n < ncol; ++n) {
                float v = 0.f;
                for (int j = 0; j < terms; ++j) {
                    const uint16_t h = raw[(((size_t)i * 2 + k / 8) * 3 * ncol + (size_t)j * ncol + n) * 8 + k % 8];
                    v += terms == 2 ? half_to_float(h) * (j ? 1.f / 2048.f : 1.f) : bf16_to_float(h);
                }
                out[((size_t)i * 16 + k) * ncol + n] = v;
            }
    return out;
}

float finish(double acc, int ch, const float* scale, const float* bias, const float* skip, long long off, int relu) {
    float o = (float)acc;
    if (scale) o *= scale[ch];
    if (bias) o += bias[ch];
    if (relu & 1) o = o > 0.f ? o : 0.f;
    if (skip) o += skip[off];
    if ((relu & MVSTER_TC3_BF16X1) && (relu & MVSTER_TC3_ROUND_OUT)) o = round_bf16(o);
    return o;
}

int conv_run(const float* x, const void* w_packed, const float* scale, const float* bias, const float* skip, float* y,
             int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int s, int relu, int block, long long block_stride) {
    using namespace mvster::tc3;
    if (!x || !w_packed || !y || !supported(Cin, Cout, kd, k, s)) return -1;
    const int terms = terms_of(relu), ncol = Cout < 16 ? 16 : (Cout > 64 ? 80 : Cout);
    int slabs[MAX_STAGES * MAX_TAPS][6];
    const int nslab = build_plan(Cin, kd, k, s, nullptr, slabs);
    const std::vector<float> mats = decode_slabs(w_packed, nslab, ncol, terms);
    // dense weights [kz][ky][kx][Cin][Cout] from the slabs
    std::vector<float> wd((size_t)kd * k * k * Cin * Cout, 0.f);
    auto put = [&](int kz, int ky, int kx, int c, int i, int row) {
        for (int n = 0; n < Cout; ++n) wd[((((size_t)kz * k + ky) * k + kx) * Cin + c) * Cout + n] = mats[((size_t)i * 16 + row) * ncol + n];
    };
    for (int i = 0; i < nslab; ++i) {
        const int kz = slabs[i][0], ky = slabs[i][1], kx = slabs[i][2], c0 = slabs[i][3], ky2 = slabs[i][4], kx2 = slabs[i][5];
        if (Cin <= 8) {
            for (int c = 0; c < Cin; ++c) put(kz, ky, kx, c, i, c);
            if (ky2 >= 0)
                for (int c = 0; c < Cin; ++c) put(kz, ky2, kx2, c, i, 8 + c);
        } else {
            for (int c = 0; c < 16 && c0 + c < Cin; ++c) put(kz, ky, kx, c0 + c, i, c);
        }
    }
    const int Ho = (H - 1) / s + 1, Wo = (W - 1) / s + 1, pad = k / 2, pz = kd / 2;
    const bool blocks = block > 0 && block < Cout;
    std::vector<double> acc(Cout);
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            for (int yo = 0; yo < Ho; ++yo)
                for (int xo = 0; xo < Wo; ++xo) {
                    std::fill(acc.begin(), acc.end(), 0.0);
                    for (int kz = 0; kz < kd; ++kz) {
                        const int dz = d + kz - pz;
                        if (dz < 0 || dz >= D) continue;
                        for (int ky = 0; ky < k; ++ky) {
                            const int yi = yo * s + ky - pad;
                            if (yi < 0 || yi >= H) continue;
                            for (int kx = 0; kx < k; ++kx) {
                                const int xi = xo * s + kx - pad;
                                if (xi < 0 || xi >= W) continue;
                                const float* px = x + ((((size_t)b * D + dz) * H + yi) * W + xi) * Cin;
                                const float* wt = wd.data() + (((size_t)kz * k + ky) * k + kx) * Cin * Cout;
                                for (int c = 0; c < Cin; ++c) {
                                    const double v = terms == 1 ? round_bf16(px[c]) : px[c];
                                    for (int n = 0; n < Cout; ++n) acc[n] += v * wt[(size_t)c * Cout + n];
                                }
                            }
                        }
                    }
                    const long long pix = (((long long)b * D + d) * Ho + yo) * Wo + xo;
                    for (int n = 0; n < Cout; ++n) {
                        if (blocks) {
                            const long long off = (long long)(n / block) * block_stride + pix * block + n % block;
                            y[off] = finish(acc[n], n, scale, bias, nullptr, 0, relu);
                        } else {
                            const long long off = pix * Cout + n;
                            y[off] = finish(acc[n], n, scale, bias, skip, off, relu);
                        }
                    }
                }
    return 0;
}

int deconv_run(const float* x, const void* w_packed, const float* scale, const float* bias, const float* skip, float* y,
               int B, int D, int H, int W, int Cin, int Cout, int rows, int relu) {
    using namespace mvster::tc3;
    if (!x || !w_packed || !y || Cin % 16 || rows < -1 || rows > 1) return -1;
    const int terms = terms_of(relu), ncls = deconv_ncls(rows), ntap = deconv_ntap(rows), ncol = ncls * Cout, kch = Cin / 16;
    const std::vector<float> mats = decode_slabs(w_packed, kch * ntap, ncol, terms);  // [chunk][tap (dy,dx)][16][class * Cout + co]
    std::vector<double> acc(ncol);
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            for (int yy = 0; yy < H; ++yy)
                for (int xx = 0; xx < W; ++xx) {
                    std::fill(acc.begin(), acc.end(), 0.0);
                    for (int kc = 0; kc < kch; ++kc)
                        for (int t = 0; t < ntap; ++t) {
                            const int yi = yy + t / 2, xi = xx + t % 2;
                            if (yi >= H || xi >= W) continue;
                            const float* px = x + ((((size_t)b * D + d) * H + yi) * W + xi) * Cin + kc * 16;
                            const float* m = mats.data() + ((size_t)kc * ntap + t) * 16 * ncol;
                            for (int c = 0; c < 16; ++c) {
                                const double v = terms == 1 ? round_bf16(px[c]) : px[c];
                                for (int n = 0; n < ncol; ++n) acc[n] += v * m[(size_t)c * ncol + n];
                            }
                        }
                    for (int cls = 0; cls < ncls; ++cls) {
                        const int py = rows < 0 ? cls / 2 : rows, pxc = rows < 0 ? cls % 2 : cls;
                        const long long pix = (((long long)b * D + d) * (2 * H) + 2 * yy + py) * (2 * W) + 2 * xx + pxc;
                        for (int n = 0; n < Cout; ++n) {
                            const long long off = pix * Cout + n;
                            y[off] = finish(acc[cls * Cout + n], n, scale, bias, skip, off, relu);
                        }
                    }
                }
    return 0;
}

unsigned f2h_sat(float f) {  // fp32 -> fp16 bits, round to nearest even, saturating (cvt.rn.satfinite.f16.f32)
    unsigned u;
    std::memcpy(&u, &f, 4);
    const unsigned sign = (u >> 16) & 0x8000u;
    u &= 0x7FFFFFFFu;
    if (u > 0x7F800000u) return sign | 0x7FFFu;
    if (u >= 0x477FF000u) return sign | 0x7BFFu;
    if (u < 0x33000001u) return sign;
    const int e = (int)(u >> 23) - 127;
    unsigned m = (u & 0x7FFFFFu) | 0x800000u;
    const int shift = e >= -14 ? 13 : 13 + (-14 - e);
    const unsigned half = 1u << (shift - 1), rest = m & ((1u << shift) - 1);
    m >>= shift;
    if (rest > half || (rest == half && (m & 1u))) ++m;
    const unsigned base = e >= -14 ? ((unsigned)(e + 15) << 10) - 0x400u : 0u;
    return sign | (base + m);
}

// octet-planar 16-bit terms [P][C/8][NT][H][W][8]  <->  fp32 [P][H][W][C]; NT = 1: bf16, NT = 2: fp16 pair a1 + 2^-11 a2
std::vector<float> unpack_pb16(const void* x, long long P, int H, int W, int C, int NT = 1) {
    const uint16_t* p = static_cast<const uint16_t*>(x);
    std::vector<float> out((size_t)P * H * W * C);
    const size_t plane = (size_t)H * W * 8;
    for (long long pl = 0; pl < P; ++pl)
        for (int o = 0; o < C / 8; ++o)
            for (long long px = 0; px < (long long)H * W; ++px)
                for (int e = 0; e < 8; ++e) {
                    const size_t i = ((size_t)pl * (C / 8) + o) * NT * plane + (size_t)px * 8 + e;
                    out[((size_t)pl * H * W + px) * C + o * 8 + e] =
                        NT == 2 ? half_to_float(p[i]) + half_to_float(p[i + plane]) * (1.f / 2048.f) : bf16_to_float(p[i]);
                }
    return out;
}
void pack_pb16(const std::vector<float>& v, void* y, long long P, int H, int W, int C, int NT = 1) {
    uint16_t* p = static_cast<uint16_t*>(y);
    const size_t plane = (size_t)H * W * 8;
    for (long long pl = 0; pl < P; ++pl)
        for (int o = 0; o < C / 8; ++o)
            for (long long px = 0; px < (long long)H * W; ++px)
                for (int e = 0; e < 8; ++e) {
                    const float val = v[((size_t)pl * H * W + px) * C + o * 8 + e];
                    const size_t i = ((size_t)pl * (C / 8) + o) * NT * plane + (size_t)px * 8 + e;
                    if (NT == 2) {
                        const unsigned h1 = f2h_sat(val);
                        p[i] = (uint16_t)h1;
                        p[i + plane] = (uint16_t)f2h_sat((val - half_to_float((uint16_t)h1)) * 2048.f);
                    } else {
                        const float f = round_bf16(val);
                        uint32_t u;
                        std::memcpy(&u, &f, 4);
                        p[i] = (uint16_t)(u >> 16);
                    }
                }
}

}  // namespace

extern "C" {
int mvster_conv3d_tc_f32(...) { return -2; }
int mvster_conv3d_tc2_f32(...) { return -2; }
int mvster_pointwise_tc2_f32(...) { return -2; }
void mvster_set_sm_budget(int) {}
static thread_local unsigned* g_flag = nullptr;
void mvster_tc3_set_overflow_flag(unsigned* p) { g_flag = p; }
unsigned* mvster_tc3_overflow_flag(void) { return g_flag; }

int mvster_conv_tc3_scaled_f32(const float* x, const void* w_packed, const float* scale, const float* bias, const float* skip,
                               float* y, int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int stride_hw, int relu,
                               mvster_stream_t) {
    return conv_run(x, w_packed, scale, bias, skip, y, B, D, H, W, Cin, Cout, kd, k, stride_hw, relu, 0, 0);
}
int mvster_conv_tc3_f32(const float* x, const void* w_packed, const float* bias, const float* skip, float* y,
                        int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int stride_hw, int relu, mvster_stream_t st) {
    return mvster_conv_tc3_scaled_f32(x, w_packed, nullptr, bias, skip, y, B, D, H, W, Cin, Cout, kd, k, stride_hw, relu, st);
}
int mvster_pointwise_tc3_blocks_ex_f32(const float* x, const void* w_packed, float* y, int N, int H, int W, int Cin, int Cout,
                                       int block, long long block_stride_floats, int flags, mvster_stream_t) {
    return conv_run(x, w_packed, nullptr, nullptr, nullptr, y, N, 1, H, W, Cin, Cout, 1, 1, 1, flags & MVSTER_TC3_FP16X2, block,
                    block_stride_floats);
}
int mvster_pointwise_tc3_blocks_f32(const float* x, const void* w_packed, float* y, int N, int H, int W, int Cin, int Cout,
                                    int block, long long block_stride_floats, mvster_stream_t st) {
    return mvster_pointwise_tc3_blocks_ex_f32(x, w_packed, y, N, H, W, Cin, Cout, block, block_stride_floats, 0, st);
}
int mvster_deconv_tc3_scaled_f32(const float* x, const void* w_packed, const float* scale, const float* bias, const float* skip,
                                 float* y, int B, int D, int H, int W, int Cin, int Cout, int rows, int relu, mvster_stream_t) {
    return deconv_run(x, w_packed, scale, bias, skip, y, B, D, H, W, Cin, Cout, rows, relu);
}
int mvster_deconv_tc3_f32(const float* x, const void* w_packed, const float* bias, const float* skip, float* y,
                          int B, int D, int H, int W, int Cin, int Cout, int rows, int relu, mvster_stream_t st) {
    return mvster_deconv_tc3_scaled_f32(x, w_packed, nullptr, bias, skip, y, B, D, H, W, Cin, Cout, rows, relu, st);
}
// packed operands: unpack -> the same direct evaluation (one bf16 term) -> pack (or fp32 out)
int mvster_conv_tc3_pb16(const void* x, const void* w_packed, const float* scale, const float* bias, const void* skip, void* y,
                         int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int s, int flags, mvster_stream_t) {
    if (!x || !y || Cin < 8) return -1;
    const int Ho = (H - 1) / s + 1, Wo = (W - 1) / s + 1, NT = (flags & MVSTER_TC3_FP16X2) ? 2 : 1;
    const long long P = (long long)B * D;
    const std::vector<float> xf = unpack_pb16(x, P, H, W, Cin, NT);
    std::vector<float> sf, yf((size_t)P * Ho * Wo * Cout);
    if (skip) sf = unpack_pb16(skip, P, Ho, Wo, Cout, NT);
    const int rc = conv_run(xf.data(), w_packed, scale, bias, skip ? sf.data() : nullptr, yf.data(), B, D, H, W, Cin, Cout, kd, k, s,
                            (flags & 1) | (NT == 2 ? MVSTER_TC3_FP16X2 : MVSTER_TC3_BF16X1), 0, 0);
    if (rc) return rc;
    if (flags & MVSTER_TC3_OUT_PB16) pack_pb16(yf, y, P, Ho, Wo, Cout, NT);
    else std::memcpy(y, yf.data(), yf.size() * 4);
    return 0;
}
int mvster_deconv_tc3_pb16(const void* x, const void* w_packed, const float* scale, const float* bias, const void* skip, void* y,
                           int B, int D, int H, int W, int Cin, int Cout, int rows, int flags, mvster_stream_t) {
    if (!x || !y) return -1;
    const long long P = (long long)B * D;
    const int NT = (flags & MVSTER_TC3_FP16X2) ? 2 : 1;
    const std::vector<float> xf = unpack_pb16(x, P, H, W, Cin, NT);
    std::vector<float> sf, yf;
    // a rows = 0 / 1 launch writes only its own output rows: start from what the output buffer already holds
    if (flags & MVSTER_TC3_OUT_PB16) yf = unpack_pb16(y, P, 2 * H, 2 * W, Cout, NT);
    else yf.assign(static_cast<const float*>(y), static_cast<const float*>(y) + (size_t)P * 4 * H * W * Cout);
    if (skip) sf = unpack_pb16(skip, P, 2 * H, 2 * W, Cout, NT);
    const int rc = deconv_run(xf.data(), w_packed, scale, bias, skip ? sf.data() : nullptr, yf.data(), B, D, H, W, Cin, Cout, rows,
                              (flags & 1) | (NT == 2 ? MVSTER_TC3_FP16X2 : MVSTER_TC3_BF16X1));
    if (rc) return rc;
    if (flags & MVSTER_TC3_OUT_PB16) pack_pb16(yf, y, P, 2 * H, 2 * W, Cout, NT);
    else std::memcpy(y, yf.data(), yf.size() * 4);
    return 0;
}
size_t mvster_conv_tc3_packed_bytes(int Cin, int Cout, int kd, int k, int s) {
    if (!mvster::tc3::supported(Cin, Cout, kd, k, s)) return 0;
    const int NC = Cout < 16 ? 16 : (Cout > 64 ? 80 : Cout);
    return (size_t)mvster::tc3::build_plan(Cin, kd, k, s, nullptr, nullptr) * 96 * NC;
}
size_t mvster_deconv_tc3_packed_bytes(int Cin, int Cout, int rows) {
    const int n = mvster::tc3::deconv_ncls(rows) * Cout;
    const bool ok = (Cin == 16 || Cin == 32 || Cin == 64) && (Cout == 8 || Cout == 16 || Cout == 32) && rows >= -1 && rows <= 1 &&
                    (n == 16 || n == 32 || n == 64);
    return ok ? (size_t)(Cin / 16) * mvster::tc3::deconv_ntap(rows) * 96 * n : 0;
}
}
