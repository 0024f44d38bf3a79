// Host-side layer plan of the generation-3 tensor-core convolution (conv_tc3.cu): tile geometry, the stage / tap enumeration
// shared by the launcher and by the weight packer (mvster_conv_tc3_plan), and the class / tap counts of the transposed form.
// Plain C++ (no CUDA): also compiled into the CPU emulation of the test suite (tests/emu/emu_stubs.cpp), which decodes the packed
// weight slabs with it and evaluates the layer by a direct convolution - so packers and drivers are testable without a GPU.
#pragma once
#include <stdint.h>

namespace mvster {
namespace tc3 {

constexpr int TW = 8, TH = 16, HW_ = TW + 2, HH_ = TH + 2, HPIX = HW_ * HH_;
constexpr int QBYTES = HPIX * 16;            // bytes per channel quad of a halo tile (2880)
constexpr int PLANE = QBYTES;                // bf16 operand plane = [180 pixels][8 channels]; 2880 = 64 (mod 128): the two
                                             // octet planes a converter half-warp writes fall into disjoint banks
constexpr int MAX_STAGES = 16, MAX_TAPS = 9;

struct Stage {
    short c0, nq, ox, oy, dz, ntap, slab0, pad;
};
struct Plan {
    Stage st[MAX_STAGES];
    // Per MMA slot of a stage, the low descriptor word to add to the tile's base: start shift = halo row * 10 + halo column of the
    // tap (16-byte units, bits 0-13) | LBO (bits 16-29).  LBO = distance between the two 8-channel K halves of the MMA: the next
    // channel-octet plane for Cin >= 16; for Cin <= 8 the second K half is a SECOND TAP of the same plane (LBO = its shift minus
    // the first tap's), so a 3x3 conv on 8 channels needs 5 MMA slots instead of 9.
    uint32_t a_desc[MAX_STAGES][MAX_TAPS];
};
inline uint32_t tap_desc(int off, int lbo) { return (uint32_t)off | ((uint32_t)lbo << 16); }

inline bool supported(int Cin, int Cout, int kd, int k, int s) {
    const bool cin_ok = Cin == 4 || Cin == 8 || Cin == 16 || Cin == 32 || Cin == 64;
    const bool cout_ok = Cout == 8 || Cout == 16 || Cout == 32 || Cout == 64 || (Cout == 72 && k == 1 && kd == 1);  // 72: N padded to 80
    const bool shape_ok =(s == 1 && (k == 1 || k == 3) && (kd == 1 || kd == 3)) || (s == 2 && (k == 3 || k == 5) && kd == 1);
    return cin_ok && cout_ok && shape_ok && kd * (s == 2 ? 4 : 1) * ((Cin + 15) / 16) <= MAX_STAGES;
}

// Stage/tap enumeration shared by the launcher and by the host-side weight packer (mvster_conv_tc3_plan).
// Stride 2: input row 2y + ky - pad = 2 (y + m) + py with parity class py in {0,1}; class (py, px) is staged as its own halo
// tile (TMA element strides 2, origin 2*y0 - 2 + py) and tap ky lands on halo row m + 1.
// slabs[i] = {kz, ky, kx, first input channel, ky2, kx2}: the weights of MMA slot i; (ky2, kx2) = the tap in the second K half
// when two taps of an <= 8-channel layer share one MMA, else (-1, -1).
inline int build_plan(int Cin, int kd, int k, int s, Plan* plan, int (*slabs)[6]) {
    int ns = 0, nslab = 0;
    const int kch = (Cin + 15) / 16, pz = kd / 2, pad = k / 2, npar = s == 2 ? 2 : 1;
    const bool pair = Cin <= 8;
    for (int kz = 0; kz < kd; ++kz)
        for (int py = 0; py < npar; ++py)
            for (int px = 0; px < npar; ++px)
                for (int kc = 0; kc < kch; ++kc, ++ns) {
                    Stage S;
                    S.c0 = (short)(kc * 16);
                    S.nq = (short)((Cin - kc * 16) / 4 < 4 ? (Cin - kc * 16) / 4 : 4);
                    S.dz = (short)(kz - pz);
                    S.ox = (short)(s == 2 ? -2 + px : -1);
                    S.oy = (short)(s == 2 ? -2 + py : -1);
                    S.slab0 = (short)nslab;
                    S.pad = 0;
                    int nt = 0, toff[25], tky[25], tkx[25];  // the stage's taps, halo offsets ascending
                    for (int ky = 0; ky < k; ++ky)
                        for (int kx = 0; kx < k; ++kx) {
                            int hy, hx;
                            if (s == 1) {
                                hy = ky - pad + 1; hx = kx - pad + 1;
                            } else {
                                const int oy = ky - pad, ox = kx - pad, cy = ((oy % 2) + 2) % 2, cx = ((ox % 2) + 2) % 2;
                                if (cy != py || cx != px) continue;
                                hy = (oy - cy) / 2 + 1; hx = (ox - cx) / 2 + 1;
                            }
                            toff[nt] = hy * HW_ + hx; tky[nt] = ky; tkx[nt] = kx;
                            ++nt;
                        }
                    int nslot = 0;
                    for (int t = 0; t < nt; ++nslot, ++nslab) {
                        // paired layers with an odd tap count: the FIRST slot is the single one, so that its second K half
                        // (LBO = 1: the next pixel, against zero weights) still reads converted data and never past the tile
                        const bool two = pair && !(t == 0 && (nt & 1));
                        // LBO: next channel-octet plane, or (paired) the second tap relative to the first
                        const int lbo = pair ? (two ? toff[t + 1] - toff[t] : 1) : (PLANE >> 4);
                        if (plan) plan->a_desc[ns][nslot] = tap_desc(toff[t], lbo);
                        if (slabs) {
                            slabs[nslab][0] = kz; slabs[nslab][1] = tky[t]; slabs[nslab][2] = tkx[t]; slabs[nslab][3] = kc * 16;
                            slabs[nslab][4] = two ? tky[t + 1] : -1; slabs[nslab][5] = two ? tkx[t + 1] : -1;
                        }
                        t += two ? 2 : 1;
                    }
                    S.ntap = (short)nslot;
                    if (plan) plan->st[ns] = S;
                }
    return nslab;
}

inline int deconv_ncls(int rows) { return rows < 0 ? 4 : 2; }
inline int deconv_ntap(int rows) { return rows == 0 ? 2 : 4; }


}  // namespace tc3
}  // namespace mvster
