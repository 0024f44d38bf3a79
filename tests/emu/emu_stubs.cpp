// Tensor-core entry points that the SIMT sources reference (reg2d's optional tcgen05 layers): unavailable on the host.
#include <cstddef>
#include <cstdint>
typedef void* mvster_stream_t;
extern "C" {
int mvster_conv3d_tc_f32(...) { return -2; }
int mvster_conv3d_tc2_f32(...) { return -2; }
int mvster_conv_tc3_f32(...) { return -2; }
int mvster_deconv_tc3_f32(...) { return -2; }
size_t mvster_conv_tc3_packed_bytes(int, int, int, int, int) { return 0; }
size_t mvster_deconv_tc3_packed_bytes(int, int, int) { return 0; }
}
