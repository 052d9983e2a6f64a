// conv_tc.cu - placeholder until the tcgen05 kernel lands (next commit).
#include "common.cuh"
int seg3d_conv_tc_supported(int, int, int, int, int, int, int, int, int) { return 0; }
int seg3d_conv_tc(int, int, const void*, int, int, const void*, const float*, void*, int, int, int, int, int, int, double*, cudaStream_t) {
  seg3d_set_error("tcgen05 conv not built"); return SEG3D_EUNSUPPORTED;
}
