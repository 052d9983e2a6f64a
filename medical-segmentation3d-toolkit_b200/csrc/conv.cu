// conv.cu - C-ABI dispatch of the convolution entry point to the SIMT or tcgen05 implementation.
#include "common.cuh"

int seg3d_conv_simt(int mode, int dtype, const void* x, int x_ld, int Cin, const void* w, const float* bias,
                    void* y, int y_ld, int Cout, int N, int D, int H, int W, double* stats, cudaStream_t st);
int seg3d_conv_tc(int mode, int dtype, const void* x, int x_ld, int Cin, const void* w, const float* bias,
                  void* y, int y_ld, int Cout, int N, int D, int H, int W, double* stats, cudaStream_t st,
                  int epi_mode, const double* gn_stats, const float* gn_gamma, const float* gn_beta, float gn_eps,
                  int split_lo_off, int out_f32_generic,
                  int xf_nch = 0, const double* xf_stats = nullptr, const float* xf_gamma = nullptr, const float* xf_beta = nullptr,
                  float xf_eps = 0.f);
int seg3d_conv_cin1_tc_supported(int dtype, int Cin, int Cout, int x_ld, int y_ld, int W);
int seg3d_conv_cin1_tc(int dtype, const void* x, const void* w, const float* bias, void* y, int y_ld,
                       int N, int D, int H, int W, double* stats, cudaStream_t st);
int seg3d_conv_tc_supported(int mode, int dtype, int Cin, int Cout, int x_ld, int y_ld, int D, int H, int W);

extern "C" int seg3d_conv3d_fwd(int mode, int dtype, int impl, const void* x, int x_ld, int Cin, const void* w,
                                const float* bias, void* y, int y_ld, int Cout, int N, int D, int H, int W,
                                double* stats, void* stream) {
  SEG3D_REQUIRE(x && w && y, "conv3d_fwd: null pointer");
  SEG3D_REQUIRE(Cin > 0 && Cout > 0 && N > 0 && D > 0 && H > 0 && W > 0, "conv3d_fwd: bad dims");
  // with SEG3D_OUT_F32 the fp32 result keeps only the first y_ld channels (zero-padded output channels are dropped)
  SEG3D_REQUIRE(x_ld >= Cin && (y_ld >= Cout || ((dtype & SEG3D_OUT_F32) && y_ld > 0)), "conv3d_fwd: pitch smaller than channel count");
  cudaStream_t st = (cudaStream_t)stream;
  // input block: Cin == 1, fp32 [27][16] weights (the SIMT layout) for both implementations
  if (impl == SEG3D_IMPL_AUTO && mode == SEG3D_CONV_K3 && seg3d_conv_cin1_tc_supported(dtype, Cin, Cout, x_ld, y_ld, W))
    return seg3d_conv_cin1_tc(dtype, x, w, bias, y, y_ld, N, D, H, W, stats, st);
  if (impl == SEG3D_IMPL_AUTO)
    impl = seg3d_conv_tc_supported(mode, dtype, Cin, Cout, x_ld, y_ld, D, H, W) ? SEG3D_IMPL_TCGEN05 : SEG3D_IMPL_SIMT;
  if (impl == SEG3D_IMPL_TCGEN05) {
    if (!seg3d_conv_tc_supported(mode, dtype, Cin, Cout, x_ld, y_ld, D, H, W)) {
      seg3d_set_error("conv3d_fwd: tcgen05 path does not take mode=%d dtype=%d Cin=%d Cout=%d", mode, dtype, Cin, Cout);
      return SEG3D_EUNSUPPORTED;
    }
    return seg3d_conv_tc(mode, dtype, x, x_ld, Cin, w, bias, y, y_ld, Cout, N, D, H, W, stats, st, 0, nullptr, nullptr, nullptr, 0.f, 0, 0);
  }
  SEG3D_REQUIRE(!(dtype & SEG3D_OUT_F32), "conv3d_fwd: SEG3D_OUT_F32 needs the tensor-core path");
  if (impl == SEG3D_IMPL_SIMT)
    return seg3d_conv_simt(mode, dtype, x, x_ld, Cin, w, bias, y, y_ld, Cout, N, D, H, W, stats, st);
  seg3d_set_error("conv3d_fwd: unknown impl %d", impl);
  return SEG3D_EINVAL;
}

// conv -> GroupNorm(1,C) -> ReLU in two launches of the same tensor-core convolution, for convolutions that are cheap to run
// twice (the HBM-bound stride-2 / transposed convs: conv_gn_relu of vnet_downblock.py:19 and vnet_upblock.py:19 without the
// raw intermediate).  pass 0: accumulate sum / sum-of-squares into `stats` (nothing is stored);  pass 1: recompute and store
// y = relu((conv - mean) * rstd * gamma + beta) from the finished statistics.
extern "C" int seg3d_conv3d_gn_relu_fwd(int mode, int dtype, int pass, const void* x, int x_ld, int Cin, const void* w,
                                        const float* bias, void* y, int y_ld, int Cout, int N, int D, int H, int W,
                                        double* stats, const float* gamma, const float* beta, float eps, void* stream) {
  SEG3D_REQUIRE(x && w && y && stats && gamma && beta, "conv3d_gn_relu_fwd: null pointer");
  SEG3D_REQUIRE(pass == 0 || pass == 1, "conv3d_gn_relu_fwd: pass must be 0 or 1");
  SEG3D_REQUIRE(Cin > 0 && Cout > 0 && N > 0 && D > 0 && H > 0 && W > 0 && x_ld >= Cin && y_ld >= Cout, "conv3d_gn_relu_fwd: bad dims");
  if (!seg3d_conv_tc_supported(mode, dtype, Cin, Cout, x_ld, y_ld, D, H, W)) {
    seg3d_set_error("conv3d_gn_relu_fwd: tcgen05 path does not take mode=%d dtype=%d Cin=%d Cout=%d", mode, dtype, Cin, Cout);
    return SEG3D_EUNSUPPORTED;
  }
  return seg3d_conv_tc(mode, dtype, x, x_ld, Cin, w, bias, y, y_ld, Cout, N, D, H, W, pass == 0 ? stats : nullptr,
                       (cudaStream_t)stream, pass == 0 ? 1 : 2, stats, gamma, beta, eps, 0, 0);
}

// Strict-parity convolution on the tensor cores (split operands): x is stored as two f16 halves per voxel row,
// x = hi + lo with hi at channel c and lo at channel lo_off + c (lo_off = x_ld / 2 for the plan's buffers); w holds
// [taps][Cout][whi(Cin) | wlo(Cin)] (T2S2: [8*Cout][whi | wlo]).  The K loop accumulates hi*whi + lo*whi + hi*wlo in fp32
// (the dropped lo*wlo term is 2^-22 relative), and y is the fp32 result + bias with pitch y_ld floats.
extern "C" int seg3d_conv3d_split_fwd(int mode, const void* x, int x_ld, int lo_off, int Cin, const void* w, const float* bias,
                                      float* y, int y_ld, int Cout, int N, int D, int H, int W, double* stats, void* stream) {
  SEG3D_REQUIRE(x && w && y, "conv3d_split_fwd: null pointer");
  SEG3D_REQUIRE(Cin > 0 && Cout > 0 && N > 0 && D > 0 && H > 0 && W > 0 && lo_off >= Cin && x_ld >= lo_off + Cin && y_ld >= Cout,
                "conv3d_split_fwd: bad dims");
  SEG3D_REQUIRE(lo_off % 8 == 0 && y_ld % 4 == 0 && ((uintptr_t)y) % 16 == 0, "conv3d_split_fwd: alignment");
  if (!seg3d_conv_tc_supported(mode, SEG3D_F16, Cin, Cout, x_ld, 8, D, H, W)) {
    seg3d_set_error("conv3d_split_fwd: tcgen05 path does not take mode=%d Cin=%d Cout=%d", mode, Cin, Cout);
    return SEG3D_EUNSUPPORTED;
  }
  return seg3d_conv_tc(mode, SEG3D_F16, x, x_ld, Cin, w, bias, y, y_ld, Cout, N, D, H, W, stats, (cudaStream_t)stream,
                       0, nullptr, nullptr, nullptr, 0.f, lo_off, 1);
}

// k3 convolution whose first gn_ch INPUT channels are still the raw result of the producing convolution: relu(GroupNorm(1, gn_ch))
// of those channels (finished sums gn_stats, as seg3d_gn_apply with relu = 1 would write them) is formed in shared memory between
// the TMA loads and the MMAs of the z-march kernel, so the apply pass of vnet_upblock.py:19 never touches HBM.  The remaining
// channels (the skip half of the concat buffer, vnet_upblock.py:21) are used as stored.  Cin must be 32, gn_ch 8, 16 or 32.
extern "C" int seg3d_conv3d_k3_gnin_fwd(int dtype, const void* x, int x_ld, int Cin, int gn_ch, const double* gn_stats,
                                        const float* gamma, const float* beta, float eps, const void* w, const float* bias,
                                        void* y, int y_ld, int Cout, int N, int D, int H, int W, double* stats, void* stream) {
  SEG3D_REQUIRE(x && w && y && gn_stats && gamma && beta, "conv3d_k3_gnin_fwd: null pointer");
  SEG3D_REQUIRE(Cin == 32 && Cout > 0 && N > 0 && D >= 4 && H > 0 && W > 0 && W % 8 == 0 && x_ld >= Cin && y_ld >= Cout, "conv3d_k3_gnin_fwd: bad dims");
  if (!seg3d_conv_tc_supported(SEG3D_CONV_K3, dtype, Cin, Cout, x_ld, y_ld, D, H, W)) {
    seg3d_set_error("conv3d_k3_gnin_fwd: tcgen05 path does not take dtype=%d Cin=%d Cout=%d", dtype, Cin, Cout);
    return SEG3D_EUNSUPPORTED;
  }
  return seg3d_conv_tc(SEG3D_CONV_K3, dtype, x, x_ld, Cin, w, bias, y, y_ld, Cout, N, D, H, W, stats, (cudaStream_t)stream,
                       0, nullptr, nullptr, nullptr, 0.f, 0, 0, gn_ch, gn_stats, gamma, beta, eps);
}
