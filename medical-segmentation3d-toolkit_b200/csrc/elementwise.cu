// elementwise.cu - GroupNorm(1,C) apply (+ReLU, +residual, concat-in-place) and the output-block
// tail (GN1+ReLU -> 1x1x1 conv -> GN2 -> channel softmax).  HBM-bound streaming kernels:
// 16-byte vector accesses, grid sized in multiples of the SM count, fp32 math.
#include "common.cuh"
#include <stdarg.h>

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message)
static thread_local char g_err[512] = "";
void seg3d_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* seg3d_last_error(void) { return g_err; }
extern "C" int seg3d_version(void) { return 100; }
extern "C" int seg3d_device_check(int device) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { seg3d_set_error("cudaGetDeviceProperties failed"); return SEG3D_ECUDA; }
  if (p.major != 10) { seg3d_set_error("device %d is sm_%d%d; this library is sm_100a only", device, p.major, p.minor); return SEG3D_EUNSUPPORTED; }
  return SEG3D_OK;
}

// ---------------------------------------------------------------------------------------------
constexpr int GN_MAXC = 512;

template <typename T, bool RELU, bool RES>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const T* __restrict__ y, int y_ld, int C, const double* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                const T* __restrict__ res, int res_ld, T* __restrict__ out, int out_ld, long long nvox) {
  __shared__ float sa[GN_MAXC], sb[GN_MAXC];
  const int n = blockIdx.y;
  float mean, rstd;
  gn_mean_rstd(stats + 2 * n, (double)nvox * C, eps, mean, rstd);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float a = rstd * gamma[c];
    sa[c] = a; sb[c] = beta[c] - mean * a;
  }
  __syncthreads();
  const int cv = C >> 3;                                   // 8-channel vectors per voxel
  const long long items = nvox * cv;
  const T* yn = y + (size_t)n * nvox * y_ld;
  const T* rn = RES ? res + (size_t)n * nvox * res_ld : nullptr;
  T* on = out + (size_t)n * nvox * out_ld;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const long long vox = it / cv; const int c0 = (int)(it - vox * cv) << 3;
    Vec8<T> v; v.load(yn + vox * y_ld + c0);
    float f[8]; v.get(f);
    float r[8];
    if (RES) { Vec8<T> rv; rv.load(rn + vox * res_ld + c0); rv.get(r); }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(f[j], sa[c0 + j], sb[c0 + j]);
      if (RES) t += r[j];
      if (RELU) t = fmaxf(t, 0.f);
      f[j] = t;
    }
    v.set(f); v.store(on + vox * out_ld + c0);
  }
}

// Split-operand strict mode: y is the fp32 conv result; res / out are [hi | lo] f16 rows (lo half lo_off channels after hi).
template <bool RELU, bool RES>
__global__ void __launch_bounds__(256)
gn_apply_split_kernel(const float* __restrict__ y, int y_ld, int C, const double* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                      const __half* __restrict__ res, int res_ld, int res_lo, __half* __restrict__ out, int out_ld, int out_lo,
                      long long nvox) {
  __shared__ float sa[GN_MAXC], sb[GN_MAXC];
  const int n = blockIdx.y;
  float mean, rstd;
  gn_mean_rstd(stats + 2 * n, (double)nvox * C, eps, mean, rstd);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float a = rstd * gamma[c];
    sa[c] = a; sb[c] = beta[c] - mean * a;
  }
  __syncthreads();
  const int cv = C >> 3;
  const long long items = nvox * cv;
  const float* yn = y + (size_t)n * nvox * y_ld;
  const __half* rn = RES ? res + (size_t)n * nvox * res_ld : nullptr;
  __half* on = out + (size_t)n * nvox * out_ld;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const long long vox = it / cv; const int c0 = (int)(it - vox * cv) << 3;
    Vec8<float> v; v.load(yn + vox * y_ld + c0);
    float f[8]; v.get(f);
    float rh[8], rl[8];
    if (RES) {
      Vec8<__half> a, b; a.load(rn + vox * res_ld + c0); b.load(rn + vox * res_ld + res_lo + c0);
      a.get(rh); b.get(rl);
    }
    float hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(f[j], sa[c0 + j], sb[c0 + j]);
      if (RES) t += rh[j] + rl[j];
      if (RELU) t = fmaxf(t, 0.f);
      const float h = __half2float(__float2half_rn(t));
      hi[j] = h; lo[j] = t - h;
    }
    Vec8<__half> o; o.set(hi); o.store(on + vox * out_ld + c0);
    o.set(lo); o.store(on + vox * out_ld + out_lo + c0);
  }
}

extern "C" int seg3d_gn_apply_split(const float* y, int y_ld, int C, const double* stats, const float* gamma, const float* beta,
                                    float eps, const void* res, int res_ld, int res_lo, void* out, int out_ld, int out_lo,
                                    int relu, int N, int64_t nvox, void* stream) {
  SEG3D_REQUIRE(y && stats && gamma && beta && out && C > 0 && C % 8 == 0 && C <= GN_MAXC && N > 0 && nvox > 0, "gn_apply_split: bad arguments");
  SEG3D_REQUIRE(y_ld % 4 == 0 && out_ld % 8 == 0 && out_lo % 8 == 0 && (!res || (res_ld % 8 == 0 && res_lo % 8 == 0)), "gn_apply_split: pitches");
  const long long items = (long long)nvox * (C / 8);
  const int sms = seg3d_num_sms();
  long long want = (items + 255) / 256;
  dim3 grid((unsigned)(want > 8ll * sms ? 8ll * sms : (want < 1 ? 1 : want)), N);
  cudaStream_t st = (cudaStream_t)stream;
  const __half* r = (const __half*)res; __half* o = (__half*)out;
#define SEG3D_GNS(RL, RS) gn_apply_split_kernel<RL, RS><<<grid, 256, 0, st>>>(y, y_ld, C, stats, gamma, beta, eps, r, res_ld, res_lo, o, out_ld, out_lo, nvox)
  if (relu) { if (res) SEG3D_GNS(true, true); else SEG3D_GNS(true, false); }
  else      { if (res) SEG3D_GNS(false, true); else SEG3D_GNS(false, false); }
#undef SEG3D_GNS
  SEG3D_CHECK_LAUNCH("gn_apply_split_kernel");
  return SEG3D_OK;
}

extern "C" int seg3d_gn_apply(int dtype, const void* y, int y_ld, int C, const double* stats,
                              const float* gamma, const float* beta, float eps, const void* res, int res_ld,
                              void* out, int out_ld, int relu, int N, int64_t nvox, void* stream) {
  SEG3D_REQUIRE(C > 0 && C % 8 == 0 && C <= GN_MAXC, "gn_apply: C=%d must be a multiple of 8 and <= %d", C, GN_MAXC);
  SEG3D_REQUIRE(y_ld % 8 == 0 && out_ld % 8 == 0 && (!res || res_ld % 8 == 0), "gn_apply: pitches must be multiples of 8");
  SEG3D_REQUIRE(N > 0 && nvox > 0 && stats && gamma && beta, "gn_apply: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long items = nvox * (C / 8);
  long long want = (items + 256 * 4 - 1) / (256 * 4);     // ~4 vectors per thread
  const int sms = seg3d_num_sms();
  int gx = (int)(want < 1 ? 1 : (want > 8LL * sms ? 8LL * sms : want));
  if (gx > sms) gx = (gx / sms) * sms;                     // whole waves
  dim3 grid(gx, N), block(256);
  SEG3D_DISPATCH_DTYPE(dtype, T, {
    const T* yp = (const T*)y; const T* rp = (const T*)res; T* op = (T*)out;
    if (relu && res)       gn_apply_kernel<T, true, true><<<grid, block, 0, st>>>(yp, y_ld, C, stats, gamma, beta, eps, rp, res_ld, op, out_ld, nvox);
    else if (relu)         gn_apply_kernel<T, true, false><<<grid, block, 0, st>>>(yp, y_ld, C, stats, gamma, beta, eps, rp, res_ld, op, out_ld, nvox);
    else if (res)          gn_apply_kernel<T, false, true><<<grid, block, 0, st>>>(yp, y_ld, C, stats, gamma, beta, eps, rp, res_ld, op, out_ld, nvox);
    else                   gn_apply_kernel<T, false, false><<<grid, block, 0, st>>>(yp, y_ld, C, stats, gamma, beta, eps, rp, res_ld, op, out_ld, nvox);
  });
  SEG3D_CHECK_LAUNCH("gn_apply_kernel");
  return SEG3D_OK;
}

// ---------------------------------------------------------------------------------------------
// Output-block tail.  C <= 16 classes (the reference's own network/vbnet_test.py builds 16); one thread per voxel keeps the
// C-vector in registers.
constexpr int TAIL_MAXC = 16;

struct TailParams {
  float a1[TAIL_MAXC], b1[TAIL_MAXC];        // GN1 folded scale/shift (filled per sample in-kernel)
};

template <typename T, int C>
__device__ __forceinline__ void tail_z(const T* p, const float* a1, const float* b1, const float* w2, const float* bias2, float* z) {
  float h[C];
#pragma unroll
  for (int c = 0; c < C; ++c) h[c] = fmaxf(fmaf(to_f32<T>(p[c]), a1[c], b1[c]), 0.f);
#pragma unroll
  for (int o = 0; o < C; ++o) {
    float t = bias2 ? bias2[o] : 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) t = fmaf(w2[o * C + c], h[c], t);
    z[o] = t;
  }
}

template <typename T, int C, bool PROBS>
__global__ void __launch_bounds__(256)
outblock_tail_kernel(const T* __restrict__ y1, int ld, const double* __restrict__ stats1,
                     const float* __restrict__ gamma1, const float* __restrict__ beta1,
                     const float* __restrict__ w2, const float* __restrict__ bias2, float eps,
                     double* __restrict__ stats2_out, const double* __restrict__ stats2_in,
                     const float* __restrict__ gamma2, const float* __restrict__ beta2,
                     float* __restrict__ probs, long long nvox) {
  __shared__ float sw2[C * C], sbias2[C], sa1[C], sb1[C], sa2[C], sb2[C];
  __shared__ float red[64];
  const int n = blockIdx.y;
  if (threadIdx.x < C * C) sw2[threadIdx.x] = w2[threadIdx.x];
  if (threadIdx.x < C) {
    float mean, rstd;
    gn_mean_rstd(stats1 + 2 * n, (double)nvox * C, eps, mean, rstd);
    const float a = rstd * gamma1[threadIdx.x];
    sa1[threadIdx.x] = a; sb1[threadIdx.x] = beta1[threadIdx.x] - mean * a;
    sbias2[threadIdx.x] = bias2 ? bias2[threadIdx.x] : 0.f;
    if (PROBS) {
      gn_mean_rstd(stats2_in + 2 * n, (double)nvox * C, eps, mean, rstd);
      const float a2 = rstd * gamma2[threadIdx.x];
      sa2[threadIdx.x] = a2; sb2[threadIdx.x] = beta2[threadIdx.x] - mean * a2;
    }
  }
  __syncthreads();
  const T* yn = y1 + (size_t)n * nvox * ld;
  float s = 0.f, ss = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
    float z[C];
    tail_z<T, C>(yn + v * ld, sa1, sb1, sw2, sbias2, z);
    if (!PROBS) {
#pragma unroll
      for (int c = 0; c < C; ++c) { s += z[c]; ss += z[c] * z[c]; }
    } else {
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < C; ++c) { z[c] = fmaf(z[c], sa2[c], sb2[c]); m = fmaxf(m, z[c]); }
      float den = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) { z[c] = expf(z[c] - m); den += z[c]; }
      const float inv = 1.f / den;
#pragma unroll
      for (int c = 0; c < C; ++c) probs[((size_t)n * C + c) * nvox + v] = z[c] * inv;
    }
  }
  if (!PROBS) block_stats_atomic(s, ss, stats2_out + 2 * n, red);
}

template <typename T, bool PROBS>
static int launch_tail(int C, dim3 grid, cudaStream_t st, const T* y1, int ld, const double* stats1, const float* g1,
                       const float* b1, const float* w2, const float* bias2, float eps, double* s2o, const double* s2i,
                       const float* g2, const float* b2, float* probs, long long nvox) {
#define TAIL_CASE(CC) case CC: outblock_tail_kernel<T, CC, PROBS><<<grid, 256, 0, st>>>(y1, ld, stats1, g1, b1, w2, bias2, eps, s2o, s2i, g2, b2, probs, nvox); break;
  switch (C) {
    TAIL_CASE(1) TAIL_CASE(2) TAIL_CASE(3) TAIL_CASE(4) TAIL_CASE(5) TAIL_CASE(6) TAIL_CASE(7) TAIL_CASE(8)
    TAIL_CASE(9) TAIL_CASE(10) TAIL_CASE(11) TAIL_CASE(12) TAIL_CASE(13) TAIL_CASE(14) TAIL_CASE(15) TAIL_CASE(16)
    default: seg3d_set_error("outblock tail: C=%d not in 1..16", C); return SEG3D_EUNSUPPORTED;
  }
#undef TAIL_CASE
  SEG3D_CHECK_LAUNCH("outblock_tail_kernel");
  return SEG3D_OK;
}

static dim3 tail_grid(int N, long long nvox) {
  const int sms = seg3d_num_sms();
  long long want = (nvox + 255) / 256;
  int gx = (int)(want < 1 ? 1 : (want > 4LL * sms ? 4LL * sms : want));
  return dim3(gx, N);
}

extern "C" int seg3d_outblock_tail_stats(int dtype, const void* y1, int ld, int C, const double* stats1,
                                         const float* gamma1, const float* beta1, const float* w2, const float* bias2,
                                         float eps, double* stats2, int N, int64_t nvox, void* stream) {
  SEG3D_REQUIRE(y1 && stats1 && gamma1 && beta1 && w2 && stats2 && N > 0 && nvox > 0 && ld >= C, "outblock_tail_stats: bad arguments");
  SEG3D_DISPATCH_DTYPE(dtype, T, return (launch_tail<T, false>(C, tail_grid(N, nvox), (cudaStream_t)stream, (const T*)y1, ld, stats1, gamma1, beta1,
                                                                w2, bias2, eps, stats2, nullptr, nullptr, nullptr, nullptr, nvox)));
  return SEG3D_OK;
}

extern "C" int seg3d_outblock_tail_probs(int dtype, const void* y1, int ld, int C, const double* stats1,
                                         const float* gamma1, const float* beta1, const float* w2, const float* bias2,
                                         const double* stats2, const float* gamma2, const float* beta2, float eps,
                                         float* probs, int N, int64_t nvox, void* stream) {
  SEG3D_REQUIRE(y1 && stats1 && gamma1 && beta1 && w2 && stats2 && gamma2 && beta2 && probs && N > 0 && nvox > 0 && ld >= C,
                "outblock_tail_probs: bad arguments");
  SEG3D_DISPATCH_DTYPE(dtype, T, return (launch_tail<T, true>(C, tail_grid(N, nvox), (cudaStream_t)stream, (const T*)y1, ld, stats1, gamma1, beta1,
                                                               w2, bias2, eps, nullptr, stats2, gamma2, beta2, probs, nvox)));
  return SEG3D_OK;
}
