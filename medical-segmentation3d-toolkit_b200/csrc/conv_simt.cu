// conv_simt.cu - CUDA-core fp32-accumulate implicit-GEMM convolution.
//
// Strict-parity path (fp32 storage, FFMA) and the fallback for shapes the tcgen05 kernel does
// not take (Cin == 1 input block, Cout < 16 output block).  One kernel covers the four
// convolution forms of the reference network (include/seg3d_b200.h: seg3d_conv_mode):
// rows of the GEMM are output voxels (input voxels for the transposed conv), columns are
// output channels (8*Cout for the transposed conv), K runs over (tap, ci).
#include "common.cuh"

struct ConvGeom {
  int mode, N, D, H, W;   // input spatial dims
  int Do, Ho, Wo;         // dims of the GEMM row space
  int Cin, Cout, x_ld, y_ld;
  int K, Ng;              // GEMM K and N
  int vps;                // rows per sample
  int tiles_per_sample;
  int vec;                // 1 if 8-channel vector loads are legal
};

template <typename T, int BN>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const ConvGeom g, const T* __restrict__ x, const float* __restrict__ w,
                 const float* __restrict__ bias, T* __restrict__ y, double* __restrict__ stats) {
  constexpr int BM = 128, BK = 16, TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ float red[64];

  const int tid = threadIdx.x;
  const int n = blockIdx.x / g.tiles_per_sample;
  const int m0 = (blockIdx.x % g.tiles_per_sample) * BM;
  const int n0 = blockIdx.y * BN;

  // ---- A-load role: one GEMM row, 8 consecutive k per thread ----
  const int lrow = tid & 127, lk = (tid >> 7) * 8;
  const int mv = m0 + lrow;
  const bool row_ok = mv < g.vps;
  int rz = 0, ry = 0, rx = 0;
  if (row_ok) { rx = mv % g.Wo; int t = mv / g.Wo; ry = t % g.Ho; rz = t / g.Ho; }
  const T* xn = x + (size_t)n * g.D * g.H * g.W * g.x_ld;

  auto src_voxel = [&](int tap, int& ok) -> size_t {
    int zi, yi, xi;
    if (g.mode == SEG3D_CONV_K3) {
      zi = rz + tap / 9 - 1; yi = ry + (tap / 3) % 3 - 1; xi = rx + tap % 3 - 1;
      ok = (zi >= 0) & (zi < g.D) & (yi >= 0) & (yi < g.H) & (xi >= 0) & (xi < g.W);
    } else if (g.mode == SEG3D_CONV_K2S2) {
      zi = 2 * rz + (tap >> 2); yi = 2 * ry + ((tap >> 1) & 1); xi = 2 * rx + (tap & 1); ok = 1;
    } else { zi = rz; yi = ry; xi = rx; ok = 1; }
    return ((size_t)(zi * g.H + yi) * g.W + xi) * g.x_ld;
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nchunks = (g.K + BK - 1) / BK;
  for (int kc = 0; kc < nchunks; ++kc) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    const int k0 = kc * BK + lk;
    if (row_ok && k0 < g.K) {
      if (g.vec) {
        const int tap = k0 / g.Cin, ci = k0 - tap * g.Cin;
        int ok; const size_t off = src_voxel(tap, ok);
        if (ok) { Vec8<T> v; v.load(xn + off + ci); v.get(a); }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + j;
          if (k < g.K) {
            const int tap = k / g.Cin, ci = k - tap * g.Cin;
            int ok; const size_t off = src_voxel(tap, ok);
            if (ok) a[j] = to_f32<T>(xn[off + ci]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) As[lk + j][lrow] = a[j];
    for (int i = tid; i < BK * BN; i += 256) {
      const int r = i / BN, c = i - r * BN;
      const int k = kc * BK + r, nn = n0 + c;
      Bs[r][c] = (k < g.K && nn < g.Ng) ? w[(size_t)k * g.Ng + nn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: + bias, GroupNorm partial sums from the fp32 values, store ----
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + ty * 8 + i;
    if (r >= g.vps) continue;
    size_t obase;
    int oz = 0, oy = 0, ox = 0;
    if (g.mode == SEG3D_CONV_T2S2) {
      ox = r % g.Wo; int t = r / g.Wo; oy = t % g.Ho; oz = t / g.Ho;
      obase = 0;
    } else {
      obase = ((size_t)n * g.vps + r) * g.y_ld;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + tx + 16 * j;
      if (col >= g.Ng) continue;
      int co = col; size_t o = obase;
      if (g.mode == SEG3D_CONV_T2S2) {
        const int tap = col / g.Cout; co = col - tap * g.Cout;
        const int zo = 2 * oz + (tap >> 2), yo = 2 * oy + ((tap >> 1) & 1), xo = 2 * ox + (tap & 1);
        o = ((((size_t)n * 2 * g.D + zo) * 2 * g.H + yo) * 2 * g.W + xo) * g.y_ld;
      }
      const float v = acc[i][j] + (bias ? bias[co] : 0.f);
      s += v; ss += v * v;
      y[o + co] = from_f32<T>(v);
    }
  }
  if (stats) block_stats_atomic(s, ss, stats + 2 * n, red);
}


// ---- input block: k3 conv with a single input channel (vnet_inblock.py:9), Cout == 16 ---------------
// Direct convolution: one thread per output voxel keeps the 16 outputs in registers; the 8x8x4 tile
// reads its 10x10x6 input halo and the 27x16 filter bank from shared memory.  HBM traffic is the
// output write (16 channels per voxel); zero padding comes from the halo fill.
template <typename T>
__global__ void __launch_bounds__(256)
conv3d_k3_cin1_kernel(const T* __restrict__ x, const float* __restrict__ w /*[27][16]*/, const float* __restrict__ bias,
                      T* __restrict__ y, int y_ld, int D, int H, int W, int ntx, int nty, int ntz, double* __restrict__ stats) {
  __shared__ float halo[6][10][10];
  __shared__ __align__(16) float sw[27][16];
  __shared__ float red[64];
  int t = blockIdx.x;
  const int x0 = (t % ntx) * 8; t /= ntx;
  const int y0 = (t % nty) * 8; t /= nty;
  const int z0 = (t % ntz) * 4; const int n = t / ntz;
  const T* xn = x + (size_t)n * D * H * W;
  for (int i = threadIdx.x; i < 600; i += 256) {
    const int hx = i % 10, hy = (i / 10) % 10, hz = i / 100;
    const int gx = x0 + hx - 1, gy = y0 + hy - 1, gz = z0 + hz - 1;
    float v = 0.f;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < D) v = to_f32<T>(xn[((size_t)gz * H + gy) * W + gx]);
    halo[hz][hy][hx] = v;
  }
  for (int i = threadIdx.x; i < 27 * 16; i += 256) sw[i / 16][i % 16] = w[i];
  __syncthreads();
  const int lx = threadIdx.x & 7, ly = (threadIdx.x >> 3) & 7, lz = threadIdx.x >> 6;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = bias ? bias[c] : 0.f;
#pragma unroll
  for (int kd = 0; kd < 3; ++kd)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float v = halo[lz + kd][ly + kh][lx + kw];
        const float4* wr = reinterpret_cast<const float4*>(&sw[(kd * 3 + kh) * 3 + kw][0]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wr[q];
          acc[4 * q] = fmaf(v, wv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, wv.w, acc[4 * q + 3]);
        }
      }
  const int gx = x0 + lx, gy = y0 + ly, gz = z0 + lz;
  float s = 0.f, ss = 0.f;
  if (gx < W && gy < H && gz < D) {
#pragma unroll
    for (int c = 0; c < 16; ++c) { s += acc[c]; ss += acc[c] * acc[c]; }
    T* dst = y + ((((size_t)n * D + gz) * H + gy) * W + gx) * y_ld;
    Vec8<T> o; o.set(acc); o.store(dst); o.set(acc + 8); o.store(dst + 8);
  }
  if (stats) block_stats_atomic(s, ss, stats + 2 * n, red);
}

template <typename T>
static int launch_cin1(const void* x, const void* w, const float* bias, void* y, int y_ld, int N, int D, int H, int W,
                       double* stats, cudaStream_t st) {
  const int ntx = (W + 7) / 8, nty = (H + 7) / 8, ntz = (D + 3) / 4;
  const long long blocks = (long long)ntx * nty * ntz * N;
  SEG3D_REQUIRE(blocks > 0 && blocks < (1ll << 31), "conv cin1: bad dims");
  conv3d_k3_cin1_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, (const float*)w, bias, (T*)y, y_ld, D, H, W, ntx, nty, ntz, stats);
  SEG3D_CHECK_LAUNCH("conv3d_k3_cin1_kernel");
  return SEG3D_OK;
}

template <typename T>
static int launch_simt(const ConvGeom& g, const void* x, const void* w, const float* bias, void* y,
                       double* stats, cudaStream_t st) {
  dim3 grid(g.N * g.tiles_per_sample, 1, 1), block(256);
  const T* xp = static_cast<const T*>(x); T* yp = static_cast<T*>(y); const float* wp = static_cast<const float*>(w);
  if (g.Ng >= 64) { grid.y = (g.Ng + 63) / 64; conv_simt_kernel<T, 64><<<grid, block, 0, st>>>(g, xp, wp, bias, yp, stats); }
  else if (g.Ng >= 32) { grid.y = (g.Ng + 31) / 32; conv_simt_kernel<T, 32><<<grid, block, 0, st>>>(g, xp, wp, bias, yp, stats); }
  else { grid.y = (g.Ng + 15) / 16; conv_simt_kernel<T, 16><<<grid, block, 0, st>>>(g, xp, wp, bias, yp, stats); }
  SEG3D_CHECK_LAUNCH("conv_simt_kernel");
  return SEG3D_OK;
}

int seg3d_conv_simt(int mode, int dtype, const void* x, int x_ld, int Cin, const void* w, const float* bias,
                    void* y, int y_ld, int Cout, int N, int D, int H, int W, double* stats, cudaStream_t st) {
  if (mode == SEG3D_CONV_K3 && Cin == 1 && Cout == 16 && x_ld == 1 && y_ld % 8 == 0 && ((uintptr_t)y) % 16 == 0) {
    SEG3D_DISPATCH_DTYPE(dtype, T, return launch_cin1<T>(x, w, bias, y, y_ld, N, D, H, W, stats, st));
  }
  ConvGeom g;
  g.mode = mode; g.N = N; g.D = D; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.x_ld = x_ld; g.y_ld = y_ld;
  g.Do = D; g.Ho = H; g.Wo = W; g.Ng = Cout;
  switch (mode) {
    case SEG3D_CONV_K3: g.K = 27 * Cin; break;
    case SEG3D_CONV_K1: g.K = Cin; break;
    case SEG3D_CONV_K2S2:
      SEG3D_REQUIRE(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "conv k2s2: odd input dims %d %d %d", D, H, W);
      g.K = 8 * Cin; g.Do = D / 2; g.Ho = H / 2; g.Wo = W / 2; break;
    case SEG3D_CONV_T2S2: g.K = Cin; g.Ng = 8 * Cout; break;
    default: seg3d_set_error("unknown conv mode %d", mode); return SEG3D_EINVAL;
  }
  const long long vps = (long long)g.Do * g.Ho * g.Wo;
  SEG3D_REQUIRE(vps > 0 && vps < (1ll << 31) && N > 0, "conv: bad dims");
  g.vps = (int)vps;
  g.tiles_per_sample = (g.vps + 127) / 128;
  const int esz = dtype == SEG3D_F32 ? 4 : 2;
  g.vec = (Cin % 8 == 0) && (x_ld % 8 == 0) && (((uintptr_t)x) % 16 == 0) && (esz == 2 || x_ld % 4 == 0);
  SEG3D_DISPATCH_DTYPE(dtype, T, return launch_simt<T>(g, x, w, bias, y, stats, st));
  return SEG3D_OK;
}
