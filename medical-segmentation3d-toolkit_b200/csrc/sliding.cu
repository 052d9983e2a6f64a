// sliding.cu - device side of the patch-partitioned sliding window (core/seg_infer.py:208-339):
// crop + normalise patches, accumulate probability patches into the volume accumulators,
// normalise by the overlap count and take the first-argmax mask.  All HBM-bound.
#include "common.cuh"

// ---- per-patch sum / sum of squares (AdaptiveNormalizer, normalizer.py:59) -------------------
__global__ void __launch_bounds__(256)
patch_stats_kernel(const float* __restrict__ vol, int Z, int Y, int X, const int32_t* __restrict__ starts,
                   int pz, int py, int px, double* __restrict__ stats) {
  __shared__ double red[16];
  const int n = blockIdx.y;
  const int x0 = starts[3 * n], y0 = starts[3 * n + 1], z0 = starts[3 * n + 2];
  const long long nv = (long long)pz * py * px;
  double s = 0.0, ss = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % px); const long long t = i / px; const int yy = (int)(t % py), z = (int)(t / py);
    const double v = (double)vol[((size_t)(z0 + z) * Y + (y0 + yy)) * X + (x0 + x)];
    s += v; ss += v * v;
  }
  s = warp_sum(s); ss = warp_sum(ss);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[warp] = s; red[8 + warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += red[w]; b += red[8 + w]; }
    atomicAdd(stats + 2 * n, a); atomicAdd(stats + 2 * n + 1, b);
  }
}

extern "C" int seg3d_patch_stats(const float* vol, int Z, int Y, int X, const int32_t* starts, int N,
                                 int pz, int py, int px, double* stats, void* stream) {
  SEG3D_REQUIRE(vol && starts && stats && N > 0 && pz > 0 && py > 0 && px > 0, "patch_stats: bad arguments");
  const long long nv = (long long)pz * py * px;
  int gx = (int)((nv + 256 * 8 - 1) / (256 * 8)); if (gx < 1) gx = 1; if (gx > 1024) gx = 1024;
  patch_stats_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(vol, Z, Y, X, starts, pz, py, px, stats);
  SEG3D_CHECK_LAUNCH("patch_stats_kernel");
  return SEG3D_OK;
}

// ---- crop + normalise (seg_infer.py:221-226, image_tools.py:221-238) -------------------------
template <typename T>
__global__ void __launch_bounds__(256)
patch_gather_kernel(const float* __restrict__ vol, int Z, int Y, int X, const int32_t* __restrict__ starts,
                    int pz, int py, int px, int norm, float mean, float stddev, int clip, float lo, float hi,
                    const double* __restrict__ stats, T* __restrict__ out, int row_pitch, int x_off) {
  const int n = blockIdx.y;
  const int x0 = starts[3 * n], y0 = starts[3 * n + 1], z0 = starts[3 * n + 2];
  const long long nv = (long long)pz * py * px;
  if (norm == SEG3D_NORM_ADAPTIVE) {
    // np.mean / np.std of the float32 crop (population std), stddev = max(std, 1e-6)
    const double m = stats[2 * n] / (double)nv;
    double var = stats[2 * n + 1] / (double)nv - m * m; if (var < 0) var = 0;
    mean = (float)m; stddev = fmaxf((float)sqrt(var), 1e-6f);
  }
  T* on = out + (size_t)n * pz * py * row_pitch + x_off;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % px); const long long t = i / px; const int yy = (int)(t % py), z = (int)(t / py);
    float v = vol[((size_t)(z0 + z) * Y + (y0 + yy)) * X + (x0 + x)];
    if (norm != SEG3D_NORM_NONE) {
      v = __fdiv_rn(__fsub_rn(v, mean), stddev);           // numpy float32: (a - mean) / std
      if (clip) { v = v < lo ? lo : v; v = v > hi ? hi : v; }
    }
    on[(size_t)(z * py + yy) * row_pitch + x] = from_f32<T>(v);
  }
}

// Row-padded output with 16-byte words: a thread owns ONE aligned word of 8 stored values of a padded row (columns 8j .. 8j+7 =
// x indices 8j - x_off ..), reads its up to 8 voxels and writes the word once - columns outside [x_off, x_off + px) are written as
// zeros, which is exactly what the padded layout wants there.  32-bit index arithmetic.  (2-byte storage types, row_pitch % 8 == 0.)
template <typename T>
__global__ void __launch_bounds__(256)
patch_gather_rows_w8_kernel(const float* __restrict__ vol, int Z, int Y, int X, const int32_t* __restrict__ starts,
                            int pz, int py, int px, int norm, float mean, float stddev, int clip, float lo, float hi,
                            const double* __restrict__ stats, T* __restrict__ out, int row_pitch, int x_off) {
  const int n = blockIdx.y;
  const int x0 = starts[3 * n], y0 = starts[3 * n + 1], z0 = starts[3 * n + 2];
  if (norm == SEG3D_NORM_ADAPTIVE) {
    const double nv = (double)pz * py * px;
    const double m = stats[2 * n] / nv;
    double var = stats[2 * n + 1] / nv - m * m; if (var < 0) var = 0;
    mean = (float)m; stddev = fmaxf((float)sqrt(var), 1e-6f);
  }
  const unsigned wpr = (unsigned)row_pitch >> 3;                    // words per row
  const unsigned nwords = (unsigned)pz * py * wpr;
  T* on = out + (size_t)n * pz * py * row_pitch;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) {
    const unsigned j = i % wpr, row = i / wpr;
    const unsigned yy = row % py, z = row / py;
    const float* src = vol + ((size_t)(z0 + z) * Y + (y0 + yy)) * X + x0;
    const int xb = (int)(8 * j) - x_off;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int x = xb + k;
      float v = 0.f;
      if (x >= 0 && x < px) {
        v = src[x];
        if (norm != SEG3D_NORM_NONE) {
          v = __fdiv_rn(__fsub_rn(v, mean), stddev);           // numpy float32: (a - mean) / std
          if (clip) { v = v < lo ? lo : v; v = v > hi ? hi : v; }
        }
      }
      f[k] = v;
    }
    Vec8<T> w; w.set(f); w.store(on + (size_t)row * row_pitch + 8 * j);
  }
}

extern "C" int seg3d_patch_gather_rows(const float* vol, int Z, int Y, int X, const int32_t* starts, int N,
                                       int pz, int py, int px, int norm, float mean, float stddev, int clip,
                                       float clip_lo, float clip_hi, const double* stats, int dtype, void* out,
                                       int row_pitch, int x_off, void* stream) {
  SEG3D_REQUIRE(vol && starts && out && N > 0 && pz > 0 && py > 0 && px > 0, "patch_gather: bad arguments");
  SEG3D_REQUIRE(pz <= Z && py <= Y && px <= X, "patch_gather: patch larger than volume");
  SEG3D_REQUIRE(norm != SEG3D_NORM_ADAPTIVE || stats, "patch_gather: adaptive normaliser needs stats");
  SEG3D_REQUIRE(norm != SEG3D_NORM_FIXED || stddev > 0.f, "patch_gather: stddev must be positive");
  SEG3D_REQUIRE(x_off >= 0 && row_pitch >= x_off + px, "patch_gather: row pitch smaller than offset + width");
  const long long nv = (long long)pz * py * px;
  if ((dtype == SEG3D_F16 || dtype == SEG3D_BF16) && x_off > 0 && row_pitch % 8 == 0 && row_pitch >= x_off + px && ((uintptr_t)out) % 16 == 0 &&
      (long long)pz * py * (row_pitch / 8) < (1ll << 31)) {
    const long long nwords = (long long)pz * py * (row_pitch / 8);
    int gxw = (int)((nwords + 255) / 256); if (gxw < 1) gxw = 1; if (gxw > 4096) gxw = 4096;
    dim3 gridw(gxw, N);
    if (dtype == SEG3D_F16)
      patch_gather_rows_w8_kernel<__half><<<gridw, 256, 0, (cudaStream_t)stream>>>(vol, Z, Y, X, starts, pz, py, px, norm, mean, stddev, clip,
                                                                                   clip_lo, clip_hi, stats, (__half*)out, row_pitch, x_off);
    else
      patch_gather_rows_w8_kernel<__nv_bfloat16><<<gridw, 256, 0, (cudaStream_t)stream>>>(vol, Z, Y, X, starts, pz, py, px, norm, mean, stddev, clip,
                                                                                          clip_lo, clip_hi, stats, (__nv_bfloat16*)out, row_pitch, x_off);
    SEG3D_CHECK_LAUNCH("patch_gather_rows_w8_kernel");
    return SEG3D_OK;
  }
  int gx = (int)((nv + 256 * 4 - 1) / (256 * 4)); if (gx < 1) gx = 1; if (gx > 2048) gx = 2048;
  dim3 grid(gx, N);
  SEG3D_DISPATCH_DTYPE(dtype, T, (patch_gather_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
      vol, Z, Y, X, starts, pz, py, px, norm, mean, stddev, clip, clip_lo, clip_hi, stats, (T*)out, row_pitch, x_off)));
  SEG3D_CHECK_LAUNCH("patch_gather_kernel");
  return SEG3D_OK;
}

extern "C" int seg3d_patch_gather(const float* vol, int Z, int Y, int X, const int32_t* starts, int N,
                                  int pz, int py, int px, int norm, float mean, float stddev, int clip,
                                  float clip_lo, float clip_hi, const double* stats, int dtype, void* out, void* stream) {
  return seg3d_patch_gather_rows(vol, Z, Y, X, starts, N, pz, py, px, norm, mean, stddev, clip, clip_lo, clip_hi, stats, dtype, out,
                                 px, 0, stream);
}

// ---- acc[c][region] += probs (add_image_region, image_tools.py:435-452) ----------------------
// Patches of one launch may overlap (stride < size, or the clamped last box), so the adds are
// red.global.add.f32; the accumulation order differs from the reference's sequential order,
// which is inside the float tolerance of the parity bar.
__global__ void __launch_bounds__(256)
blend_accumulate_kernel(const float* __restrict__ probs, int C, int pz, int py, int px,
                        const int32_t* __restrict__ starts, float* __restrict__ acc, int Z, int Y, int X) {
  const int n = blockIdx.y;
  const int x0 = starts[3 * n], y0 = starts[3 * n + 1], z0 = starts[3 * n + 2];
  const long long nv = (long long)pz * py * px, tot = nv * C;
  const float* pn = probs + (size_t)n * tot;
  const size_t vsz = (size_t)Z * Y * X;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / nv); const long long r = i - (long long)c * nv;
    const int x = (int)(r % px); const long long t = r / px; const int yy = (int)(t % py), z = (int)(t / py);
    atomicAdd(acc + c * vsz + ((size_t)(z0 + z) * Y + (y0 + yy)) * X + (x0 + x), pn[i]);
  }
}

// four x-consecutive voxels per thread: one 16-byte load and one red.global.add.v4.f32 (px, X and every x0 multiples of 4,
// 16-byte aligned pointers), 32-bit index arithmetic
__global__ void __launch_bounds__(256)
blend_accumulate_v4_kernel(const float4* __restrict__ probs, int C, int pz, int py, int px4,
                           const int32_t* __restrict__ starts, float* __restrict__ acc, int Z, int Y, int X) {
  const int n = blockIdx.y;
  const int x0 = starts[3 * n], y0 = starts[3 * n + 1], z0 = starts[3 * n + 2];
  const unsigned tot4 = (unsigned)C * pz * py * px4;
  const float4* pn = probs + (size_t)n * tot4;
  const size_t vsz = (size_t)Z * Y * X;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < tot4; i += gridDim.x * blockDim.x) {
    const unsigned x4 = i % px4; unsigned t = i / px4;
    const unsigned yy = t % py; t /= py;
    const unsigned z = t % pz; const unsigned c = t / pz;
    const float4 v = pn[i];
    float* dst = acc + c * vsz + ((size_t)(z0 + z) * Y + (y0 + yy)) * X + (x0 + 4 * x4);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

extern "C" int seg3d_blend_accumulate(const float* probs, int N, int C, int pz, int py, int px,
                                      const int32_t* starts, float* acc, int Z, int Y, int X, int starts_x_mult4, void* stream) {
  SEG3D_REQUIRE(probs && starts && acc && N > 0 && C > 0, "blend_accumulate: bad arguments");
  const long long tot = (long long)pz * py * px * C;
  const bool v4 = starts_x_mult4 && px % 4 == 0 && X % 4 == 0 && ((uintptr_t)probs) % 16 == 0 && ((uintptr_t)acc) % 16 == 0 && tot / 4 < (1ll << 31);
  if (v4) {
    int gx = (int)((tot / 4 + 256 * 4 - 1) / (256 * 4)); if (gx < 1) gx = 1; if (gx > 4096) gx = 4096;
    blend_accumulate_v4_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>((const float4*)probs, C, pz, py, px / 4, starts, acc, Z, Y, X);
  } else {
    int gx = (int)((tot + 256 * 4 - 1) / (256 * 4)); if (gx < 1) gx = 1; if (gx > 4096) gx = 4096;
    blend_accumulate_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(probs, C, pz, py, px, starts, acc, Z, Y, X);
  }
  SEG3D_CHECK_LAUNCH("blend_accumulate_kernel");
  return SEG3D_OK;
}

// ---- acc *= float32(1/count); mask = first argmax (seg_infer.py:325-327,336-338) --------------
template <int VEC>
__global__ void __launch_bounds__(256)
blend_finalize_kernel(float* __restrict__ acc, int C, int Z, int Y, int X, int z0, int z1, const int32_t* __restrict__ cx,
                      const int32_t* __restrict__ cy, const int32_t* __restrict__ cz, int8_t* __restrict__ mask) {
  const size_t vsz = (size_t)Z * Y * X;
  const size_t vbeg = (size_t)z0 * Y * X;
  const size_t nvec = (size_t)(z1 - z0) * Y * X / VEC;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const size_t v0 = vbeg + i * VEC;
    const int x0 = (int)(v0 % X); const size_t t = v0 / X; const int y = (int)(t % Y), z = (int)(t / Y);
    const float cyz = (float)(cy[y] * cz[z]);
    float best[VEC]; int8_t arg[VEC];
    for (int c = 0; c < C; ++c) {
      float p[VEC];
      if (VEC == 4) { const float4 q = *reinterpret_cast<const float4*>(acc + c * vsz + v0); p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
      else p[0] = acc[c * vsz + v0];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float cnt = cyz * (float)cx[x0 + j];
        p[j] = p[j] * __fdiv_rn(1.0f, cnt);                 // Cast(1.0 / count, float32) then multiply
        if (c == 0 || p[j] > best[j]) { best[j] = p[j]; arg[j] = (int8_t)c; }   // strict >: lowest index wins ties
      }
      if (VEC == 4) *reinterpret_cast<float4*>(acc + c * vsz + v0) = make_float4(p[0], p[1], p[2], p[3]);
      else acc[c * vsz + v0] = p[0];
    }
    if (mask) {
      if (VEC == 4) *reinterpret_cast<char4*>(mask + v0) = make_char4(arg[0], arg[1], arg[2], arg[3]);
      else mask[v0] = arg[0];
    }
  }
}

extern "C" int seg3d_blend_finalize_argmax_z(float* acc, int C, int Z, int Y, int X, int z0, int z1, const int32_t* cx,
                                             const int32_t* cy, const int32_t* cz, int8_t* mask, void* stream) {
  SEG3D_REQUIRE(acc && cx && cy && cz && C > 0 && C <= 127 && Z > 0 && Y > 0 && X > 0, "blend_finalize: bad arguments");
  SEG3D_REQUIRE(0 <= z0 && z0 < z1 && z1 <= Z, "blend_finalize: bad z range [%d,%d) of %d", z0, z1, Z);
  const size_t vsz = (size_t)(z1 - z0) * Y * X;
  const int sms = seg3d_num_sms();
  const bool v4 = (X % 4 == 0) && (((uintptr_t)acc) % 16 == 0) && (!mask || ((uintptr_t)mask) % 4 == 0);
  const size_t nvec = v4 ? vsz / 4 : vsz;
  size_t want = (nvec + 255) / 256; int gx = (int)(want > (size_t)16 * sms ? (size_t)16 * sms : (want < 1 ? 1 : want));
  if (v4) blend_finalize_kernel<4><<<gx, 256, 0, (cudaStream_t)stream>>>(acc, C, Z, Y, X, z0, z1, cx, cy, cz, mask);
  else    blend_finalize_kernel<1><<<gx, 256, 0, (cudaStream_t)stream>>>(acc, C, Z, Y, X, z0, z1, cx, cy, cz, mask);
  SEG3D_CHECK_LAUNCH("blend_finalize_kernel");
  return SEG3D_OK;
}

extern "C" int seg3d_blend_finalize_argmax(float* acc, int C, int Z, int Y, int X, const int32_t* cx,
                                           const int32_t* cy, const int32_t* cz, int8_t* mask, void* stream) {
  return seg3d_blend_finalize_argmax_z(acc, C, Z, Y, X, 0, Z, cx, cy, cz, mask, stream);
}
