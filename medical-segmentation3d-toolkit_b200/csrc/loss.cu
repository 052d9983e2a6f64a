// loss.cu - fused Dice / focal reductions on probabilities (loss/multi_dice_loss.py,
// loss/binary_dice_loss.py, loss/focal_loss.py).  One pass over probs + labels per direction.
#include "common.cuh"

// terms[b][c] = { sum q t, sum q q, sum t t },  q = p [p > 1/C],  t = [target == c]
__global__ void __launch_bounds__(256)
dice_terms_kernel(const float* __restrict__ probs, const float* __restrict__ target, int C, long long n,
                  double* __restrict__ terms) {
  __shared__ double red[3][8];
  const int b = blockIdx.y, c = blockIdx.z;
  const float thr = 1.0f / (float)C;                        // multi_dice_loss.py:36: 1.0 / num_class (+ zeros, float32)
  const float* p = probs + ((size_t)b * C + c) * n;
  const float* t = target + (size_t)b * n;
  float I = 0.f, A = 0.f, T = 0.f;
  double dI = 0, dA = 0, dT = 0;
  int k = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float q = pv > thr ? pv : 0.f;                    // max over [1/C, p]: ties -> index 0 -> label 0
    const float tv = (t[i] == (float)c) ? 1.f : 0.f;
    I += q * tv; A += q * q; T += tv;
    if (++k == 64) { dI += I; dA += A; dT += T; I = A = T = 0.f; k = 0; }
  }
  dI += I; dA += A; dT += T;
  dI = warp_sum(dI); dA = warp_sum(dA); dT = warp_sum(dT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = dI; red[1][warp] = dA; red[2][warp] = dT; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0; for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    atomicAdd(terms + ((size_t)b * C + c) * 3 + threadIdx.x, s);
  }
}

extern "C" int seg3d_dice_terms(const float* probs, const float* target, int B, int C, int64_t n, double* terms, void* stream) {
  SEG3D_REQUIRE(probs && target && terms && B > 0 && C > 0 && n > 0, "dice_terms: bad arguments");
  int gx = (int)((n + 256 * 16 - 1) / (256 * 16)); if (gx < 1) gx = 1; if (gx > 592) gx = 592;
  dice_terms_kernel<<<dim3(gx, B, C), 256, 0, (cudaStream_t)stream>>>(probs, target, C, n, terms);
  SEG3D_CHECK_LAUNCH("dice_terms_kernel");
  return SEG3D_OK;
}

// grad[b][c][i] = m * (coef0 * t + coef1 * p)   with m = [p > 1/C]
__global__ void __launch_bounds__(256)
dice_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ target, int C, long long n,
                const float* __restrict__ coef, float* __restrict__ grad) {
  const int b = blockIdx.y, c = blockIdx.z;
  const float thr = 1.0f / (float)C;
  const float c0 = coef[((size_t)b * C + c) * 2], c1 = coef[((size_t)b * C + c) * 2 + 1];
  const float* p = probs + ((size_t)b * C + c) * n;
  const float* t = target + (size_t)b * n;
  float* g = grad + ((size_t)b * C + c) * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float tv = (t[i] == (float)c) ? 1.f : 0.f;
    g[i] = pv > thr ? fmaf(c1, pv, c0 * tv) : 0.f;
  }
}

extern "C" int seg3d_dice_bwd(const float* probs, const float* target, int B, int C, int64_t n,
                              const float* coef, float* grad, void* stream) {
  SEG3D_REQUIRE(probs && target && coef && grad && B > 0 && C > 0 && n > 0, "dice_bwd: bad arguments");
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8)); if (gx < 1) gx = 1; if (gx > 1184) gx = 1184;
  dice_bwd_kernel<<<dim3(gx, B, C), 256, 0, (cudaStream_t)stream>>>(probs, target, C, n, coef, grad);
  SEG3D_CHECK_LAUNCH("dice_bwd_kernel");
  return SEG3D_OK;
}

// focal_loss.py:45-59: p_t = p[target] + 1e-10; loss_i = -alpha[t] (1-p_t)^gamma log p_t
// partial[0] += sum loss_i;  partial[1] += number of voxels whose label lies outside [0, C) (they contribute nothing)
__global__ void __launch_bounds__(256)
focal_fwd_kernel(const float* __restrict__ probs, const float* __restrict__ target, int C, long long n,
                 const float* __restrict__ alpha, float gamma, double* __restrict__ partial) {
  __shared__ double red[8];
  const int b = blockIdx.y;
  const float* t = target + (size_t)b * n;
  double acc = 0;
  int bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cls = (int)(long long)t[i];                    // target.long()
    if (cls < 0 || cls >= C) { ++bad; continue; }            // the reference's one-hot gather raises here: counted, reported by the caller
    const float pt = probs[((size_t)b * C + cls) * n + i] + 1e-10f;
    const float lg = logf(pt);
    const float w = gamma > 0.f ? powf(1.f - pt, gamma) : 1.f;
    acc += (double)(-alpha[cls] * w * lg);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < 8; ++w) s += red[w]; atomicAdd(partial, s); }
  if (bad) atomicAdd(partial + 1, (double)bad);
}

extern "C" int seg3d_focal_fwd(const float* probs, const float* target, int B, int C, int64_t n,
                               const float* alpha, float gamma, double* partial, void* stream) {
  SEG3D_REQUIRE(probs && target && alpha && partial && B > 0 && C > 0 && n > 0, "focal_fwd: bad arguments");
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8)); if (gx < 1) gx = 1; if (gx > 592) gx = 592;
  focal_fwd_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(probs, target, C, n, alpha, gamma, partial);
  SEG3D_CHECK_LAUNCH("focal_fwd_kernel");
  return SEG3D_OK;
}

// d/dp_t [ -a (1-p)^g log p ] = a ( g (1-p)^(g-1) log p - (1-p)^g / p ), zero for the other classes
__global__ void __launch_bounds__(256)
focal_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ target, int C, long long n,
                 const float* __restrict__ alpha, float gamma, float scale, float* __restrict__ grad) {
  const int b = blockIdx.y;
  const float* t = target + (size_t)b * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cls = (int)(long long)t[i];                    // a label outside [0, C) matches no class: zero gradient
    for (int c = 0; c < C; ++c) {
      float g = 0.f;
      if (c == cls) {
        const float pt = probs[((size_t)b * C + c) * n + i] + 1e-10f;
        const float om = 1.f - pt;
        if (gamma > 0.f) g = alpha[c] * (gamma * powf(om, gamma - 1.f) * logf(pt) - powf(om, gamma) / pt);
        else g = -alpha[c] / pt;
      }
      grad[((size_t)b * C + c) * n + i] = g * scale;
    }
  }
}

extern "C" int seg3d_focal_bwd(const float* probs, const float* target, int B, int C, int64_t n,
                               const float* alpha, float gamma, float scale, float* grad, void* stream) {
  SEG3D_REQUIRE(probs && target && alpha && grad && B > 0 && C > 0 && n > 0, "focal_bwd: bad arguments");
  int gx = (int)((n + 256 * 4 - 1) / (256 * 4)); if (gx < 1) gx = 1; if (gx > 1184) gx = 1184;
  focal_bwd_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(probs, target, C, n, alpha, gamma, scale, grad);
  SEG3D_CHECK_LAUNCH("focal_bwd_kernel");
  return SEG3D_OK;
}

// ---- cross entropy on the network output (loss/cross_entropy_loss.py:5-18 -> nn.CrossEntropyLoss: the reference feeds the
// probabilities in as logits).  loss_i = w[t_i] * (logsumexp_c x_c - x_t); voxels with t_i == ignore_index (or a label
// outside [0, C)) contribute nothing.  partial[0] += sum loss_i, partial[1] += sum w[t_i]  (double).
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ target, int C, long long n,
              const float* __restrict__ weight, int ignore_index, double* __restrict__ partial, float* __restrict__ loss_map) {
  __shared__ double red[2][8];
  const int b = blockIdx.y;
  const float* t = target + (size_t)b * n;
  const float* xb = x + (size_t)b * C * n;
  double acc = 0, wacc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cls = (int)(long long)t[i];                    // target.long()
    float l = 0.f;
    if (cls != ignore_index && cls >= 0 && cls < C) {
      float m = xb[i];
      for (int c = 1; c < C; ++c) m = fmaxf(m, xb[(size_t)c * n + i]);
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += expf(xb[(size_t)c * n + i] - m);
      const float w = weight ? weight[cls] : 1.f;
      l = w * (m + logf(s) - xb[(size_t)cls * n + i]);
      acc += (double)l; wacc += (double)w;
    }
    if (loss_map) loss_map[(size_t)b * n + i] = l;
  }
  acc = warp_sum(acc); wacc = warp_sum(wacc);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc; red[1][threadIdx.x >> 5] = wacc; }
  __syncthreads();
  if (threadIdx.x < 2) { double s = 0; for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w]; atomicAdd(partial + threadIdx.x, s); }
}

extern "C" int seg3d_ce_fwd(const float* logits, const float* target, int B, int C, int64_t n, const float* weight,
                            int ignore_index, double* partial, float* loss_map, void* stream) {
  SEG3D_REQUIRE(logits && target && partial && B > 0 && C > 0 && n > 0, "ce_fwd: bad arguments");
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8)); if (gx < 1) gx = 1; if (gx > 592) gx = 592;
  ce_fwd_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(logits, target, C, n, weight, ignore_index, partial, loss_map);
  SEG3D_CHECK_LAUNCH("ce_fwd_kernel");
  return SEG3D_OK;
}

// grad[b][c][i] = scale * g_i * w[t_i] * (softmax_c(x) - [c == t_i]),  g_i = grad_map[b][i] (or 1), 0 for ignored voxels
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ target, int C, long long n,
              const float* __restrict__ weight, int ignore_index, float scale, const float* __restrict__ grad_map,
              float* __restrict__ grad) {
  const int b = blockIdx.y;
  const float* t = target + (size_t)b * n;
  const float* xb = x + (size_t)b * C * n;
  float* gb = grad + (size_t)b * C * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cls = (int)(long long)t[i];
    if (cls == ignore_index || cls < 0 || cls >= C) {
      for (int c = 0; c < C; ++c) gb[(size_t)c * n + i] = 0.f;
      continue;
    }
    float m = xb[i];
    for (int c = 1; c < C; ++c) m = fmaxf(m, xb[(size_t)c * n + i]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(xb[(size_t)c * n + i] - m);
    const float k = scale * (weight ? weight[cls] : 1.f) * (grad_map ? grad_map[(size_t)b * n + i] : 1.f);
    const float inv = 1.f / s;
    for (int c = 0; c < C; ++c) {
      const float p = expf(xb[(size_t)c * n + i] - m) * inv;
      gb[(size_t)c * n + i] = k * (p - (c == cls ? 1.f : 0.f));
    }
  }
}

extern "C" int seg3d_ce_bwd(const float* logits, const float* target, int B, int C, int64_t n, const float* weight,
                            int ignore_index, float scale, const float* grad_map, float* grad, void* stream) {
  SEG3D_REQUIRE(logits && target && grad && B > 0 && C > 0 && n > 0, "ce_bwd: bad arguments");
  int gx = (int)((n + 256 * 4 - 1) / (256 * 4)); if (gx < 1) gx = 1; if (gx > 1184) gx = 1184;
  ce_bwd_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(logits, target, C, n, weight, ignore_index, scale, grad_map, grad);
  SEG3D_CHECK_LAUNCH("ce_bwd_kernel");
  return SEG3D_OK;
}
