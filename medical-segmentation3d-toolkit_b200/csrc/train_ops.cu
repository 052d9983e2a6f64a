// train_ops.cu - the parameter side of the training step (reference core/seg_train.py:83,119-127):
//   seg3d_gather_pack : table-driven strided gather with a cast - ONE launch re-packs every convolution weight of the network
//                       from the reference's fp32 OIDHW / IODHW parameters into the kernels' layouts (tensor-core [tap][Cout][Cin]
//                       in the storage type, SIMT [tap][Cin][Cout], split hi/lo halves, the folded narrow-output layout, the
//                       flipped / transposed data-gradient weights), and one more launch turns the kernel-layout weight
//                       gradients back into parameter layout;
//   seg3d_adam_step   : torch.optim.Adam's update (the reference's optimiser, lr / betas / eps / weight decay) over one flat
//                       fp32 range holding every parameter of the network.
// Both are pure streaming kernels over ~58 MB; they replace ~250 tiny framework launches per step.
#include "common.cuh"

namespace {

__device__ __forceinline__ void store_as(void* dst, long long i, int dtype, float v) {
  if (dtype == SEG3D_F32) reinterpret_cast<float*>(dst)[i] = v;
  else if (dtype == SEG3D_F16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
gather_pack_kernel(const seg3d_pack_entry* __restrict__ table, int n_entries) {
  const seg3d_pack_entry e = table[blockIdx.y];
  const long long n = (long long)e.size[0] * e.size[1] * e.size[2] * e.size[3] * e.size[4];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i, so = e.src_base, d_o = e.dst_base;
    bool inside = true;
#pragma unroll
    for (int d = 4; d >= 0; --d) {
      const int idx = (int)(r % e.size[d]); r /= e.size[d];
      inside = inside && idx < e.limit[d];
      so += (long long)idx * e.src_stride[d];
      d_o += (long long)idx * e.dst_stride[d];
    }
    float v = inside ? e.src[so] : 0.f;
    if (e.kind != SEG3D_PACK_PLAIN) {              // split operands: hi = f16(w), lo = f16(w - hi)
      const float hi = __half2float(__float2half_rn(v));
      v = e.kind == SEG3D_PACK_SPLIT_HI ? hi : v - hi;
    }
    store_as(e.dst, d_o, e.dtype, v);
  }
}

// torch/optim/adam.py::_single_tensor_adam (no amsgrad, no maximize):
//   g += wd * p;  m.lerp_(g, 1 - b1);  v = b2 * v + (1 - b2) * g * g;  p -= step_size * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                 float b1, float b2, float eps, float wd, float step_size, float bc2_sqrt) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = wd != 0.f ? fmaf(wd, pa[j], ga[j]) : ga[j];
      ma[j] = ma[j] + (1.f - b1) * (gr - ma[j]);
      va[j] = fmaf(1.f - b2, gr * gr, b2 * va[j]);
      pa[j] -= step_size * (ma[j] / (sqrtf(va[j]) / bc2_sqrt + eps));
    }
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gr = wd != 0.f ? fmaf(wd, p[i], g[i]) : g[i];
    const float mi = m[i] + (1.f - b1) * (gr - m[i]);
    const float vi = fmaf(1.f - b2, gr * gr, b2 * v[i]);
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

}  // namespace

extern "C" int seg3d_gather_pack(const seg3d_pack_entry* table, int n_entries, int64_t max_elems, void* stream) {
  SEG3D_REQUIRE(table && n_entries > 0 && max_elems > 0, "gather_pack: bad arguments");
  long long want = (max_elems + 256 * 4 - 1) / (256 * 4);
  const int cap = 2 * seg3d_num_sms();
  const int gx = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  gather_pack_kernel<<<dim3(gx, n_entries), 256, 0, (cudaStream_t)stream>>>(table, n_entries);
  SEG3D_CHECK_LAUNCH("gather_pack_kernel");
  return SEG3D_OK;
}

extern "C" int seg3d_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                               float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  SEG3D_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad arguments");
  SEG3D_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, "adam_step: pointers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  long long want = ((n >> 2) + 255) / 256;
  const int cap = 8 * seg3d_num_sms();
  const int gx = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  adam_step_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, weight_decay, step_size, bc2_sqrt);
  SEG3D_CHECK_LAUNCH("adam_step_kernel");
  return SEG3D_OK;
}
