// Probe: does cuTensorMapEncodeTiled accept a map whose dimension-1 stride (16 B) is smaller than the extent of dimension 0
// (16 halfs = 32 B), i.e. overlapping windows, and does a box load deliver the expected elements?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/probes/tma_overlap_probe.cu -lcuda && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, __half* out, int c1, int c2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(16 * 4 * 6 * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(&map), "r"(b), "r"(0), "r"(c1), "r"(c2) : "memory");
  }
  __syncthreads();
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b) : "memory");
  for (int i = threadIdx.x; i < 16 * 4 * 6; i += blockDim.x) out[i] = reinterpret_cast<__half*>(smem)[i];
}

int main() {
  const int W = 32, H = 8, Wp = W + 16;
  std::vector<__half> h(H * Wp);
  for (int y = 0; y < H; ++y) for (int i = 0; i < Wp; ++i) h[y * Wp + i] = __float2half((float)(y * 100 + i));
  __half *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 16 * 4 * 6 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t dims[3] = {16, (cuuint64_t)W / 8, (cuuint64_t)H};
  cuuint64_t strides[2] = {16, (cuuint64_t)Wp * 2};
  cuuint32_t box[3] = {16, 4, 6}, estr[3] = {1, 1, 1};
  cuInit(0);
  CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d + 8, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (overlapping windows, stride 16 B < extent 32 B): CUresult %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  probe<<<1, 128, 4096>>>(map, o, 1, -1);      // segments 1..4 (segment 4 is out of bounds), rows -1..4 (row -1 out of bounds)
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<__half> res(16 * 4 * 6);
  cudaMemcpy(res.data(), o, res.size() * 2, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < 6; ++y) for (int s = 0; s < 4; ++s) for (int i = 0; i < 16; ++i) {
    const int gy = y - 1, gs = s + 1;
    float want = (gy < 0 || gy >= H || gs >= W / 8) ? 0.f : (float)(gy * 100 + 8 + 8 * gs + i);
    float got = __half2float(res[(y * 4 + s) * 16 + i]);
    if (want != got) { if (bad < 8) printf("mismatch y %d s %d i %d: want %g got %g\n", y, s, i, want, got); ++bad; }
  }
  printf("mismatches: %d\n", bad);
  return bad != 0;
}
