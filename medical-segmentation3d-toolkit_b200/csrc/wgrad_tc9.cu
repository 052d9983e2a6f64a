// wgrad_tc9.cu - weight gradient of the k3 convolution for NARROW outputs (Cout = 16 or 32) on the tensor cores,
// nine in-plane taps per MMA.
//
//     dW[kd][kh][kw][ci][co] = sum_v dy[v][co] * x[v + (kd-1, kh-1, kw-1)][ci]
//                            = sum_u dy[u - (kh-1) e_y][co] * x[u + (kw-1) e_x + (kd-1) e_z][ci]          (u = v + (kh-1) e_y)
//
// is a GEMM over the voxels u (K) with both operands voxel-major / channel-contiguous, i.e. MN-major UMMA operands
// exactly as TMA delivers NDHWC rows.  With Cout <= 32 one M=128 MMA has room for several M-atoms: the kh shift goes
// on the A side (dy tile with a one-line y halo, M-atom stride = one 8-voxel y-line) and the kw shift on the B side
// (x tile with a one-voxel x halo, N-atom stride = one voxel row), so ONE MMA per K=16 step yields
//     D[(kh, co)][(kw, ci)]     for the nine taps of one kd
// instead of the three MMAs (one per kh) of conv3d_k3_wgrad_tc_kernel, whose M rows beyond Cout were idle.
// A CTA owns one kd and a K split of the 8 x 16 plane tiles (persistent, one wave), accumulates in TMEM over all its
// tiles and adds to the fp32 gradient with atomics once.
#include "tc_ptx.cuh"

namespace {

struct Wg9Params {
  int Cin, Cout, nblk, n_ci_blk;
  int D, H, W, N, ntx, nty, ntiles;
  int stages, a_bytes, b_bytes, a_tx, b_tx, tmem_cols;
  uint32_t idesc, a_step, b_step;          // descriptor advance per K=16 step (two y-lines), bytes >> 4
  uint32_t a_sbo, a_lbo, a_layout, b_sbo, b_lbo, b_layout;
};

template <typename T>
__global__ void __launch_bounds__(TC_THREADS)
conv3d_k3_wgrad9_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                           const Wg9Params p, float* __restrict__ dw) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + p.stages * p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + p.stages * (p.a_bytes + p.b_bytes) + 1024);   // +1 KB: shifted atoms over-read
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = full_bar + 8 * p.stages;
  const uint32_t done_bar = empty_bar + 8 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  // blockIdx.y -> (ci block, kd)
  const int kd = blockIdx.y % 3, cib = blockIdx.y / 3;
  const int ci0 = cib * p.nblk;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int t = tile;
      const int x0 = (t % p.ntx) * 8; t /= p.ntx;
      const int y0 = (t % p.nty) * 16; t /= p.nty;
      const int z = t % p.D; const int n = t / p.D;
      mbar_wait(empty_bar + 8 * stage, phase ^ 1);
      mbar_expect_tx_e(full_bar + 8 * stage, (uint32_t)(p.a_tx + p.b_tx));
      tma_load_5d_e(a_base + stage * p.a_bytes, &map_dy, full_bar + 8 * stage, 0, x0, y0 - 1, z, n);              // 8 x 18 lines
      tma_load_5d_e(b_base + stage * p.b_bytes, &map_x, full_bar + 8 * stage, ci0, x0 - 1, y0, z + kd - 1, n);    // 10 x 16
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0;
    const uint32_t hi_a = desc_hi(p.a_sbo, p.a_layout), hi_b = desc_hi(p.b_sbo, p.b_layout);
    const uint32_t lbo_a = ((p.a_lbo >> 4) & 0x3FFFu) << 16, lbo_b = ((p.b_lbo >> 4) & 0x3FFFu) << 16;
    uint32_t accumulate = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      mbar_wait(full_bar + 8 * stage, phase);
      tc_fence_after();
      const uint32_t lo_a = ((a_base + stage * p.a_bytes) >> 4) | lbo_a;
      const uint32_t lo_b = ((b_base + stage * p.b_bytes) >> 4) | lbo_b;
#pragma unroll
      for (int k = 0; k < 8; ++k)        // 128 voxels = 8 steps of K = 16 (two 8-voxel y-lines)
        tc_mma_f16_e(tmem_base, desc_pack(hi_a, lo_a + (uint32_t)k * p.a_step), desc_pack(hi_b, lo_b + (uint32_t)k * p.b_step),
                     p.idesc, k == 0 ? accumulate : 1u);
      accumulate = 1;
      tc_commit_e(empty_bar + 8 * stage);
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    tc_commit_e(done_bar);
  } else {
    // epilogue (once): TMEM lane = (M-atom j, co) with kh = 2 - j; column = kw * nblk + ci
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int j = row / p.Cout, co = row - j * p.Cout;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (blockIdx.x < p.ntiles) {
      for (int kw = 0; kw < 3; ++kw) {
        const int tap = (kd * 3 + (2 - j)) * 3 + kw;
        for (int c0 = 0; c0 < p.nblk; c0 += 16) {
          uint32_t v[16];
          tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kw * p.nblk + c0), v);
          tc_wait_ld();
          if (j < 3) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int ci = ci0 + c0 + jj;
              if (ci < p.Cin) atomicAdd(dw + ((size_t)tap * p.Cin + ci) * p.Cout + co, __uint_as_float(v[jj]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

}  // namespace

// dW (fp32 [27][Cin][Cout], accumulated).  Returns SEG3D_EUNSUPPORTED for shapes it does not take.
int seg3d_wgrad_tc9(int dtype, const void* x, int x_ld, int Cin, const void* dy, int dy_ld, int Cout, float* dw,
                    int N, int D, int H, int W, cudaStream_t st) {
  if (dtype != SEG3D_F16 && dtype != SEG3D_BF16) return SEG3D_EUNSUPPORTED;
  if ((Cout != 16 && Cout != 32) || Cin % 16 || x_ld % 8 || dy_ld % 8) return SEG3D_EUNSUPPORTED;
  if (env_int("SEG3D_TC_WGRAD9", 1) == 0) return SEG3D_EUNSUPPORTED;
  EncodeTiledFn encode = get_encode();
  if (!encode) return SEG3D_EUNSUPPORTED;
  Wg9Params p;
  memset(&p, 0, sizeof(p));
  p.Cin = Cin; p.Cout = Cout; p.D = D; p.H = H; p.W = W; p.N = N;
  p.nblk = Cin >= 64 ? 64 : Cin;
  if (Cin % p.nblk || (p.nblk != 16 && p.nblk != 32 && p.nblk != 64)) return SEG3D_EUNSUPPORTED;
  p.n_ci_blk = Cin / p.nblk;
  p.ntx = (W + 7) / 8; p.nty = (H + 15) / 16;
  const long long ntiles = (long long)p.ntx * p.nty * D * N;
  if (ntiles <= 0 || ntiles >= (1ll << 31)) return SEG3D_EUNSUPPORTED;
  p.ntiles = (int)ntiles;
  const int rba = Cout * 2, rbb = p.nblk * 2;               // bytes of one voxel row of A (dy) / B (x)
  p.a_tx = 8 * 18 * rba;
  p.b_tx = 10 * 16 * rbb;
  // M = 128 spans 128 / Cout M-atoms (one y-line apart); atoms beyond the third read lines past the 18 loaded ones:
  // valid shared memory (the next slab / the 1 KB tail), values only reach accumulator rows that are never read
  p.a_bytes = ((8 * (16 + 128 / Cout) * rba) + 1023) & ~1023;
  p.b_bytes = ((10 * 16 + 2) * rbb + 1023) & ~1023;
  p.stages = (200 * 1024) / (p.a_bytes + p.b_bytes); if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) return SEG3D_EUNSUPPORTED;
  p.a_sbo = 8 * rba; p.a_lbo = 8 * rba;                     // K-atom stride = M-atom stride = one 8-voxel y-line
  p.b_sbo = 10 * rbb; p.b_lbo = rbb;                        // K-atom stride = one 10-voxel halo line; N-atom stride = one voxel
  p.a_layout = rba == 64 ? 4u : 6u;
  p.b_layout = rbb == 128 ? 2u : (rbb == 64 ? 4u : 6u);
  p.a_step = (uint32_t)(2 * 8 * rba) >> 4;
  p.b_step = (uint32_t)(2 * 10 * rbb) >> 4;
  const int cols = 3 * p.nblk;
  p.tmem_cols = cols <= 64 ? 64 : (cols <= 128 ? 128 : 256);
  const uint32_t fmt = dtype == SEG3D_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const CUtensorMapDataType tdt = dtype == SEG3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap map_dy, map_x;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)dy_ld * 2, (cuuint64_t)W * dy_ld * 2, (cuuint64_t)H * W * dy_ld * 2, (cuuint64_t)D * H * W * dy_ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)Cout, 8, 18, 1, 1};
    CUresult r = encode(&map_dy, tdt, 5, const_cast<void*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        rba == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seg3d_set_error("wgrad_tc9: cuTensorMapEncodeTiled(dy) failed with %d", (int)r); return SEG3D_ECUDA; }
  }
  {
    const CUtensorMapSwizzle sw = rbb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rbb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2, (cuuint64_t)D * H * W * x_ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.nblk, 10, 16, 1, 1};
    CUresult r = encode(&map_x, tdt, 5, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seg3d_set_error("wgrad_tc9: cuTensorMapEncodeTiled(x) failed with %d", (int)r); return SEG3D_ECUDA; }
  }
  const int combos = p.n_ci_blk * 3;
  long long ksplit = (long long)seg3d_num_sms() / combos;        // one CTA per SM, one wave
  if (ksplit < 1) ksplit = 1;
  if (ksplit > ntiles) ksplit = ntiles;
  const size_t smem = 1024 + (size_t)p.stages * (p.a_bytes + p.b_bytes) + 1024 + (2 * p.stages + 1) * 8 + 64;
  dim3 grid((unsigned)ksplit, (unsigned)combos);
  cudaError_t e;
  if (dtype == SEG3D_BF16) {
    e = cudaFuncSetAttribute(conv3d_k3_wgrad9_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv3d_k3_wgrad9_tc_kernel<__nv_bfloat16><<<grid, TC_THREADS, smem, st>>>(map_dy, map_x, p, dw);
  } else {
    e = cudaFuncSetAttribute(conv3d_k3_wgrad9_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv3d_k3_wgrad9_tc_kernel<__half><<<grid, TC_THREADS, smem, st>>>(map_dy, map_x, p, dw);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { seg3d_set_error("conv3d_k3_wgrad9_tc_kernel launch failed: %s", cudaGetErrorString(e)); return SEG3D_ECUDA; }
  return SEG3D_OK;
}
