// conv_tc_cin1t.cu - input block (reference vnet_inblock.py:9: Conv3d(1, 16, k3, p1) + GroupNorm(1,16) + ReLU) as a
// banded-Toeplitz GEMM fed entirely by TMA: no thread ever touches an input voxel.
//
// With ONE input channel the implicit GEMM of conv_tc.cu has K = 27 and nothing for TMA to deliver as a K-major row; the
// first version of this layer (conv_tc_cin1.cu) built 64-byte im2col rows with four builder warps and ran at 0.28 of the
// HBM roofline, bound by those warps.  Here the GEMM row is a SEGMENT of 8 consecutive output voxels along x:
//
//     D[segment][(xo, co)]  =  sum over (kd, kh)  A_(kd,kh)[segment][xi = 0..15]  *  T_(kd,kh)[xi][(xo, co)]
//
//   A_(kd,kh)[segment]  = the 16 consecutive input values x[z+kd-1][y+kh-1][8s-1 .. 8s+14]        (K = 16: ONE k-step)
//   T_(kd,kh)[xi][xo,co] = w[kd][kh][kw = xi - xo][co] if 0 <= xi - xo <= 2 else 0                 (banded Toeplitz, N = 128)
//
// so the A operand of a (kd,kh) tap is a [128 segments x 32 bytes] K-major slab that one TMA box load delivers from a tensor
// map with OVERLAPPING windows (dimension 0 = 16 elements, dimension 1 = segments with a stride of 8 elements).  The input
// lives in a row-padded buffer [N][D][H][W+16] with x index i at column i+9, which makes every window start 16-byte aligned
// and supplies the left/right zero halo; the y/z halo is the TMA unit's out-of-bounds zero fill.  A CTA owns a 32(x) x 32(y)
// column tile (rows = (y, segment-in-tile), so the kh tap is a row shift of 4 = 128 bytes of the descriptor start address),
// marches along z with a ring of halo planes (each loaded once: 4.3 KB for 1024 voxels), and issues 9 MMAs
// 128 x 128 x 16 per output plane (weights rounded to the storage type, as in every other layer; SEG3D_CIN1_LO=1 splits them into
// two half-precision terms hi(T) + lo(T) accumulated into the same 128 TMEM columns: fp32-accurate weights at twice the MMAs).  Eight epilogue warps drain four rotating
// accumulators: every thread owns half a segment and writes its 4 voxels x 16 channels as full 32-byte sectors.  The banded
// matrix wastes tensor flops (82 GFLOP issued for 15 real ones at batch 20) to buy a layer without a single per-voxel
// instruction outside the epilogue: algorithmic bytes = 2 B read + 32 B written per voxel.
//
// Epilogue modes: 0 = store conv + bias and accumulate the GroupNorm sums; 1 = sums only (nothing stored);
// 2 = store relu(GroupNorm(conv + bias)) from the finished sums.  Modes 1 + 2 are the inference schedule: the layer is run
// twice (its input is 1/16 of its output) and the raw tensor plus the whole GroupNorm-apply pass never touch HBM.
#include "tc_ptx.cuh"

namespace {

constexpr int CT_THREADS = 320;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue
constexpr int CT_PLANE_ROWS = 136;       // (32 + 2 halo) y rows x 4 segments
constexpr int CT_PLANE_BYTES = 4608;     // 136 x 32 B = 4352, padded to a multiple of the 256-byte swizzle period
constexpr int CT_W_SLAB = 8192;          // one (kd,kh) Toeplitz slab: 256 rows x 32 B
constexpr int CT_MAXRING = 10;
constexpr int CT_NB = 4;               // TMEM accumulators of 128 columns

struct CtParams {
  int D, H, W, N, y_ld;
  int ntx, nty, nseg, lseg, nitems, ring;
  int wide, epi_mode, lo;            // lo: N = 256 = [hi(T) | lo(T)] (fp32-accurate weights); 0: N = 128, weights rounded to T
  uint32_t idesc;
  float gn_eps;
  double gn_count;
};

template <typename T, int EPI>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv3d_k3_cin1_toeplitz_kernel(const __grid_constant__ CUtensorMap map_x, const float* __restrict__ w /*[27][16]*/,
                               const float* __restrict__ bias, T* __restrict__ y, const CtParams p, double* __restrict__ stats,
                               const float* __restrict__ gamma, const float* __restrict__ beta) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ float sw_f[27 * 16 + 16];                                   // the fp32 weights and the bias, staged once
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_base = smem_base;                                     // [9][256 rows][32 B], 32B-swizzled
  const uint32_t a_base = smem_base + 9 * CT_W_SLAB;                     // [ring][136 rows][32 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + 9 * CT_W_SLAB + p.ring * CT_PLANE_BYTES);
  const uint32_t full_bar = smem_u32(bars);                              // [ring]
  const uint32_t empty_bar = full_bar + 8 * CT_MAXRING;                  // [ring]
  const uint32_t tfull_bar = empty_bar + 8 * CT_MAXRING;                 // [CT_NB]
  const uint32_t tempty_bar = tfull_bar + 8 * CT_NB;                     // [CT_NB]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CT_MAXRING + 2 * CT_NB);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < p.ring; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int b = 0; b < CT_NB; ++b) { mbar_init(tfull_bar + 8 * b, 1); mbar_init(tempty_bar + 8 * b, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 27 * 16 + 16; i += CT_THREADS) sw_f[i] = i < 432 ? w[i] : (bias ? bias[i - 432] : 0.f);
  __syncthreads();
  // Toeplitz weight slabs, built once per CTA: slab (kd,kh), row n = half*128 + xo*16 + co, 16 K values xi
  for (int idx = threadIdx.x; idx < 9 * 256; idx += CT_THREADS) {
    const int slab = idx >> 8, row = idx & 255;
    const int lo_half = row >> 7, xo = (row >> 4) & 7, co = row & 15;
    uint32_t h[16];
#pragma unroll
    for (int xi = 0; xi < 16; ++xi) {
      const int kw = xi - xo;
      const float wv = (kw >= 0 && kw <= 2) ? sw_f[(slab * 3 + kw) * 16 + co] : 0.f;
      T hv = from_f32<T>(wv);
      if (lo_half) hv = from_f32<T>(wv - to_f32<T>(hv));
      h[xi] = (uint32_t)(*reinterpret_cast<unsigned short*>(&hv));
    }
    uint8_t* rowp = smem_al + slab * CT_W_SLAB + row * 32;
    const int sw = (row >> 2) & 1;                                         // 32B swizzle: 16-byte chunk ^= address bit 7
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint4 v;
      v.x = h[8 * c] | (h[8 * c + 1] << 16); v.y = h[8 * c + 2] | (h[8 * c + 3] << 16);
      v.z = h[8 * c + 4] | (h[8 * c + 5] << 16); v.w = h[8 * c + 6] | (h[8 * c + 7] << 16);
      *reinterpret_cast<uint4*>(rowp + ((c ^ sw) << 4)) = v;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===== TMA producer: halo planes zs-1 .. zs+L of every work item, in ring order =====
    int slot = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      int t = item;
      const int seg = t % p.nseg; t /= p.nseg;
      const int tx = t % p.ntx; t /= p.ntx;
      const int ty = t % p.nty; const int n = t / p.nty;
      const int zs = seg * p.lseg;
      const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
      for (int ip = 0; ip < L + 2; ++ip) {
        mbar_wait(empty_bar + 8 * slot, phase ^ 1);
        mbar_expect_tx_e(full_bar + 8 * slot, (uint32_t)(CT_PLANE_ROWS * 32));
        tma_load_5d_e(a_base + slot * CT_PLANE_BYTES, &map_x, full_bar + 8 * slot, 0, 4 * tx, 32 * ty - 1, zs - 1 + ip, n);
        if (++slot == p.ring) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: output plane zl takes planes ip = zl, zl+1, zl+2 (kd = 0,1,2), kh = row shift of 4; the hi and the
    // lo Toeplitz halves go into the SAME 128 accumulator columns (two N = 128 MMAs), so the epilogue reads one value =====
    const uint32_t hi_d = desc_hi(256, 6);            // 8-row groups 256 B apart, 32B swizzle (A and B alike)
    const uint32_t w16 = w_base >> 4;
    // ring position / phase of the plane ip = zl of the current tile, advanced incrementally (no divisions on this warp:
    // it is the one thread of the CTA whose instruction count is on the critical path)
    int s0 = 0; uint32_t ph0 = 0;
    int oc = 0;
    const int ring = p.ring;
    const bool lo = p.lo != 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int seg = item % p.nseg;
      const int zs = seg * p.lseg;
      const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
      for (int zl = 0; zl < L; ++zl, ++oc) {
        const int buf = oc & (CT_NB - 1);
        int s1 = s0 + 1; uint32_t ph1 = ph0; if (s1 == ring) { s1 = 0; ph1 ^= 1; }
        int s2 = s1 + 1; uint32_t ph2 = ph1; if (s2 == ring) { s2 = 0; ph2 ^= 1; }
        mbar_wait(tempty_bar + 8 * buf, ((oc / CT_NB) & 1) ^ 1);
        if (zl == 0) { mbar_wait(full_bar + 8 * s0, ph0); mbar_wait(full_bar + 8 * s1, ph1); }
        mbar_wait(full_bar + 8 * s2, ph2);
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(buf * 128);
        const uint32_t a0 = (a_base + (uint32_t)s0 * CT_PLANE_BYTES) >> 4, a1 = (a_base + (uint32_t)s1 * CT_PLANE_BYTES) >> 4,
                       a2 = (a_base + (uint32_t)s2 * CT_PLANE_BYTES) >> 4;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const uint32_t a16 = kd == 0 ? a0 : (kd == 1 ? a1 : a2);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint64_t ad = desc_pack(hi_d, a16 + (uint32_t)(kh * 8));
            const uint32_t b16 = w16 + (uint32_t)((kd * 3 + kh) * (CT_W_SLAB >> 4));
            tc_mma_f16_e(dcol, ad, desc_pack(hi_d, b16), p.idesc, (kd | kh) != 0);
            if (lo) tc_mma_f16_e(dcol, ad, desc_pack(hi_d, b16 + (uint32_t)(4096 >> 4)), p.idesc, 1);
          }
        }
        tc_commit_e(tfull_bar + 8 * buf);
        tc_commit_e(empty_bar + 8 * s0);                                   // plane ip = zl is not needed again
        if (zl == L - 1) {                                                 // nor are the item's last two planes
          tc_commit_e(empty_bar + 8 * s1);
          tc_commit_e(empty_bar + 8 * s2);
          s0 = s2 + 1; ph0 = ph2; if (s0 == ring) { s0 = 0; ph0 ^= 1; }
        } else {
          s0 = s1; ph0 = ph1;
        }
      }
    }
  } else {
    // ===== epilogue: TMEM lane = segment (y = lane / 4, segment-in-tile = lane % 4); two warps per lane quarter =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;                 // voxels xo = 4*half .. 4*half+3 of the segment
    const int r = q * 32 + lane;
    const int sx = r & 3, yl = r >> 2;
    // EPI 0 / 1: per-channel sums of the bias-free accumulators (the bias enters the GroupNorm sums in closed form at the
    // flush: 2 instructions per value);  EPI 0 stores acc + bias;  EPI 2: relu(fma(acc, ga, gb)) with the bias folded in gb
    float ga[16], gb[16];                             // EPI 0: gb = bias;  EPI 1: unused;  EPI 2: GroupNorm scale / shift
    float sc[16], ssc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { ga[c] = 1.f; gb[c] = sw_f[432 + c]; sc[c] = 0.f; ssc[c] = 0.f; }
    float cnt = 0.f;
    int cur_n = -1, oc = 0;
    const bool wide = p.wide != 0;
    auto flush = [&](int n) {
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float b = sw_f[432 + c];
        s += sc[c] + cnt * b;
        ss += ssc[c] + b * (2.f * sc[c] + cnt * b);
        sc[c] = 0.f; ssc[c] = 0.f;
      }
      cnt = 0.f;
      s = warp_sum(s); ss = warp_sum(ss);
      if (lane == 0) { atomicAdd(stats + 2 * n, (double)s); atomicAdd(stats + 2 * n + 1, (double)ss); }
    };
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      int t = item;
      const int seg = t % p.nseg; t /= p.nseg;
      const int tx = t % p.ntx; t /= p.ntx;
      const int ty = t % p.nty; const int n = t / p.nty;
      const int zs = seg * p.lseg;
      const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
      if (n != cur_n) {
        if (EPI != 2 && stats && cur_n >= 0) flush(cur_n);
        if (EPI == 2) {
          float mean, rstd;
          gn_mean_rstd(stats + 2 * n, p.gn_count, p.gn_eps, mean, rstd);
#pragma unroll
          for (int c = 0; c < 16; ++c) { ga[c] = rstd * gamma[c]; gb[c] = fmaf(sw_f[432 + c] - mean, ga[c], beta[c]); }
        }
        cur_n = n;
      }
      const int gx = 32 * tx + 8 * sx + 4 * half, gy = 32 * ty + yl;
      const bool valid = (gx < p.W) && (gy < p.H);
      for (int zl = 0; zl < L; ++zl, ++oc) {
        const int buf = oc & (CT_NB - 1);
        mbar_wait(tfull_bar + 8 * buf, (oc / CT_NB) & 1);
        tc_fence_after();
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + half * 64);
        const size_t vox = (((size_t)n * p.D + (zs + zl)) * p.H + gy) * p.W + gx;
        T* dst = y + vox * p.y_ld;
        uint32_t v[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j) tc_ld16(tcol + (uint32_t)(j * 16), v[j]);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + 8 * buf);
        if (valid) {
          if (EPI != 2) cnt += 4.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float a = __uint_as_float(v[j][c]);
              if (EPI != 2) { sc[c] += a; ssc[c] = fmaf(a, a, ssc[c]); }
              f[c] = EPI == 2 ? fmaxf(fmaf(a, ga[c], gb[c]), 0.f) : a + gb[c];
            }
            if (EPI != 1) store16<T>(dst + (size_t)j * p.y_ld, f, wide);
          }
        }
      }
    }
    if (EPI != 2 && stats && cur_n >= 0) flush(cur_n);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient of the input block with the same Toeplitz operands:
//     G[(xo, co)][(kd, kh, xi)]  =  sum over segments  dY[segment][(xo, co)] * A_(kd,kh)[segment][xi]
//     dW[kd][kh][kw][co]         =  sum over xo        G[(xo, co)][(kd, kh, xi = xo + kw)]
// K = segments, and both operands are MN-major exactly as TMA delivers them: a segment of dy is 8 voxels x 16 channels =
// 256 contiguous bytes (two 128-byte rows of a 128B-swizzled tile: M-atoms 0 and 1), a segment of the padded input is the
// 32-byte window row of the forward pass.  The three kh taps are three N-atoms one y-row (4 window rows = 128 bytes) apart, so
// ONE MMA 128 x 48 x 16 per kd and K step covers (kh, xi); the three kd taps are three accumulators fed from the three resident
// halo planes.  A CTA accumulates over all its tiles in TMEM and folds the band of G into the 27 x 16 gradient once.
struct CwParams {
  int D, H, W, N;
  int ntx, nty, nseg, lseg, nitems, ring, dstages;
  uint32_t idesc;
};
constexpr int CW_THREADS = 192;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: final fold
constexpr int CW_DY_BYTES = 32768;       // one dy plane tile: 2 M-atoms x [128 segments x 128 B]

template <typename T>
__global__ void __launch_bounds__(CW_THREADS, 1)
conv3d_k3_cin1_wgrad_toeplitz_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                                     const CwParams p, float* __restrict__ dw) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ float red_s[27 * 16];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t d_base = smem_base;                                     // [dstages][2][128 rows][128 B], 128B-swizzled
  const uint32_t a_base = smem_base + p.dstages * CW_DY_BYTES;           // [ring][136 rows][32 B], 32B-swizzled
  float* stage = reinterpret_cast<float*>(smem_al + p.dstages * CW_DY_BYTES + p.ring * CT_PLANE_BYTES);   // [128][17]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 128 * 17);
  const uint32_t full_bar = smem_u32(bars);                              // [ring]
  const uint32_t empty_bar = full_bar + 8 * CT_MAXRING;                  // [ring]
  const uint32_t dfull_bar = empty_bar + 8 * CT_MAXRING;                 // [dstages <= 4]
  const uint32_t dempty_bar = dfull_bar + 32;                            // [dstages]
  const uint32_t done_bar = dempty_bar + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CT_MAXRING + 9);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 27 * 16; i += CW_THREADS) red_s[i] = 0.f;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    for (int s = 0; s < p.ring; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < p.dstages; ++s) { mbar_init(dfull_bar + 8 * s, 1); mbar_init(dempty_bar + 8 * s, 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===== TMA producer: per item the halo planes zs-1, zs, then for every output plane the next halo plane + the dy tile =====
    int slot = 0; uint32_t phase = 0; int ds = 0; uint32_t dphase = 0;
    auto load_plane = [&](int tx, int ty, int z, int n) {
      mbar_wait(empty_bar + 8 * slot, phase ^ 1);
      mbar_expect_tx_e(full_bar + 8 * slot, (uint32_t)(CT_PLANE_ROWS * 32));
      tma_load_5d_e(a_base + slot * CT_PLANE_BYTES, &map_x, full_bar + 8 * slot, 0, 4 * tx, 32 * ty - 1, z, n);
      if (++slot == p.ring) { slot = 0; phase ^= 1; }
    };
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      int t = item;
      const int seg = t % p.nseg; t /= p.nseg;
      const int tx = t % p.ntx; t /= p.ntx;
      const int ty = t % p.nty; const int n = t / p.nty;
      const int zs = seg * p.lseg;
      const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
      load_plane(tx, ty, zs - 1, n);
      load_plane(tx, ty, zs, n);
      for (int zl = 0; zl < L; ++zl) {
        load_plane(tx, ty, zs + zl + 1, n);
        mbar_wait(dempty_bar + 8 * ds, dphase ^ 1);
        mbar_expect_tx_e(dfull_bar + 8 * ds, (uint32_t)CW_DY_BYTES);
        tma_load_5d_e(d_base + ds * CW_DY_BYTES, &map_dy, dfull_bar + 8 * ds, 0, 4 * tx, 32 * ty, zs + zl, n);
        tma_load_5d_e(d_base + ds * CW_DY_BYTES + 16384, &map_dy, dfull_bar + 8 * ds, 64, 4 * tx, 32 * ty, zs + zl, n);
        if (++ds == p.dstages) { ds = 0; dphase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t hi_a = desc_hi(1024, 2);           // dy: K-atoms (8 segments x 128 B) 1 KB apart, 128B swizzle
    const uint32_t hi_b = desc_hi(256, 6);            // x windows: K-atoms (8 segments x 32 B) 256 B apart, 32B swizzle
    const uint32_t lbo_a = ((16384u >> 4) & 0x3FFFu) << 16;       // M-atom 1 = the second half of every segment
    const uint32_t lbo_b = ((128u >> 4) & 0x3FFFu) << 16;         // N-atoms = the kh taps, one y row (4 segments) apart
    int s0 = 0; uint32_t ph0 = 0; int ds = 0; uint32_t dphase = 0;
    uint32_t accumulate = 0;
    const int ring = p.ring;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int seg = item % p.nseg;
      const int zs = seg * p.lseg;
      const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
      for (int zl = 0; zl < L; ++zl) {
        int s1 = s0 + 1; uint32_t ph1 = ph0; if (s1 == ring) { s1 = 0; ph1 ^= 1; }
        int s2 = s1 + 1; uint32_t ph2 = ph1; if (s2 == ring) { s2 = 0; ph2 ^= 1; }
        if (zl == 0) { mbar_wait(full_bar + 8 * s0, ph0); mbar_wait(full_bar + 8 * s1, ph1); }
        mbar_wait(full_bar + 8 * s2, ph2);
        mbar_wait(dfull_bar + 8 * ds, dphase);
        tc_fence_after();
        const uint32_t da = ((d_base + (uint32_t)ds * CW_DY_BYTES) >> 4) | lbo_a;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const int sk = kd == 0 ? s0 : (kd == 1 ? s1 : s2);
          const uint32_t db = ((a_base + (uint32_t)sk * CT_PLANE_BYTES) >> 4) | lbo_b;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            tc_mma_f16_e(tmem_base + (uint32_t)(kd * 48), desc_pack(hi_a, da + (uint32_t)(ks * (2048 >> 4))),
                         desc_pack(hi_b, db + (uint32_t)(ks * (512 >> 4))), p.idesc, ks == 0 ? accumulate : 1u);
        }
        accumulate = 1;
        tc_commit_e(dempty_bar + 8 * ds);
        tc_commit_e(empty_bar + 8 * s0);
        if (++ds == p.dstages) { ds = 0; dphase ^= 1; }
        if (zl == L - 1) {
          tc_commit_e(empty_bar + 8 * s1);
          tc_commit_e(empty_bar + 8 * s2);
          s0 = s2 + 1; ph0 = ph2; if (s0 == ring) { s0 = 0; ph0 ^= 1; }
        } else {
          s0 = s1; ph0 = ph1;
        }
      }
    }
    tc_commit_e(done_bar);
  } else {
    // ===== fold (once): TMEM lane = (xo, co), column = kd*48 + kh*16 + xi; keep xi = xo + kw =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int xo = r >> 4, co = r & 15;
    float* my = stage + r * 17;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    for (int g = 0; g < 9; ++g) {                     // g = kd*3 + kh
      uint32_t v[16];
      tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 16), v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) my[j] = __uint_as_float(v[j]);
      __syncwarp();
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) atomicAdd(&red_s[(g * 3 + kw) * 16 + co], my[xo + kw]);
      __syncwarp();
    }
    named_bar_sync(1, 128);
    for (int i = threadIdx.x - 64; i < 27 * 16; i += 128) atomicAdd(dw + i, red_s[i]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace

// xpad: [N][D][H][W + 16] dtype, x index i of a row at column i + 9, columns 0..8 and W+9..W+15 zero.
// w: fp32 [27][16] (the SIMT layout [taps][Cin=1][Cout]); y: [N,D,H,W,16] pitch y_ld.
extern "C" int seg3d_conv3d_cin1_fwd(int dtype, int epi_mode, const void* xpad, int x_pitch, const float* w, const float* bias,
                                     void* y, int y_ld, int N, int D, int H, int W, double* stats,
                                     const float* gamma, const float* beta, float eps, void* stream) {
  SEG3D_REQUIRE(dtype == SEG3D_F16 || dtype == SEG3D_BF16, "conv3d_cin1_fwd: dtype must be f16 or bf16");
  SEG3D_REQUIRE(epi_mode >= 0 && epi_mode <= 2, "conv3d_cin1_fwd: epi_mode must be 0, 1 or 2");
  SEG3D_REQUIRE(xpad && w && N > 0 && D > 0 && H > 0 && W > 0 && W % 8 == 0, "conv3d_cin1_fwd: bad dims (W must be a multiple of 8)");
  SEG3D_REQUIRE(x_pitch == W + SEG3D_CIN1_PAD, "conv3d_cin1_fwd: x_pitch must be W + SEG3D_CIN1_PAD");
  SEG3D_REQUIRE(((uintptr_t)xpad) % 16 == 0, "conv3d_cin1_fwd: xpad must be 16-byte aligned");
  SEG3D_REQUIRE(epi_mode == 1 || (y && y_ld >= 16 && y_ld % 8 == 0 && ((uintptr_t)y) % 16 == 0), "conv3d_cin1_fwd: bad output");
  SEG3D_REQUIRE(epi_mode == 0 || stats, "conv3d_cin1_fwd: statistics pointer required");
  SEG3D_REQUIRE(epi_mode != 2 || (gamma && beta), "conv3d_cin1_fwd: gamma / beta required");
  EncodeTiledFn encode = get_encode();
  if (!encode) { seg3d_set_error("conv3d_cin1_fwd: cuTensorMapEncodeTiled entry point not available"); return SEG3D_ECUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  CtParams p;
  memset(&p, 0, sizeof(p));
  p.D = D; p.H = H; p.W = W; p.N = N; p.y_ld = y_ld;
  p.ntx = (W + 31) / 32; p.nty = (H + 31) / 32;
  p.wide = (y && wide_ok(y, y_ld, 2) && env_int("SEG3D_WIDE_ST", 1)) ? 1 : 0;
  p.epi_mode = epi_mode; p.gn_eps = eps; p.gn_count = (double)D * H * W * 16.0;
  p.ring = env_int("SEG3D_CIN1_RING", CT_MAXRING);
  if (p.ring < 4) p.ring = 4;
  if (p.ring > CT_MAXRING) p.ring = CT_MAXRING;
  // z segments: ~8 work items per CTA for balance; each segment re-loads two halo planes
  const long long cols = (long long)N * p.ntx * p.nty;
  const long long want = 8ll * seg3d_num_sms();
  int nseg = (int)((want + cols - 1) / cols);
  if (nseg < 1) nseg = 1;
  int lseg = (D + nseg - 1) / nseg;
  if (lseg < 8) lseg = D < 8 ? D : 8;
  p.lseg = lseg; p.nseg = (D + lseg - 1) / lseg;
  const long long nitems = cols * p.nseg;
  SEG3D_REQUIRE(nitems > 0 && nitems < (1ll << 31), "conv3d_cin1_fwd: work-item count out of range");
  p.nitems = (int)nitems;
  const uint32_t fmt = dtype == SEG3D_BF16 ? 1u : 0u;
  // weights: 0 (default) = rounded to the storage type like every other layer's tensor-core weights (9 MMAs per plane);
  // 1 = split hi + lo, fp32-accurate (18 MMAs).  The layer is tensor-bound on the band's wasted flops and the step is power-limited:
  // the second half of the MMAs cost 2.5 % of the whole sliding-window step for no measurable parity change.
  p.lo = env_int("SEG3D_CIN1_LO", 0) ? 1 : 0;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  // overlapping-window tensor map: element (i, s, y, z, n) = xpad[n][z][y][8 + 8 s + i], i.e. x index 8 s - 1 + i
  CUtensorMap map_x;
  {
    const CUtensorMapDataType tdt = dtype == SEG3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t dims[5] = {16, (cuuint64_t)W / 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {16, (cuuint64_t)x_pitch * 2, (cuuint64_t)H * x_pitch * 2, (cuuint64_t)D * H * x_pitch * 2};
    cuuint32_t box[5] = {16, 4, 34, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = const_cast<void*>(static_cast<const void*>(static_cast<const uint8_t*>(xpad) + 16));
    CUresult r = encode(&map_x, tdt, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seg3d_set_error("conv3d_cin1_fwd: cuTensorMapEncodeTiled failed with %d", (int)r); return SEG3D_ECUDA; }
  }
  // the CTA allocates all 512 TMEM columns: ask for more than half of the SM's shared memory so two CTAs never share an SM
  size_t smem = 1024 + (size_t)9 * CT_W_SLAB + (size_t)p.ring * CT_PLANE_BYTES + (2 * CT_MAXRING + 2 * CT_NB) * 8 + 64;
  if (smem < 120 * 1024) smem = 120 * 1024;
  const long long max_grid = seg3d_num_sms();
  dim3 grid((unsigned)(nitems < max_grid ? nitems : max_grid));
  cudaError_t e = cudaSuccess;
#define SEG3D_LAUNCH_CT(TT, EPI)                                                                                                      \
  { e = cudaFuncSetAttribute(conv3d_k3_cin1_toeplitz_kernel<TT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    if (e == cudaSuccess) conv3d_k3_cin1_toeplitz_kernel<TT, EPI><<<grid, CT_THREADS, smem, st>>>(map_x, w, bias, (TT*)y, p, stats, gamma, beta); }
  if (dtype == SEG3D_BF16) {
    if (epi_mode == 0) SEG3D_LAUNCH_CT(__nv_bfloat16, 0) else if (epi_mode == 1) SEG3D_LAUNCH_CT(__nv_bfloat16, 1) else SEG3D_LAUNCH_CT(__nv_bfloat16, 2)
  } else {
    if (epi_mode == 0) SEG3D_LAUNCH_CT(__half, 0) else if (epi_mode == 1) SEG3D_LAUNCH_CT(__half, 1) else SEG3D_LAUNCH_CT(__half, 2)
  }
#undef SEG3D_LAUNCH_CT
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { seg3d_set_error("conv3d_k3_cin1_toeplitz_kernel launch failed: %s", cudaGetErrorString(e)); return SEG3D_ECUDA; }
  return SEG3D_OK;
}

// dw (fp32 [27][16], accumulated; the caller zeroes it) += xpad (*) dy.  xpad as in seg3d_conv3d_cin1_fwd; dy: [N,D,H,W,16]
// DENSE (pitch 16: a segment of 8 voxels is 256 contiguous bytes).
extern "C" int seg3d_conv3d_cin1_wgrad(int dtype, const void* xpad, int x_pitch, const void* dy, int dy_ld, float* dw,
                                       int N, int D, int H, int W, void* stream) {
  SEG3D_REQUIRE(dtype == SEG3D_F16 || dtype == SEG3D_BF16, "conv3d_cin1_wgrad: dtype must be f16 or bf16");
  SEG3D_REQUIRE(xpad && dy && dw && N > 0 && D > 0 && H > 0 && W > 0 && W % 8 == 0, "conv3d_cin1_wgrad: bad dims (W must be a multiple of 8)");
  SEG3D_REQUIRE(x_pitch == W + SEG3D_CIN1_PAD && dy_ld == 16, "conv3d_cin1_wgrad: x_pitch must be W + SEG3D_CIN1_PAD and dy dense (pitch 16)");
  SEG3D_REQUIRE(((uintptr_t)xpad) % 16 == 0 && ((uintptr_t)dy) % 16 == 0, "conv3d_cin1_wgrad: misaligned pointer");
  EncodeTiledFn encode = get_encode();
  if (!encode) { seg3d_set_error("conv3d_cin1_wgrad: cuTensorMapEncodeTiled entry point not available"); return SEG3D_ECUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  CwParams p;
  memset(&p, 0, sizeof(p));
  p.D = D; p.H = H; p.W = W; p.N = N;
  p.ntx = (W + 31) / 32; p.nty = (H + 31) / 32;
  p.ring = 8; p.dstages = 4;
  const long long cols = (long long)N * p.ntx * p.nty;
  const long long want = 6ll * seg3d_num_sms();
  int nseg = (int)((want + cols - 1) / cols);
  if (nseg < 1) nseg = 1;
  int lseg = (D + nseg - 1) / nseg;
  if (lseg < 8) lseg = D < 8 ? D : 8;
  p.lseg = lseg; p.nseg = (D + lseg - 1) / lseg;
  const long long nitems = cols * p.nseg;
  SEG3D_REQUIRE(nitems > 0 && nitems < (1ll << 31), "conv3d_cin1_wgrad: work-item count out of range");
  p.nitems = (int)nitems;
  const uint32_t fmt = dtype == SEG3D_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(48 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const CUtensorMapDataType tdt = dtype == SEG3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap map_x, map_dy;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  {
    cuuint64_t dims[5] = {16, (cuuint64_t)W / 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {16, (cuuint64_t)x_pitch * 2, (cuuint64_t)H * x_pitch * 2, (cuuint64_t)D * H * x_pitch * 2};
    cuuint32_t box[5] = {16, 4, 34, 1, 1};
    void* base = const_cast<void*>(static_cast<const void*>(static_cast<const uint8_t*>(xpad) + 16));
    CUresult r = encode(&map_x, tdt, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seg3d_set_error("conv3d_cin1_wgrad: cuTensorMapEncodeTiled(x) failed with %d", (int)r); return SEG3D_ECUDA; }
  }
  {
    cuuint64_t dims[5] = {128, (cuuint64_t)W / 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {256, (cuuint64_t)W * 32, (cuuint64_t)H * W * 32, (cuuint64_t)D * H * W * 32};
    cuuint32_t box[5] = {64, 4, 32, 1, 1};
    CUresult r = encode(&map_dy, tdt, 5, const_cast<void*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seg3d_set_error("conv3d_cin1_wgrad: cuTensorMapEncodeTiled(dy) failed with %d", (int)r); return SEG3D_ECUDA; }
  }
  const size_t smem = 1024 + (size_t)p.dstages * CW_DY_BYTES + (size_t)p.ring * CT_PLANE_BYTES + 128 * 17 * 4 + (2 * CT_MAXRING + 9) * 8 + 64;
  const long long max_grid = seg3d_num_sms();
  dim3 grid((unsigned)(nitems < max_grid ? nitems : max_grid));
  cudaError_t e;
  if (dtype == SEG3D_BF16) {
    e = cudaFuncSetAttribute(conv3d_k3_cin1_wgrad_toeplitz_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv3d_k3_cin1_wgrad_toeplitz_kernel<__nv_bfloat16><<<grid, CW_THREADS, smem, st>>>(map_x, map_dy, p, dw);
  } else {
    e = cudaFuncSetAttribute(conv3d_k3_cin1_wgrad_toeplitz_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv3d_k3_cin1_wgrad_toeplitz_kernel<__half><<<grid, CW_THREADS, smem, st>>>(map_x, map_dy, p, dw);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { seg3d_set_error("conv3d_k3_cin1_wgrad_toeplitz_kernel launch failed: %s", cudaGetErrorString(e)); return SEG3D_ECUDA; }
  return SEG3D_OK;
}
