// conv_tc_cin1.cu - input block: k3 s1 p1 convolution with ONE input channel and 16 output channels
// (reference vnet_inblock.py:9) on the tensor cores.
//
// The implicit GEMM has K = 27 taps (padded to 32): there is no channel dimension for TMA to deliver, so the
// A tile is built in shared memory by the CTA itself.  Four builder warps own the 128 voxels of an 8 x 16
// (x,y) plane tile and march along z: each thread keeps the 3 x 9 input window of its voxel packed in registers, has
// the loads of two more planes in flight, loads only the 9 values of one new plane per step, packs the 27 taps into one 64-byte K-major row and
// stores it into the 64B-swizzled UMMA layout (fence.proxy.async, mbarrier).  One thread issues two
// 128 x 32 x 16 MMAs per tile; four epilogue warps drain TMEM, add bias, take the GroupNorm sums, store NDHWC.
//
// The weights stay fp32-accurate: B holds 32 rows = [hi(W) ; lo(W)] with hi = round_T(W), lo = round_T(W - hi);
// the epilogue adds accumulator columns c and 16+c.  At M=128 the MMA cost is set by the A rows, so the split is free.
#include "tc_ptx.cuh"

namespace {

constexpr int C1_THREADS = 288;      // warp 0: MMA issuer, warps 1-4: builders, warps 5-8: epilogue
constexpr int C1_STAGES = 4;         // A-tile ring (8 KB each)
constexpr int C1_NB = 4;             // TMEM accumulators (32 columns each)

struct Cin1Params {
  int D, H, W, N, y_ld;
  int ntx, nty, nseg, lseg, nitems, wide;
  uint32_t idesc;
};

template <typename T>
__global__ void __launch_bounds__(C1_THREADS)
conv3d_k3_cin1_tc_kernel(const T* __restrict__ x, const float* __restrict__ w /*[27][16]*/, const float* __restrict__ bias,
                         T* __restrict__ y, const Cin1Params p, double* __restrict__ stats) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;                                   // [C1_STAGES][128 rows][64 B]
  const uint32_t b_base = smem_base + C1_STAGES * 8192;                // [32 rows][64 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + C1_STAGES * 8192 + 2048);
  const uint32_t afull_bar = smem_u32(bars);                           // [C1_STAGES] one arrival per builder warp
  const uint32_t aempty_bar = afull_bar + 8 * C1_STAGES;               // [C1_STAGES] tcgen05.commit
  const uint32_t tfull_bar = aempty_bar + 8 * C1_STAGES;               // [C1_NB]
  const uint32_t tempty_bar = tfull_bar + 8 * C1_NB;                   // [C1_NB] 4 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C1_STAGES + 2 * C1_NB);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C1_STAGES; ++s) { mbar_init(afull_bar + 8 * s, 4); mbar_init(aempty_bar + 8 * s, 1); }
    for (int b = 0; b < C1_NB; ++b) { mbar_init(tfull_bar + 8 * b, 1); mbar_init(tempty_bar + 8 * b, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(C1_NB * 32)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===== MMA issuer =====
    {   // whole warp, converged: see elect_one()
      const uint32_t hi = desc_hi(512, 4);          // 8-row groups 512 B apart, 64B swizzle
      int stage = 0; uint32_t phase = 0; int oc = 0;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const int seg = item % p.nseg;
        const int zs = seg * p.lseg;
        const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
        for (int zl = 0; zl < L; ++zl, ++oc) {
          const int buf = oc % C1_NB;
          mbar_wait(tempty_bar + 8 * buf, ((oc / C1_NB) & 1) ^ 1);
          mbar_wait(afull_bar + 8 * stage, phase);
          tc_fence_after();
          const uint32_t lo_a = (a_base + stage * 8192) >> 4, lo_b = b_base >> 4;
          const uint32_t dcol = tmem_base + (uint32_t)(buf * 32);
          tc_mma_f16_e(dcol, desc_pack(hi, lo_a), desc_pack(hi, lo_b), p.idesc, 0);
          tc_mma_f16_e(dcol, desc_pack(hi, lo_a + 2), desc_pack(hi, lo_b + 2), p.idesc, 1);
          tc_commit_e(aempty_bar + 8 * stage);
          tc_commit_e(tfull_bar + 8 * buf);
          if (++stage == C1_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp <= 4) {
    // ===== A-tile builders =====
    const int r = (warp - 1) * 32 + lane;           // GEMM row = voxel (lx, ly) of the plane tile
    const int lx = r & 7, ly = r >> 3;
    const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
    if (r < 32) {                                    // B rows: 0-15 hi(W[co]), 16-31 lo(W[co])
      const int co = r & 15;
      uint32_t h[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float wv = k < 27 ? w[k * 16 + co] : 0.f;
        T hv = from_f32<T>(wv);
        if (r >= 16) hv = from_f32<T>(wv - to_f32<T>(hv));
        h[k] = (uint32_t)(*reinterpret_cast<unsigned short*>(&hv));
      }
      uint8_t* row = smem_al + C1_STAGES * 8192 + r * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 v;
        v.x = h[8 * c] | (h[8 * c + 1] << 16); v.y = h[8 * c + 2] | (h[8 * c + 3] << 16);
        v.z = h[8 * c + 4] | (h[8 * c + 5] << 16); v.w = h[8 * c + 6] | (h[8 * c + 7] << 16);
        *reinterpret_cast<uint4*>(row + ((c ^ ((r >> 1) & 3)) << 4)) = v;
      }
    }
    uint8_t* arow = smem_al + r * 64;
    const int sw = (r >> 1) & 3;
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      int t = item;
      const int seg = t % p.nseg; t /= p.nseg;
      const int x0 = (t % p.ntx) * 8; t /= p.ntx;
      const int y0 = (t % p.nty) * 16; const int n = t / p.nty;
      const int zs = seg * p.lseg;
      const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
      const int gx = x0 + lx, gy = y0 + ly;
      // in-plane offsets and bounds of the 9 (kh,kw) neighbours
      int off[9]; bool ok[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int yy = gy + i / 3 - 1, xx = gx + i % 3 - 1;
        ok[i] = (yy >= 0) && (yy < p.H) && (xx >= 0) && (xx < p.W);
        off[i] = yy * p.W + xx;
      }
      const unsigned short* xn = xs + (size_t)n * p.D * p.H * p.W;
      auto load_plane = [&](int gz, uint32_t* dst) {
        const bool zok = gz >= 0 && gz < p.D;
        const unsigned short* pl = xn + (size_t)(zok ? gz : 0) * p.H * p.W;
#pragma unroll
        for (int i = 0; i < 9; ++i) dst[i] = (zok && ok[i]) ? (uint32_t)__ldg(pl + off[i]) : 0u;
      };
      // Register window: the three planes a row needs are kept PACKED (nine 16-bit values in five registers), and two
      // more planes are in flight as raw loads - plane s+1 is packed two full steps after its loads were issued, so the
      // builder does not wait a global-load latency per plane.  Step s: pack plane s+1 (loaded at step s-2), issue the
      // loads of plane s+3 into the buffer that just became free, build the row from planes s-1, s, s+1.
      uint32_t wp[3][5], raw[2][9];
      auto pack = [&](const uint32_t* r9, uint32_t* w5) {
#pragma unroll
        for (int q = 0; q < 4; ++q) w5[q] = r9[2 * q] | (r9[2 * q + 1] << 16);
        w5[4] = r9[8];
      };
      load_plane(zs - 1, raw[0]); load_plane(zs, raw[1]);
      pack(raw[0], wp[0]); pack(raw[1], wp[1]);
      load_plane(zs + 1, raw[1]); load_plane(zs + 2, raw[0]);
      for (int zl = 0; zl < L; zl += 6) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          if (zl + j < L) {
            pack(raw[(j + 1) % 2], wp[(j + 2) % 3]);              // plane s+1
            load_plane(zs + zl + j + 3, raw[(j + 1) % 2]);        // plane s+3
            uint32_t hw[16];
#pragma unroll
            for (int wd = 0; wd < 14; ++wd) {                      // word wd = taps (2wd, 2wd+1); tap k = plane k/9, value k%9
              const int k0 = 2 * wd, k1 = 2 * wd + 1;
              const uint32_t ra = wp[(j + k0 / 9) % 3][(k0 % 9) >> 1];
              const uint32_t rb = k1 < 27 ? wp[(j + k1 / 9) % 3][(k1 % 9) >> 1] : 0u;
              const int ha = (k0 % 9) & 1, hb = k1 < 27 ? ((k1 % 9) & 1) : 0;
              hw[wd] = __byte_perm(ra, rb, (2 * ha) | ((2 * ha + 1) << 4) | ((4 + 2 * hb) << 8) | ((5 + 2 * hb) << 12));
            }
            hw[14] = 0u; hw[15] = 0u;
            mbar_wait(aempty_bar + 8 * stage, phase ^ 1);
            uint8_t* row = arow + stage * 8192;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(afull_bar + 8 * stage);
            if (++stage == C1_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int lx = r & 7, ly = r >> 3;
    float bv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) bv[c] = bias ? bias[c] : 0.f;
    float s = 0.f, ss = 0.f;
    int cur_n = -1, oc = 0;
    const bool wide = p.wide != 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      int t = item;
      const int seg = t % p.nseg; t /= p.nseg;
      const int x0 = (t % p.ntx) * 8; t /= p.ntx;
      const int y0 = (t % p.nty) * 16; const int n = t / p.nty;
      const int zs = seg * p.lseg;
      const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
      if (stats && n != cur_n) {
        if (cur_n >= 0) {
          s = warp_sum(s); ss = warp_sum(ss);
          if (lane == 0) { atomicAdd(stats + 2 * cur_n, (double)s); atomicAdd(stats + 2 * cur_n + 1, (double)ss); }
        }
        s = 0.f; ss = 0.f; cur_n = n;
      }
      const int gx = x0 + lx, gy = y0 + ly;
      const bool valid = (gx < p.W) && (gy < p.H);
      for (int zl = 0; zl < L; ++zl, ++oc) {
        const int buf = oc % C1_NB;
        mbar_wait(tfull_bar + 8 * buf, (oc / C1_NB) & 1);
        tc_fence_after();
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 32);
        uint32_t vh[16], vl[16];
        tc_ld16(tcol, vh);
        tc_ld16(tcol + 16u, vl);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + 8 * buf);
        if (valid) {
          float f[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            f[c] = (__uint_as_float(vh[c]) + __uint_as_float(vl[c])) + bv[c];
            s += f[c]; ss += f[c] * f[c];
          }
          const size_t vox = (((size_t)n * p.D + (zs + zl)) * p.H + gy) * p.W + gx;
          T* dst = y + vox * p.y_ld;
          store16<T>(dst, f, wide);
        }
      }
    }
    if (stats && cur_n >= 0) {
      s = warp_sum(s); ss = warp_sum(ss);
      if (lane == 0) { atomicAdd(stats + 2 * cur_n, (double)s); atomicAdd(stats + 2 * cur_n + 1, (double)ss); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(C1_NB * 32)) : "memory");
  }
}

}  // namespace

int seg3d_conv_cin1_tc_supported(int dtype, int Cin, int Cout, int x_ld, int y_ld, int W) {
  return (dtype == SEG3D_F16 || dtype == SEG3D_BF16) && Cin == 1 && Cout == 16 && x_ld == 1 && y_ld % 8 == 0 && W % 8 == 0 &&
         env_int("SEG3D_CIN1_TC", 1) != 0;
}

// x: [N,D,H,W] dtype; w: fp32 [27][16] (the SIMT layout [taps][Cin=1][Cout]); y: [N,D,H,W,16] pitch y_ld
int seg3d_conv_cin1_tc(int dtype, const void* x, const void* w, const float* bias, void* y, int y_ld,
                       int N, int D, int H, int W, double* stats, cudaStream_t st) {
  SEG3D_REQUIRE(((uintptr_t)y) % 16 == 0 && ((uintptr_t)x) % 2 == 0, "conv_cin1_tc: misaligned pointer");
  Cin1Params p;
  memset(&p, 0, sizeof(p));
  p.D = D; p.H = H; p.W = W; p.N = N; p.y_ld = y_ld;
  p.ntx = W / 8; p.nty = (H + 15) / 16;
  p.wide = (wide_ok(y, y_ld, 2) && env_int("SEG3D_WIDE_ST", 1)) ? 1 : 0;
  const size_t smem = 1024 + (size_t)C1_STAGES * 8192 + 2048 + (2 * C1_STAGES + 2 * C1_NB) * 8 + 64;
  // persistent grid = CTAs that are actually co-resident: registers (allocated per warp in units of 256) cap it at
  // 2 per SM; forcing 3 with launch bounds spills and is 45 % slower
  int occ = 1;
  {
    cudaFuncAttributes fa;
    cudaError_t e0 = dtype == SEG3D_BF16 ? cudaFuncGetAttributes(&fa, conv3d_k3_cin1_tc_kernel<__nv_bfloat16>)
                                         : cudaFuncGetAttributes(&fa, conv3d_k3_cin1_tc_kernel<__half>);
    if (e0 != cudaSuccess) { seg3d_set_error("conv_cin1_tc: cudaFuncGetAttributes failed: %s", cudaGetErrorString(e0)); return SEG3D_ECUDA; }
    const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
    occ = 65536 / (regs_per_warp * (C1_THREADS / 32));
    const int by_smem = (int)((220 * 1024) / (smem + 1024));
    if (occ > by_smem) occ = by_smem;
    if (occ < 1) occ = 1;
  }
  int ctas_per_sm = env_int("SEG3D_CIN1_CTAS_PER_SM", 4);
  if (ctas_per_sm > occ) ctas_per_sm = occ;
  if (ctas_per_sm * C1_NB * 32 > 512) ctas_per_sm = 512 / (C1_NB * 32);
  const long long cols = (long long)N * p.ntx * p.nty;
  const long long want = 8ll * ctas_per_sm * seg3d_num_sms();
  int nseg = (int)((want + cols - 1) / cols);
  if (nseg < 1) nseg = 1;
  int lseg = (D + nseg - 1) / nseg;
  if (lseg < 8) lseg = D < 8 ? D : 8;
  p.lseg = lseg; p.nseg = (D + lseg - 1) / lseg;
  const long long nitems = cols * p.nseg;
  SEG3D_REQUIRE(nitems > 0 && nitems < (1ll << 31), "conv_cin1_tc: work-item count out of range");
  p.nitems = (int)nitems;
  const uint32_t fmt = dtype == SEG3D_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  const long long max_grid = (long long)ctas_per_sm * seg3d_num_sms();
  dim3 grid((unsigned)(nitems < max_grid ? nitems : max_grid));
  cudaError_t e;
  if (dtype == SEG3D_BF16) {
    e = cudaFuncSetAttribute(conv3d_k3_cin1_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv3d_k3_cin1_tc_kernel<__nv_bfloat16><<<grid, C1_THREADS, smem, st>>>((const __nv_bfloat16*)x, (const float*)w, bias, (__nv_bfloat16*)y, p, stats);
  } else {
    e = cudaFuncSetAttribute(conv3d_k3_cin1_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv3d_k3_cin1_tc_kernel<__half><<<grid, C1_THREADS, smem, st>>>((const __half*)x, (const float*)w, bias, (__half*)y, p, stats);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { seg3d_set_error("conv3d_k3_cin1_tc_kernel launch failed: %s", cudaGetErrorString(e)); return SEG3D_ECUDA; }
  return SEG3D_OK;
}
