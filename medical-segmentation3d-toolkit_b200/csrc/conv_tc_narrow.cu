// conv_tc_narrow.cu - k3 s1 p1 convolution with a NARROW output (Cout = number of classes, out_block.conv1,
// reference vnet_outblock.py:13) on the tensor cores with the nine in-plane taps folded into the GEMM N dimension.
//
// At M=128 the smallest UMMA N is 16 and every MMA re-reads its 128 x K A rows from shared memory, so a plain
// implicit GEMM spends 27 * Cin/16 A-feed-bound MMAs per 128 voxels to produce 2 useful columns.  Here the GEMM is
//
//     P[r][(kh,kw,co)]  =  sum_kd  sum_ci  X_{z+kd-1}[r][ci] * W[kd][(kh,kw,co)][ci]
//
// where r runs over the FLAT rows of a 10 x 12 (x,y) halo plane (row pitch 10 voxels) and N = 9*Cout (padded to a
// multiple of 16): 3 * Cin/16 MMAs per plane instead of 27 * Cin/16.  The epilogue finishes the convolution with
//
//     y[ly][lx][co]  =  sum_{kh,kw}  P[(ly+kh)*10 + lx+kw][(kh,kw,co)]
//
// through a shared-memory staging tile (TMEM lane = flat row, so the nine partials of one voxel sit in nine
// different lanes).  128 flat rows hold an 8 x 10 block of outputs (rows up to (9+2)*10+7+2 = 119).
// As in the z-march kernel a CTA walks along z, every halo plane is loaded once by TMA and feeds three output
// planes held in four rotating TMEM accumulators.  Output: fp32 [N][D][H][W][Cout] dense + GroupNorm partial sums.
#include "tc_ptx.cuh"

namespace {

struct FoldParams {
  int C, NP, row_bytes;             // real output channels, padded GEMM N, bytes of one voxel row (Cin * 2)
  int D, H, W, N;
  int ntx, nty, nseg, lseg, nitems;
  int ring, plane_bytes, w_slab, plane_tx, w_tx;
  int acc_cols, tmem_cols, pitch;   // staging-tile pitch in floats (odd: conflict-free scalar stores and gathers; a
                                    // 16-byte-vector variant with an even pitch measured slower, the gather conflicts)
  uint32_t idesc, sbo, layout_type;
  // FUSE_GN: the input is relu(gn(raw) + res) formed in shared memory from two TMA-loaded planes
  float gn_eps;
  double gn_count;                  // Cin * D * H * W
  const double* gn_stats;           // [N][2] finished sums of the producing convolution
  const float* gn_gamma;
  const float* gn_beta;
  // the first rg_nch channels of the RESIDUAL are themselves a raw convolution result: res = relu(gn(res_raw)) for those
  // (the up-conv half of the concat buffer whose GroupNorm-apply pass was skipped, see seg3d_conv3d_k3_gnin_fwd)
  int rg_nch;
  float rg_eps;
  double rg_count;
  const double* rg_stats;
  const float* rg_gamma;
  const float* rg_beta;
};
constexpr int FD_NB = 4;            // rotating TMEM accumulators
constexpr int FD_TX = 8, FD_TY = 10, FD_HX = 10, FD_HY = 12;
#ifndef SEG3D_FDG_MINB
#define SEG3D_FDG_MINB 2
#endif
#ifndef SEG3D_FDG_XW
#define SEG3D_FDG_XW 4
#endif
constexpr int FDG_XW = SEG3D_FDG_XW;                 // transform warps of the fused variant (4 or 8)
constexpr int FDG_THREADS = 192 + 32 * FDG_XW;       // fused variant: TMA, MMA, 4 epilogue warps + the transform warps
constexpr int FDG_NSLOT = 512 / (32 * FDG_XW);       // 16-byte slots of a plane per transform thread
constexpr int FDG_MINB = SEG3D_FDG_MINB;   // co-resident CTAs the fused variant is compiled for (register cap 65536 / (320 * MINB))

// SPLIT (strict-parity mode): a voxel row is [hi(KC/2) | lo(KC/2)], a weight row [whi | wlo]; the k loop runs hi*whi, lo*whi, hi*wlo.
template <typename T, int KC, bool FUSE_GN, bool SPLIT = false>
__global__ void __launch_bounds__(FUSE_GN ? FDG_THREADS : TC_THREADS, FUSE_GN ? FDG_MINB : 1)
conv3d_k3_fold_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_r,
                      const FoldParams p, const float* __restrict__ bias, float* __restrict__ y, double* __restrict__ stats) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_base = smem_base;
  const uint32_t a_base = smem_base + 3 * p.w_slab;
  // one ring slot = the plane the MMA reads (+ the residual plane behind it when FUSE_GN)
  const int slot_bytes = FUSE_GN ? 2 * p.plane_bytes : p.plane_bytes;
  float* stage = reinterpret_cast<float*>(smem_al + 3 * p.w_slab + p.ring * slot_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 128 * p.pitch);   // 512 * pitch bytes: 8-byte aligned
  const uint32_t full_bar = smem_u32(bars);                   // [ring]
  const uint32_t empty_bar = full_bar + 8 * p.ring;           // [ring]
  const uint32_t tfull_bar = empty_bar + 8 * p.ring;          // [FD_NB]
  const uint32_t tempty_bar = tfull_bar + 8 * FD_NB;          // [FD_NB]
  const uint32_t wfull_bar = tempty_bar + 8 * FD_NB;          // [1]
  const uint32_t xf_bar = wfull_bar + 8;                      // [ring] transform done (FUSE_GN)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * p.ring + 2 * FD_NB + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    if (FUSE_GN) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
    for (int s = 0; s < p.ring; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); mbar_init(xf_bar + 8 * s, FUSE_GN ? FDG_XW : 4); }
    for (int b = 0; b < FD_NB; ++b) { mbar_init(tfull_bar + 8 * b, 1); mbar_init(tempty_bar + 8 * b, 4); }
    mbar_init(wfull_bar, 1);
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
    {   // whole warp, converged: see elect_one()
      mbar_expect_tx_e(wfull_bar, (uint32_t)p.w_tx);
      for (int kd = 0; kd < 3; ++kd) tma_load_2d_e(w_base + kd * p.w_slab, &map_w, wfull_bar, 0, kd * p.NP);
      int stage_i = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        int t = item;
        const int seg = t % p.nseg; t /= p.nseg;
        const int x0 = (t % p.ntx) * FD_TX; t /= p.ntx;
        const int y0 = (t % p.nty) * FD_TY; const int n = t / p.nty;
        const int zs = seg * p.lseg;
        const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
        for (int ip = 0; ip < L + 2; ++ip) {
          mbar_wait(empty_bar + 8 * stage_i, phase ^ 1);
          mbar_expect_tx_e(full_bar + 8 * stage_i, (uint32_t)(FUSE_GN ? 2 * p.plane_tx : p.plane_tx));
          tma_load_5d_e(a_base + stage_i * slot_bytes, &map_x, full_bar + 8 * stage_i, 0, x0 - 1, y0 - 1, zs - 1 + ip, n);
          if (FUSE_GN)
            tma_load_5d_e(a_base + stage_i * slot_bytes + p.plane_bytes, &map_r, full_bar + 8 * stage_i, 0, x0 - 1, y0 - 1, zs - 1 + ip, n);
          if (++stage_i == p.ring) { stage_i = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, converged: see elect_one()
      mbar_wait(wfull_bar, 0);
      int stage_i = 0; uint32_t phase = 0; int oc = 0;
      constexpr int KSTEPS = KC / 16;
      const uint32_t hi = desc_hi(p.sbo, p.layout_type);
      const uint32_t slab16 = (uint32_t)p.w_slab >> 4, w16 = w_base >> 4;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const int seg = item % p.nseg;
        const int zs = seg * p.lseg;
        const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
        for (int ip = 0; ip < L + 2; ++ip) {
          mbar_wait((FUSE_GN ? xf_bar : full_bar) + 8 * stage_i, phase);
          tc_fence_after();
          const uint32_t lo_a = (a_base + stage_i * slot_bytes) >> 4;
#pragma unroll
          for (int kd = 2; kd >= 0; --kd) {
            const int zl = ip - kd;
            if (zl < 0 || zl >= L) continue;
            const int ocz = oc + zl, buf = ocz % FD_NB;
            if (kd == 0) { mbar_wait(tempty_bar + 8 * buf, ((ocz / FD_NB) & 1) ^ 1); tc_fence_after(); }
            const uint32_t dcol = tmem_base + (uint32_t)(buf * p.acc_cols);
            const uint32_t lo_w = w16 + (uint32_t)kd * slab16;
#pragma unroll
            for (int k = 0; k < (SPLIT ? KSTEPS / 2 * 3 : KSTEPS); ++k) {
              constexpr int KH = KSTEPS / 2;
              const int ka = SPLIT ? (k < 2 * KH ? k : k - 2 * KH) : k;            // hi, lo, hi
              const int kb = SPLIT ? (k < KH ? k : k - KH) : k;                    // whi, whi, wlo
              tc_mma_f16_e(dcol, desc_pack(hi, lo_a + ((ka * 32) >> 4)), desc_pack(hi, lo_w + ((kb * 32) >> 4)), p.idesc, (kd | k) != 0);
            }
          }
          tc_commit_e(empty_bar + 8 * stage_i);
          if (ip >= 2) tc_commit_e(tfull_bar + 8 * ((oc + ip - 2) % FD_NB));
          if (++stage_i == p.ring) { stage_i = 0; phase ^= 1; }
        }
        oc += L;
      }
    }
  } else if (FUSE_GN && warp >= 6) {
    // ===== transform: A = relu(gn(raw) + res), in place in the raw plane (same swizzled addresses), 64-byte rows =====
    if constexpr (FUSE_GN) {
      // thread t owns the 16-byte slots t, t+128, t+256, t+384 of a plane: their position inside the 64-byte row and the
      // swizzle phase of their rows are the same, so the 8 channels (and their scale / shift) are fixed per thread
      const int t = (warp - 6) * 32 + lane;
      const int c8 = ((t & 3) ^ ((t >> 3) & 3)) << 3;
      int hxk[FDG_NSLOT], hyk[FDG_NSLOT];
#pragma unroll
      for (int k = 0; k < FDG_NSLOT; ++k) {
        const int row = (t >> 2) + 8 * FDG_XW * k;
        hyk[k] = row < FD_HX * FD_HY ? row / FD_HX : -1000;
        hxk[k] = row - (row / FD_HX) * FD_HX;
      }
      float sa[8], sb[8], ra[8], rb[8];
      const bool rg = c8 < p.rg_nch;
      int stage_i = 0; uint32_t phase = 0; int cur_n = -1;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        int tt = item;
        const int seg = tt % p.nseg; tt /= p.nseg;
        const int x0 = (tt % p.ntx) * FD_TX; tt /= p.ntx;
        const int y0 = (tt % p.nty) * FD_TY; const int n = tt / p.nty;
        const int zs = seg * p.lseg;
        const int L = (p.D - zs) < p.lseg ? (p.D - zs) : p.lseg;
        if (n != cur_n) {
          float mean, rstd;
          gn_mean_rstd(p.gn_stats + 2 * n, p.gn_count, p.gn_eps, mean, rstd);
#pragma unroll
          for (int j = 0; j < 8; ++j) { sa[j] = rstd * p.gn_gamma[c8 + j]; sb[j] = p.gn_beta[c8 + j] - mean * sa[j]; }
          if (rg) {
            gn_mean_rstd(p.rg_stats + 2 * n, p.rg_count, p.rg_eps, mean, rstd);
#pragma unroll
            for (int j = 0; j < 8; ++j) { ra[j] = rstd * p.rg_gamma[c8 + j]; rb[j] = p.rg_beta[c8 + j] - mean * ra[j]; }
          }
          cur_n = n;
        }
        bool inb[FDG_NSLOT];
#pragma unroll
        for (int k = 0; k < FDG_NSLOT; ++k) {
          const int gx = x0 - 1 + hxk[k], gy = y0 - 1 + hyk[k];
          inb[k] = gx >= 0 && gx < p.W && gy >= 0 && gy < p.H;      // zero padding stays zero (TMA zero fill); rows >= 120 never
        }
        for (int ip = 0; ip < L + 2; ++ip) {
          mbar_wait_relaxed(full_bar + 8 * stage_i, phase);
          const int gz = zs - 1 + ip;
          if (gz >= 0 && gz < p.D) {
            uint8_t* raw = smem_al + 3 * p.w_slab + stage_i * slot_bytes + t * 16;
            const uint8_t* res = raw + p.plane_bytes;
#pragma unroll
            for (int k0 = 0; k0 < FDG_NSLOT; k0 += 2) {       // two slots in flight
              Vec8<T> a[2], r[2];
#pragma unroll
              for (int k = 0; k < 2; ++k)
                if (inb[k0 + k]) { a[k].load(reinterpret_cast<const T*>(raw + (k0 + k) * (512 * FDG_XW))); r[k].load(reinterpret_cast<const T*>(res + (k0 + k) * (512 * FDG_XW))); }
#pragma unroll
              for (int k = 0; k < 2; ++k)
                if (inb[k0 + k]) {
                  float fa[8], fr[8];
                  a[k].get(fa); r[k].get(fr);
                  if (rg) {            // the value seg3d_gn_apply would have stored: relu(fma(raw, a, b)) rounded to T
#pragma unroll
                    for (int j = 0; j < 8; ++j) fr[j] = to_f32<T>(from_f32<T>(fmaxf(fmaf(fr[j], ra[j], rb[j]), 0.f)));
                  }
#pragma unroll
                  for (int j = 0; j < 8; ++j) fa[j] = fmaxf(fmaf(fa[j], sa[j], sb[j]) + fr[j], 0.f);
                  a[k].set(fa);
                  a[k].store(reinterpret_cast<T*>(raw + (k0 + k) * (512 * FDG_XW)));
                }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(xf_bar + 8 * stage_i);
          if (++stage_i == p.ring) { stage_i = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;                  // TMEM lane = flat halo row
    const int lx = r & 7, ly = r >> 3;            // output voxel of this thread (r < 80)
    const bool is_out = r < FD_TX * FD_TY;
    const int C = p.C, nine_c = 9 * p.C, pitch = p.pitch;
    float* my_row = stage + r * pitch;
    const float* gather = stage + (r + 2 * ly) * pitch;       // flat row of (ly, lx) at tap (0,0)
    float s = 0.f, ss = 0.f;
    int cur_n = -1, oc = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      int t = item;
      const int seg = t % p.nseg; t /= p.nseg;
      const int x0 = (t % p.ntx) * FD_TX; t /= p.ntx;
      const int y0 = (t % p.nty) * FD_TY; const int n = t / p.nty;
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
      const bool valid = is_out && (gx < p.W) && (gy < p.H);
      for (int zl = 0; zl < L; ++zl) {
        const int ocz = oc + zl, buf = ocz % FD_NB;
        mbar_wait_relaxed(tfull_bar + 8 * buf, (ocz / FD_NB) & 1);
        tc_fence_after();
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.acc_cols);
        for (int c0 = 0; c0 < p.NP; c0 += 16) {
          uint32_t v[16];
          tc_ld16(tcol + (uint32_t)c0, v);
          tc_wait_ld();
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) if (c0 + jj < nine_c) my_row[c0 + jj] = __uint_as_float(v[jj]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + 8 * buf);
        named_bar_sync(1, 128);
        if (valid) {
          const size_t vox = (((size_t)n * p.D + (zs + zl)) * p.H + gy) * p.W + gx;
          float* yo = y + vox * C;
          if (C == 2) {
            float a0 = bias ? bias[0] : 0.f, a1 = bias ? bias[1] : 0.f;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const float* g = gather + (kh * FD_HX + kw) * pitch + (kh * 3 + kw) * 2;
                a0 += g[0]; a1 += g[1];
              }
            s += a0 + a1; ss += a0 * a0 + a1 * a1;
            *reinterpret_cast<float2*>(yo) = make_float2(a0, a1);
          } else {
            for (int co = 0; co < C; ++co) {
              float a = bias ? bias[co] : 0.f;
#pragma unroll
              for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
                  a += gather[(kh * FD_HX + kw) * pitch + (kh * 3 + kw) * C + co];
              s += a; ss += a * a;
              yo[co] = a;
            }
          }
        }
        named_bar_sync(1, 128);
      }
      oc += L;
    }
    if (stats && cur_n >= 0) {
      s = warp_sum(s); ss = warp_sum(ss);
      if (lane == 0) { atomicAdd(stats + 2 * cur_n, (double)s); atomicAdd(stats + 2 * cur_n + 1, (double)ss); }
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

extern "C" int seg3d_conv3d_k3_narrow_np(int C) { return C >= 1 && 9 * C <= 64 ? ((9 * C + 15) / 16) * 16 : 0; }

static int launch_fold(int dtype, const void* x, int x_ld, int Cin, const void* w, const float* bias,
                       float* y, int C, int N, int D, int H, int W, double* stats, void* stream,
                       const void* res, int res_ld, const double* gn_stats, const float* gn_gamma, const float* gn_beta, float gn_eps,
                       bool split = false, int rg_nch = 0, const double* rg_stats = nullptr, const float* rg_gamma = nullptr,
                       const float* rg_beta = nullptr, float rg_eps = 0.f) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool fuse = res != nullptr;
  if (split) {       // rows [hi(Cin) | lo(Cin)]: one K block of 2*Cin channels
    SEG3D_REQUIRE(!fuse && dtype == SEG3D_F16 && Cin == 32 && x_ld >= 64, "conv3d_k3_narrow_split_fwd: Cin must be 32 with rows [hi | lo] of 64 f16 values");
    Cin = 64;
  }
  if (fuse) {
    SEG3D_REQUIRE(Cin == 32, "conv3d_k3_narrow_gn_fwd: Cin must be 32 (64-byte rows), got %d", Cin);
    SEG3D_REQUIRE(gn_stats && gn_gamma && gn_beta && res_ld % 8 == 0 && ((uintptr_t)res) % 16 == 0, "conv3d_k3_narrow_gn_fwd: bad GroupNorm / residual arguments");
  }
  SEG3D_REQUIRE(x && w && y, "conv3d_k3_narrow_fwd: null pointer");
  SEG3D_REQUIRE(dtype == SEG3D_F16 || dtype == SEG3D_BF16, "conv3d_k3_narrow_fwd: dtype must be f16 or bf16");
  SEG3D_REQUIRE(Cin == 16 || Cin == 32 || Cin == 64, "conv3d_k3_narrow_fwd: Cin must be 16, 32 or 64 (got %d)", Cin);
  const int NP = seg3d_conv3d_k3_narrow_np(C);
  SEG3D_REQUIRE(NP > 0, "conv3d_k3_narrow_fwd: 1 <= Cout <= 7 (got %d)", C);
  SEG3D_REQUIRE(W % 8 == 0 && x_ld % 8 == 0 && N > 0 && D > 0 && H > 0, "conv3d_k3_narrow_fwd: bad dims / pitch");
  SEG3D_REQUIRE(((uintptr_t)x) % 16 == 0 && ((uintptr_t)w) % 16 == 0 && ((uintptr_t)y) % 8 == 0, "conv3d_k3_narrow_fwd: misaligned pointer");
  EncodeTiledFn encode = get_encode();
  if (!encode) { seg3d_set_error("conv3d_k3_narrow_fwd: cuTensorMapEncodeTiled entry point not available"); return SEG3D_ECUDA; }

  FoldParams p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.NP = NP; p.row_bytes = Cin * 2;
  p.D = D; p.H = H; p.W = W; p.N = N;
  p.w_slab = (NP * p.row_bytes + 1023) & ~1023;
  p.w_tx = 3 * NP * p.row_bytes;
  p.plane_tx = FD_HX * FD_HY * p.row_bytes;
  p.plane_bytes = (128 * p.row_bytes + 1023) & ~1023;        // the MMA reads 128 flat rows; rows 120..127 only reach unused accumulator rows
  p.pitch = (9 * C) | 1;
  p.ntx = W / FD_TX; p.nty = (H + FD_TY - 1) / FD_TY;
  p.acc_cols = NP <= 32 ? 32 : 64;
  p.tmem_cols = FD_NB * p.acc_cols;
  int ctas_per_sm = env_int("SEG3D_FD_CTAS_PER_SM", fuse ? FDG_MINB : 4);
  if (fuse && ctas_per_sm > FDG_MINB) ctas_per_sm = FDG_MINB;       // the fused variant is compiled for FDG_MINB co-resident CTAs
  if (ctas_per_sm * p.tmem_cols > 512) ctas_per_sm = 512 / p.tmem_cols;
  const int fixed = 3 * p.w_slab + 128 * p.pitch * 4 + 1024;
  int ring = 0;
  const int slot_bytes = fuse ? 2 * p.plane_bytes : p.plane_bytes;
  for (; ctas_per_sm >= 1; --ctas_per_sm) {                  // as many co-resident CTAs as leave a ring of >= 4 planes (3 fused)
    ring = ((200 * 1024 / ctas_per_sm) - fixed) / slot_bytes;
    if (ring >= (fuse ? 3 : 4) || ctas_per_sm == 1) break;
  }
  if (ring > 8) ring = 8;
  { const int e = env_int("SEG3D_FD_RING", 0); if (e >= 3) ring = e; }
  SEG3D_REQUIRE(ring >= 3, "conv3d_k3_narrow_fwd: plane ring does not fit in shared memory");
  p.ring = ring;
  const long long cols = (long long)N * p.ntx * p.nty;
  const long long want = 4ll * ctas_per_sm * seg3d_num_sms();
  int nseg = (int)((want + cols - 1) / cols);
  if (nseg < 1) nseg = 1;
  int lseg = (D + nseg - 1) / nseg;
  if (lseg < 8) lseg = D < 8 ? D : 8;
  { const int e = env_int("SEG3D_FD_LSEG", 0); if (e >= 1) lseg = e; }
  p.lseg = lseg; p.nseg = (D + lseg - 1) / lseg;
  const long long nitems = cols * p.nseg;
  SEG3D_REQUIRE(nitems > 0 && nitems < (1ll << 31), "conv3d_k3_narrow_fwd: work-item count out of range");
  p.nitems = (int)nitems;
  p.gn_eps = gn_eps; p.gn_count = (double)Cin * D * H * W; p.gn_stats = gn_stats; p.gn_gamma = gn_gamma; p.gn_beta = gn_beta;
  SEG3D_REQUIRE(rg_nch == 0 || (fuse && rg_nch % 8 == 0 && rg_nch <= Cin && rg_stats && rg_gamma && rg_beta), "conv3d_k3_narrow_gn2_fwd: bad residual GroupNorm arguments");
  p.rg_nch = rg_nch; p.rg_eps = rg_eps; p.rg_count = (double)rg_nch * D * H * W; p.rg_stats = rg_stats; p.rg_gamma = rg_gamma; p.rg_beta = rg_beta;
  p.sbo = 8 * p.row_bytes;
  p.layout_type = p.row_bytes == 128 ? 2u : (p.row_bytes == 64 ? 4u : 6u);
  const uint32_t fmt = dtype == SEG3D_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  const CUtensorMapDataType tdt = dtype == SEG3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap map_x, map_w, map_r;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)x_ld * 2, (cuuint64_t)W * x_ld * 2, (cuuint64_t)H * W * x_ld * 2, (cuuint64_t)D * H * W * x_ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)Cin, FD_HX, FD_HY, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map_x, tdt, 5, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seg3d_set_error("conv3d_k3_narrow_fwd: cuTensorMapEncodeTiled(x) failed with %d", (int)r); return SEG3D_ECUDA; }
    map_r = map_x;
    if (fuse) {
      cuuint64_t rstr[4] = {(cuuint64_t)res_ld * 2, (cuuint64_t)W * res_ld * 2, (cuuint64_t)H * W * res_ld * 2, (cuuint64_t)D * H * W * res_ld * 2};
      r = encode(&map_r, tdt, 5, const_cast<void*>(res), dims, rstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { seg3d_set_error("conv3d_k3_narrow_gn_fwd: cuTensorMapEncodeTiled(res) failed with %d", (int)r); return SEG3D_ECUDA; }
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)3 * NP};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)Cin, (cuuint32_t)NP};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map_w, tdt, 2, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seg3d_set_error("conv3d_k3_narrow_fwd: cuTensorMapEncodeTiled(w) failed with %d", (int)r); return SEG3D_ECUDA; }
  }
  const size_t smem = 1024 + (size_t)3 * p.w_slab + (size_t)ring * slot_bytes + (size_t)(128 * p.pitch + 1) * 4 +
                      (3 * ring + 2 * FD_NB + 1) * 8 + 64 + 512;
  const long long max_grid = (long long)ctas_per_sm * seg3d_num_sms();
  dim3 grid((unsigned)(nitems < max_grid ? nitems : max_grid));
  cudaError_t e = cudaSuccess;
#define SEG3D_LAUNCH_F(TT, KCV)                                                                                                 \
  { e = cudaFuncSetAttribute(conv3d_k3_fold_kernel<TT, KCV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    if (e == cudaSuccess) { conv3d_k3_fold_kernel<TT, KCV, false><<<grid, TC_THREADS, smem, st>>>(map_x, map_w, map_r, p, bias, y, stats); e = cudaGetLastError(); } }
#define SEG3D_LAUNCH_FG(TT)                                                                                                     \
  { e = cudaFuncSetAttribute(conv3d_k3_fold_kernel<TT, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e == cudaSuccess) { conv3d_k3_fold_kernel<TT, 32, true><<<grid, FDG_THREADS, smem, st>>>(map_x, map_w, map_r, p, bias, y, stats); e = cudaGetLastError(); } }
  if (split) {
    e = cudaFuncSetAttribute(conv3d_k3_fold_kernel<__half, 64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) { conv3d_k3_fold_kernel<__half, 64, false, true><<<grid, TC_THREADS, smem, st>>>(map_x, map_w, map_r, p, bias, y, stats); e = cudaGetLastError(); }
  } else if (fuse) {
    if (dtype == SEG3D_BF16) SEG3D_LAUNCH_FG(__nv_bfloat16) else SEG3D_LAUNCH_FG(__half)
  } else if (dtype == SEG3D_BF16) {
    if (Cin == 64) SEG3D_LAUNCH_F(__nv_bfloat16, 64) else if (Cin == 32) SEG3D_LAUNCH_F(__nv_bfloat16, 32) else SEG3D_LAUNCH_F(__nv_bfloat16, 16)
  } else {
    if (Cin == 64) SEG3D_LAUNCH_F(__half, 64) else if (Cin == 32) SEG3D_LAUNCH_F(__half, 32) else SEG3D_LAUNCH_F(__half, 16)
  }
#undef SEG3D_LAUNCH_F
#undef SEG3D_LAUNCH_FG
  if (e != cudaSuccess) { seg3d_set_error("conv3d_k3_fold_kernel launch failed: %s", cudaGetErrorString(e)); return SEG3D_ECUDA; }
  return SEG3D_OK;
}

extern "C" int seg3d_conv3d_k3_narrow_fwd(int dtype, const void* x, int x_ld, int Cin, const void* w, const float* bias,
                                          float* y, int C, int N, int D, int H, int W, double* stats, void* stream) {
  return launch_fold(dtype, x, x_ld, Cin, w, bias, y, C, N, D, H, W, stats, stream, nullptr, 0, nullptr, nullptr, nullptr, 0.f);
}

// Same convolution on x = relu(GroupNorm(raw) + res): the last GroupNorm + residual + ReLU of the network
// (residual_block3.py:24 of up_32.rblock) is formed in shared memory between the TMA loads and the MMAs instead of being
// streamed through HBM by seg3d_gn_apply.  raw / res: [N,D,H,W,32] f16/bf16; gn_stats: finished sums of raw.
extern "C" int seg3d_conv3d_k3_narrow_gn_fwd(int dtype, const void* raw, int raw_ld, const void* res, int res_ld, int Cin,
                                             const double* gn_stats, const float* gamma, const float* beta, float eps,
                                             const void* w, const float* bias, float* y, int C, int N, int D, int H, int W,
                                             double* stats, void* stream) {
  SEG3D_REQUIRE(res != nullptr, "conv3d_k3_narrow_gn_fwd: null residual");
  return launch_fold(dtype, raw, raw_ld, Cin, w, bias, y, C, N, D, H, W, stats, stream, res, res_ld, gn_stats, gamma, beta, eps);
}

// Strict-parity variant (split operands, see seg3d_conv3d_split_fwd): x rows are [hi(32) | lo(32)] f16 (pitch x_ld >= 64, the lo
// half directly behind the hi half), w is [3 kd][NP][whi(32) | wlo(32)] f16; the MMAs accumulate hi*whi + lo*whi + hi*wlo.
extern "C" int seg3d_conv3d_k3_narrow_split_fwd(const void* x, int x_ld, int Cin, const void* w, const float* bias,
                                                float* y, int C, int N, int D, int H, int W, double* stats, void* stream) {
  return launch_fold(SEG3D_F16, x, x_ld, Cin, w, bias, y, C, N, D, H, W, stats, stream, nullptr, 0, nullptr, nullptr, nullptr, 0.f, true);
}

// seg3d_conv3d_k3_narrow_gn_fwd whose residual's first res_gn_ch channels are still a raw convolution result: the kernel uses
// relu(GroupNorm(1, res_gn_ch)(res[:, :res_gn_ch])) for them (finished sums res_gn_stats), rounded to the storage type exactly as
// seg3d_gn_apply would have stored it, so the result is bit-identical to the unfused sequence.
extern "C" int seg3d_conv3d_k3_narrow_gn2_fwd(int dtype, const void* raw, int raw_ld, const void* res, int res_ld, int Cin,
                                              const double* gn_stats, const float* gamma, const float* beta, float eps,
                                              int res_gn_ch, const double* res_gn_stats, const float* res_gamma, const float* res_beta,
                                              const void* w, const float* bias, float* y, int C, int N, int D, int H, int W,
                                              double* stats, void* stream) {
  SEG3D_REQUIRE(res != nullptr, "conv3d_k3_narrow_gn2_fwd: null residual");
  return launch_fold(dtype, raw, raw_ld, Cin, w, bias, y, C, N, D, H, W, stats, stream, res, res_ld, gn_stats, gamma, beta, eps,
                     false, res_gn_ch, res_gn_stats, res_gamma, res_beta, eps);
}
