// resample.cu - trilinear / nearest-neighbour resampling between two grids that share origin and direction
// (reference utils/image_tools.py:329-377: sitk.Resample with an identity transform, called by
// core/seg_infer.py:267 to bring the scan to the model spacing and by :330-333 to bring every class probability map
// back to the scan's grid).
//
// With a shared origin and direction the output index i maps to the continuous input index c = i * (spacing_out /
// spacing_in) per axis.  Restated ITK semantics (ResampleImageFilter + LinearInterpolateImageFunction /
// NearestNeighborInterpolateImageFunction, double-precision coordinates and weights):
//   * c is inside the buffer when -0.5 <= c < size - 0.5 on every axis, otherwise the default pixel value is written;
//   * linear: base = floor(c), weights from c - base, the upper neighbour index is clamped to size - 1;
//   * nearest: index = floor(c + 0.5).
// HBM-bound streaming kernel: one thread per output voxel, x fastest (coalesced writes, cached gathers).
#include "common.cuh"

namespace {

template <int LINEAR>
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ src, int sz, int sy, int sx, float* __restrict__ dst, int dz, int dy, int dx,
                double rz, double ry, double rx, float dflt) {
  const size_t total = (size_t)dz * dy * dx;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % dx); const size_t t = i / dx; const int y = (int)(t % dy), z = (int)(t / dy);
    const double cx = x * rx, cy = y * ry, cz = z * rz;
    float out = dflt;
    if (cx < sx - 0.5 && cy < sy - 0.5 && cz < sz - 0.5) {          // c >= 0 > -0.5 always
      if (LINEAR) {
        const double fx = floor(cx), fy = floor(cy), fz = floor(cz);
        const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
        const int x1 = x0 + 1 < sx ? x0 + 1 : sx - 1, y1 = y0 + 1 < sy ? y0 + 1 : sy - 1, z1 = z0 + 1 < sz ? z0 + 1 : sz - 1;
        const double wx = cx - fx, wy = cy - fy, wz = cz - fz;
        const size_t r00 = ((size_t)z0 * sy + y0) * sx, r01 = ((size_t)z0 * sy + y1) * sx;
        const size_t r10 = ((size_t)z1 * sy + y0) * sx, r11 = ((size_t)z1 * sy + y1) * sx;
        const double v000 = src[r00 + x0], v001 = src[r00 + x1], v010 = src[r01 + x0], v011 = src[r01 + x1];
        const double v100 = src[r10 + x0], v101 = src[r10 + x1], v110 = src[r11 + x0], v111 = src[r11 + x1];
        const double a00 = v000 + (v001 - v000) * wx, a01 = v010 + (v011 - v010) * wx;
        const double a10 = v100 + (v101 - v100) * wx, a11 = v110 + (v111 - v110) * wx;
        const double b0 = a00 + (a01 - a00) * wy, b1 = a10 + (a11 - a10) * wy;
        out = (float)(b0 + (b1 - b0) * wz);
      } else {
        const int xn = (int)floor(cx + 0.5), yn = (int)floor(cy + 0.5), zn = (int)floor(cz + 0.5);
        out = src[((size_t)zn * sy + yn) * sx + xn];
      }
    }
    dst[i] = out;
  }
}

// Crop with resampling (reference utils/image_tools.py:111-146: sitk.Resample onto a grid with its own origin and spacing
// and the volume's direction): output index i reads the continuous input index c = o + i * r per axis, o = (crop origin -
// volume origin) expressed in input voxels.  Same ITK semantics as above, with c allowed to be negative: inside means
// -0.5 <= c < size - 0.5, and BOTH neighbours of the linear interpolation are clamped to [0, size - 1].
template <int LINEAR>
__global__ void __launch_bounds__(256)
crop_resample_kernel(const float* __restrict__ src, int sz, int sy, int sx, float* __restrict__ dst, int dz, int dy, int dx,
                     double oz, double oy, double ox, double rz, double ry, double rx, float dflt) {
  const size_t total = (size_t)dz * dy * dx;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % dx); const size_t t = i / dx; const int y = (int)(t % dy), z = (int)(t / dy);
    // product and sum rounded separately (no FMA contraction): the host restatement computes o + i * r in two double
    // operations, and at an exact half-voxel tie a fused product would pick the other nearest neighbour
    const double cx = __dadd_rn(ox, __dmul_rn((double)x, rx)), cy = __dadd_rn(oy, __dmul_rn((double)y, ry)), cz = __dadd_rn(oz, __dmul_rn((double)z, rz));
    float out = dflt;
    if (cx >= -0.5 && cy >= -0.5 && cz >= -0.5 && cx < sx - 0.5 && cy < sy - 0.5 && cz < sz - 0.5) {
      if (LINEAR) {
        const double fx = floor(cx), fy = floor(cy), fz = floor(cz);
        const int bx = (int)fx, by = (int)fy, bz = (int)fz;
        const int x0 = bx < 0 ? 0 : bx, y0 = by < 0 ? 0 : by, z0 = bz < 0 ? 0 : bz;          // bx >= -1, bx <= size - 1
        const int x1 = bx + 1 < sx ? bx + 1 : sx - 1, y1 = by + 1 < sy ? by + 1 : sy - 1, z1 = bz + 1 < sz ? bz + 1 : sz - 1;
        const double wx = cx - fx, wy = cy - fy, wz = cz - fz;
        const size_t r00 = ((size_t)z0 * sy + y0) * sx, r01 = ((size_t)z0 * sy + y1) * sx;
        const size_t r10 = ((size_t)z1 * sy + y0) * sx, r11 = ((size_t)z1 * sy + y1) * sx;
        const double v000 = src[r00 + x0], v001 = src[r00 + x1], v010 = src[r01 + x0], v011 = src[r01 + x1];
        const double v100 = src[r10 + x0], v101 = src[r10 + x1], v110 = src[r11 + x0], v111 = src[r11 + x1];
        const double a00 = v000 + (v001 - v000) * wx, a01 = v010 + (v011 - v010) * wx;
        const double a10 = v100 + (v101 - v100) * wx, a11 = v110 + (v111 - v110) * wx;
        const double b0 = a00 + (a01 - a00) * wy, b1 = a10 + (a11 - a10) * wy;
        out = (float)(b0 + (b1 - b0) * wz);
      } else {
        const int xn = (int)floor(cx + 0.5), yn = (int)floor(cy + 0.5), zn = (int)floor(cz + 0.5);   // in [0, size - 1]
        out = src[((size_t)zn * sy + yn) * sx + xn];
      }
    }
    dst[i] = out;
  }
}

}  // namespace

extern "C" int seg3d_crop_resample(const float* src, int sz, int sy, int sx, float* dst, int dz, int dy, int dx,
                                   double oz, double oy, double ox, double rz, double ry, double rx,
                                   int linear, float default_value, void* stream) {
  SEG3D_REQUIRE(src && dst && sz > 0 && sy > 0 && sx > 0 && dz > 0 && dy > 0 && dx > 0, "crop_resample: bad arguments");
  SEG3D_REQUIRE(rz > 0.0 && ry > 0.0 && rx > 0.0, "crop_resample: spacing ratios must be positive");
  const size_t total = (size_t)dz * dy * dx;
  const int sms = seg3d_num_sms();
  size_t want = (total + 255) / 256;
  const int gx = (int)(want > (size_t)32 * sms ? (size_t)32 * sms : want);
  if (linear) crop_resample_kernel<1><<<gx, 256, 0, (cudaStream_t)stream>>>(src, sz, sy, sx, dst, dz, dy, dx, oz, oy, ox, rz, ry, rx, default_value);
  else        crop_resample_kernel<0><<<gx, 256, 0, (cudaStream_t)stream>>>(src, sz, sy, sx, dst, dz, dy, dx, oz, oy, ox, rz, ry, rx, default_value);
  SEG3D_CHECK_LAUNCH("crop_resample_kernel");
  return SEG3D_OK;
}

extern "C" int seg3d_resample(const float* src, int sz, int sy, int sx, float* dst, int dz, int dy, int dx,
                              double rz, double ry, double rx, int linear, float default_value, void* stream) {
  SEG3D_REQUIRE(src && dst && sz > 0 && sy > 0 && sx > 0 && dz > 0 && dy > 0 && dx > 0, "resample: bad arguments");
  SEG3D_REQUIRE(rz > 0.0 && ry > 0.0 && rx > 0.0, "resample: spacing ratios must be positive");
  const size_t total = (size_t)dz * dy * dx;
  const int sms = seg3d_num_sms();
  size_t want = (total + 255) / 256;
  const int gx = (int)(want > (size_t)32 * sms ? (size_t)32 * sms : want);
  if (linear) resample_kernel<1><<<gx, 256, 0, (cudaStream_t)stream>>>(src, sz, sy, sx, dst, dz, dy, dx, rz, ry, rx, default_value);
  else        resample_kernel<0><<<gx, 256, 0, (cudaStream_t)stream>>>(src, sz, sy, sx, dst, dz, dy, dx, rz, ry, rx, default_value);
  SEG3D_CHECK_LAUNCH("resample_kernel");
  return SEG3D_OK;
}
