// On-GPU ray generation: the reference builds every batch on the host with numpy
// inside an 8-worker DataLoader (Z/train.py:111-118) -- camera rays by
// camera_utils.pixels_to_rays / cast_ray_batch (Z/internal/camera_utils.py:454-617), LiDAR rays by
// lidar_utils.cast_lidar_ray_batch (Z/internal/lidar_utils.py:8-33) over a direction table from
// get_directions (:559-568).  At > 1 M rays/s that loader is the bottleneck; here the cameras and
// the LiDAR tables stay resident in HBM and one thread produces one ray.
//
// numpy computes these in float64 and the batch is cast to float32 at the very end
// (datasets.py: `torch.from_numpy(v.copy()).float()`); the kernels do the same -- float64
// arithmetic in the reference's operation order, one rounding to float32 at the store -- so the
// outputs equal the reference's up to the last-bit freedom of its BLAS / pairwise sums.
#include "common.cuh"
#include "../../include/nlb200.h"

namespace nlb {
namespace raygen {

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 matvec3(const double* __restrict__ A, int row_stride, D3 v) {
  // matmul(A, b[..., None])[..., 0]: sum over j in index order
  D3 o;
  o.x = A[0] * v.x + A[1] * v.y + A[2] * v.z;
  o.y = A[row_stride] * v.x + A[row_stride + 1] * v.y + A[row_stride + 2] * v.z;
  o.z = A[2 * row_stride] * v.x + A[2 * row_stride + 1] * v.y + A[2 * row_stride + 2] * v.z;
  return o;
}
__device__ __forceinline__ double norm3(D3 v) { return sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
__device__ __forceinline__ void store3(float* __restrict__ p, size_t i, D3 v) {
  p[3 * i] = (float)v.x; p[3 * i + 1] = (float)v.y; p[3 * i + 2] = (float)v.z;
}

// camera_utils.pixels_to_rays, perspective camera, no distortion, no NDC (the nuScenes loader,
// datasets.py:1183-1234): one thread per pixel.
__global__ void k_camera_rays(const int32_t* __restrict__ pix_x, const int32_t* __restrict__ pix_y,
                              const int32_t* __restrict__ cam_idx, const double* __restrict__ pixtocams, int n_p2c,
                              const double* __restrict__ camtoworlds, int n_c2w, int64_t n, nlb_ray_out_t o) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cam = cam_idx ? cam_idx[i] : 0;
  // `batch_index = lambda arr: arr if arr.ndim == 2 else arr[cam_idx]` (camera_utils.py:591): a single
  // matrix is shared by every ray
  const double* P = pixtocams + (n_p2c > 1 ? (size_t)cam * 9 : 0);
  const double* C = camtoworlds + (n_c2w > 1 ? (size_t)cam * 12 : 0);
  const double x = (double)pix_x[i], y = (double)pix_y[i];
  // pixel centre and its +x / +y neighbours (the cone radii need them), through the inverse
  // intrinsics, OpenCV -> OpenGL (y and z negated), then the camera rotation
  D3 dir[3];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    D3 p = {x + (s == 1 ? 1.0 : 0.0) + 0.5, y + (s == 2 ? 1.0 : 0.0) + 0.5, 1.0};
    D3 c = matvec3(P, 3, p);
    c.y = -c.y; c.z = -c.z;
    if (s == 0 && o.imageplane) { o.imageplane[2 * i] = (float)c.x; o.imageplane[2 * i + 1] = (float)c.y; }
    dir[s] = matvec3(C, 4, c);
  }
  const D3 d = dir[0];
  const double dn = norm3(d);
  const D3 ex = {dir[1].x - d.x, dir[1].y - d.y, dir[1].z - d.z};
  const D3 ey = {dir[2].x - d.x, dir[2].y - d.y, dir[2].z - d.z};
  const double nx = norm3(ex), ny = norm3(ey);
  store3(o.origins, i, D3{C[3], C[7], C[11]});
  store3(o.directions, i, d);
  store3(o.viewdirs, i, D3{d.x / dn, d.y / dn, d.z / dn});
  store3(o.base_x, i, D3{ex.x / nx, ex.y / nx, ex.z / nx});
  store3(o.base_y, i, D3{ey.x / ny, ey.y / ny, ey.z / ny});
  // half the mean distance to the neighbours, scaled to the variance of a uniform pixel footprint
  o.radii[i] = (float)((0.5 * (nx + ny)) * 2.0 / sqrt(12.0));
}

// lidar_utils.get_directions: beam-major table, float64 trigonometry, float32 result
__global__ void k_lidar_directions(const double* __restrict__ elev_deg, int n_beams, const double* __restrict__ azim_rad,
                                   int width, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_beams * width) return;
  const double theta = elev_deg[i / width] / 180.0 * 3.141592653589793;
  const double phi = azim_rad[i % width];
  const double ct = cos(theta);
  out[3 * i] = (float)(ct * sin(phi));
  out[3 * i + 1] = (float)(ct * cos(phi));
  out[3 * i + 2] = (float)sin(theta);
}

// sum of squares of the whole direction array (np.linalg.norm without an axis: the GLOBAL Frobenius norm)
__global__ void k_sumsq(const float* __restrict__ d, int64_t count, double* __restrict__ acc) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)d[i];
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}

// lidar_utils.cast_lidar_ray_batch: origins / directions pass through, viewdirs = directions / the global
// norm, radii = 5e-4, base_x = base_y = directions
__global__ void k_lidar_rays(const float* __restrict__ origins, const float* __restrict__ directions, int64_t n,
                             const double* __restrict__ sumsq, nlb_ray_out_t o) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double inv = 1.0 / sqrt(*sumsq);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float dv = directions[3 * i + c];
    o.origins[3 * i + c] = origins[3 * i + c];
    o.directions[3 * i + c] = dv;
    o.viewdirs[3 * i + c] = (float)((double)dv * inv);
    o.base_x[3 * i + c] = dv;
    o.base_y[3 * i + c] = dv;
  }
  o.radii[i] = 0.0005f;
  if (o.imageplane) { o.imageplane[2 * i] = 0.f; o.imageplane[2 * i + 1] = 0.f; }
}

static bool outputs_ok(const nlb_ray_out_t* o) {
  return o && o->origins && o->directions && o->viewdirs && o->radii && o->base_x && o->base_y;
}

}  // namespace raygen
}  // namespace nlb

using namespace nlb;

extern "C" int nlb_camera_rays(const int32_t* pix_x, const int32_t* pix_y, const int32_t* cam_idx, const double* pixtocams,
                               int n_pixtocams, const double* camtoworlds, int n_camtoworlds, int64_t n,
                               const nlb_ray_out_t* out, void* stream) {
  if (n == 0) return NLB_OK;
  if (n < 0 || !pix_x || !pix_y || !pixtocams || !camtoworlds || !raygen::outputs_ok(out)) {
    nlb_set_error("camera_rays: null pointer or negative ray count");
    return NLB_EINVAL;
  }
  if (n_pixtocams < 1 || n_camtoworlds < 1) { nlb_set_error("camera_rays: at least one camera matrix is required"); return NLB_EINVAL; }
  if (!cam_idx && (n_pixtocams > 1 || n_camtoworlds > 1)) {
    nlb_set_error("camera_rays: cam_idx is required with more than one camera");
    return NLB_EINVAL;
  }
  const int threads = 128;
  raygen::k_camera_rays<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      pix_x, pix_y, cam_idx, pixtocams, n_pixtocams, camtoworlds, n_camtoworlds, n, *out);
  return nlb_check_launch("camera_rays");
}

extern "C" int nlb_lidar_directions(const double* elev_deg, int n_beams, const double* azim_rad, int width, float* out,
                                    void* stream) {
  if (n_beams == 0 || width == 0) return NLB_OK;
  if (n_beams < 0 || width < 0 || !elev_deg || !azim_rad || !out) { nlb_set_error("lidar_directions: bad argument"); return NLB_EINVAL; }
  const int total = n_beams * width, threads = 128;
  raygen::k_lidar_directions<<<(total + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(elev_deg, n_beams, azim_rad,
                                                                                                      width, out);
  return nlb_check_launch("lidar_directions");
}

extern "C" int nlb_lidar_rays(const float* origins, const float* directions, int64_t n, double* workspace,
                              const nlb_ray_out_t* out, void* stream) {
  if (n == 0) return NLB_OK;
  if (n < 0 || !origins || !directions || !workspace || !raygen::outputs_ok(out)) {
    nlb_set_error("lidar_rays: null pointer or negative ray count");
    return NLB_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(workspace, 0, sizeof(double), st);
  const int threads = 256;
  int blocks = (int)((3 * n + threads - 1) / threads);
  if (blocks > 592) blocks = 592;
  raygen::k_sumsq<<<blocks, threads, 0, st>>>(directions, 3 * n, workspace);
  raygen::k_lidar_rays<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(origins, directions, n, workspace, *out);
  return nlb_check_launch("lidar_rays");
}
