// Ray generation and geometric lifts (K13 / K14 of SURVEY.md §2.3).  HBM-trivial elementwise kernels.
#include "svs_common.cuh"

namespace svs {

struct Cam {
  float fx, fy, cx, cy, sk;
  float r[9];
  float t[3];
};

__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
  float n = sqrtf(x * x + y * y + z * z);
  n = fmaxf(n, 1e-12f);  // F.normalize eps (rend_util.py:93)
  x /= n;
  y /= n;
  z /= n;
}

// rend_util.get_camera_params + lift (rend_util.py:60-95,143-156), evaluated for the real pose and for
// the identity pose (depth_scale, network.py:216-217) in one pass.
__global__ void raygen_kernel(const float* __restrict__ uv, const float* __restrict__ pose,
                              const float* __restrict__ K, int64_t R, float* __restrict__ dirs,
                              float* __restrict__ cam_loc, float* __restrict__ depth_scale) {
  __shared__ Cam c;
  if (threadIdx.x == 0) {
    c.fx = K[0];
    c.sk = K[1];
    c.cx = K[2];
    c.fy = K[5];
    c.cy = K[6];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) c.r[i * 3 + j] = pose[i * 4 + j];
      c.t[i] = pose[i * 4 + 3];
    }
  }
  __syncthreads();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R) return;
  float u = uv[2 * i], v = uv[2 * i + 1];
  float xl = (u - c.cx + c.cy * c.sk / c.fy - c.sk * v / c.fy) / c.fx;
  float yl = (v - c.cy) / c.fy;
  // world = R p + t ; dir = normalize(world - cam_loc)
  float wx = c.r[0] * xl + c.r[1] * yl + c.r[2] + c.t[0];
  float wy = c.r[3] * xl + c.r[4] * yl + c.r[5] + c.t[1];
  float wz = c.r[6] * xl + c.r[7] * yl + c.r[8] + c.t[2];
  float dx = wx - c.t[0], dy = wy - c.t[1], dz = wz - c.t[2];
  normalize3(dx, dy, dz);
  dirs[3 * i] = dx;
  dirs[3 * i + 1] = dy;
  dirs[3 * i + 2] = dz;
  cam_loc[3 * i] = c.t[0];
  cam_loc[3 * i + 1] = c.t[1];
  cam_loc[3 * i + 2] = c.t[2];
  float cxn = xl, cyn = yl, czn = 1.0f;
  normalize3(cxn, cyn, czn);
  depth_scale[i] = czn;
}

// rend_util.get_sphere_intersections (rend_util.py:200-216)
__global__ void sphere_kernel(const float* __restrict__ cam, const float* __restrict__ dirs, int64_t R,
                              float radius, float* __restrict__ nf, int32_t* __restrict__ bad) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R) return;
  float ox = cam[3 * i], oy = cam[3 * i + 1], oz = cam[3 * i + 2];
  float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
  float dot = dx * ox + dy * oy + dz * oz;
  float on = sqrtf(ox * ox + oy * oy + oz * oz);
  float under = dot * dot - (on * on - radius * radius);
  if (!(under > 0.0f)) atomicExch(bad, 1);
  float s = sqrtf(under);
  nf[2 * i] = fmaxf(-s - dot, 0.0f);
  nf[2 * i + 1] = fmaxf(s - dot, 0.0f);
}

// points = cam_loc + z * dir (network.py:227-228)
__global__ void points_kernel(const float* __restrict__ cam, const float* __restrict__ dirs,
                              const float* __restrict__ z, int64_t R, int S, int ldz,
                              float* __restrict__ pts) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  int64_t r = i / S;
  int s = (int)(i - r * S);
  float zz = z[r * ldz + s];
  pts[3 * i] = cam[3 * r] + zz * dirs[3 * r];
  pts[3 * i + 1] = cam[3 * r + 1] + zz * dirs[3 * r + 1];
  pts[3 * i + 2] = cam[3 * r + 2] + zz * dirs[3 * r + 2];
}

// NeRF++ inverted-sphere lift (network_bg.py:182-214): rotate the sphere exit point towards the ray
// direction by phi - theta (Rodrigues) and append 1/r as the 4th coordinate.
__global__ void depth2pts_kernel(const float* __restrict__ cam, const float* __restrict__ dirs,
                                 const float* __restrict__ depth, int64_t R, int S, float radius,
                                 float* __restrict__ pts, float* __restrict__ depth_real) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  int64_t r = i / S;
  float ox = cam[3 * r], oy = cam[3 * r + 1], oz = cam[3 * r + 2];
  float dx = dirs[3 * r], dy = dirs[3 * r + 1], dz = dirs[3 * r + 2];
  float inv_r = depth[i];
  float od = dx * ox + dy * oy + dz * oz;
  float under = od * od - ((ox * ox + oy * oy + oz * oz) - radius * radius);
  float dsph = sqrtf(under) - od;
  float px = ox + dsph * dx, py = oy + dsph * dy, pz = oz + dsph * dz;  // exit point on the sphere
  float mx = ox - od * dx, my = oy - od * dy, mz = oz - od * dz;        // closest point to the origin
  float mn = sqrtf(mx * mx + my * my + mz * mz);
  float ax = oy * pz - oz * py, ay = oz * px - ox * pz, az = ox * py - oy * px;  // cross(o, p_sphere)
  float an = sqrtf(ax * ax + ay * ay + az * az);
  ax /= an;
  ay /= an;
  az /= an;
  float phi = asinf(mn / radius);
  float theta = asinf(mn * inv_r);
  float ang = phi - theta;
  float ca = cosf(ang), sa = sinf(ang);
  float kx = ay * pz - az * py, ky = az * px - ax * pz, kz = ax * py - ay * px;  // cross(axis, p)
  float adp = ax * px + ay * py + az * pz;
  float nx = px * ca + kx * sa + ax * adp * (1.0f - ca);
  float ny = py * ca + ky * sa + ay * adp * (1.0f - ca);
  float nz = pz * ca + kz * sa + az * adp * (1.0f - ca);
  float nn = sqrtf(nx * nx + ny * ny + nz * nz);
  pts[4 * i] = nx / nn;
  pts[4 * i + 1] = ny / nn;
  pts[4 * i + 2] = nz / nn;
  pts[4 * i + 3] = inv_r;
  float dd = dx * dx + dy * dy + dz * dz;
  float d1 = -od / dd;
  float dcos = 1.0f / sqrtf(dd);
  depth_real[i] = 1.0f / (inv_r + 1e-6f) * cosf(theta) * dcos + d1;
}

}  // namespace svs

using namespace svs;

extern "C" int svs_raygen(const float* uv, const float* pose, const float* intrinsics, int64_t R,
                          float* ray_dirs, float* cam_loc, float* depth_scale, void* stream) {
  SVS_CHECK_ARG(R >= 0 && uv && pose && intrinsics && ray_dirs && cam_loc && depth_scale, "svs_raygen: null/neg");
  if (R == 0) return SVS_OK;
  raygen_kernel<<<(unsigned)cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(uv, pose, intrinsics, R, ray_dirs,
                                                                          cam_loc, depth_scale);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_sphere_intersections(const float* cam_loc, const float* ray_dirs, int64_t R, float radius,
                                        float* near_far, int32_t* bad_flag, void* stream) {
  SVS_CHECK_ARG(R >= 0 && cam_loc && ray_dirs && near_far && bad_flag, "svs_sphere_intersections: null/neg");
  if (R == 0) return SVS_OK;
  sphere_kernel<<<(unsigned)cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(cam_loc, ray_dirs, R, radius,
                                                                          near_far, bad_flag);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_ray_points(const float* cam_loc, const float* ray_dirs, const float* z, int64_t R,
                              int32_t S, int32_t ldz, float* points, void* stream) {
  SVS_CHECK_ARG(R >= 0 && S > 0 && ldz >= S && cam_loc && ray_dirs && z && points, "svs_ray_points: bad args");
  if (R == 0) return SVS_OK;
  points_kernel<<<(unsigned)cdiv(R * S, 256), 256, 0, (cudaStream_t)stream>>>(cam_loc, ray_dirs, z, R, S, ldz,
                                                                              points);
  SVS_LAUNCH_OK();
  return SVS_OK;
}

extern "C" int svs_depth2pts_outside(const float* cam_loc, const float* ray_dirs, const float* depth,
                                     int64_t R, int32_t S, float radius, float* pts, float* depth_real,
                                     void* stream) {
  SVS_CHECK_ARG(R >= 0 && S > 0 && cam_loc && ray_dirs && depth && pts && depth_real, "svs_depth2pts_outside: bad args");
  if (R == 0) return SVS_OK;
  depth2pts_kernel<<<(unsigned)cdiv(R * S, 256), 256, 0, (cudaStream_t)stream>>>(cam_loc, ray_dirs, depth, R,
                                                                                 S, radius, pts, depth_real);
  SVS_LAUNCH_OK();
  return SVS_OK;
}
