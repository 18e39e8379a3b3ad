// Evaluation helpers of the registration path ("next" row N3): Jacobian determinant of x -> x + disp(x) on the device,
// with the number of folded voxels (det J <= 0) counted in the same pass.
//
// Reference: voxelmorph/py/utils.py:473-516 (numpy, np.gradient of disp + identity grid).  np.gradient's unit-spacing
// rule is replayed: central differences (f[i+1] - f[i-1]) / 2 in the interior, one-sided f[1] - f[0] / f[n-1] - f[n-2]
// at the two ends; the determinant is expanded exactly as py/utils.py:497-508.  The field stays in the module layout
// (B, nd, D, H, W) that `VxmDense(..., registration=True)` returns, so register.py's warp never leaves the GPU to be
// checked.  HBM bound: 4*nd bytes read (neighbours come from L1/L2) + 4 bytes written per voxel.
#include "common.cuh"

namespace vxm {

struct JacGeom {
  int B, D, H, W, nd;
};

// d/d(axis) of channel plane p at index i of n (stride s), np.gradient edge_order=1, unit spacing
__device__ __forceinline__ float grad1(const float* __restrict__ p, int i, int n, size_t s) {
  if (i == 0) return __ldg(p + s) - __ldg(p);
  if (i == n - 1) return __ldg(p) - __ldg(p - s);
  return (__ldg(p + s) - __ldg(p - s)) * 0.5f;
}

__global__ void __launch_bounds__(256) jacdet_kernel(const float* __restrict__ disp, float* __restrict__ det, unsigned long long* __restrict__ folds,
                                                     JacGeom g) {
  const size_t HW = (size_t)g.H * g.W, DHW = HW * g.D, n = DHW * g.B;
  unsigned local = 0;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
    const size_t b = q / DHW, p = q - b * DHW;
    const int z = (int)(p / HW), r = (int)(p - (size_t)z * HW), y = r / g.W, x = r - y * g.W;
    const float* f = disp + b * g.nd * DHW + p;
    float d;
    if (g.nd == 3) {
      // J[a][c] = d(x_c + disp_c) / d x_a ; a, c in (z, y, x) order = the reference's (dx, dy, dz) naming of axes 0, 1, 2
      float J[3][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* pc = f + (size_t)c * DHW;
        J[0][c] = grad1(pc, z, g.D, HW) + (c == 0 ? 1.f : 0.f);
        J[1][c] = grad1(pc, y, g.H, (size_t)g.W) + (c == 1 ? 1.f : 0.f);
        J[2][c] = grad1(pc, x, g.W, 1) + (c == 2 ? 1.f : 0.f);
      }
      const float d0 = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]);
      const float d1 = J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]);
      const float d2 = J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
      d = d0 - d1 + d2;
    } else {
      const float* p0 = f;
      const float* p1 = f + DHW;
      const float a00 = grad1(p0, y, g.H, (size_t)g.W) + 1.f, a01 = grad1(p1, y, g.H, (size_t)g.W);
      const float a10 = grad1(p0, x, g.W, 1), a11 = grad1(p1, x, g.W, 1) + 1.f;
      d = a00 * a11 - a10 * a01;
    }
    if (det) det[q] = d;
    local += d <= 0.f ? 1u : 0u;
  }
  if (folds) {
    // integer count: exact and order independent
    const unsigned w = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(folds, (unsigned long long)w);
  }
}

}  // namespace vxm

using namespace vxm;

extern "C" int vxm_jacdet(const float* disp, float* det, unsigned long long* folds, int B, int D, int H, int W, int nd, void* stream) {
  VXM_REQUIRE(disp && (det || folds), "jacdet: null pointer");
  VXM_REQUIRE(nd == 2 || nd == 3, "jacdet: flow has to be 2D or 3D");
  VXM_REQUIRE(B > 0 && H > 1 && W > 1 && (nd == 2 ? D == 1 : D > 1), "jacdet: every spatial axis needs at least 2 samples (2-D fields are passed with D == 1)");
  JacGeom g{B, D, H, W, nd};
  cudaStream_t st = as_stream(stream);
  if (folds) VXM_CUDA(cudaMemsetAsync(folds, 0, sizeof(unsigned long long), st));
  const size_t n = (size_t)B * D * H * W;
  size_t blocks = (n + 255) / 256, cap = (size_t)sm_count() * 16;
  jacdet_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(disp, det, folds, g);
  return check_launch("jacdet");
}
