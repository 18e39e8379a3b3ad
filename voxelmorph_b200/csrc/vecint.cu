// VecInt — scaling and squaring (reference voxelmorph/torch/layers.py:51-68).
//
// The reference issues ~7 launches per squaring step (add grid, normalise x3, permute, index,
// grid_sample, add).  Here ALL nsteps squarings run in ONE cooperative launch: a persistent grid
// walks the field, `grid.sync()` separates the steps, and the (B,nd,D,H,W) field (10.3 MB at
// 80x96x112) stays L2-resident between steps.  The backward runs the reversed chain in one
// cooperative launch as well.
//
// Algorithmic bytes (fp32) per voxel per step: forward 4*nd read + 4*nd write;
// backward 4*nd (v_k) + 4*nd (g_{k+1}) read + 4*nd (g_k) write.
#include <cooperative_groups.h>

#include "sampler.cuh"

namespace cg = cooperative_groups;

namespace vxm {

struct VecGeom {
  Vol vol;
  AxisNorm ax, ay, az;
  int B, nd;
  size_t N;  // B * nd * DHW
};

template <bool IS3D, int ARITH>
__device__ __forceinline__ void field_coords(const float* __restrict__ fb, size_t p, int x, int y, int z,
                                             const VecGeom& g, float fv[3], float& cx, float& cy, float& cz) {
  // fv[i]: the field's own value at p (channel order: D,H,W for 3-D; H,W for 2-D)
  if (IS3D) {
    fv[0] = fb[p]; fv[1] = fb[p + g.vol.DHW]; fv[2] = fb[p + 2 * g.vol.DHW];
    cz = sample_coord<ARITH>((float)z, fv[0], g.az);
    cy = sample_coord<ARITH>((float)y, fv[1], g.ay);
    cx = sample_coord<ARITH>((float)x, fv[2], g.ax);
  } else {
    fv[0] = fb[p]; fv[1] = fb[p + g.vol.DHW]; fv[2] = 0.f;
    cz = 0.f;
    cy = sample_coord<ARITH>((float)y, fv[0], g.ay);
    cx = sample_coord<ARITH>((float)x, fv[1], g.ax);
  }
}

// step == -1: scaling pass; 0 <= step < nsteps: one squaring
template <bool IS3D, int ARITH>
__global__ void __launch_bounds__(256) vecint_fwd_kernel(const float* __restrict__ vel, float* out,
                                                         float* states, float* work, VecGeom g,
                                                         int nsteps, int step_begin, int step_end,
                                                         float scale, int coop) {
  cg::grid_group grid = cg::this_grid();
  const size_t tid0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nvox = (size_t)g.B * g.vol.DHW;
  auto buf = [&](int k) -> float* {
    if (k == nsteps) return out;
    if (states) return states + (size_t)k * g.N;
    return ((nsteps - k) & 1) ? work : out;
  };
  for (int s = step_begin; s < step_end; ++s) {
    if (s < 0) {
      float* dst = buf(0);
      for (size_t i = tid0; i < g.N; i += stride) dst[i] = __fmul_rn(vel[i], scale);
    } else {
      const float* cur = buf(s);
      float* nxt = buf(s + 1);
      for (size_t q = tid0; q < nvox; q += stride) {
        int b = (int)(q / g.vol.DHW);
        size_t p = q - (size_t)b * g.vol.DHW;
        int z = (int)(p / g.vol.HW);
        int r = (int)(p - (size_t)z * g.vol.HW);
        int y = r / g.vol.W, x = r - y * g.vol.W;
        const float* fb = cur + (size_t)b * g.nd * g.vol.DHW;
        float fv[3], cx, cy, cz;
        field_coords<IS3D, ARITH>(fb, p, x, y, z, g, fv, cx, cy, cz);
        Stencil8 st;
        make_stencil8<IS3D>(cx, cy, cz, g.vol.D, g.vol.H, g.vol.W, st);
        float* ob = nxt + (size_t)b * g.nd * g.vol.DHW + p;
#pragma unroll
        for (int c = 0; c < (IS3D ? 3 : 2); ++c) {
          // plain (coherent) loads: `cur` was written earlier in this same launch
          ob[(size_t)c * g.vol.DHW] = __fadd_rn(fv[c], sample8<IS3D, false>(fb + (size_t)c * g.vol.DHW, st));
        }
      }
    }
    if (coop && s + 1 < step_end) grid.sync();
  }
}

// phase 0: G_cur = g_next + (d warp / d flow)^T g_next     (gather, plain stores)
// phase 1: G_cur += (d warp / d src)^T g_next               (scatter, fp32 atomics)
// phase 2 (after step 0): grad_vel = G_0 * scale
template <bool IS3D, int ARITH>
__global__ void __launch_bounds__(256) vecint_bwd_kernel(const float* __restrict__ gout,
                                                         const float* __restrict__ states,
                                                         float* grad_vel, float* work, VecGeom g,
                                                         int nsteps, int item_begin, int item_end,
                                                         float scale, float mz, float my, float mx, int coop) {
  cg::grid_group grid = cg::this_grid();
  const size_t tid0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nvox = (size_t)g.B * g.vol.DHW;
  constexpr int NCH = IS3D ? 3 : 2;
  constexpr int NC = IS3D ? 8 : 4;
  // work items: for k = nsteps-1 .. 0: (phase 0, phase 1); then the final scaling pass
  // item index it = 2*(nsteps-1-k) + phase, final = 2*nsteps
  for (int it = item_begin; it < item_end; ++it) {
    if (it == 2 * nsteps) {
      const float* G0 = nsteps == 0 ? gout : (((nsteps - 1) & 1) ? work + g.N : work);
      for (size_t i = tid0; i < g.N; i += stride) grad_vel[i] = G0[i] * scale;
    } else {
      int j = it >> 1, phase = it & 1;     // j-th processed step, k = nsteps-1-j
      int k = nsteps - 1 - j;
      const float* gn = (j == 0) ? gout : (((j - 1) & 1) ? work + g.N : work);
      float* gc = (j & 1) ? work + g.N : work;
      const float* v = states + (size_t)k * g.N;
      for (size_t q = tid0; q < nvox; q += stride) {
        int b = (int)(q / g.vol.DHW);
        size_t p = q - (size_t)b * g.vol.DHW;
        int z = (int)(p / g.vol.HW);
        int r = (int)(p - (size_t)z * g.vol.HW);
        int y = r / g.vol.W, x = r - y * g.vol.W;
        const float* fb = v + (size_t)b * g.nd * g.vol.DHW;
        float fv[3], cx, cy, cz;
        field_coords<IS3D, ARITH>(fb, p, x, y, z, g, fv, cx, cy, cz);
        Stencil st = make_stencil<IS3D>(cx, cy, cz, g.vol);
        ptrdiff_t base = corner_offset(st, 0, g.vol);
        const float* gnb = gn + (size_t)b * g.nd * g.vol.DHW + p;
        float go[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) go[c] = gnb[(size_t)c * g.vol.DHW];
        if (phase == 0) {
          float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
          for (int kk = 0; kk < NC; ++kk) {
            if (st.mask & (1u << kk)) {
              ptrdiff_t off = base + ((kk >> 2) & 1) * (ptrdiff_t)g.vol.HW + ((kk >> 1) & 1) * (ptrdiff_t)g.vol.W + (kk & 1);
              float wx = (kk & 1) ? st.wx1 : st.wx0, wy = (kk & 2) ? st.wy1 : st.wy0;
              float wz = IS3D ? ((kk & 4) ? st.wz1 : st.wz0) : 1.0f;
              float vg = 0.f;
#pragma unroll
              for (int c = 0; c < NCH; ++c) vg += __ldg(fb + (size_t)c * g.vol.DHW + off) * go[c];
              gx += ((kk & 1) ? vg : -vg) * wy * wz;
              gy += ((kk & 2) ? vg : -vg) * wx * wz;
              if (IS3D) gz += ((kk & 4) ? vg : -vg) * wx * wy;
            }
          }
          float* gcb = gc + (size_t)b * g.nd * g.vol.DHW + p;
          if (IS3D) {
            gcb[0] = go[0] + gz * mz;
            gcb[g.vol.DHW] = go[1] + gy * my;
            gcb[2 * g.vol.DHW] = go[2] + gx * mx;
          } else {
            gcb[0] = go[0] + gy * my;
            gcb[g.vol.DHW] = go[1] + gx * mx;
          }
        } else {
          float* gcb = gc + (size_t)b * g.nd * g.vol.DHW;
#pragma unroll
          for (int kk = 0; kk < NC; ++kk) {
            if (st.mask & (1u << kk)) {
              ptrdiff_t off = base + ((kk >> 2) & 1) * (ptrdiff_t)g.vol.HW + ((kk >> 1) & 1) * (ptrdiff_t)g.vol.W + (kk & 1);
              float w = corner_weight<IS3D>(st, kk);
#pragma unroll
              for (int c = 0; c < NCH; ++c) atomicAdd(gcb + (size_t)c * g.vol.DHW + off, w * go[c]);
            }
          }
        }
      }
    }
    if (coop && it + 1 < item_end) grid.sync();
  }
}

static VecGeom make_vgeom(int B, int D, int H, int W, int nd) {
  VecGeom g;
  g.vol = make_vol(D, H, W);
  g.ax = make_axis(W, W);
  g.ay = make_axis(H, H);
  g.az = make_axis(D, D);
  g.B = B; g.nd = nd;
  g.N = (size_t)B * nd * g.vol.DHW;
  return g;
}

template <typename K>
static int coop_grid(K kernel, int threads, size_t work_items, int* grid_out) {
  int dev = 0, nsm = 0, coop = 0, per_sm = 0;
  VXM_CUDA(cudaGetDevice(&dev));
  VXM_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  VXM_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  VXM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
  if (!coop || per_sm < 1) {
    set_error("vecint: cooperative launch unavailable on this device");
    return VXM_ERR_UNSUPPORTED;
  }
  size_t need = (work_items + threads - 1) / threads;
  size_t cap = (size_t)nsm * per_sm;
  *grid_out = (int)(need < cap ? (need ? need : 1) : cap);
  return VXM_OK;
}

}  // namespace vxm

using namespace vxm;

extern "C" size_t vxm_vecint_workspace_bytes(int B, int D, int H, int W, int nd, int nsteps) {
  (void)nsteps;
  return (size_t)B * nd * D * H * W * sizeof(float);
}

template <bool IS3D, int ARITH>
static int vecint_fwd_launch(const float* vel, float* out, float* states, float* work, VecGeom g,
                             int nsteps, float scale, cudaStream_t st) {
  auto kern = vecint_fwd_kernel<IS3D, ARITH>;
  int grid = 0;
  int rc = coop_grid(kern, 256, (size_t)g.B * g.vol.DHW, &grid);
  if (rc) return rc;
  int sb = -1, se = nsteps, coop = 1;
  void* args[] = {(void*)&vel, (void*)&out, (void*)&states, (void*)&work, (void*)&g,
                  (void*)&nsteps, (void*)&sb, (void*)&se, (void*)&scale, (void*)&coop};
  VXM_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(256), args, 0, st));
  return check_launch("vecint_fwd");
}

extern "C" int vxm_vecint_fwd(const float* vel, float* out, float* states, void* work, int B, int D,
                              int H, int W, int nd, int nsteps, int arith, void* stream) {
  VXM_REQUIRE(nd == 2 || nd == 3, "vecint: nd must be 2 or 3");
  VXM_REQUIRE(nsteps >= 0 && nsteps < 31, "vecint: nsteps should be >= 0, found: %d", nsteps);  // layers.py:59
  VXM_REQUIRE(vel && out, "vecint_fwd: null pointer");
  VXM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "vecint: non-positive dimension");
  VXM_REQUIRE(nd == 3 || D == 1, "vecint: a 2-D problem must be passed with D == 1");
  VXM_REQUIRE(states || work || nsteps == 0, "vecint_fwd: need `states` or `work`");
  VecGeom g = make_vgeom(B, D, H, W, nd);
  float scale = 1.0f / (float)(1u << nsteps);
  cudaStream_t st = as_stream(stream);
  bool rec = arith == VXM_ARITH_RECIPROCAL;
  if (nd == 3) return rec ? vecint_fwd_launch<true, 1>(vel, out, states, (float*)work, g, nsteps, scale, st)
                          : vecint_fwd_launch<true, 0>(vel, out, states, (float*)work, g, nsteps, scale, st);
  return rec ? vecint_fwd_launch<false, 1>(vel, out, states, (float*)work, g, nsteps, scale, st)
             : vecint_fwd_launch<false, 0>(vel, out, states, (float*)work, g, nsteps, scale, st);
}

template <bool IS3D, int ARITH>
static int vecint_bwd_launch(const float* gout, const float* states, float* grad_vel, float* work,
                             VecGeom g, int nsteps, float scale, cudaStream_t st) {
  auto kern = vecint_bwd_kernel<IS3D, ARITH>;
  int grid = 0;
  int rc = coop_grid(kern, 256, (size_t)g.B * g.vol.DHW, &grid);
  if (rc) return rc;
  float mx = (g.ax.src_sm1 * 0.5f) * 2.0f / g.ax.sm1;
  float my = (g.ay.src_sm1 * 0.5f) * 2.0f / g.ay.sm1;
  float mz = IS3D ? (g.az.src_sm1 * 0.5f) * 2.0f / g.az.sm1 : 0.f;
  int ib = 0, ie = 2 * nsteps + 1, coop = 1;
  void* args[] = {(void*)&gout, (void*)&states, (void*)&grad_vel, (void*)&work, (void*)&g, (void*)&nsteps,
                  (void*)&ib, (void*)&ie, (void*)&scale, (void*)&mz, (void*)&my, (void*)&mx, (void*)&coop};
  VXM_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(256), args, 0, st));
  return check_launch("vecint_bwd");
}

extern "C" int vxm_vecint_bwd(const float* grad_out, const float* states, float* grad_vel, void* work,
                              int B, int D, int H, int W, int nd, int nsteps, int arith, void* stream) {
  VXM_REQUIRE(nd == 2 || nd == 3, "vecint: nd must be 2 or 3");
  VXM_REQUIRE(nsteps >= 0 && nsteps < 31, "vecint: nsteps should be >= 0, found: %d", nsteps);
  VXM_REQUIRE(grad_out && grad_vel, "vecint_bwd: null pointer");
  VXM_REQUIRE(nsteps == 0 || (states && work), "vecint_bwd: need `states` and `work`");
  VXM_REQUIRE(nd == 3 || D == 1, "vecint: a 2-D problem must be passed with D == 1");
  VecGeom g = make_vgeom(B, D, H, W, nd);
  float scale = 1.0f / (float)(1u << nsteps);
  cudaStream_t st = as_stream(stream);
  bool rec = arith == VXM_ARITH_RECIPROCAL;
  if (nd == 3) return rec ? vecint_bwd_launch<true, 1>(grad_out, states, grad_vel, (float*)work, g, nsteps, scale, st)
                          : vecint_bwd_launch<true, 0>(grad_out, states, grad_vel, (float*)work, g, nsteps, scale, st);
  return rec ? vecint_bwd_launch<false, 1>(grad_out, states, grad_vel, (float*)work, g, nsteps, scale, st)
             : vecint_bwd_launch<false, 0>(grad_out, states, grad_vel, (float*)work, g, nsteps, scale, st);
}
