// SpatialTransformer forward / backward (reference voxelmorph/torch/layers.py:30-48).
//
// One thread per output voxel, x fastest (coalesced flow reads / output writes); the 8-corner
// gather goes through the read-only path and is served by L1/L2 (neighbouring threads touch
// neighbouring source lines because registration flows are smooth).  The identity grid the
// reference materialises as a buffer (layers.py:17-28, 82.6 MB at 160x192x224) is never built:
// p is the thread's own index.
//
// Algorithmic HBM bytes per output voxel (fp32): 4*C (src) + 4*nd (flow) + 4*C (out).
#include "sampler.cuh"

namespace vxm {

constexpr int TX = 32, TY = 8;

struct WarpGeom {
  Vol src, dst;
  AxisNorm ax, ay, az;
  int B, C, nd;
};

template <bool IS3D, int MODE, int ARITH>
__global__ void __launch_bounds__(TX* TY) warp_fwd_kernel(const float* __restrict__ src,
                                                          const float* __restrict__ flow,
                                                          float* __restrict__ out, WarpGeom g) {
  int x = blockIdx.x * TX + threadIdx.x;
  int y = blockIdx.y * TY + threadIdx.y;
  int zb = blockIdx.z;
  int z = zb % g.dst.D, b = zb / g.dst.D;
  if (x >= g.dst.W || y >= g.dst.H) return;
  size_t p = ((size_t)z * g.dst.H + y) * g.dst.W + x;
  const float* fb = flow + (size_t)b * g.nd * g.dst.DHW + p;
  float cz = 0.f, cy, cx;
  if (IS3D) {
    cz = sample_coord<ARITH>((float)z, __ldg(fb), g.az);
    cy = sample_coord<ARITH>((float)y, __ldg(fb + g.dst.DHW), g.ay);
    cx = sample_coord<ARITH>((float)x, __ldg(fb + 2 * g.dst.DHW), g.ax);
  } else {
    cy = sample_coord<ARITH>((float)y, __ldg(fb), g.ay);
    cx = sample_coord<ARITH>((float)x, __ldg(fb + g.dst.DHW), g.ax);
  }
  const float* sb = src + (size_t)b * g.C * g.src.DHW;
  float* ob = out + (size_t)b * g.C * g.dst.DHW + p;
  if (MODE == VXM_MODE_NEAREST) {
    ptrdiff_t idx = nearest_index<IS3D>(cx, cy, cz, g.src);
    for (int c = 0; c < g.C; ++c) ob[(size_t)c * g.dst.DHW] = idx >= 0 ? __ldg(sb + (size_t)c * g.src.DHW + idx) : 0.0f;
  } else {
    Stencil8 st;
    make_stencil8<IS3D>(cx, cy, cz, g.src.D, g.src.H, g.src.W, st);
    for (int c = 0; c < g.C; ++c)
      ob[(size_t)c * g.dst.DHW] = sample8<IS3D, true>(sb + (size_t)c * g.src.DHW, st);
  }
}

template <bool IS3D, int MODE, int ARITH>
__global__ void __launch_bounds__(TX* TY) warp_bwd_kernel(const float* __restrict__ gout,
                                                          const float* __restrict__ src,
                                                          const float* __restrict__ flow,
                                                          float* __restrict__ gsrc,
                                                          float* __restrict__ gflow, WarpGeom g,
                                                          float mz, float my, float mx) {
  int x = blockIdx.x * TX + threadIdx.x;
  int y = blockIdx.y * TY + threadIdx.y;
  int zb = blockIdx.z;
  int z = zb % g.dst.D, b = zb / g.dst.D;
  if (x >= g.dst.W || y >= g.dst.H) return;
  size_t p = ((size_t)z * g.dst.H + y) * g.dst.W + x;
  const float* fb = flow + (size_t)b * g.nd * g.dst.DHW + p;
  float cz = 0.f, cy, cx;
  if (IS3D) {
    cz = sample_coord<ARITH>((float)z, __ldg(fb), g.az);
    cy = sample_coord<ARITH>((float)y, __ldg(fb + g.dst.DHW), g.ay);
    cx = sample_coord<ARITH>((float)x, __ldg(fb + 2 * g.dst.DHW), g.ax);
  } else {
    cy = sample_coord<ARITH>((float)y, __ldg(fb), g.ay);
    cx = sample_coord<ARITH>((float)x, __ldg(fb + g.dst.DHW), g.ax);
  }
  const float* sb = src + (size_t)b * g.C * g.src.DHW;
  float* gsb = gsrc ? gsrc + (size_t)b * g.C * g.src.DHW : nullptr;
  const float* gob = gout + (size_t)b * g.C * g.dst.DHW + p;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  if (MODE == VXM_MODE_NEAREST) {
    ptrdiff_t idx = nearest_index<IS3D>(cx, cy, cz, g.src);
    if (gsb && idx >= 0)
      for (int c = 0; c < g.C; ++c) atomicAdd(gsb + (size_t)c * g.src.DHW + idx, __ldg(gob + (size_t)c * g.dst.DHW));
  } else {
    constexpr int NC = IS3D ? 8 : 4;
    Stencil st = make_stencil<IS3D>(cx, cy, cz, g.src);
    ptrdiff_t base = corner_offset(st, 0, g.src);
    for (int c = 0; c < g.C; ++c) {
      float go = __ldg(gob + (size_t)c * g.dst.DHW);
      const float* plane = sb + (size_t)c * g.src.DHW;
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        if (st.mask & (1u << k)) {
          ptrdiff_t off = base + ((k >> 2) & 1) * (ptrdiff_t)g.src.HW + ((k >> 1) & 1) * (ptrdiff_t)g.src.W + (k & 1);
          float wx = (k & 1) ? st.wx1 : st.wx0, wy = (k & 2) ? st.wy1 : st.wy0;
          float wz = IS3D ? ((k & 4) ? st.wz1 : st.wz0) : 1.0f;
          if (gsb) atomicAdd(gsb + (size_t)c * g.src.DHW + off, corner_weight<IS3D>(st, k) * go);
          if (gflow) {
            float v = __ldg(plane + off) * go;
            gx += ((k & 1) ? v : -v) * wy * wz;
            gy += ((k & 2) ? v : -v) * wx * wz;
            if (IS3D) gz += ((k & 4) ? v : -v) * wx * wy;
          }
        }
      }
    }
  }
  if (gflow) {
    float* gf = gflow + (size_t)b * g.nd * g.dst.DHW + p;
    if (IS3D) {
      gf[0] = gz * mz;
      gf[g.dst.DHW] = gy * my;
      gf[2 * g.dst.DHW] = gx * mx;
    } else {
      gf[0] = gy * my;
      gf[g.dst.DHW] = gx * mx;
    }
  }
}

static int check_geom(int B, int C, int Ds, int Hs, int Ws, int D, int H, int W, int nd) {
  VXM_REQUIRE(nd == 2 || nd == 3, "warp: nd must be 2 or 3 (reference layers.py:41-46), got %d", nd);
  VXM_REQUIRE(B > 0 && C > 0 && Ds > 0 && Hs > 0 && Ws > 0 && D > 0 && H > 0 && W > 0, "warp: non-positive dimension");
  VXM_REQUIRE(nd == 3 || (D == 1 && Ds == 1), "warp: a 2-D problem must be passed with D == 1");
  VXM_REQUIRE((size_t)D * B <= 65535u, "warp: D*B exceeds the launch grid limit (65535)");
  VXM_REQUIRE((size_t)Ds * Hs * Ws < (1u << 31) && (size_t)D * H * W < (1u << 31), "warp: volume exceeds 2^31 voxels");
  return VXM_OK;
}

static WarpGeom make_geom(int B, int C, int Ds, int Hs, int Ws, int D, int H, int W, int nd) {
  WarpGeom g;
  g.src = make_vol(Ds, Hs, Ws);
  g.dst = make_vol(D, H, W);
  g.ax = make_axis(W, Ws);
  g.ay = make_axis(H, Hs);
  g.az = make_axis(D, Ds);
  g.B = B; g.C = C; g.nd = nd;
  return g;
}

#define WARP_DISPATCH(KERNEL, ...)                                                                  \
  do {                                                                                              \
    bool is3d = (nd == 3);                                                                          \
    bool near = (mode == VXM_MODE_NEAREST);                                                         \
    bool rec = (arith == VXM_ARITH_RECIPROCAL);                                                     \
    if (is3d && !near && !rec) KERNEL<true, 0, 0><<<grid, block, 0, st>>>(__VA_ARGS__);             \
    else if (is3d && !near && rec) KERNEL<true, 0, 1><<<grid, block, 0, st>>>(__VA_ARGS__);         \
    else if (is3d && near && !rec) KERNEL<true, 1, 0><<<grid, block, 0, st>>>(__VA_ARGS__);         \
    else if (is3d && near && rec) KERNEL<true, 1, 1><<<grid, block, 0, st>>>(__VA_ARGS__);          \
    else if (!is3d && !near && !rec) KERNEL<false, 0, 0><<<grid, block, 0, st>>>(__VA_ARGS__);      \
    else if (!is3d && !near && rec) KERNEL<false, 0, 1><<<grid, block, 0, st>>>(__VA_ARGS__);       \
    else if (!is3d && near && !rec) KERNEL<false, 1, 0><<<grid, block, 0, st>>>(__VA_ARGS__);       \
    else KERNEL<false, 1, 1><<<grid, block, 0, st>>>(__VA_ARGS__);                                  \
  } while (0)

}  // namespace vxm

using namespace vxm;

extern "C" int vxm_warp_fwd(const float* src, const float* flow, float* out, int B, int C, int Ds,
                            int Hs, int Ws, int D, int H, int W, int nd, int mode, int arith,
                            void* stream) {
  int rc = check_geom(B, C, Ds, Hs, Ws, D, H, W, nd);
  if (rc) return rc;
  VXM_REQUIRE(src && flow && out, "warp_fwd: null pointer");
  VXM_REQUIRE(mode == VXM_MODE_LINEAR || mode == VXM_MODE_NEAREST, "warp_fwd: bad mode %d", mode);
  WarpGeom g = make_geom(B, C, Ds, Hs, Ws, D, H, W, nd);
  dim3 block(TX, TY, 1), grid((W + TX - 1) / TX, (H + TY - 1) / TY, D * B);
  cudaStream_t st = as_stream(stream);
  WARP_DISPATCH(warp_fwd_kernel, src, flow, out, g);
  return check_launch("warp_fwd");
}

extern "C" int vxm_warp_bwd(const float* grad_out, const float* src, const float* flow,
                            float* grad_src, float* grad_flow, int B, int C, int Ds, int Hs, int Ws,
                            int D, int H, int W, int nd, int mode, int arith, void* stream) {
  int rc = check_geom(B, C, Ds, Hs, Ws, D, H, W, nd);
  if (rc) return rc;
  VXM_REQUIRE(grad_out && src && flow, "warp_bwd: null pointer");
  WarpGeom g = make_geom(B, C, Ds, Hs, Ws, D, H, W, nd);
  // d coord / d flow = ((Ssrc-1)/2) * 2 / (Sflow-1)   (GridSampler.h:45-47 and layers.py:37)
  float mx = (g.ax.src_sm1 * 0.5f) * 2.0f / g.ax.sm1;
  float my = (g.ay.src_sm1 * 0.5f) * 2.0f / g.ay.sm1;
  float mz = nd == 3 ? (g.az.src_sm1 * 0.5f) * 2.0f / g.az.sm1 : 0.f;
  dim3 block(TX, TY, 1), grid((W + TX - 1) / TX, (H + TY - 1) / TY, D * B);
  cudaStream_t st = as_stream(stream);
  WARP_DISPATCH(warp_bwd_kernel, grad_out, src, flow, grad_src, grad_flow, g, mz, my, mx);
  return check_launch("warp_bwd");
}
