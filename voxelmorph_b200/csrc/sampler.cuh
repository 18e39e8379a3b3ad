// Trilinear / nearest gather primitives shared by the warp and VecInt kernels.
//
// Reference semantics: voxelmorph/torch/layers.py:30-48 -> F.grid_sample(align_corners=True,
// padding_mode='zeros').  Corner weights and accumulation order follow ATen's
// grid_sampler_3d (weights are products (x1-x)(y1-y)(z1-z) ...; corners outside the volume
// contribute nothing).  Products and sums use explicit _rn intrinsics so the result does not
// depend on FMA contraction: the linear path reproduces the torch CPU reference to the bit on
// every input we have tried, and the nearest path is bit-exact by construction.
#pragma once
#include "common.cuh"

namespace vxm {

struct Vol {
  int D, H, W;
  size_t HW, DHW;
};
__host__ __device__ inline Vol make_vol(int D, int H, int W) {
  Vol v;
  v.D = D; v.H = H; v.W = W;
  v.HW = (size_t)H * W;
  v.DHW = (size_t)D * H * W;
  return v;
}

// Corner stencil of one sampling position: base index (may be out of range), per-corner
// weights and validity mask, in ATen corner order: bit2 = z+1, bit1 = y+1, bit0 = x+1.
struct Stencil {
  int x0, y0, z0;
  float wx0, wx1, wy0, wy1, wz0, wz1;
  unsigned mask;  // bit k set <=> corner k inside the volume
};

template <bool IS3D>
__device__ __forceinline__ Stencil make_stencil(float cx, float cy, float cz, const Vol& s) {
  Stencil st;
  float fx = floorf(cx), fy = floorf(cy);
  st.x0 = f2i(fx);
  st.y0 = f2i(fy);
  st.wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), cx);
  st.wx1 = __fsub_rn(cx, fx);
  st.wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), cy);
  st.wy1 = __fsub_rn(cy, fy);
  bool x0ok = (unsigned)st.x0 < (unsigned)s.W, x1ok = (unsigned)(st.x0 + 1) < (unsigned)s.W;
  bool y0ok = (unsigned)st.y0 < (unsigned)s.H, y1ok = (unsigned)(st.y0 + 1) < (unsigned)s.H;
  bool z0ok = true, z1ok = false;
  if (IS3D) {
    float fz = floorf(cz);
    st.z0 = f2i(fz);
    st.wz0 = __fsub_rn(__fadd_rn(fz, 1.0f), cz);
    st.wz1 = __fsub_rn(cz, fz);
    z0ok = (unsigned)st.z0 < (unsigned)s.D;
    z1ok = (unsigned)(st.z0 + 1) < (unsigned)s.D;
  } else {
    st.z0 = 0; st.wz0 = 1.0f; st.wz1 = 0.0f;
  }
  unsigned m = 0;
  m |= (z0ok && y0ok && x0ok) ? 1u : 0u;
  m |= (z0ok && y0ok && x1ok) ? 2u : 0u;
  m |= (z0ok && y1ok && x0ok) ? 4u : 0u;
  m |= (z0ok && y1ok && x1ok) ? 8u : 0u;
  m |= (z1ok && y0ok && x0ok) ? 16u : 0u;
  m |= (z1ok && y0ok && x1ok) ? 32u : 0u;
  m |= (z1ok && y1ok && x0ok) ? 64u : 0u;
  m |= (z1ok && y1ok && x1ok) ? 128u : 0u;
  st.mask = m;
  return st;
}

// weight of corner k: (x-term * y-term) * z-term, each product rounded (ATen order)
template <bool IS3D>
__device__ __forceinline__ float corner_weight(const Stencil& st, int k) {
  float wx = (k & 1) ? st.wx1 : st.wx0;
  float wy = (k & 2) ? st.wy1 : st.wy0;
  float w = __fmul_rn(wx, wy);
  if (IS3D) {
    float wz = (k & 4) ? st.wz1 : st.wz0;
    w = __fmul_rn(w, wz);
  }
  return w;
}

__device__ __forceinline__ ptrdiff_t corner_offset(const Stencil& st, int k, const Vol& s) {
  return ((ptrdiff_t)(st.z0 + ((k >> 2) & 1)) * s.H + (st.y0 + ((k >> 1) & 1))) * (ptrdiff_t)s.W +
         (st.x0 + (k & 1));
}

// value of one channel plane at the stencil
template <bool IS3D>
__device__ __forceinline__ float sample_linear(const float* __restrict__ plane, const Stencil& st,
                                               const Vol& s) {
  constexpr int NC = IS3D ? 8 : 4;
  float acc = 0.0f;
  ptrdiff_t base = corner_offset(st, 0, s);
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    if (st.mask & (1u << k)) {
      ptrdiff_t off = base + ((k >> 2) & 1) * (ptrdiff_t)s.HW + ((k >> 1) & 1) * (ptrdiff_t)s.W + (k & 1);
      float v = __ldg(plane + off);
      acc = __fadd_rn(acc, __fmul_rn(v, corner_weight<IS3D>(st, k)));
    }
  }
  return acc;
}

// nearest index (round half to even) or -1 when outside the volume
template <bool IS3D>
__device__ __forceinline__ ptrdiff_t nearest_index(float cx, float cy, float cz, const Vol& s) {
  int x = f2i(rintf(cx)), y = f2i(rintf(cy));
  int z = IS3D ? f2i(rintf(cz)) : 0;
  bool ok = (unsigned)x < (unsigned)s.W && (unsigned)y < (unsigned)s.H && (unsigned)z < (unsigned)s.D;
  return ok ? ((ptrdiff_t)z * s.H + y) * (ptrdiff_t)s.W + x : (ptrdiff_t)-1;
}

}  // namespace vxm
