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

// Branch-free form of the same stencil for the hot forward kernels: 32-bit clamped offsets and weights that are
// exactly 0 for out-of-volume corners.  acc + v*0 leaves acc unchanged, so the result equals the predicated form
// (for finite inputs) bit for bit, with ~3x fewer instructions.
struct Stencil8 {
  int off[8];
  float w[8];
};
template <bool IS3D>
__device__ __forceinline__ void make_stencil8(float cx, float cy, float cz, int D, int H, int W, Stencil8& s) {
  const float fx = floorf(cx), fy = floorf(cy);
  const int x0 = f2i(fx), y0 = f2i(fy);
  float wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), cx), wx1 = __fsub_rn(cx, fx);
  float wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), cy), wy1 = __fsub_rn(cy, fy);
  wx0 = (unsigned)x0 < (unsigned)W ? wx0 : 0.f;
  wx1 = (unsigned)(x0 + 1) < (unsigned)W ? wx1 : 0.f;
  wy0 = (unsigned)y0 < (unsigned)H ? wy0 : 0.f;
  wy1 = (unsigned)(y0 + 1) < (unsigned)H ? wy1 : 0.f;
  const int xa = min(max(x0, 0), W - 1), xb = min(max(x0, -1) + 1, W - 1);
  const int ya = min(max(y0, 0), H - 1) * W, yb = (min(max(y0, -1) + 1, H - 1)) * W;
  const float w00 = __fmul_rn(wx0, wy0), w10 = __fmul_rn(wx1, wy0), w01 = __fmul_rn(wx0, wy1), w11 = __fmul_rn(wx1, wy1);
  if (IS3D) {
    const float fz = floorf(cz);
    const int z0 = f2i(fz);
    float wz0 = __fsub_rn(__fadd_rn(fz, 1.0f), cz), wz1 = __fsub_rn(cz, fz);
    wz0 = (unsigned)z0 < (unsigned)D ? wz0 : 0.f;
    wz1 = (unsigned)(z0 + 1) < (unsigned)D ? wz1 : 0.f;
    const int za = min(max(z0, 0), D - 1) * (H * W), zb = (min(max(z0, -1) + 1, D - 1)) * (H * W);
    s.off[0] = za + ya + xa; s.off[1] = za + ya + xb; s.off[2] = za + yb + xa; s.off[3] = za + yb + xb;
    s.off[4] = zb + ya + xa; s.off[5] = zb + ya + xb; s.off[6] = zb + yb + xa; s.off[7] = zb + yb + xb;
    s.w[0] = __fmul_rn(w00, wz0); s.w[1] = __fmul_rn(w10, wz0); s.w[2] = __fmul_rn(w01, wz0); s.w[3] = __fmul_rn(w11, wz0);
    s.w[4] = __fmul_rn(w00, wz1); s.w[5] = __fmul_rn(w10, wz1); s.w[6] = __fmul_rn(w01, wz1); s.w[7] = __fmul_rn(w11, wz1);
  } else {
    s.off[0] = ya + xa; s.off[1] = ya + xb; s.off[2] = yb + xa; s.off[3] = yb + xb;
    s.w[0] = w00; s.w[1] = w10; s.w[2] = w01; s.w[3] = w11;
  }
}
template <bool IS3D, bool READONLY>
__device__ __forceinline__ float sample8(const float* __restrict__ plane, const Stencil8& s) {
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < (IS3D ? 8 : 4); ++k) {
    const float v = READONLY ? __ldg(plane + s.off[k]) : plane[s.off[k]];
    acc = __fadd_rn(acc, __fmul_rn(v, s.w[k]));
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
