// ResizeTransform (reference voxelmorph/torch/layers.py:85-97): align_corners=True linear
// resampling of a flow field fused with its rescaling,  out = post * lerp(pre * x).
//
// Index / weight arithmetic follows ATen UpSample.h:271-296 (area_pixel_compute_scale /
// _source_index) and :442-475 (guard_index_and_lambda); the lerp is evaluated innermost axis
// first like ATen's generic N-d kernel.  The backward is a deterministic gather-form adjoint
// (the reference's autograd reaches ATen's atomicAdd-based upsample_trilinear3d_backward).
//
// Algorithmic bytes (fp32): 4*C*(V_in + V_out) forward, the same backward.
#include "common.cuh"

namespace vxm {

struct AxisMap {
  int in, out;
  float ratio;  // (in-1)/(out-1), 0 when out == 1
};

__host__ inline AxisMap make_map(int in, int out) {
  AxisMap m;
  m.in = in; m.out = out;
  m.ratio = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f;
  return m;
}

__device__ __forceinline__ void src_index(const AxisMap& m, int o, int& i0, int& i1, float& l0, float& l1) {
  if (m.in == m.out) {  // UpSample.h:452-458
    i0 = i1 = o; l0 = 1.0f; l1 = 0.0f;
    return;
  }
  float real = __fmul_rn(m.ratio, (float)o);
  i0 = min((int)real, m.in - 1);
  float lam = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.0f), 1.0f);
  i1 = i0 + (i0 < m.in - 1 ? 1 : 0);
  l1 = lam;
  l0 = __fsub_rn(1.0f, lam);
}

struct ResizeGeom {
  AxisMap mz, my, mx;
  int BC;
  float pre, post;
};

constexpr int RS_ZPB = 4;
__global__ void __launch_bounds__(256) resize_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, ResizeGeom g) {
  int ox = blockIdx.x * 32 + threadIdx.x;
  int oy = blockIdx.y * 8 + threadIdx.y;
  const int nzb = (g.mz.out + RS_ZPB - 1) / RS_ZPB;     // RS_ZPB output slices per block share the x / y index arithmetic
  const int oz0 = (blockIdx.z % nzb) * RS_ZPB, bc = blockIdx.z / nzb;
  if (ox >= g.mx.out || oy >= g.my.out) return;
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  src_index(g.my, oy, y0, y1, ly0, ly1);
  src_index(g.mx, ox, x0, x1, lx0, lx1);
  const float* xb = x + (size_t)bc * g.mz.in * g.my.in * g.mx.in;
  size_t sH = (size_t)g.mx.in, sD = (size_t)g.my.in * g.mx.in;
  float pre = g.pre;
  auto row = [&](int z, int y) {
    const float* r = xb + z * sD + y * sH;
    return __fmul_rn(__ldg(r + x0), pre) * lx0 + __fmul_rn(__ldg(r + x1), pre) * lx1;
  };
#pragma unroll
  for (int zz = 0; zz < RS_ZPB; ++zz) {
    const int oz = oz0 + zz;
    if (oz >= g.mz.out) break;
    int z0, z1;
    float lz0, lz1;
    src_index(g.mz, oz, z0, z1, lz0, lz1);
    float p0 = row(z0, y0) * ly0 + row(z0, y1) * ly1;
    float v;
    if (z1 != z0 || g.mz.in != g.mz.out) {
      float p1 = row(z1, y0) * ly0 + row(z1, y1) * ly1;
      v = p0 * lz0 + p1 * lz1;
    } else {
      v = p0;  // l0 = 1, l1 = 0
    }
    out[(((size_t)bc * g.mz.out + oz) * g.my.out + oy) * g.mx.out + ox] = v * g.post;
  }
}

// adjoint weight of output index o on input index i along one axis
__device__ __forceinline__ float adj_weight(const AxisMap& m, int i, int o) {
  int i0, i1;
  float l0, l1;
  src_index(m, o, i0, i1, l0, l1);
  float w = 0.f;
  if (i0 == i) w += l0;
  if (i1 == i) w += l1;
  return w;
}

// candidate output range [lo, hi] whose stencil can touch input index i
__device__ __forceinline__ void adj_range(const AxisMap& m, int i, int& lo, int& hi) {
  if (m.in == m.out) { lo = hi = i; return; }
  if (m.ratio <= 0.f) { lo = 0; hi = m.out - 1; return; }
  float inv = 1.0f / m.ratio;
  lo = max(0, (int)floorf((float)(i - 1) * inv) - 1);
  hi = min(m.out - 1, (int)ceilf((float)(i + 1) * inv) + 1);
}

// Table-driven adjoint: each block first tabulates, for its 32 x 8 x 1 input indices, the (output index, weight)
// pairs of the 1-D adjoint along every axis (at most ADJ_MAX per index), then every thread runs a pure
// multiply-add loop over the outer product of its three short lists.
constexpr int ADJ_MAX = 8;
struct AdjList {
  int n;
  int o[ADJ_MAX];
  float w[ADJ_MAX];
};

constexpr int ADJ_ZPB = 8;   // input slices per block: the x / y tables are built once and reused
__global__ void __launch_bounds__(256) resize_bwd_table_kernel(const float* __restrict__ gout, float* __restrict__ gx, ResizeGeom g) {
  __shared__ AdjList lx[32], ly[8], lz[ADJ_ZPB];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int ix = blockIdx.x * 32 + threadIdx.x;
  const int iy = blockIdx.y * 8 + threadIdx.y;
  const int nzb = (g.mz.in + ADJ_ZPB - 1) / ADJ_ZPB;
  const int iz0 = (blockIdx.z % nzb) * ADJ_ZPB, bc = blockIdx.z / nzb;
  if (tid < 40 + ADJ_ZPB) {
    const AxisMap& m = tid < 32 ? g.mx : (tid < 40 ? g.my : g.mz);
    const int i = tid < 32 ? blockIdx.x * 32 + tid : (tid < 40 ? blockIdx.y * 8 + (tid - 32) : iz0 + (tid - 40));
    AdjList& L = tid < 32 ? lx[tid] : (tid < 40 ? ly[tid - 32] : lz[tid - 40]);
    L.n = 0;
    if (i < m.in) {
      int lo, hi;
      adj_range(m, i, lo, hi);
      for (int o = lo; o <= hi; ++o) {
        float w = adj_weight(m, i, o);
        if (w != 0.f && L.n < ADJ_MAX) { L.o[L.n] = o; L.w[L.n] = w; ++L.n; }
      }
    }
  }
  __syncthreads();
  if (ix >= g.mx.in || iy >= g.my.in) return;
  const AdjList& X = lx[threadIdx.x];
  const AdjList& Y = ly[threadIdx.y];
  const float* gb = gout + (size_t)bc * g.mz.out * g.my.out * g.mx.out;
  for (int zz = 0; zz < ADJ_ZPB && iz0 + zz < g.mz.in; ++zz) {
    const AdjList& Z = lz[zz];
    float acc = 0.f;
    for (int a = 0; a < Z.n; ++a) {
      for (int b = 0; b < Y.n; ++b) {
        const float* r = gb + ((size_t)Z.o[a] * g.my.out + Y.o[b]) * g.mx.out;
        float racc = 0.f;
        for (int c = 0; c < X.n; ++c) racc += X.w[c] * __ldg(r + X.o[c]);
        acc += Z.w[a] * Y.w[b] * racc;
      }
    }
    gx[(((size_t)bc * g.mz.in + iz0 + zz) * g.my.in + iy) * g.mx.in + ix] = acc * (g.pre * g.post);
  }
}

__global__ void __launch_bounds__(256) resize_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gx, ResizeGeom g) {
  int ix = blockIdx.x * 32 + threadIdx.x;
  int iy = blockIdx.y * 8 + threadIdx.y;
  int iz = blockIdx.z % g.mz.in, bc = blockIdx.z / g.mz.in;
  if (ix >= g.mx.in || iy >= g.my.in) return;
  int zlo, zhi, ylo, yhi, xlo, xhi;
  adj_range(g.mz, iz, zlo, zhi);
  adj_range(g.my, iy, ylo, yhi);
  adj_range(g.mx, ix, xlo, xhi);
  const float* gb = gout + (size_t)bc * g.mz.out * g.my.out * g.mx.out;
  float acc = 0.f;
  for (int oz = zlo; oz <= zhi; ++oz) {
    float wz = adj_weight(g.mz, iz, oz);
    if (wz == 0.f) continue;
    for (int oy = ylo; oy <= yhi; ++oy) {
      float wy = adj_weight(g.my, iy, oy);
      if (wy == 0.f) continue;
      const float* r = gb + ((size_t)oz * g.my.out + oy) * g.mx.out;
      float racc = 0.f;
      for (int ox = xlo; ox <= xhi; ++ox) {
        float wx = adj_weight(g.mx, ix, ox);
        if (wx != 0.f) racc += wx * __ldg(r + ox);
      }
      acc += wz * wy * racc;
    }
  }
  gx[(((size_t)bc * g.mz.in + iz) * g.my.in + iy) * g.mx.in + ix] = acc * (g.pre * g.post);
}

}  // namespace vxm

using namespace vxm;

static int resize_check(int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo) {
  VXM_REQUIRE(B > 0 && C > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "resize: non-positive dimension");
  VXM_REQUIRE((size_t)B * C * (Di > Do ? Di : Do) <= 65535u, "resize: B*C*D exceeds the launch grid limit");
  return VXM_OK;
}

extern "C" int vxm_resize_fwd(const float* x, float* out, int B, int C, int Di, int Hi, int Wi, int Do,
                              int Ho, int Wo, float pre, float post, void* stream) {
  int rc = resize_check(B, C, Di, Hi, Wi, Do, Ho, Wo);
  if (rc) return rc;
  VXM_REQUIRE(x && out, "resize_fwd: null pointer");
  ResizeGeom g{make_map(Di, Do), make_map(Hi, Ho), make_map(Wi, Wo), B * C, pre, post};
  dim3 block(32, 8, 1), grid((Wo + 31) / 32, (Ho + 7) / 8, ((Do + RS_ZPB - 1) / RS_ZPB) * B * C);
  resize_fwd_kernel<<<grid, block, 0, as_stream(stream)>>>(x, out, g);
  return check_launch("resize_fwd");
}

extern "C" int vxm_resize_bwd(const float* grad_out, float* grad_x, int B, int C, int Di, int Hi, int Wi,
                              int Do, int Ho, int Wo, float pre, float post, void* stream) {
  int rc = resize_check(B, C, Di, Hi, Wi, Do, Ho, Wo);
  if (rc) return rc;
  VXM_REQUIRE(grad_out && grad_x, "resize_bwd: null pointer");
  ResizeGeom g{make_map(Di, Do), make_map(Hi, Ho), make_map(Wi, Wo), B * C, pre, post};
  dim3 block(32, 8, 1), grid((Wi + 31) / 32, (Hi + 7) / 8, Di * B * C);
  auto fits = [](const AxisMap& m) { return m.in == m.out || (m.ratio > 0.f && 2.0f / m.ratio + 5.0f <= (float)ADJ_MAX + 2.f); };
  if (fits(g.mz) && fits(g.my) && fits(g.mx)) {
    dim3 gridt((Wi + 31) / 32, (Hi + 7) / 8, ((Di + ADJ_ZPB - 1) / ADJ_ZPB) * B * C);
    resize_bwd_table_kernel<<<gridt, block, 0, as_stream(stream)>>>(grad_out, grad_x, g);
  }
  else resize_bwd_kernel<<<grid, block, 0, as_stream(stream)>>>(grad_out, grad_x, g);
  return check_launch("resize_bwd");
}
