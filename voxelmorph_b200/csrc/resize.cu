// ResizeTransform (reference voxelmorph/torch/layers.py:85-97): align_corners=True linear
// resampling of a flow field fused with its rescaling,  out = post * lerp(pre * x).
//
// Index / weight arithmetic follows ATen UpSample.h:271-296 (area_pixel_compute_scale /
// _source_index) and :442-475 (guard_index_and_lambda).  The backward is a deterministic gather-form
// adjoint (the reference's autograd reaches ATen's atomicAdd-based upsample_trilinear3d_backward).
//
// Both directions are "column marching" kernels: a thread owns one (x, y) position, walks a chunk of z
// and carries every channel of the field (the index / weight arithmetic of x and y is done once per thread,
// of z once per slice, and is shared by the channels):
//   forward  — the x/y-interpolated values of the two source planes a slice needs stay in registers; when
//              upsampling, consecutive output slices share a source plane, so a new slice costs ~one plane
//              (4 loads per channel) instead of two;
//   backward — the thread walks the OUTPUT-gradient planes that touch its z chunk, reduces each over its
//              short x / y adjoint lists (register resident, contiguous ranges) and adds the result to the
//              (at most two) input slices the plane interpolates from, held in two running accumulators.
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
  int BC, zchunk, nzc;
  float scale;   // pre * post
};

constexpr int RS_ZCHUNK = 16;

// x/y-interpolated value of one source plane for NC channels
template <int NC>
__device__ __forceinline__ void plane_xy(const float* __restrict__ p, size_t cstride, int o00, int o01, int o10, int o11,
                                         float lx0, float lx1, float ly0, float ly1, float (&v)[NC]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float* q = p + (size_t)c * cstride;
    const float r0 = __ldg(q + o00) * lx0 + __ldg(q + o01) * lx1;
    const float r1 = __ldg(q + o10) * lx0 + __ldg(q + o11) * lx1;
    v[c] = r0 * ly0 + r1 * ly1;
  }
}

template <int NC>
__global__ void __launch_bounds__(256) resize_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, ResizeGeom g) {
  const int ox = blockIdx.x * 32 + threadIdx.x;
  const int oy = blockIdx.y * 8 + threadIdx.y;
  const int zc = blockIdx.z % g.nzc, bc0 = (blockIdx.z / g.nzc) * NC;
  if (ox >= g.mx.out || oy >= g.my.out) return;
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  src_index(g.my, oy, y0, y1, ly0, ly1);
  src_index(g.mx, ox, x0, x1, lx0, lx1);
  const size_t sD = (size_t)g.my.in * g.mx.in, cin = sD * g.mz.in;
  const size_t oHW = (size_t)g.my.out * g.mx.out, cout = oHW * g.mz.out;
  const float* xb = x + (size_t)bc0 * cin;
  float* ob = out + (size_t)bc0 * cout + (size_t)oy * g.mx.out + ox;
  const int o00 = y0 * g.mx.in + x0, o01 = y0 * g.mx.in + x1, o10 = y1 * g.mx.in + x0, o11 = y1 * g.mx.in + x1;
  const int oz_begin = zc * g.zchunk, oz_end = min(oz_begin + g.zchunk, g.mz.out);
  int cz0 = -2;          // source plane held in P0 (P1 holds cz0 + 1 when cz1 says so)
  int cz1 = -2;
  float P0[NC], P1[NC];
  for (int oz = oz_begin; oz < oz_end; ++oz) {
    int z0, z1;
    float lz0, lz1;
    src_index(g.mz, oz, z0, z1, lz0, lz1);
    if (z0 != cz0) {
      if (z0 == cz1) {
#pragma unroll
        for (int c = 0; c < NC; ++c) P0[c] = P1[c];
      } else {
        plane_xy<NC>(xb + (size_t)z0 * sD, cin, o00, o01, o10, o11, lx0, lx1, ly0, ly1, P0);
      }
      cz0 = z0;
      cz1 = -2;
    }
    if (z1 != z0 && z1 != cz1) {
      plane_xy<NC>(xb + (size_t)z1 * sD, cin, o00, o01, o10, o11, lx0, lx1, ly0, ly1, P1);
      cz1 = z1;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float v = (z1 != z0) ? P0[c] * lz0 + P1[c] * lz1 : P0[c];
      ob[(size_t)c * cout + (size_t)oz * oHW] = v * g.scale;
    }
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

// The outputs whose stencil touches input index i form a contiguous range: first output + up to ADJ_L weights.
constexpr int ADJ_L = 6;
struct AdjList {
  int lo, n;
  float w[ADJ_L];
};
__device__ __forceinline__ void make_adj(const AxisMap& m, int i, AdjList& L) {
  int lo, hi;
  adj_range(m, i, lo, hi);
  L.lo = lo; L.n = 0;
#pragma unroll
  for (int k = 0; k < ADJ_L; ++k) L.w[k] = 0.f;
  bool started = false;
  for (int o = lo; o <= hi; ++o) {
    const float w = adj_weight(m, i, o);
    if (!started) {
      if (w == 0.f) continue;
      started = true;
      L.lo = o;
    }
    const int k = o - L.lo;
    if (k < ADJ_L) {
      // static indexing keeps the list in registers
#pragma unroll
      for (int q = 0; q < ADJ_L; ++q) if (q == k) L.w[q] = w;
      if (w != 0.f) L.n = k + 1;
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(256) resize_bwd_march_kernel(const float* __restrict__ gout, float* __restrict__ gx, ResizeGeom g) {
  const int ix = blockIdx.x * 32 + threadIdx.x;
  const int iy = blockIdx.y * 8 + threadIdx.y;
  const int zc = blockIdx.z % g.nzc, bc0 = (blockIdx.z / g.nzc) * NC;
  if (ix >= g.mx.in || iy >= g.my.in) return;
  AdjList X, Y;
  make_adj(g.mx, ix, X);
  make_adj(g.my, iy, Y);
  const size_t oHW = (size_t)g.my.out * g.mx.out, cout = oHW * g.mz.out;
  const size_t iHW = (size_t)g.my.in * g.mx.in, cin = iHW * g.mz.in;
  const float* gb = gout + (size_t)bc0 * cout + (size_t)Y.lo * g.mx.out + X.lo;
  float* ob = gx + (size_t)bc0 * cin + (size_t)iy * g.mx.in + ix;
  const int iz_begin = zc * g.zchunk, iz_end = min(iz_begin + g.zchunk, g.mz.in);
  int oz_lo, oz_hi, t;
  adj_range(g.mz, iz_begin, oz_lo, t);
  adj_range(g.mz, iz_end - 1, t, oz_hi);
  float A0[NC], A1[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) A0[c] = A1[c] = 0.f;
  int zcur = iz_begin;           // A0 accumulates input slice zcur, A1 slice zcur + 1
  auto flush = [&]() {
    if (zcur < iz_end) {
#pragma unroll
      for (int c = 0; c < NC; ++c) ob[(size_t)c * cin + (size_t)zcur * iHW] = A0[c] * g.scale;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) { A0[c] = A1[c]; A1[c] = 0.f; }
    ++zcur;
  };
  for (int oz = oz_lo; oz <= oz_hi; ++oz) {
    int z0, z1;
    float lz0, lz1;
    src_index(g.mz, oz, z0, z1, lz0, lz1);
    if (z1 < iz_begin || z0 >= iz_end) continue;
    while (zcur < z0 && zcur < iz_end) flush();
    // x / y reduction of this output plane over the adjoint lists
    float R[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) R[c] = 0.f;
    const float* pl = gb + (size_t)oz * oHW;
#pragma unroll
    for (int b = 0; b < ADJ_L; ++b) {
      if (b < Y.n) {
        float racc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) racc[c] = 0.f;
#pragma unroll
        for (int a = 0; a < ADJ_L; ++a) {
          if (a < X.n) {
#pragma unroll
            for (int c = 0; c < NC; ++c) racc[c] += X.w[a] * __ldg(pl + (size_t)c * cout + b * g.mx.out + a);
          }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) R[c] += Y.w[b] * racc[c];
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (z0 == zcur) A0[c] += lz0 * R[c];
      if (z1 == zcur) A0[c] += lz1 * R[c];
      else if (z1 == zcur + 1) A1[c] += lz1 * R[c];
    }
  }
  while (zcur < iz_end) flush();
}

// Upsampling adjoint (the backward of `fullsize`, reading the FULL-resolution gradient), "shared-memory column marching":
// the plain marching kernel's lanes read the gradient with stride ~2 and ~25 loads per plane and channel (L1-wavefront
// bound).  Here a CTA owns a 32 x 8 (x, y) tile of the INPUT grid and a chunk of its z range, and walks the output planes
// that touch the chunk.  Per plane the rows the tile touches (<= 22 rows of <= 72 columns per channel) are fetched with
// coalesced row loads — prefetched into registers one plane ahead — and staged in shared memory; the x adjoint (lane =
// input column, list in registers) and the y adjoint (warp = input row) are applied there, and the z adjoint runs in two
// register accumulators exactly as in the marching kernel.  Same lists, same summation order per axis: deterministic.
constexpr int RT_X = 32, RT_Y = 8, RT_MX = 72, RT_MY = 22;
constexpr int RC_ZCHUNK = 8;

__device__ __forceinline__ float adj_dot(const AdjList& L, const float* __restrict__ p, int stride) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < ADJ_L; ++k)
    if (k < L.n) acc += L.w[k] * p[k * stride];
  return acc;
}

template <int NC>
__global__ void __launch_bounds__(256) resize_bwd_colsm_kernel(const float* __restrict__ gout, float* __restrict__ gx, ResizeGeom g) {
  constexpr int RMAX = (NC * RT_MY + 7) / 8;          // rows per warp and plane
  __shared__ float G[NC * RT_MY * RT_MX];            // staged rows of one output plane
  __shared__ float X1[NC * RT_MY * RT_X];            // after the x adjoint
  __shared__ AdjList ends[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bx = blockIdx.x * RT_X, by = blockIdx.y * RT_Y;
  const int zc = blockIdx.z % g.nzc, bc0 = (blockIdx.z / g.nzc) * NC;
  if (tid < 4) {
    const AxisMap& m = tid < 2 ? g.mx : g.my;
    const int b0 = tid < 2 ? bx : by, T = tid < 2 ? RT_X : RT_Y;
    make_adj(m, (tid & 1) ? min(b0 + T - 1, m.in - 1) : b0, ends[tid]);
  }
  __syncthreads();
  const int xlo = ends[0].lo, ylo = ends[2].lo;
  const int EX = min(ends[1].lo + ends[1].n - xlo, RT_MX), EY = min(ends[3].lo + ends[3].n - ylo, RT_MY);
  const int nrows = NC * EY;
  const int ix = bx + lane, iy = by + warp;
  AdjList X, Y;
  if (ix < g.mx.in) make_adj(g.mx, ix, X); else { X.lo = xlo; X.n = 0; }
  if (iy < g.my.in) make_adj(g.my, iy, Y); else { Y.lo = ylo; Y.n = 0; }
  const int xo = X.lo - xlo, yo = Y.lo - ylo;
  const size_t oHW = (size_t)g.my.out * g.mx.out, cout = oHW * g.mz.out;
  const size_t iHW = (size_t)g.my.in * g.mx.in, cin = iHW * g.mz.in;
  const float* gb = gout + (size_t)bc0 * cout + (size_t)ylo * g.mx.out + xlo;
  float* ob = gx + (size_t)bc0 * cin + (size_t)iy * g.mx.in + ix;
  const bool owner = ix < g.mx.in && iy < g.my.in;
  const int iz_begin = zc * g.zchunk, iz_end = min(iz_begin + g.zchunk, g.mz.in);
  int oz_lo, oz_hi, t;
  adj_range(g.mz, iz_begin, oz_lo, t);
  adj_range(g.mz, iz_end - 1, t, oz_hi);
  // row r of a plane: channel r / EY, output row ylo + r % EY; this warp stages rows warp, warp + 8, ...
  int roff[RMAX];
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    const int r = warp + 8 * j;
    const int c = r / EY, y = r - c * EY;
    roff[j] = r < nrows ? (int)((size_t)c * cout + (size_t)y * g.mx.out) : -1;     // < 2^31: checked on the host
  }
  float pre[RMAX][3];
  auto fetch = [&](int oz) {
    const float* pl = gb + (size_t)oz * oHW;
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int x = lane + 32 * q;
        pre[j][q] = (roff[j] >= 0 && x < EX) ? __ldg(pl + roff[j] + x) : 0.f;
      }
    }
  };
  float A0[NC], A1[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) A0[c] = A1[c] = 0.f;
  int zcur = iz_begin;           // A0 accumulates input slice zcur, A1 slice zcur + 1
  auto flush = [&]() {
    if (zcur < iz_end && owner) {
#pragma unroll
      for (int c = 0; c < NC; ++c) ob[(size_t)c * cin + (size_t)zcur * iHW] = A0[c] * g.scale;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) { A0[c] = A1[c]; A1[c] = 0.f; }
    ++zcur;
  };
  // the planes that touch the chunk are a contiguous range: trim both ends (block-uniform)
  while (oz_lo <= oz_hi) { int z0, z1; float a, b; src_index(g.mz, oz_lo, z0, z1, a, b); if (z1 < iz_begin) ++oz_lo; else break; }
  while (oz_hi >= oz_lo) { int z0, z1; float a, b; src_index(g.mz, oz_hi, z0, z1, a, b); if (z0 >= iz_end) --oz_hi; else break; }
  if (oz_lo <= oz_hi) fetch(oz_lo);
  for (int oz = oz_lo; oz <= oz_hi; ++oz) {
    int z0, z1;
    float lz0, lz1;
    src_index(g.mz, oz, z0, z1, lz0, lz1);
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      const int r = warp + 8 * j;
      if (r < nrows) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int x = lane + 32 * q;
          if (x < RT_MX) G[r * RT_MX + x] = pre[j][q];
        }
      }
    }
    __syncthreads();
    if (oz < oz_hi) fetch(oz + 1);                 // next plane's loads fly during this plane's arithmetic
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      const int r = warp + 8 * j;
      if (r < nrows) X1[r * RT_X + lane] = adj_dot(X, G + r * RT_MX + xo, 1);
    }
    __syncthreads();
    while (zcur < z0 && zcur < iz_end) flush();
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float R = adj_dot(Y, X1 + (c * EY + yo) * RT_X + lane, RT_X);
      if (z0 == zcur) A0[c] += lz0 * R;
      if (z1 == zcur) A0[c] += lz1 * R;
      else if (z1 == zcur + 1) A1[c] += lz1 * R;
    }
  }
  while (zcur < iz_end) flush();
}

// largest number of outputs any T-wide input tile touches along one axis (same fp32 index arithmetic as the device)
static int host_tile_extent(const AxisMap& m, int T) {
  if (m.in == m.out) return T;
  int worst = 0;
  const int ntile = (m.in + T - 1) / T;
  int o = 0;
  for (int t = 0; t < ntile; ++t) {
    const int b0 = t * T, b1 = (b0 + T - 1 < m.in - 1) ? b0 + T - 1 : m.in - 1;
    int first = -1, last = -1;
    // outputs are monotone in their source index: walk from a little before the previous tile's end
    for (int q = (o > 4 ? o - 4 : 0); q < m.out; ++q) {
      const float real = m.ratio * (float)q;
      int i0 = (int)real;
      if (i0 > m.in - 1) i0 = m.in - 1;
      const int i1 = i0 + (i0 < m.in - 1 ? 1 : 0);
      if (i1 < b0) continue;
      if (i0 > b1) break;
      if (first < 0) first = q;
      last = q;
    }
    if (first >= 0) { if (last - first + 1 > worst) worst = last - first + 1; o = last; }
  }
  return worst;
}

// generic fallback (any ratio): one thread per input voxel, weights recomputed in the loops
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
  gx[(((size_t)bc * g.mz.in + iz) * g.my.in + iy) * g.mx.in + ix] = acc * g.scale;
}

}  // namespace vxm

using namespace vxm;

static int resize_check(int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo) {
  VXM_REQUIRE(B > 0 && C > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "resize: non-positive dimension");
  VXM_REQUIRE((size_t)B * C * (Di > Do ? Di : Do) <= 65535u, "resize: B*C*D exceeds the launch grid limit");
  VXM_REQUIRE((size_t)Di * Hi * Wi < (1u << 31) && (size_t)Do * Ho * Wo < (1u << 31), "resize: volume exceeds 2^31 voxels");
  return VXM_OK;
}

static int channels_per_thread(int BC) { return BC % 3 == 0 ? 3 : (BC % 2 == 0 ? 2 : 1); }

extern "C" int vxm_resize_fwd(const float* x, float* out, int B, int C, int Di, int Hi, int Wi, int Do,
                              int Ho, int Wo, float pre, float post, void* stream) {
  int rc = resize_check(B, C, Di, Hi, Wi, Do, Ho, Wo);
  if (rc) return rc;
  VXM_REQUIRE(x && out, "resize_fwd: null pointer");
  ResizeGeom g{make_map(Di, Do), make_map(Hi, Ho), make_map(Wi, Wo), B * C, RS_ZCHUNK, (Do + RS_ZCHUNK - 1) / RS_ZCHUNK, pre * post};
  const int nc = channels_per_thread(B * C);
  dim3 block(32, 8, 1), grid((Wo + 31) / 32, (Ho + 7) / 8, g.nzc * (B * C / nc));
  cudaStream_t st = as_stream(stream);
  if (nc == 3) resize_fwd_kernel<3><<<grid, block, 0, st>>>(x, out, g);
  else if (nc == 2) resize_fwd_kernel<2><<<grid, block, 0, st>>>(x, out, g);
  else resize_fwd_kernel<1><<<grid, block, 0, st>>>(x, out, g);
  return check_launch("resize_fwd");
}

extern "C" int vxm_resize_bwd(const float* grad_out, float* grad_x, int B, int C, int Di, int Hi, int Wi,
                              int Do, int Ho, int Wo, float pre, float post, void* stream) {
  int rc = resize_check(B, C, Di, Hi, Wi, Do, Ho, Wo);
  if (rc) return rc;
  VXM_REQUIRE(grad_out && grad_x, "resize_bwd: null pointer");
  ResizeGeom g{make_map(Di, Do), make_map(Hi, Ho), make_map(Wi, Wo), B * C, RS_ZCHUNK, (Di + RS_ZCHUNK - 1) / RS_ZCHUNK, pre * post};
  cudaStream_t st = as_stream(stream);
  // an input index is touched by at most ~2/ratio + 1 outputs; the marching kernel keeps ADJ_L of them in registers
  auto fits = [](const AxisMap& m) { return m.in == m.out || (m.ratio > 0.f && 2.0f / m.ratio + 1.5f <= (float)ADJ_L); };
  const char* ek = getenv("VXM_B200_RESIZE_BWD");      // "march": A/B switch
  const bool up = Do > Di && Ho > Hi && Wo > Wi && !(ek && ek[0] == 'm');
  if (up && fits(g.mz) && fits(g.my) && fits(g.mx) && host_tile_extent(g.mx, RT_X) <= RT_MX && host_tile_extent(g.my, RT_Y) <= RT_MY) {
    const int nc = (B * C) % 3 == 0 ? 3 : 1;
    g.zchunk = RC_ZCHUNK;
    g.nzc = (Di + RC_ZCHUNK - 1) / RC_ZCHUNK;
    VXM_REQUIRE((size_t)g.nzc * (B * C / nc) <= 65535u, "resize_bwd: B*C*D exceeds the launch grid limit");
    VXM_REQUIRE((size_t)nc * Do * Ho * Wo < (1ull << 31), "resize_bwd: gradient exceeds 2^31 elements per channel group");
    dim3 grid((Wi + RT_X - 1) / RT_X, (Hi + RT_Y - 1) / RT_Y, g.nzc * (B * C / nc));
    if (nc == 3) resize_bwd_colsm_kernel<3><<<grid, 256, 0, st>>>(grad_out, grad_x, g);
    else resize_bwd_colsm_kernel<1><<<grid, 256, 0, st>>>(grad_out, grad_x, g);
  } else if (fits(g.mz) && fits(g.my) && fits(g.mx)) {
    const int nc = channels_per_thread(B * C);
    dim3 block(32, 8, 1), grid((Wi + 31) / 32, (Hi + 7) / 8, g.nzc * (B * C / nc));
    if (nc == 3) resize_bwd_march_kernel<3><<<grid, block, 0, st>>>(grad_out, grad_x, g);
    else if (nc == 2) resize_bwd_march_kernel<2><<<grid, block, 0, st>>>(grad_out, grad_x, g);
    else resize_bwd_march_kernel<1><<<grid, block, 0, st>>>(grad_out, grad_x, g);
  } else {
    dim3 block(32, 8, 1), grid((Wi + 31) / 32, (Hi + 7) / 8, Di * B * C);
    resize_bwd_kernel<<<grid, block, 0, st>>>(grad_out, grad_x, g);
  }
  return check_launch("resize_bwd");
}
