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

// Tiled adjoint for upsampling ratios (the backward of `fullsize`, reading the FULL-resolution gradient): the marching
// kernel's lanes read that gradient with stride ~2 (~3 L1 wavefronts per load, 16-25 loads per plane and channel) and are
// bound by L1 wavefronts (143 us at 160x192x224).  Here a CTA stages the output-gradient region its 4 x 8 x 32 input tile
// touches in shared memory with coalesced row loads, then applies the three 1-D adjoints separably in shared memory
// (x: lane = input column with its list in registers; y: warp = input row; z: 4 slices), ~45 shared loads per result.
constexpr int RT_X = 32, RT_Y = 8, RT_Z = 4, RT_MX = 72, RT_MY = 22, RT_MZ = 12;
constexpr size_t RT_SMEM = (size_t)(RT_MZ * RT_MY * RT_MX + RT_MZ * RT_MY * RT_X) * sizeof(float);

__device__ __forceinline__ float adj_dot(const AdjList& L, const float* __restrict__ p, int stride) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < ADJ_L; ++k)
    if (k < L.n) acc += L.w[k] * p[k * stride];
  return acc;
}

__global__ void __launch_bounds__(256, 2) resize_bwd_tile_kernel(const float* __restrict__ gout, float* __restrict__ gx, ResizeGeom g, int nzt) {
  extern __shared__ __align__(16) float rsm[];
  float* G = rsm;                                   // [EZ][EY][RT_MX]  staged output gradient
  float* X1 = rsm + RT_MZ * RT_MY * RT_MX;          // [EZ][EY][32]     after the x adjoint
  float* Y1 = rsm;                                  // [EZ][8][32]      after the y adjoint (reuses G)
  __shared__ AdjList ends[6];                       // first / last list of the tile along x, y, z
  __shared__ AdjList zl[RT_Z];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bx = blockIdx.x * RT_X, by = blockIdx.y * RT_Y;
  const int bz = (blockIdx.z % nzt) * RT_Z, bc = blockIdx.z / nzt;
  if (tid < 6) {
    const AxisMap& m = tid < 2 ? g.mx : (tid < 4 ? g.my : g.mz);
    const int b0 = tid < 2 ? bx : (tid < 4 ? by : bz), T = tid < 2 ? RT_X : (tid < 4 ? RT_Y : RT_Z);
    make_adj(m, (tid & 1) ? min(b0 + T - 1, m.in - 1) : b0, ends[tid]);
  } else if (tid >= 32 && tid < 32 + RT_Z) {
    const int iz = bz + (tid - 32);
    if (iz < g.mz.in) make_adj(g.mz, iz, zl[tid - 32]); else zl[tid - 32].n = 0;
  }
  __syncthreads();
  const int xlo = ends[0].lo, ylo = ends[2].lo, zlo = ends[4].lo;
  const int EX = min(ends[1].lo + ends[1].n - xlo, RT_MX), EY = min(ends[3].lo + ends[3].n - ylo, RT_MY), EZ = min(ends[5].lo + ends[5].n - zlo, RT_MZ);
  const size_t oHW = (size_t)g.my.out * g.mx.out, cout = oHW * g.mz.out;
  const float* gb = gout + (size_t)bc * cout + (size_t)zlo * oHW + (size_t)ylo * g.mx.out + xlo;
  for (int row = warp; row < EZ * EY; row += 8) {
    const int z = row / EY, y = row - z * EY;
    const float* src = gb + (size_t)z * oHW + (size_t)y * g.mx.out;
    float* dst = G + (z * RT_MY + y) * RT_MX;
    for (int x = lane; x < EX; x += 32) dst[x] = __ldg(src + x);
  }
  __syncthreads();
  const int ix = bx + lane, iy = by + warp;
  {
    AdjList X;
    if (ix < g.mx.in) make_adj(g.mx, ix, X); else { X.lo = xlo; X.n = 0; }
    const int xo = X.lo - xlo;
    for (int row = warp; row < EZ * EY; row += 8) {
      const int z = row / EY, y = row - z * EY;
      X1[(z * RT_MY + y) * RT_X + lane] = adj_dot(X, G + (z * RT_MY + y) * RT_MX + xo, 1);
    }
  }
  __syncthreads();
  {
    AdjList Y;
    if (iy < g.my.in) make_adj(g.my, iy, Y); else { Y.lo = ylo; Y.n = 0; }
    const int yo = Y.lo - ylo;
    for (int z = 0; z < EZ; ++z) Y1[(z * RT_Y + warp) * RT_X + lane] = adj_dot(Y, X1 + (z * RT_MY + yo) * RT_X + lane, RT_X);
  }
  __syncthreads();
  if (ix < g.mx.in && iy < g.my.in) {
    float* ob = gx + (((size_t)bc * g.mz.in + bz) * g.my.in + iy) * g.mx.in + ix;
#pragma unroll
    for (int j = 0; j < RT_Z; ++j) {
      if (bz + j < g.mz.in) {
        const AdjList& Z = zl[j];
        ob[(size_t)j * g.my.in * g.mx.in] = adj_dot(Z, Y1 + ((Z.lo - zlo) * RT_Y + warp) * RT_X + lane, RT_Y * RT_X) * g.scale;
      }
    }
  }
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
  if (up && fits(g.mz) && fits(g.my) && fits(g.mx) && host_tile_extent(g.mx, RT_X) <= RT_MX && host_tile_extent(g.my, RT_Y) <= RT_MY &&
      host_tile_extent(g.mz, RT_Z) <= RT_MZ) {
    const int nzt = (Di + RT_Z - 1) / RT_Z;
    VXM_REQUIRE((size_t)nzt * B * C <= 65535u, "resize_bwd: B*C*D exceeds the launch grid limit");
    VXM_CUDA(cudaFuncSetAttribute(resize_bwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM));
    dim3 grid((Wi + RT_X - 1) / RT_X, (Hi + RT_Y - 1) / RT_Y, nzt * B * C);
    resize_bwd_tile_kernel<<<grid, 256, RT_SMEM, st>>>(grad_out, grad_x, g, nzt);
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
