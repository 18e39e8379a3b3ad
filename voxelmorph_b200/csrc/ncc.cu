// Local normalised cross-correlation loss (reference voxelmorph/torch/losses.py:15-67).
//
// The reference forms five zero-padded box sums with five dense F.conv3d calls against a
// ones(1,1,9,9,9) filter (50 GFLOP of multiply-by-one at 160x192x224) plus ~15 elementwise
// kernels.  Here one kernel does everything: each CTA owns a (16 x 32) column of the volume and
// marches along D; per input slice it stages I and J (with halo) in shared memory, forms the
// products, runs the W pass and the H pass through shared memory, keeps the last `wd` slice sums
// in a register ring for the D pass, evaluates cc exactly as losses.py:57-65 does and reduces.
// Box sums are direct sums (no running prefix), zero padded, divisor prod(win) everywhere.
//
// Backward (closed form, see oracle/spec_np.ncc_grad_pred): with den = Ivar*Jvar + 1e-5,
//   A = 2 cross/den,  Bq = -cross^2 Ivar/den^2,  T = A u_I + 2 Bq u_J
//   dL/dJ = -(1/N) [ I S(A) + 2 J S(Bq) - S(T) ]
// The forward stores A, Bq, T (3 fields); the backward box-sums them with the same machinery.
//
// Algorithmic bytes (fp32): forward 8 B/voxel (read I, J); backward 12 B/voxel (I, J, dJ)
// (+ 12 B/voxel written and read again for the three saved fields in training).
#include "common.cuh"

namespace vxm {

constexpr int NTH = 16, NTW = 32, NHALO = 4;        // tile and max halo (window <= 9)
constexpr int NIH = NTH + 2 * NHALO, NIW = NTW + 2 * NHALO;
// Depth chunk per CTA.  Every chunk re-reads wd - 1 halo slices, and the grid should fill the 2-CTA-per-SM slots in
// whole waves: pick the chunk count that minimises waves x (slices per chunk + halo).  VXM_B200_NCC_ZCHUNK overrides.
static int ncc_zchunk(int D, long long tiles, int wd) {
  const char* e = getenv("VXM_B200_NCC_ZCHUNK");
  if (e && atoi(e) >= 4) return atoi(e);
  const long long slots = 2LL * sm_count();
  int best = D;
  double best_cost = 1e300;
  for (int nch = 1; nch <= D; ++nch) {
    const int zc = (D + nch - 1) / nch;
    if (zc < 8 && nch > 1) break;
    const long long ctas = tiles * ((D + zc - 1) / zc);
    const double cost = (double)((ctas + slots - 1) / slots) * (zc + wd - 1);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = zc; }
  }
  return best;
}

struct NccArgs {
  const float* I;
  const float* J;
  const float* saved_in;   // bwd: A, Bq, T
  float* saved_out;        // fwd: A, Bq, T (may be null)
  float* out;              // fwd: loss scalar ; bwd: grad_J
  const float* grad_loss;  // bwd
  ReduceWork rw;
  int B, D, H, W, wd, wh, ww, zchunk;
  float nwin;              // prod(win)
  double scale;            // -1 / (B*D*H*W)
};

template <int MODE, int WD>
__global__ void __launch_bounds__(256, 2) ncc_kernel(NccArgs a) {   // 2 CTAs per SM: the per-slice barriers of one overlap the loads of the other
  constexpr int NF = MODE == 0 ? 2 : 3;
  constexpr int NS = MODE == 0 ? 5 : 3;
  __shared__ __align__(16) float s_in[NF][NIH][NIW];
  __shared__ __align__(16) float s_w[NS][NIH + 2][NTW];
  __shared__ double s_red[32];

  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;  // ty in 0..7 -> rows ty and ty+8
  const int w0 = blockIdx.x * NTW, h0 = blockIdx.y * NTH;
  const int nchunks = (a.D + a.zchunk - 1) / a.zchunk;
  const int chunk = blockIdx.z % nchunks, b = blockIdx.z / nchunks;
  const int z0 = chunk * a.zchunk, z1 = min(z0 + a.zchunk, a.D);
  const int pd = WD / 2, ph = a.wh / 2, pw = a.ww / 2;
  const size_t HW = (size_t)a.H * a.W, DHW = HW * a.D;
  const float* f0 = (MODE == 0 ? a.I : a.saved_in) + (size_t)b * (MODE == 0 ? 1 : 3) * DHW;
  const float* f1 = MODE == 0 ? a.J + (size_t)b * DHW : f0 + DHW;
  const float* f2 = MODE == 0 ? nullptr : f0 + 2 * DHW;

  float ring[2][WD][NS];
#pragma unroll
  for (int o = 0; o < 2; ++o)
#pragma unroll
    for (int j = 0; j < WD; ++j)
#pragma unroll
      for (int s = 0; s < NS; ++s) ring[o][j][s] = 0.f;

  double local = 0.0;
  const int rows_in = NTH + 2 * ph, cols_in = NTW + 2 * pw;

  for (int base = z0 - pd; base < z1 + pd; base += WD) {
#pragma unroll
    for (int j = 0; j < WD; ++j) {
      const int zi = base + j;
      if (zi < z1 + pd) {  // block-uniform
        const bool inside = zi >= 0 && zi < a.D;
        if (inside) {
          // ---- stage the slice (zero padded) ----
          for (int idx = tid; idx < rows_in * cols_in; idx += 256) {
            int r = idx / cols_in, c = idx - r * cols_in;
            int h = h0 - ph + r, w = w0 - pw + c;
            bool ok = h >= 0 && h < a.H && w >= 0 && w < a.W;
            size_t off = (size_t)zi * HW + (size_t)h * a.W + w;
            s_in[0][r][c] = ok ? __ldg(f0 + off) : 0.f;
            s_in[1][r][c] = ok ? __ldg(f1 + off) : 0.f;
            if (NF == 3) s_in[2][r][c] = ok ? __ldg(f2 + off) : 0.f;
          }
          __syncthreads();
          // ---- W pass: each thread forms 4 adjacent window sums of one row from 12 staged values (register blocked) ----
          if (tid < rows_in * (NTW / 4)) {
            const int r = tid >> 3, c4 = (tid & 7) * 4;
            float x0[12], x1[12], x2[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const float4 a4 = *reinterpret_cast<const float4*>(&s_in[0][r][c4 + 4 * q]);
              const float4 b4 = *reinterpret_cast<const float4*>(&s_in[1][r][c4 + 4 * q]);
              x0[4 * q] = a4.x; x0[4 * q + 1] = a4.y; x0[4 * q + 2] = a4.z; x0[4 * q + 3] = a4.w;
              x1[4 * q] = b4.x; x1[4 * q + 1] = b4.y; x1[4 * q + 2] = b4.z; x1[4 * q + 3] = b4.w;
              if (MODE == 1) {
                const float4 c4v = *reinterpret_cast<const float4*>(&s_in[NF - 1][r][c4 + 4 * q]);
                x2[4 * q] = c4v.x; x2[4 * q + 1] = c4v.y; x2[4 * q + 2] = c4v.z; x2[4 * q + 3] = c4v.w;
              }
            }
            float acc[NS][4];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
              for (int o = 0; o < 4; ++o) acc[s][o] = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
              if (k < a.ww) {
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                  const float u = x0[o + k], v = x1[o + k];
                  if (MODE == 0) {
                    acc[0][o] += u; acc[1][o] += v; acc[2][o] += u * u; acc[3][o] += v * v; acc[4][o] += u * v;
                  } else {
                    acc[0][o] += u; acc[1][o] += v; acc[2][o] += x2[o + k];
                  }
                }
              }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s)
              *reinterpret_cast<float4*>(&s_w[s][r][c4]) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
          }
          __syncthreads();
          // ---- H pass -> ring slot j: two ADJACENT rows per thread share their 10 loads ----
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            float col[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) col[k] = s_w[s][2 * ty + k][tx];
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
              if (k < a.wh) { a0 += col[k]; a1 += col[k + 1]; }
            }
            ring[0][j][s] = a0;
            ring[1][j][s] = a1;
          }
        } else {
#pragma unroll
          for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int s = 0; s < NS; ++s) ring[o][j][s] = 0.f;
        }
        // ---- D pass + pointwise for output slice zo ----
        const int zo = zi - pd;
        if (zo >= z0) {
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            int h = h0 + 2 * ty + o, w = w0 + tx;
            if (h < a.H && w < a.W) {
              float S[NS];
#pragma unroll
              for (int s = 0; s < NS; ++s) {
                float acc = 0.f;
#pragma unroll
                for (int q = 0; q < WD; ++q) acc += ring[o][q][s];
                S[s] = acc;
              }
              size_t off = (size_t)zo * HW + (size_t)h * a.W + w;
              if (MODE == 0) {
                // losses.py:57-65, evaluated left to right with one rounding per op
                float Is = S[0], Js = S[1], I2s = S[2], J2s = S[3], IJs = S[4];
                float uI = __fdiv_rn(Is, a.nwin), uJ = __fdiv_rn(Js, a.nwin);
                float cross = __fadd_rn(__fsub_rn(__fsub_rn(IJs, __fmul_rn(uJ, Is)), __fmul_rn(uI, Js)),
                                        __fmul_rn(__fmul_rn(uI, uJ), a.nwin));
                float Ivar = __fadd_rn(__fsub_rn(I2s, __fmul_rn(__fmul_rn(2.f, uI), Is)),
                                       __fmul_rn(__fmul_rn(uI, uI), a.nwin));
                float Jvar = __fadd_rn(__fsub_rn(J2s, __fmul_rn(__fmul_rn(2.f, uJ), Js)),
                                       __fmul_rn(__fmul_rn(uJ, uJ), a.nwin));
                float den = __fadd_rn(__fmul_rn(Ivar, Jvar), 1e-5f);
                float cc = __fdiv_rn(__fmul_rn(cross, cross), den);
                local += (double)cc;
                if (a.saved_out) {
                  float A = 2.f * cross / den;
                  float Bq = -(cross * cross) * Ivar / (den * den);
                  float T = A * uI + 2.f * Bq * uJ;
                  float* so = a.saved_out + (size_t)b * 3 * DHW + off;
                  so[0] = A; so[DHW] = Bq; so[2 * DHW] = T;
                }
              } else {
                float Iv = __ldg(a.I + (size_t)b * DHW + off), Jv = __ldg(a.J + (size_t)b * DHW + off);
                float gl = __ldg(a.grad_loss) * (float)a.scale;
                a.out[(size_t)b * DHW + off] = gl * (Iv * S[0] + 2.f * Jv * S[1] - S[2]);
              }
            }
          }
        }
        __syncthreads();  // s_in / s_w reuse in the next slice
      }
    }
  }
  if (MODE == 0) {
    double tot = block_sum<double>(local, s_red);
    int nblocks = gridDim.x * gridDim.y * gridDim.z;
    int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    finish_reduce(tot, a.rw, nblocks, bid, a.scale, a.out, s_red);
  }
}

template <int MODE>
static int ncc_launch(const NccArgs& a, dim3 grid, cudaStream_t st) {
  switch (a.wd) {
    case 1: ncc_kernel<MODE, 1><<<grid, 256, 0, st>>>(a); break;
    case 3: ncc_kernel<MODE, 3><<<grid, 256, 0, st>>>(a); break;
    case 5: ncc_kernel<MODE, 5><<<grid, 256, 0, st>>>(a); break;
    case 7: ncc_kernel<MODE, 7><<<grid, 256, 0, st>>>(a); break;
    case 9: ncc_kernel<MODE, 9><<<grid, 256, 0, st>>>(a); break;
    default: set_error("ncc: unsupported window depth %d", a.wd); return VXM_ERR_UNSUPPORTED;
  }
  return check_launch(MODE == 0 ? "ncc_fwd" : "ncc_bwd");
}

static int ncc_check(int B, int D, int H, int W, int wd, int wh, int ww, dim3* grid, int* zchunk) {
  VXM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "ncc: non-positive dimension");
  auto okw = [](int w) { return w >= 1 && w <= 9 && (w & 1); };
  if (!(okw(wd) && okw(wh) && okw(ww))) {
    set_error("ncc: window (%d,%d,%d) unsupported (odd sizes 1..9 only)", wd, wh, ww);
    return VXM_ERR_UNSUPPORTED;
  }
  const int zc = ncc_zchunk(D, (long long)B * ((W + NTW - 1) / NTW) * ((H + NTH - 1) / NTH), wd);
  *zchunk = zc;
  int nchunks = (D + zc - 1) / zc;
  *grid = dim3((W + NTW - 1) / NTW, (H + NTH - 1) / NTH, nchunks * B);
  VXM_REQUIRE((size_t)grid->x * grid->y * grid->z <= (size_t)kMaxReduceBlocks && grid->z <= 65535u,
              "ncc: volume too large for the reduction workspace");
  return VXM_OK;
}

}  // namespace vxm

using namespace vxm;

extern "C" size_t vxm_reduce_workspace_bytes(void) { return sizeof(double) * kMaxReduceBlocks + 256; }
extern "C" size_t vxm_ncc_workspace_bytes(int, int, int, int) { return vxm_reduce_workspace_bytes(); }

namespace vxm {
ReduceWork as_reduce_work(void* work) {
  ReduceWork rw;
  rw.counter = reinterpret_cast<unsigned int*>(work);
  rw.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(work) + 256);
  return rw;
}
}  // namespace vxm

extern "C" int vxm_ncc_fwd(const float* I, const float* J, float* loss, float* saved, void* work, int B,
                           int D, int H, int W, int wd, int wh, int ww, void* stream) {
  dim3 grid;
  int zchunk = 0;
  int rc = ncc_check(B, D, H, W, wd, wh, ww, &grid, &zchunk);
  if (rc) return rc;
  VXM_REQUIRE(I && J && loss && work, "ncc_fwd: null pointer");
  NccArgs a{};
  a.I = I; a.J = J; a.saved_out = saved; a.out = loss; a.rw = as_reduce_work(work);
  a.B = B; a.D = D; a.H = H; a.W = W; a.wd = wd; a.wh = wh; a.ww = ww;
  a.nwin = (float)(wd * wh * ww); a.zchunk = zchunk;
  a.scale = -1.0 / ((double)B * D * H * W);
  return ncc_launch<0>(a, grid, as_stream(stream));
}

extern "C" int vxm_ncc_bwd(const float* I, const float* J, const float* saved, const float* grad_loss,
                           float* grad_J, int B, int D, int H, int W, int wd, int wh, int ww, void* stream) {
  dim3 grid;
  int zchunk = 0;
  int rc = ncc_check(B, D, H, W, wd, wh, ww, &grid, &zchunk);
  if (rc) return rc;
  VXM_REQUIRE(I && J && saved && grad_loss && grad_J, "ncc_bwd: null pointer");
  NccArgs a{};
  a.I = I; a.J = J; a.saved_in = saved; a.out = grad_J; a.grad_loss = grad_loss;
  a.B = B; a.D = D; a.H = H; a.W = W; a.wd = wd; a.wh = wh; a.ww = ww;
  a.nwin = (float)(wd * wh * ww); a.zchunk = zchunk;
  a.scale = -1.0 / ((double)B * D * H * W);
  return ncc_launch<1>(a, grid, as_stream(stream));
}
