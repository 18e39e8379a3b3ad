// Weight gradient of the k=3 convolution on tcgen05 tensor cores:
//     gw[tap][ci][co] = sum_{b,v} x[b, v + tap - 1, ci] * gz[b, v, co]          (gz = grad wrt the conv output)
// (autograd of nn.Conv3d at reference voxelmorph/torch/networks.py:299,211).
//
// Per (tap, 16 voxels) one MMA  D[64 x N] += A^T[64 ch x 16 vox] * B[16 vox x N]  with both operands
// "MN-major" straight out of the channels-last slabs the forward kernel also uses:
//   A = x slab  [Cin/8][180 rows][8 ch]  (halo'd 18 x 10 voxels of one input slice; the tap is a start-address
//       offset; K = voxels: 8 consecutive w are 16 B apart = one core matrix, the next K group is the next h row)
//   B = gz tile [Cout/8][128 rows][8 ch] (16 x 8 voxels, no halo)
// M = 64 accumulator tiles use 16 of every 32 TMEM lanes, so two taps share one column range (lane offset 16):
// all 27 tap accumulators (27 x 64 x N fp32) stay resident in TMEM for the CTA's whole lifetime; each CTA
// streams its share of the volume, then writes ONE partial [27][64][N]; a second kernel reduces the partials
// over CTAs in fixed order (deterministic) into the fp32 (Cout,Cin,kd,3,3) gradient.
#include "tc_common.cuh"

namespace vxm {
namespace tc {

constexpr int WTH = 16, WTW = 8, WSW = WTW + 2, WSH = WTH + 2;
constexpr int WROWS = WSH * WSW, WPLANE = WROWS * 16;   // x slab plane: 180 rows
constexpr int GPLANE = 128 * 16;                        // gz tile plane: 128 rows
constexpr int WMAXSLOT = 8, WNG = 4, WKMAX = 12;
constexpr int WNLOADER = 128, WNTHREADS = 288;

struct WgradTcArgs {
  const __nv_bfloat16* xa; const __nv_bfloat16* xb;
  const float* xf[4]; long long xf_bs[4]; int nplanar_x;
  int Ca, Cb, up, upd;
  const __nv_bfloat16* gz; int Cg;
  const float* gf[4]; long long gf_bs[4]; int nplanar_g;
  float* partial;
  float* bias_partial;   // [grid][NP]
  int B, D, H, W, KD, NP;
  int tiles_h, tiles_w, dchunk, nchunks, nitems, nslot;
};

// STK > 0 (= channel chunks per slab, 1/2/4; KD == 3 only): the three kd taps are STACKED along M — the slabs of
// consecutive input slices are adjacent in shared memory, so one operand of 3*STK channel chunks (stride = one plane)
// spans slices d-1, d, d+1 and a single MMA accumulates all three kd taps: 72 MMAs per 128-voxel tile instead of 216.
// The ring carries two mirror slots (copies of slots 0 and 1) so that a 3-slab window never wraps.
template <int KD, int NP, int STK>
__global__ void __launch_bounds__(WNTHREADS, 1) wgrad_tc_kernel(const WgradTcArgs a) {
  constexpr int SM = STK == 0 ? 64 : (3 * STK <= 8 ? 64 : 128);   // MMA M
  constexpr int NMIRROR = STK ? 2 : 0;
  extern __shared__ __align__(128) uint8_t smem[];
  const bool px = a.nplanar_x > 0, pg = a.nplanar_g > 0;
  const int nc8 = px ? 1 : (a.Ca + a.Cb) / 8;
  const int ncg = pg ? 1 : a.Cg / 8;
  const uint32_t slab_bytes = (uint32_t)nc8 * WPLANE;
  const uint32_t gt_bytes = (uint32_t)ncg * GPLANE;
  uint8_t* s_slab = smem;
  const int WNSLOT = a.nslot;
  // tail padding so that the M=64 operand (8 channel planes) never reads past the allocation
  uint8_t* s_g = s_slab + (WNSLOT + NMIRROR) * slab_bytes + 16 * WPLANE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_g + WNG * gt_bytes + 4 * GPLANE);
  uint64_t* xfull = bars;
  uint64_t* xempty = bars + WMAXSLOT;
  uint64_t* gfull = bars + 2 * WMAXSLOT;
  uint64_t* gempty = gfull + WNG;
  uint64_t* done = gempty + WNG;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = KD * 9, npairs = STK ? 5 : (T + 1) / 2;
  const uint32_t need_cols = (STK && SM == 128) ? 9u * NP : (uint32_t)npairs * NP;
  const uint32_t tmem_cols = need_cols <= 32 ? 32u : need_cols <= 64 ? 64u : need_cols <= 128 ? 128u : need_cols <= 256 ? 256u : 512u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WNSLOT; ++i) { mbar_init(&xfull[i], WNLOADER); mbar_init(&xempty[i], 1); }
    for (int i = 0; i < WNG; ++i) { mbar_init(&gfull[i], WNLOADER); mbar_init(&gempty[i], 1 + 128); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int HW_tiles = a.tiles_h * a.tiles_w;
  const bool has_work = blockIdx.x < a.nitems;

  if (warp >= 5) {
    // ================================ LOADER ================================
    const int lt = threadIdx.x - 5 * 32;
    uint32_t xcnt = 0, gcnt = 0;
    const int Da = a.upd ? a.D >> 1 : a.D, Ha = a.up ? a.H >> 1 : a.H, Wa = a.up ? a.W >> 1 : a.W;
    const int nca8 = a.Ca >> 3;
    const int nchunk = nc8 * WROWS, ngchunk = ncg * 128;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int wt = item % a.tiles_w, ht = (item / a.tiles_w) % a.tiles_h;
      const int ch = (item / HW_tiles) % a.nchunks, b = item / (HW_tiles * a.nchunks);
      const int h0 = ht * WTH, w0 = wt * WTW, d0 = ch * a.dchunk, d1 = min(d0 + a.dchunk, a.D);
      const int s_begin = KD == 3 ? d0 - 1 : d0, s_end = KD == 3 ? d1 + 1 : d1;
      int soff[WKMAX];
      uint32_t doff[WKMAX];
      if (!px) {
#pragma unroll
        for (int k = 0; k < WKMAX; ++k) {
          const int id = lt + k * WNLOADER;
          soff[k] = -1;
          doff[k] = 0;
          if (id < nchunk) {
            const int c8 = id % nc8, row = id / nc8;
            const int r = row / WSW, c = row - r * WSW;
            const int h = h0 - 1 + r, w = w0 - 1 + c;
            doff[k] = (uint32_t)c8 * WPLANE + (uint32_t)row * 16u;
            if (h >= 0 && h < a.H && w >= 0 && w < a.W) {
              if (c8 < nca8) soff[k] = (((a.up ? h >> 1 : h) * Wa + (a.up ? w >> 1 : w)) * a.Ca + c8 * 8) << 1;
              else soff[k] = (((h * a.W + w) * a.Cb + (c8 - nca8) * 8) << 1) | 1;
            }
          }
        }
      }
      int goff[4];
      uint32_t gdoff[4];
      if (!pg) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int id = lt + k * WNLOADER;
          goff[k] = -1;
          gdoff[k] = 0;
          if (id < ngchunk) {
            const int c8 = id % ncg, row = id / ncg;
            const int h = h0 + (row >> 3), w = w0 + (row & 7);
            gdoff[k] = (uint32_t)c8 * GPLANE + (uint32_t)row * 16u;
            if (h < a.H && w < a.W) goff[k] = (h * a.W + w) * a.Cg + c8 * 8;
          }
        }
      }
      for (int ds = s_begin; ds < s_end; ++ds) {
        // ---- x slab of input slice ds ----
        const int slot = xcnt % WNSLOT;
        mbar_wait(&xempty[slot], ((xcnt / WNSLOT) & 1) ^ 1);
        uint8_t* slab = s_slab + (size_t)slot * slab_bytes;
        const bool dok = ds >= 0 && ds < a.D;
        if (!px) {
          const __nv_bfloat16* baseA = a.xa ? a.xa + (((size_t)b * Da + (dok ? (a.upd ? ds >> 1 : ds) : 0)) * Ha * Wa) * a.Ca : nullptr;
          const __nv_bfloat16* baseB = a.xb ? a.xb + (((size_t)b * a.D + (dok ? ds : 0)) * a.H * a.W) * a.Cb : nullptr;
          const __nv_bfloat16* dummy = a.xa ? a.xa : a.xb;
#pragma unroll
          for (int k = 0; k < WKMAX; ++k) {
            if (lt + k * WNLOADER < nchunk) {
              const bool ok = dok && soff[k] >= 0;
              const __nv_bfloat16* src = ok ? ((soff[k] & 1) ? baseB : baseA) + (soff[k] >> 1) : dummy;
              cp_async16(slab + doff[k], src, ok ? 16u : 0u);
              if (STK && slot < NMIRROR) cp_async16(slab + (size_t)WNSLOT * slab_bytes + doff[k], src, ok ? 16u : 0u);
            }
          }
          cp_async_arrive_noinc(&xfull[slot]);
        } else {
          for (int row = lt; row < WROWS; row += WNLOADER) {
            const int r = row / WSW, c = row - r * WSW;
            const int h = h0 - 1 + r, w = w0 - 1 + c;
            const bool ok = dok && h >= 0 && h < a.H && w >= 0 && w < a.W;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (ok) {
              const size_t off = ((size_t)ds * a.H + h) * a.W + w;
              for (int p = 0; p < a.nplanar_x; ++p) v[p] = __ldg(a.xf[p] + (size_t)b * a.xf_bs[p] + off);
            }
            const uint4 q = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), 0u, 0u);
            *reinterpret_cast<uint4*>(slab + row * 16) = q;
            if (STK && slot < NMIRROR) *reinterpret_cast<uint4*>(slab + (size_t)WNSLOT * slab_bytes + row * 16) = q;
          }
          fence_proxy_async();
          mbar_arrive(&xfull[slot]);
        }
        ++xcnt;
        // ---- gz tile of OUTPUT slice dg (the slice whose window this x slab completes) ----
        const int dg = KD == 3 ? ds - 1 : ds;
        if (dg >= d0 && dg < d1) {
          const int gslot = gcnt % WNG;
          mbar_wait(&gempty[gslot], ((gcnt / WNG) & 1) ^ 1);
          uint8_t* gt = s_g + (size_t)gslot * gt_bytes;
          if (!pg) {
            const __nv_bfloat16* baseG = a.gz + (((size_t)b * a.D + dg) * a.H * a.W) * a.Cg;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (lt + k * WNLOADER < ngchunk) {
                const bool ok = goff[k] >= 0;
                cp_async16(gt + gdoff[k], ok ? baseG + goff[k] : a.gz, ok ? 16u : 0u);
              }
            }
            cp_async_arrive_noinc(&gfull[gslot]);
          } else {
            for (int row = lt; row < 128; row += WNLOADER) {
              const int h = h0 + (row >> 3), w = w0 + (row & 7);
              float v[4] = {0.f, 0.f, 0.f, 0.f};
              if (h < a.H && w < a.W) {
                const size_t off = ((size_t)dg * a.H + h) * a.W + w;
                for (int p = 0; p < a.nplanar_g; ++p) v[p] = __ldg(a.gf[p] + (size_t)b * a.gf_bs[p] + off);
              }
              *reinterpret_cast<uint4*>(gt + row * 16) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), 0u, 0u);
            }
            fence_proxy_async();
            mbar_arrive(&gfull[gslot]);
          }
          ++gcnt;
        }
      }
    }
  } else if (warp == 4) {
    // ================================ MMA ISSUER ================================
    if (has_work) {   // whole warp, warp-uniform; one elected lane issues
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
      const uint32_t slab_u32 = smem_u32(s_slab), g_u32 = smem_u32(s_g);
      uint32_t xbase = 0, gcnt = 0;
      bool first_tile = true;
      for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int ch = (item / HW_tiles) % a.nchunks;
        const int d0 = ch * a.dchunk, d1 = min(d0 + a.dchunk, a.D);
        const int nd = d1 - d0;
        for (int j = 0; j < nd; ++j) {
          if (KD == 3) {
            if (j == 0) for (int q = 0; q < 2; ++q) { uint32_t c = xbase + q; mbar_wait(&xfull[c % WNSLOT], (c / WNSLOT) & 1); }
            uint32_t c = xbase + j + 2;
            mbar_wait(&xfull[c % WNSLOT], (c / WNSLOT) & 1);
          } else {
            uint32_t c = xbase + j;
            mbar_wait(&xfull[c % WNSLOT], (c / WNSLOT) & 1);
          }
          const uint32_t gs = gcnt % WNG;
          mbar_wait(&gfull[gs], (gcnt / WNG) & 1);
          tc_fence_after();
          const uint64_t bdesc0 = make_desc_mnmajor_noswz(g_u32 + gs * gt_bytes, 128u, (uint32_t)GPLANE);
          uint64_t adesc_kd[KD];
#pragma unroll
          for (int kd = 0; kd < KD; ++kd)
            adesc_kd[kd] = make_desc_mnmajor_noswz(slab_u32 + ((xbase + j + kd) % WNSLOT) * slab_bytes, (uint32_t)WSW * 16u, (uint32_t)WPLANE);
          const uint32_t acc0 = first_tile ? 0u : 1u;
          if constexpr (STK > 0) {
            constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(SM >> 4) << 24);
            // window = slabs of slices j, j+1, j+2, contiguous thanks to the mirror slots
            const uint64_t adesc_w = make_desc_mnmajor_noswz(slab_u32 + ((xbase + j) % WNSLOT) * slab_bytes, (uint32_t)WSW * 16u, (uint32_t)WPLANE);
            if (elect_one()) {
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                  const int t9 = kh * 3 + kw;
                  const uint32_t tmem_d = SM == 128 ? tmem_base + (uint32_t)(t9 * NP)
                                                    : tmem_base + ((uint32_t)((t9 & 1) * 16) << 16) + (uint32_t)((t9 >> 1) * NP);
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const uint64_t adesc = adesc_w + (uint64_t)((kh + 2 * i) * WSW + kw);
                    const uint64_t bdesc = bdesc0 + (uint64_t)(2 * i * 128 / 16);
                    umma_f16(tmem_d, adesc, bdesc, idesc_s, i == 0 ? acc0 : 1u);
                  }
                }
              }
              umma_commit(&xempty[(xbase + j) % WNSLOT]);
              umma_commit(&gempty[gs]);
            }
          } else
          if (elect_one()) {
#pragma unroll
            for (int kd = 0; kd < KD; ++kd) {
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                  const int tap = (kd * 3 + kh) * 3 + kw;
                  const uint32_t tmem_d = tmem_base + ((uint32_t)((tap & 1) * 16) << 16) + (uint32_t)((tap >> 1) * NP);
#pragma unroll
                  for (int i = 0; i < 8; ++i) {   // 8 x 16 voxels = the 128-voxel tile; 16-byte address units
                    const uint64_t adesc = adesc_kd[kd] + (uint64_t)((kh + 2 * i) * WSW + kw);
                    const uint64_t bdesc = bdesc0 + (uint64_t)(2 * i * 128 / 16);
                    umma_f16(tmem_d, adesc, bdesc, idesc, i == 0 ? acc0 : 1u);
                  }
                }
              }
            }
            umma_commit(&xempty[(xbase + j) % WNSLOT]);
            umma_commit(&gempty[gs]);
          }
          __syncwarp();
          first_tile = false;
          ++gcnt;
        }
        if (KD == 3) {
          if (elect_one()) {
            umma_commit(&xempty[(xbase + nd) % WNSLOT]);
            umma_commit(&xempty[(xbase + nd + 1) % WNSLOT]);
          }
          __syncwarp();
          xbase += nd + 2;
        } else {
          xbase += nd;
        }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    }
  } else {
    // ================================ EPILOGUE WARPS ================================
    // While the tensor core works they fold the bias gradient (sum of gz over voxels) out of the staged gz tiles:
    // thread r owns tile row r and accumulates its NP channels in registers.
    float bsum[NP];
#pragma unroll
    for (int c = 0; c < NP; ++c) bsum[c] = 0.f;
    if (has_work) {
      uint32_t gcnt = 0;
      const int rowi = warp * 32 + lane;
      for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int ch = (item / HW_tiles) % a.nchunks;
        const int nd = min(ch * a.dchunk + a.dchunk, a.D) - ch * a.dchunk;
        for (int j = 0; j < nd; ++j) {
          const uint32_t gs = gcnt % WNG;
          mbar_wait(&gfull[gs], (gcnt / WNG) & 1);
          const uint8_t* gt = s_g + (size_t)gs * gt_bytes + rowi * 16;
#pragma unroll
          for (int c8 = 0; c8 < NP / 8; ++c8) {
            const uint4 q = *reinterpret_cast<const uint4*>(gt + (size_t)c8 * GPLANE);
            const __nv_bfloat162* hq = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(hq[e]);
              bsum[c8 * 8 + 2 * e] += f.x;
              bsum[c8 * 8 + 2 * e + 1] += f.y;
            }
          }
          mbar_arrive(&gempty[gs]);
          ++gcnt;
        }
      }
    }
    float* part = a.partial + (size_t)blockIdx.x * T * 64 * NP;
    const int ci = warp * 16 + (lane & 15);
    if (has_work && STK > 0 && SM == 128) {
      // rows = kd * (8*STK) + ci on TMEM lanes 0..127; columns t9 * NP
      mbar_wait(done, 0);
      tc_fence_after();
      const int rowm = warp * 32 + lane, kd = rowm / (8 * STK), cis = rowm % (8 * STK);
      for (int t9 = 0; t9 < 9; ++t9) {
        for (int c0 = 0; c0 < NP; c0 += 8) {
          uint32_t r[8];
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)t9 * NP + c0;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                       : "r"(taddr) : "memory");
          tmem_ld_wait();
          if (kd < 3) {
            float4* o = reinterpret_cast<float4*>(part + ((size_t)(kd * 9 + t9) * 64 + cis) * NP + c0);
            o[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
            o[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
          }
        }
      }
    } else if (has_work && STK > 0) {
      // M = 64: rows kd * (8*STK) + ci on lanes (m % 16) + 32 * (m / 16); tap pairs share columns (lane offset 16)
      mbar_wait(done, 0);
      tc_fence_after();
      const int rowm = warp * 16 + (lane & 15), kd = rowm / (8 * STK), cis = rowm % (8 * STK);
      for (int p = 0; p < 5; ++p) {
        const int t9 = 2 * p + (lane >> 4);
        for (int c0 = 0; c0 < NP; c0 += 8) {
          uint32_t r[8];
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)p * NP + c0;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                       : "r"(taddr) : "memory");
          tmem_ld_wait();
          if (t9 < 9 && kd < 3) {
            float4* o = reinterpret_cast<float4*>(part + ((size_t)(kd * 9 + t9) * 64 + cis) * NP + c0);
            o[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
            o[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
          }
        }
      }
    } else if (has_work) {
      mbar_wait(done, 0);
      tc_fence_after();
      for (int p = 0; p < npairs; ++p) {
        const int tap = 2 * p + (lane >> 4);
        for (int c0 = 0; c0 < NP; c0 += 8) {
          uint32_t r[8];
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)p * NP + c0;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                       : "r"(taddr) : "memory");
          tmem_ld_wait();
          if (tap < T) {
            float4* o = reinterpret_cast<float4*>(part + ((size_t)tap * 64 + ci) * NP + c0);
            o[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
            o[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < T * 64 * NP; i += 128) part[i] = 0.f;
    }
    // bias partial of this CTA: deterministic reduction over the 128 rows through shared memory (the slab ring is idle now)
    {
      float* s_b = reinterpret_cast<float*>(s_slab);
      const int rowi = warp * 32 + lane;
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int c = 0; c < NP; ++c) s_b[rowi * (NP + 1) + c] = bsum[c];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (rowi < NP) {
        float t = 0.f;
        for (int r2 = 0; r2 < 128; ++r2) t += s_b[r2 * (NP + 1) + rowi];
        a.bias_partial[(size_t)blockIdx.x * NP + rowi] = t;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// gw[co][ci][tap] (+)= sum_cta partial[cta][tap][ci][co]   (fixed order -> deterministic).  Threads walk the partial
// layout (co fastest) so the ncta reads per element are coalesced; the small transposed write is scattered.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ gw, int ncta, int T, int NP, int Cout, int Cin,
                                    const float* __restrict__ bias_partial, float* __restrict__ gb, int accumulate) {
  if (gb && blockIdx.x == 0 && threadIdx.x < Cout) {
    float acc = accumulate ? gb[threadIdx.x] : 0.f;
    for (int c = 0; c < ncta; ++c) acc += bias_partial[(size_t)c * NP + threadIdx.x];
    gb[threadIdx.x] = acc;
  }
  const int per_cta = T * 64 * NP;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < per_cta; j += gridDim.x * blockDim.x) {
    const int co = j % NP, ci = (j / NP) % 64, tap = j / (NP * 64);
    if (co >= Cout || ci >= Cin) continue;
    float acc = 0.f;
    for (int c = 0; c < ncta; ++c) acc += partial[(size_t)c * per_cta + j];
    float* dst = gw + ((size_t)co * Cin + ci) * T + tap;
    *dst = (accumulate ? *dst : 0.f) + acc;
  }
}

// bias gradient from a bf16 NDHWC tensor: gb[c] = sum_v gz[v][c]; two-stage, deterministic
__global__ void __launch_bounds__(256) bias_grad_partial_kernel(const __nv_bfloat16* __restrict__ gz, float* __restrict__ part, size_t V, int C) {
  // thread t handles channel t % C of voxels t / C, t / C + stride...
  __shared__ float s[256];
  const int c = threadIdx.x % C, lane_v = threadIdx.x / C, vper = 256 / C;
  float acc = 0.f;
  if (lane_v < vper)
    for (size_t v = (size_t)blockIdx.x * vper + lane_v; v < V; v += (size_t)gridDim.x * vper) acc += __bfloat162float(gz[v * C + c]);
  s[threadIdx.x] = lane_v < vper ? acc : 0.f;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int k = 0; k < vper; ++k) t += s[k * C + threadIdx.x];
    part[(size_t)blockIdx.x * C + threadIdx.x] = t;
  }
}
__global__ void bias_grad_final_kernel(const float* __restrict__ part, float* __restrict__ gb, int nblocks, int C, int Creal) {
  const int c = threadIdx.x;
  if (c >= Creal) return;
  float acc = 0.f;
  for (int b = 0; b < nblocks; ++b) acc += part[(size_t)b * C + c];
  gb[c] = acc;
}

}  // namespace tc
}  // namespace vxm

using namespace vxm;
using namespace vxm::tc;

namespace vxm {
namespace tcw {   // conv3d_tc_wgrad2.cu: the kw-stacked Toeplitz formulation (channels-last bf16 sources only)
bool wgrad2_supported(int Ca, int Cb, int Cg);
struct ReduceDesc;
int wgrad2_launch(const void* x, int Cx, int up, const void* gz, int Cg, float* grad_w, float* grad_b, void* work, int B, int D, int H, int W,
                  int kd, int Cout_real, int Cin_total, int ci_off, int ci_cnt, int accumulate, cudaStream_t st, ReduceDesc* defer = nullptr,
                  size_t* work_used = nullptr, bool khm = false);
}
}

static int wgrad_np(int cg) { return cg <= 8 ? 8 : cg <= 16 ? 16 : 32; }

extern "C" size_t vxm_conv3d_tc_wgrad_workspace_bytes(int kd) {
  // worst case: 256 CTAs x 27 taps x 64 rows x 32 cols fp32, + bias partials
  return (size_t)256 * kd * 9 * 64 * 32 * sizeof(float) + 1024 * 32 * sizeof(float);
}

extern "C" int vxm_conv3d_tc_wgrad(const void* xa, const void* xb, const float* const* xf, const long long* xf_bs, int nplanar_x,
                                   const void* gz, const float* const* gf, const long long* gf_bs, int nplanar_g,
                                   float* grad_w, float* grad_b, void* work, int B, int D, int H, int W, int Ca, int Cb, int up,
                                   int Cin_real, int Cg, int Cout_real, int kd, int accumulate, void* stream) {
  VXM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && grad_w && work, "conv3d_tc_wgrad: bad argument");
  VXM_REQUIRE(kd == 1 || kd == 3, "conv3d_tc_wgrad: kd must be 1 or 3");
  WgradTcArgs a{};
  int Cin;
  if (nplanar_x > 0) {
    VXM_REQUIRE(nplanar_x <= 4 && xf && xf_bs, "conv3d_tc_wgrad: at most 4 planar x sources");
    for (int i = 0; i < nplanar_x; ++i) { a.xf[i] = xf[i]; a.xf_bs[i] = xf_bs[i]; }
    a.nplanar_x = nplanar_x;
    Cin = 8;
  } else {
    VXM_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0 && Ca + Cb >= 8 && Ca + Cb <= 64, "conv3d_tc_wgrad: channel counts (%d,%d) unsupported", Ca, Cb);
    VXM_REQUIRE((Ca == 0 || xa) && (Cb == 0 || xb), "conv3d_tc_wgrad: missing source tensor");
    Cin = Ca + Cb;
  }
  VXM_REQUIRE(Cin_real <= Cin && Cin_real > 0, "conv3d_tc_wgrad: Cin_real out of range");
  if (nplanar_g > 0) {
    VXM_REQUIRE(nplanar_g <= 4 && gf && gf_bs, "conv3d_tc_wgrad: at most 4 planar gz sources");
    for (int i = 0; i < nplanar_g; ++i) { a.gf[i] = gf[i]; a.gf_bs[i] = gf_bs[i]; }
    a.nplanar_g = nplanar_g;
    Cg = 8;
  } else {
    VXM_REQUIRE(gz && Cg % 8 == 0 && Cg >= 8 && Cg <= 32, "conv3d_tc_wgrad: gz channels %d unsupported", Cg);
  }
  VXM_REQUIRE(Cout_real > 0 && Cout_real <= Cg, "conv3d_tc_wgrad: Cout_real out of range");
  if (nplanar_x == 0 && nplanar_g == 0 && tcw::wgrad2_supported(Ca, Cb, Cg)) {
    // one launch per source tensor of the (virtual) concatenation: xa -> weights [0, Ca), xb -> [Ca, Ca + Cb)
    cudaStream_t st2 = as_stream(stream);
    int rc = 0;
    if (Ca) {
      const int cnt = Cin_real < Ca ? Cin_real : Ca;
      rc = tcw::wgrad2_launch(xa, Ca, up, gz, Cg, grad_w, grad_b, work, B, D, H, W, kd, Cout_real, Cin_real, 0, cnt, accumulate, st2);
      if (rc) return rc;
    }
    if (Cb && Cin_real > Ca) {
      const int cnt = Cin_real - Ca < Cb ? Cin_real - Ca : Cb;
      rc = tcw::wgrad2_launch(xb, Cb, 0, gz, Cg, grad_w, Ca ? nullptr : grad_b, work, B, D, H, W, kd, Cout_real, Cin_real, Ca, cnt, accumulate, st2);
    }
    return rc;
  }
  a.xa = (const __nv_bfloat16*)xa; a.xb = (const __nv_bfloat16*)xb; a.gz = (const __nv_bfloat16*)gz;
  a.Ca = Ca; a.Cb = Cb; a.up = up; a.upd = (up && kd == 3) ? 1 : 0; a.Cg = Cg;
  a.B = B; a.D = D; a.H = H; a.W = W; a.KD = kd; a.NP = wgrad_np(Cg);
  VXM_REQUIRE(a.NP == Cg, "conv3d_tc_wgrad: gz channels must be 8, 16 or 32");
  a.tiles_h = (H + WTH - 1) / WTH; a.tiles_w = (W + WTW - 1) / WTW;
  int nsm = sm_count();
  int dchunk = D;
  auto items = [&](int dc) { return (long long)B * a.tiles_h * a.tiles_w * ((D + dc - 1) / dc); };
  while (items(dchunk) < 4LL * nsm && dchunk > 8) dchunk = (dchunk + 1) / 2;
  a.dchunk = dchunk; a.nchunks = (D + dchunk - 1) / dchunk;
  a.nitems = (int)items(dchunk);
  int grid = a.nitems < nsm ? a.nitems : nsm;
  if (grid > 256) grid = 256;
  a.partial = (float*)work;
  a.bias_partial = (float*)work + (size_t)256 * kd * 9 * 64 * 32;
  int nc8 = nplanar_x > 0 ? 1 : Cin / 8, ncg = nplanar_g > 0 ? 1 : Cg / 8;
  VXM_REQUIRE(nc8 * WROWS <= WKMAX * WNLOADER && ncg * 128 <= 4 * WNLOADER, "conv3d_tc_wgrad: tile too large for the loader table");
  const int stk = (kd == 3 && (nc8 == 1 || nc8 == 2 || nc8 == 4)) ? nc8 : 0;
  size_t fixed = 16 * WPLANE + (size_t)(stk ? 2 : 0) * nc8 * WPLANE + (size_t)WNG * ncg * GPLANE + 4 * GPLANE + 512;
  int nslot = (int)((200 * 1024 - fixed) / ((size_t)nc8 * WPLANE));
  if (nslot > WMAXSLOT) nslot = WMAXSLOT;
  VXM_REQUIRE(nslot >= 4, "conv3d_tc_wgrad: not enough shared memory for the slab ring");
  a.nslot = nslot;
  size_t smem = fixed + (size_t)nslot * nc8 * WPLANE;
  cudaStream_t st = as_stream(stream);
#define VXM_WG_LAUNCH(KD_, NP_, STK_)                                                                                        \
  do {                                                                                                                       \
    VXM_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<KD_, NP_, STK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    wgrad_tc_kernel<KD_, NP_, STK_><<<grid, WNTHREADS, smem, st>>>(a);                                                        \
  } while (0)
#define VXM_WG_NP(KD_, STK_)                                                                                            \
  do { if (a.NP == 8) VXM_WG_LAUNCH(KD_, 8, STK_); else if (a.NP == 16) VXM_WG_LAUNCH(KD_, 16, STK_); else VXM_WG_LAUNCH(KD_, 32, STK_); } while (0)
  if (kd == 3) {
    if (stk == 1) VXM_WG_NP(3, 1); else if (stk == 2) VXM_WG_NP(3, 2); else if (stk == 4) VXM_WG_NP(3, 4); else VXM_WG_NP(3, 0);
  } else {
    VXM_WG_NP(1, 0);
  }
  int rc = check_launch("conv3d_tc_wgrad");
  if (rc) return rc;
  int T = kd * 9;
  int per_cta = T * 64 * a.NP;
  wgrad_reduce_kernel<<<(per_cta + 255) / 256, 256, 0, st>>>(a.partial, grad_w, grid, T, a.NP, Cout_real, Cin_real, a.bias_partial, grad_b,
                                                             accumulate);
  return check_launch("conv3d_tc_wgrad_reduce");
}
