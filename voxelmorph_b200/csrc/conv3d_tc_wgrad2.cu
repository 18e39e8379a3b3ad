// Weight gradient of the k=3 convolution on tcgen05 tensor cores, "kw-stacked Toeplitz" formulation:
//     gw[kd][kh][kw][ci][co] = sum_{b,v} x[b, v + (kd,kh,kw) - 1, ci] * gz[b, v, co]
// (autograd of nn.Conv3d at reference voxelmorph/torch/networks.py:299,211; gz = grad wrt the conv output).
//
// Both operands are staged exactly like the forward kernel's A operand (conv3d_tc_s.cu): one shared-memory row per
// voxel, the channels of the row contiguous (32 or 64 bytes) and XOR-swizzled, rows of a 32-voxel-wide (h, w) tile
// in linear order.  Read "MN-major" (MN = channels, K = voxels) the same bytes are a valid UMMA operand, and a
// one-voxel shift of the window is a one-row shift of the start address.  One MMA then covers 16 voxels (K) and
//   M = (kd, ci) : the x slabs of input slices d-1, d, d+1 are adjacent in the ring, so the three kd taps are three
//                  MN atoms one slab apart (leading byte offset = slab pitch);
//   N = (kw, co) : the three kw taps are three MN atoms ONE ROW apart (leading byte offset = one row): the B operand
//                  is a Toeplitz view of the single gz slab, nothing is copied;
//   kh           : three accumulators, the A window start moves by one 32-voxel tile row.
// 24 MMAs per (tile, slice) instead of 72-216 in conv3d_tc_wgrad.cu, each with N = 48 or 96 instead of 16 or 32.
// The x slab keeps its halo columns, the gz slab has ZERO halo columns (and zero pad rows before and after), so the
// products that pair a voxel with a neighbour across the tile-row wrap vanish.
// All 3 (kh) x [M x 3*GOUT] fp32 accumulators stay in TMEM for the CTA's whole lifetime; each CTA writes ONE partial
// [27][G][GOUT]; wgrad2_reduce_kernel sums the partials in fixed order (deterministic).  The otherwise idle epilogue
// warps fold the bias gradient (sum of gz) out of the staged gz rows.
#include <stdlib.h>

#include "tc_common.cuh"

namespace vxm {
namespace tcw {

using namespace vxm::tc;

constexpr int TH = 4, TWR = 32, TUSE = 30;
constexpr int XROWS = (TH + 2) * TWR;                       // 192 voxel rows per x slab
constexpr int GPAD = 16, GROWS = TH * TWR, GSROWS = GROWS + 2 * GPAD;   // gz slab: 16 zero rows, 128 rows, 16 zero rows
constexpr int MAXSLOT = 8, NGS = 4;
constexpr int NLOADER = 128, NTHREADS = 288;   // warps 0-3 bias + final epilogue, 4 MMA issuer, 5-8 loader

struct Wgrad2Args {
  const __nv_bfloat16* x; int Cx, up, upd;   // (B, Dx, Hx, Wx, Cx) bf16, Cx in {8,16,32}; up: nearest x2 (H, W), upd: also D
  const __nv_bfloat16* gz; int Cg;           // (B, D, H, W, Cg) bf16, Cg in {8,16,32}
  float* partial;                            // [grid][T][G][GOUT]
  float* bias_partial;                       // [grid][GOUT] or null
  int B, D, H, W;
  int tiles_h, tiles_w, dchunk, nchunks, nitems, nslot;
  int dbg;      // profiling only (VXM_B200_WGRAD_DBG): 1 = no MMAs issued, 2 = no slab copies, 4 = no bias sums
};

__host__ __device__ inline uint32_t swz(uint32_t off, uint32_t width) { return off ^ (((off >> 7) & (width / 16 - 1)) << 4); }

// MN-major swizzled operand: rows of WIDTH bytes (one voxel, WIDTH/2 channels = one MN atom), 8-row K groups contiguous
// (SBO = 8 * WIDTH), MN atoms `lbo_bytes` apart.  cute make_umma_desc<Major::MN>: LBO = atom stride, SBO = K-group stride.
template <int WIDTH>
__device__ __forceinline__ uint64_t make_desc_mn_swz(uint32_t saddr, uint32_t lbo_bytes, uint32_t boff) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(((8u * WIDTH) >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(boff & 7u) << 49;
  d |= (uint64_t)(WIDTH == 128 ? 2 : (WIDTH == 64 ? 4 : 6)) << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}

// KHM ("kh in M", KD == 1 only): the three kh taps become three MN atoms of the A operand ONE TILE ROW (32 voxel rows) apart —
// the same Toeplitz trick the B operand plays for kw — so a (tile, slice) costs 8 MMAs with one accumulator instead of 24
// with three.  Used for the layers whose kd taps are folded into the channels of one operand (first layer: 2 image planes
// x 3 slices; flow head: 3 flow-gradient planes x 3 slices), i.e. exactly the layers that otherwise pad 2 or 3 real
// channels to 16.
template <int KD, int G, int GOUT, bool KHM = false>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad2_kernel(const Wgrad2Args a) {
  static_assert(!KHM || (KD == 1 && G == 16), "kh-in-M needs a 2-D operand with 16-channel rows");
  constexpr int WA = 2 * G, WG = 2 * GOUT;                  // row bytes
  constexpr uint32_t XSLAB = XROWS * WA, GSLAB = GSROWS * WG;
  constexpr int MM = (KD * G > 64) ? 128 : 64;              // KHM: M = (kh, ci) = 48 -> 64
  constexpr int NN = 3 * GOUT;
  constexpr int NMIRROR = KD == 3 ? 2 : 0, NPADSLAB = KD == 3 ? 1 : 0;
  constexpr int T = KD * 9;
  constexpr int KX = XROWS * (G / 8) / NLOADER, KG = GROWS * (GOUT / 8) / NLOADER;
  static_assert(XROWS * (G / 8) % NLOADER == 0 && GROWS * (GOUT / 8) % NLOADER == 0, "loader tables");
  constexpr uint32_t need_cols = KHM ? NN : (MM == 128 ? 3 * NN : 2 * NN);
  constexpr uint32_t tmem_cols = need_cols <= 128 ? 128u : (need_cols <= 256 ? 256u : 512u);

  extern __shared__ __align__(1024) uint8_t smem[];
  const int NS = a.nslot;
  uint8_t* s_x = smem;
  uint8_t* s_g = s_x + (size_t)(NS + NMIRROR + NPADSLAB) * XSLAB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_g + NGS * GSLAB);
  uint64_t* xfull = bars;
  uint64_t* xempty = bars + MAXSLOT;
  uint64_t* gfull = bars + 2 * MAXSLOT;
  uint64_t* gempty = gfull + NGS;
  uint64_t* done = gempty + NGS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // the gz slabs' pad rows (and everything else there) start as zeros; the loader only ever writes rows GPAD..GPAD+127
  for (uint32_t i = threadIdx.x * 16u; i < NGS * GSLAB; i += NTHREADS * 16u) *reinterpret_cast<uint4*>(s_g + i) = make_uint4(0u, 0u, 0u, 0u);
  if (NPADSLAB)   // the M atom past the kd window reads one more slab: keep it finite
    for (uint32_t i = threadIdx.x * 16u; i < XSLAB; i += NTHREADS * 16u)
      *reinterpret_cast<uint4*>(s_x + (size_t)(NS + NMIRROR) * XSLAB + i) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&xfull[i], NLOADER); mbar_init(&xempty[i], 1); }
    for (int i = 0; i < NGS; ++i) { mbar_init(&gfull[i], NLOADER); mbar_init(&gempty[i], 1 + 128); }
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
    // ================================ LOADER (128 threads) ================================
    const int lt = threadIdx.x - 5 * 32;
    uint32_t xslot = 0, xphase = 1, gslot = 0, gphase = 1;    // producer side: first lap passes on the fresh barriers
    const int Dx = a.upd ? a.D >> 1 : a.D, Hx = a.up ? a.H >> 1 : a.H, Wx = a.up ? a.W >> 1 : a.W;
    const int ncx = a.Cx >> 3, ncg = a.Cg >> 3;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int wt = item % a.tiles_w, ht = (item / a.tiles_w) % a.tiles_h;
      const int ch = (item / HW_tiles) % a.nchunks, b = item / (HW_tiles * a.nchunks);
      const int h0 = ht * TH, w0 = wt * TUSE, d0 = ch * a.dchunk, d1 = min(d0 + a.dchunk, a.D);
      const int s_begin = KD == 3 ? d0 - 1 : d0, s_end = KD == 3 ? d1 + 1 : d1;
      int soff[KX];
      uint32_t doff[KX];
#pragma unroll
      for (int k = 0; k < KX; ++k) {
        const int id = lt + k * NLOADER;
        const int c8 = id % (G / 8), row = id / (G / 8);
        const int h = h0 - 1 + (row >> 5), w = w0 - 1 + (row & 31);
        doff[k] = swz((uint32_t)row * WA + (uint32_t)c8 * 16u, WA);
        soff[k] = -1;
        if (h >= 0 && h < a.H && w >= 0 && w < a.W && c8 < ncx) soff[k] = ((a.up ? h >> 1 : h) * Wx + (a.up ? w >> 1 : w)) * a.Cx + c8 * 8;
      }
      int goff[KG];
      uint32_t gdoff[KG];
#pragma unroll
      for (int k = 0; k < KG; ++k) {
        const int id = lt + k * NLOADER;
        const int c8 = id % (GOUT / 8), row = id / (GOUT / 8);
        const int jj = row & 31, h = h0 + (row >> 5), w = w0 - 1 + jj;
        gdoff[k] = swz((uint32_t)(GPAD + row) * WG + (uint32_t)c8 * 16u, WG);
        goff[k] = -1;
        if (jj >= 1 && jj <= TUSE && h < a.H && w < a.W && c8 < ncg) goff[k] = (h * a.W + w) * a.Cg + c8 * 8;
      }
      for (int ds = s_begin; ds < s_end; ++ds) {
        // ---- x slab of input slice ds ----
        mbar_wait(&xempty[xslot], xphase);
        uint8_t* slab = s_x + (size_t)xslot * XSLAB;
        const bool dok = ds >= 0 && ds < a.D;
        const __nv_bfloat16* base = a.x + (((size_t)b * Dx + (dok ? (a.upd ? ds >> 1 : ds) : 0)) * Hx * Wx) * a.Cx;
#pragma unroll
        for (int k = 0; k < KX; ++k) {
          const bool ok = dok && soff[k] >= 0;
          const __nv_bfloat16* src = ok ? base + soff[k] : a.x;
          if (a.dbg & 2) continue;
          cp_async16(slab + doff[k], src, ok ? 16u : 0u);
          if (NMIRROR && xslot < (uint32_t)NMIRROR) cp_async16(slab + (size_t)NS * XSLAB + doff[k], src, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(&xfull[xslot]);
        if (++xslot == (uint32_t)NS) { xslot = 0; xphase ^= 1; }
        // ---- gz slab of OUTPUT slice dg (the slice whose kd window this x slab completes) ----
        const int dg = KD == 3 ? ds - 1 : ds;
        if (dg >= d0 && dg < d1) {
          mbar_wait(&gempty[gslot], gphase);
          uint8_t* gt = s_g + (size_t)gslot * GSLAB;
          const __nv_bfloat16* baseG = a.gz + (((size_t)b * a.D + dg) * a.H * a.W) * a.Cg;
#pragma unroll
          for (int k = 0; k < KG; ++k) {
            const bool ok = goff[k] >= 0;
            if (a.dbg & 2) continue;
            cp_async16(gt + gdoff[k], ok ? baseG + goff[k] : a.gz, ok ? 16u : 0u);
          }
          cp_async_arrive_noinc(&gfull[gslot]);
          if (++gslot == NGS) { gslot = 0; gphase ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    // ================================ MMA ISSUER (whole warp, one elected lane) ================================
    if (has_work) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);
      const uint32_t x_u32 = smem_u32(s_x), g_u32 = smem_u32(s_g);
      uint32_t wslot = 0, wphase = 0;   // next x slab to wait for
      uint32_t hslot = 0;               // head of the kd window
      uint32_t gs = 0, gph = 0;
      uint32_t acc0 = 0;                // 0 only for the very first (tile, slice) of this CTA
      for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int ch = (item / HW_tiles) % a.nchunks;
        const int nd = min(ch * a.dchunk + a.dchunk, a.D) - ch * a.dchunk;
        for (int j = 0; j < nd; ++j) {
          const int nwait = (KD == 3 && j == 0) ? 3 : 1;
          for (int q = 0; q < nwait; ++q) {
            mbar_wait(&xfull[wslot], wphase);
            if (++wslot == (uint32_t)NS) { wslot = 0; wphase ^= 1; }
          }
          mbar_wait(&gfull[gs], gph);
          tc_fence_after();
          const uint32_t a_start = x_u32 + hslot * XSLAB;
          const uint32_t b_start = g_u32 + gs * GSLAB + (uint32_t)(GPAD - 1) * WG;
          const uint64_t adesc0 = make_desc_mn_swz<WA>(a_start, KHM ? (uint32_t)(TWR * WA) : (KD == 3 ? XSLAB : 0u), 0u);
          const uint64_t bdesc0 = make_desc_mn_swz<WG>(b_start, (uint32_t)WG, 0u);   // base offset 0: the swizzle is a function of the absolute address (verified on B200: the matrix-base-offset form gives wrong sums)
          if (KHM) {
            if (elect_one()) {
#pragma unroll
              for (int i = 0; i < GROWS / 16; ++i) {
                if (a.dbg & 1) break;
                const uint64_t adesc = adesc0 + (uint64_t)(((16 * i) * WA) >> 4);
                const uint64_t bdesc = bdesc0 + (uint64_t)((16 * i * WG) >> 4);
                umma_f16(tmem_base, adesc, bdesc, idesc, i == 0 ? acc0 : 1u);
              }
              umma_commit(&xempty[hslot]);
              umma_commit(&gempty[gs]);
            }
          } else if (elect_one()) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              if (a.dbg & 1) break;
              const uint32_t tmem_d = MM == 128 ? tmem_base + (uint32_t)(kh * NN)
                                                : tmem_base + ((uint32_t)((kh & 1) * 16) << 16) + (uint32_t)((kh >> 1) * NN);
#pragma unroll
              for (int i = 0; i < GROWS / 16; ++i) {
                const uint64_t adesc = adesc0 + (uint64_t)(((kh * TWR + 16 * i) * WA) >> 4);
                const uint64_t bdesc = bdesc0 + (uint64_t)((16 * i * WG) >> 4);
                umma_f16(tmem_d, adesc, bdesc, idesc, i == 0 ? acc0 : 1u);
              }
            }
            umma_commit(&xempty[hslot]);
            umma_commit(&gempty[gs]);
          }
          __syncwarp();
          acc0 = 1u;
          if (++hslot == (uint32_t)NS) hslot = 0;
          if (++gs == NGS) { gs = 0; gph ^= 1; }
        }
        if (KD == 3) {
          if (elect_one()) {
            umma_commit(&xempty[hslot]);
            umma_commit(&xempty[hslot + 1 == (uint32_t)NS ? 0 : hslot + 1]);
          }
          __syncwarp();
          hslot = hslot + 2 >= (uint32_t)NS ? hslot + 2 - NS : hslot + 2;
        }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    }
  } else {
    // ================================ BIAS + FINAL EPILOGUE (warps 0-3) ================================
    // thread r owns gz slab row r: sums its GOUT channels over all staged slabs (halo / out-of-volume rows are zero)
    float bsum[GOUT];
#pragma unroll
    for (int c = 0; c < GOUT; ++c) bsum[c] = 0.f;
    const int rowi = warp * 32 + lane;
    if (has_work && a.bias_partial && !(a.dbg & 4)) {
      uint32_t gs = 0, gph = 0;
      uint32_t roff[GOUT / 8];
#pragma unroll
      for (int c8 = 0; c8 < GOUT / 8; ++c8) roff[c8] = swz((uint32_t)(GPAD + rowi) * WG + (uint32_t)c8 * 16u, WG);
      for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int ch = (item / HW_tiles) % a.nchunks;
        const int nd = min(ch * a.dchunk + a.dchunk, a.D) - ch * a.dchunk;
        for (int j = 0; j < nd; ++j) {
          mbar_wait(&gfull[gs], gph);
          const uint8_t* gt = s_g + (size_t)gs * GSLAB;
#pragma unroll
          for (int c8 = 0; c8 < GOUT / 8; ++c8) {
            const uint4 q = *reinterpret_cast<const uint4*>(gt + roff[c8]);
            const __nv_bfloat162* hq = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(hq[e]);
              bsum[c8 * 8 + 2 * e] += f.x;
              bsum[c8 * 8 + 2 * e + 1] += f.y;
            }
          }
          mbar_arrive(&gempty[gs]);
          if (++gs == NGS) { gs = 0; gph ^= 1; }
        }
      }
    } else if (has_work) {
      // no bias wanted: still release the gz slabs
      uint32_t gs = 0, gph = 0;
      for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int ch = (item / HW_tiles) % a.nchunks;
        const int nd = min(ch * a.dchunk + a.dchunk, a.D) - ch * a.dchunk;
        for (int j = 0; j < nd; ++j) {
          mbar_wait(&gfull[gs], gph);
          mbar_arrive(&gempty[gs]);
          if (++gs == NGS) { gs = 0; gph ^= 1; }
        }
      }
    }
    float* part = a.partial + (size_t)blockIdx.x * T * G * GOUT;
    if (has_work) {
      mbar_wait(done, 0);
      tc_fence_after();
      if (KHM) {
        // accumulator row m = kh * 16 + ci on TMEM lane (m % 16) + 32 * (m / 16): warp = kh, lanes 0..15 = ci; columns q * GOUT + co
        const int kh = warp, ci = lane & 15;
#pragma unroll 1
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int c0 = 0; c0 < GOUT; c0 += 8) {
            uint32_t r[8];
            tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(q * GOUT + c0), r);
            tmem_ld_wait();
            if (kh < 3 && lane < 16) {
              const int tap = kh * 3 + (2 - q);
              float4* o = reinterpret_cast<float4*>(part + ((size_t)tap * G + ci) * GOUT + c0);
              o[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
              o[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
            }
          }
      } else if (MM == 128) {
        // accumulator row m = kd * G + ci on TMEM lane m; columns kh * NN + q * GOUT + co, q = 2 - kw
        const int m = warp * 32 + lane, kd = m / G, ci = m % G;
#pragma unroll 1
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll 1
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int c0 = 0; c0 < GOUT; c0 += 8) {
              uint32_t r[8];
              tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kh * NN + q * GOUT + c0), r);
              tmem_ld_wait();
              if (kd < KD) {
                const int tap = (kd * 3 + kh) * 3 + (2 - q);
                float4* o = reinterpret_cast<float4*>(part + ((size_t)tap * G + ci) * GOUT + c0);
                o[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                o[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
              }
            }
      } else {
        // M = 64: row m on lane (m % 16) + 32 * (m / 16); kh = 0 / 1 share columns [0, NN) at lane offsets 0 / 16, kh = 2 in [NN, 2NN)
        const int m = warp * 16 + (lane & 15), kd = m / G, ci = m % G, half = lane >> 4;
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll 1
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int c0 = 0; c0 < GOUT; c0 += 8) {
              uint32_t r[8];
              tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(blk * NN + q * GOUT + c0), r);
              tmem_ld_wait();
              const int kh = blk == 0 ? half : 2;
              if (kd < KD && !(blk == 1 && half == 1)) {
                const int tap = (kd * 3 + kh) * 3 + (2 - q);
                float4* o = reinterpret_cast<float4*>(part + ((size_t)tap * G + ci) * GOUT + c0);
                o[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                o[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
              }
            }
      }
    } else {
      for (int i = threadIdx.x; i < T * G * GOUT; i += 128) part[i] = 0.f;
    }
    if (a.bias_partial) {
      // bias partial of this CTA: fixed-order reduction over the 128 rows through shared memory (the x ring is idle now)
      float* s_b = reinterpret_cast<float*>(s_x);
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int c = 0; c < GOUT; ++c) s_b[rowi * (GOUT + 1) + c] = bsum[c];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (rowi < GOUT) {
        float t = 0.f;
        for (int r2 = 0; r2 < 128; ++r2) t += s_b[r2 * (GOUT + 1) + rowi];
        a.bias_partial[(size_t)blockIdx.x * GOUT + rowi] = t;
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

// gw[co][ci_off + ci][tap] (+)= sum_cta partial[cta][tap][ci][co]   (fixed order -> deterministic).  Block = 64 elements
// x 4 quarters of the CTA range: threads follow the partial layout (co fastest) so the reads coalesce, and the four
// quarter sums (combined in fixed order through shared memory) keep 4x more loads in flight than one serial loop.
__global__ void __launch_bounds__(256) wgrad2_reduce_kernel(const float* __restrict__ partial, float* __restrict__ gw, int ncta, int T, int G,
                                                            int GOUT, int Cout, int Cin_total, int ci_off, int ci_cnt,
                                                            const float* __restrict__ bias_partial, float* __restrict__ gb, int accumulate) {
  __shared__ float sh[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  if (gb && blockIdx.x == 0 && threadIdx.x < Cout) {
    float acc = accumulate ? gb[threadIdx.x] : 0.f;
    for (int c = 0; c < ncta; ++c) acc += bias_partial[(size_t)c * GOUT + threadIdx.x];
    gb[threadIdx.x] = acc;
  }
  const int per_cta = T * G * GOUT;
  const int j = blockIdx.x * 64 + tx;
  const int q = (ncta + 3) / 4, c0 = ty * q, c1 = min(c0 + q, ncta);
  float acc = 0.f;
  if (j < per_cta)
    for (int c = c0; c < c1; ++c) acc += partial[(size_t)c * per_cta + j];
  sh[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < per_cta) {
    const int co = j % GOUT, ci = (j / GOUT) % G, tap = j / (GOUT * G);
    if (co < Cout && ci < ci_cnt) {
      const float tot = ((sh[0][tx] + sh[1][tx]) + sh[2][tx]) + sh[3][tx];
      float* dst = gw + ((size_t)co * Cin_total + ci_off + ci) * T + tap;
      *dst = (accumulate ? *dst : 0.f) + tot;
    }
  }
}

// All reductions of a backward pass in ONE launch (16 reduce launches per training step otherwise).  The descriptors
// travel by value in the kernel parameters (no device table to keep alive or re-upload inside a captured graph).
struct ReduceDesc {
  const float* partial; float* gw; const float* bias_partial; float* gb;
  int ncta, T, G, GOUT, Cout, Cin_total, ci_off, ci_cnt, accumulate, blk_begin;
};
constexpr int MAXRED = 40;
struct ReduceBatch {
  ReduceDesc d[MAXRED];
  int n, total_blocks;
};
__global__ void __launch_bounds__(256) wgrad2_reduce_multi_kernel(const ReduceBatch rb) {
  __shared__ float sh[4][64];
  int k = 0;
  for (int i = 1; i < rb.n; ++i)
    if ((int)blockIdx.x >= rb.d[i].blk_begin) k = i;
  const ReduceDesc& r = rb.d[k];
  const int blk = blockIdx.x - r.blk_begin;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  if (r.gb && blk == 0 && threadIdx.x < r.Cout) {
    float acc = r.accumulate ? r.gb[threadIdx.x] : 0.f;
    for (int c = 0; c < r.ncta; ++c) acc += r.bias_partial[(size_t)c * r.GOUT + threadIdx.x];
    r.gb[threadIdx.x] = acc;
  }
  const int per_cta = r.T * r.G * r.GOUT;
  const int j = blk * 64 + tx;
  const int q = (r.ncta + 3) / 4, c0 = ty * q, c1 = min(c0 + q, r.ncta);
  float acc = 0.f;
  if (j < per_cta)
    for (int c = c0; c < c1; ++c) acc += r.partial[(size_t)c * per_cta + j];
  sh[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < per_cta) {
    const int co = j % r.GOUT, ci = (j / r.GOUT) % r.G, tap = j / (r.GOUT * r.G);
    if (co < r.Cout && ci < r.ci_cnt) {
      const float tot = ((sh[0][tx] + sh[1][tx]) + sh[2][tx]) + sh[3][tx];
      float* dst = r.gw + ((size_t)co * r.Cin_total + r.ci_off + ci) * r.T + tap;
      *dst = (r.accumulate ? *dst : 0.f) + tot;
    }
  }
}

}  // namespace tcw
}  // namespace vxm

using namespace vxm;
using namespace vxm::tcw;

namespace vxm {
namespace tcw {

static bool chan_ok(int c) { return c == 8 || c == 16 || c == 32; }

bool wgrad2_supported(int Ca, int Cb, int Cg) {
  const char* e = getenv("VXM_B200_WGRAD");
  if (e && e[0] == 'o') return false;   // "old": conv3d_tc_wgrad.cu
  return (Ca == 0 || chan_ok(Ca)) && (Cb == 0 || chan_ok(Cb)) && Ca + Cb > 0 && chan_ok(Cg);
}

// one source tensor (C channels, optionally nearest-x2 upsampled) against gz; weights [ci_off, ci_off + ci_cnt) of Cin_total
int wgrad2_launch(const void* x, int Cx, int up, const void* gz, int Cg, float* grad_w, float* grad_b, void* work, int B, int D, int H, int W,
                  int kd, int Cout_real, int Cin_total, int ci_off, int ci_cnt, int accumulate, cudaStream_t st, ReduceDesc* defer,
                  size_t* work_used, bool khm) {
  Wgrad2Args a{};
  a.x = (const __nv_bfloat16*)x; a.Cx = Cx; a.up = up; a.upd = (up && kd == 3) ? 1 : 0;
  a.gz = (const __nv_bfloat16*)gz; a.Cg = Cg;
  a.B = B; a.D = D; a.H = H; a.W = W;
  a.tiles_h = (H + TH - 1) / TH; a.tiles_w = (W + TUSE - 1) / TUSE;
  {
    const char* de = getenv("VXM_B200_WGRAD_DBG");
    a.dbg = de ? atoi(de) : 0;
  }
  const int G = Cx <= 16 ? 16 : 32, GOUT = Cg <= 16 ? 16 : 32;
  const int nsm = sm_count();
  // depth chunking: balance the persistent CTAs (waves of nsm items) against the 2 halo slabs every chunk re-loads
  const long long tiles = (long long)B * a.tiles_h * a.tiles_w;
  int best_nch = 1;
  double best_cost = 1e300;
  for (int nch = 1; nch <= 32 && nch <= D; ++nch) {
    const int dc = (D + nch - 1) / nch;
    const long long items = tiles * ((D + dc - 1) / dc);
    const long long waves = (items + nsm - 1) / nsm;
    const double cost = (double)waves * (dc + (kd == 3 ? 2.5 : 0.5));
    if (cost < best_cost - 1e-9) { best_cost = cost; best_nch = nch; }
  }
  a.dchunk = (D + best_nch - 1) / best_nch; a.nchunks = (D + a.dchunk - 1) / a.dchunk;
  a.nitems = (int)(tiles * a.nchunks);
  int grid = a.nitems < nsm ? a.nitems : nsm;
  if (grid > 256) grid = 256;
  a.partial = (float*)work;
  if (defer) {
    // deferred reduction: this launch owns exactly [work, work + *work_used): partials, then bias partials
    const size_t npart = (size_t)grid * kd * 9 * G * GOUT;
    a.bias_partial = grad_b ? (float*)work + npart : nullptr;
    *work_used = (npart + (grad_b ? (size_t)grid * GOUT : 0)) * sizeof(float);
    *work_used = (*work_used + 255) & ~(size_t)255;
  } else {
    a.bias_partial = grad_b ? (float*)work + (size_t)256 * kd * 9 * 64 * 32 : nullptr;
  }
  const size_t xslab = (size_t)XROWS * 2 * G, gslab = (size_t)GSROWS * 2 * GOUT;
  const int extra = kd == 3 ? 3 : 0;
  int nslot = (int)((200 * 1024 - NGS * gslab - 512) / xslab) - extra;
  if (nslot > MAXSLOT) nslot = MAXSLOT;
  VXM_REQUIRE(nslot >= 4, "conv3d_tc_wgrad: not enough shared memory for the slab ring");
  a.nslot = nslot;
  const size_t smem = (size_t)(nslot + extra) * xslab + NGS * gslab + 512;
#define VXM_W2_LAUNCH(KD_, G_, GO_)                                                                                          \
  do {                                                                                                                       \
    VXM_CUDA(cudaFuncSetAttribute(wgrad2_kernel<KD_, G_, GO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    wgrad2_kernel<KD_, G_, GO_><<<grid, NTHREADS, smem, st>>>(a);                                                             \
  } while (0)
#define VXM_W2_G(KD_)                                                                                                        \
  do {                                                                                                                       \
    if (G == 16 && GOUT == 16) VXM_W2_LAUNCH(KD_, 16, 16); else if (G == 16) VXM_W2_LAUNCH(KD_, 16, 32);                     \
    else if (GOUT == 16) VXM_W2_LAUNCH(KD_, 32, 16); else VXM_W2_LAUNCH(KD_, 32, 32);                                        \
  } while (0)
  if (khm) {
    VXM_REQUIRE(kd == 1 && G == 16 && GOUT == 16 && !up, "conv3d_tc_wgrad2 (kh in M): needs kd = 1 and at most 16 channels on both sides");
    VXM_CUDA(cudaFuncSetAttribute(wgrad2_kernel<1, 16, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad2_kernel<1, 16, 16, true><<<grid, NTHREADS, smem, st>>>(a);
  } else if (kd == 3) VXM_W2_G(3); else VXM_W2_G(1);
  int rc = check_launch("conv3d_tc_wgrad2");
  if (rc) return rc;
  const int T = kd * 9, per_cta = T * G * GOUT;
  if (defer) {
    defer->partial = a.partial; defer->gw = grad_w; defer->bias_partial = a.bias_partial; defer->gb = grad_b;
    defer->ncta = grid; defer->T = T; defer->G = G; defer->GOUT = GOUT; defer->Cout = Cout_real; defer->Cin_total = Cin_total;
    defer->ci_off = ci_off; defer->ci_cnt = ci_cnt; defer->accumulate = accumulate; defer->blk_begin = (per_cta + 63) / 64;   // block count, turned into an offset by the flush
    return VXM_OK;
  }
  wgrad2_reduce_kernel<<<(per_cta + 63) / 64, 256, 0, st>>>(a.partial, grad_w, grid, T, G, GOUT, Cout_real, Cin_total, ci_off, ci_cnt,
                                                              a.bias_partial, grad_b, accumulate);
  return check_launch("conv3d_tc_wgrad2_reduce");
}

}  // namespace tcw
}  // namespace vxm

// ---- deferred reduction API -------------------------------------------------------------------------------------------
// vxm_conv3d_tc_wgrad2_partial launches the weight-gradient kernel(s) of one layer into caller-provided workspace and
// appends the pending reductions to a HOST descriptor buffer; vxm_conv3d_tc_wgrad2_flush then runs every pending
// reduction of the backward pass in one launch (fixed summation order -> deterministic, as before).
extern "C" size_t vxm_conv3d_tc_wgrad2_desc_bytes(void) { return sizeof(ReduceDesc); }
extern "C" int vxm_conv3d_tc_wgrad2_max_pending(void) { return MAXRED; }
extern "C" size_t vxm_conv3d_tc_wgrad2_partial_bytes(int kd) {
  // upper bound for one layer (two sources): 2 x 256 CTAs x 27 x 32 x 32 floats + bias partials
  return 2 * ((size_t)256 * kd * 9 * 32 * 32 + 256 * 32) * sizeof(float) + 1024;
}

extern "C" int vxm_conv3d_tc_wgrad2_partial(const void* xa, const void* xb, const void* gz, float* grad_w, float* grad_b, void* work,
                                            size_t work_bytes, size_t* work_used, void* descs_host, int* ndesc, int B, int D, int H, int W,
                                            int Ca, int Cb, int up, int Cin_real, int Cg, int Cout_real, int kd, int accumulate, void* stream) {
  VXM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && grad_w && work && work_used && descs_host && ndesc && gz, "conv3d_tc_wgrad2_partial: bad argument");
  VXM_REQUIRE(kd == 1 || kd == 3, "conv3d_tc_wgrad2_partial: kd must be 1 or 3");
  VXM_REQUIRE(wgrad2_supported(Ca, Cb, Cg), "conv3d_tc_wgrad2_partial: channel counts (%d,%d | %d) unsupported", Ca, Cb, Cg);
  VXM_REQUIRE((Ca == 0 || xa) && (Cb == 0 || xb), "conv3d_tc_wgrad2_partial: missing source tensor");
  VXM_REQUIRE(*ndesc + 2 <= MAXRED, "conv3d_tc_wgrad2_partial: too many pending reductions (flush first)");
  VXM_REQUIRE(Cout_real > 0 && Cout_real <= Cg, "conv3d_tc_wgrad2_partial: Cout_real out of range");
  ReduceDesc* d = (ReduceDesc*)descs_host;
  cudaStream_t st = as_stream(stream);
  size_t used = 0, total = 0;
  char* wp = (char*)work;
  if (Ca) {
    const int cnt = Cin_real < Ca ? Cin_real : Ca;
    int rc = wgrad2_launch(xa, Ca, up, gz, Cg, grad_w, grad_b, wp, B, D, H, W, kd, Cout_real, Cin_real, 0, cnt, accumulate, st, &d[*ndesc], &used, false);
    if (rc) return rc;
    VXM_REQUIRE(used <= work_bytes, "conv3d_tc_wgrad2_partial: workspace too small");
    ++*ndesc; wp += used; total += used;
  }
  if (Cb && Cin_real > Ca) {
    const int cnt = Cin_real - Ca < Cb ? Cin_real - Ca : Cb;
    int rc = wgrad2_launch(xb, Cb, 0, gz, Cg, grad_w, Ca ? nullptr : grad_b, wp, B, D, H, W, kd, Cout_real, Cin_real, Ca, cnt, accumulate, st,
                           &d[*ndesc], &used, false);
    if (rc) return rc;
    VXM_REQUIRE(total + used <= work_bytes, "conv3d_tc_wgrad2_partial: workspace too small");
    ++*ndesc; total += used;
  }
  *work_used = total;
  return VXM_OK;
}

// Weight gradient of a kd-folded layer (see vxm_planar_fold_kd_bf16): x (B, D, H, W, Cx <= 16) against gz (B, D, H, W, Cg <= 16) as a
// 2-D problem per slice with the kh taps stacked in M.  grad_w receives the 2-D layout (Cout_real, Cin_real, 1, 3, 3).
extern "C" int vxm_conv3d_tc_wgrad2_partial_khm(const void* x, const void* gz, float* grad_w, float* grad_b, void* work, size_t work_bytes,
                                                size_t* work_used, void* descs_host, int* ndesc, int B, int D, int H, int W, int Cx,
                                                int Cin_real, int Cg, int Cout_real, int accumulate, void* stream) {
  VXM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && x && gz && grad_w && work && work_used && descs_host && ndesc, "conv3d_tc_wgrad2_partial_khm: bad argument");
  VXM_REQUIRE((Cx == 8 || Cx == 16) && (Cg == 8 || Cg == 16), "conv3d_tc_wgrad2_partial_khm: channel counts (%d | %d) unsupported", Cx, Cg);
  VXM_REQUIRE(Cin_real > 0 && Cin_real <= Cx && Cout_real > 0 && Cout_real <= Cg, "conv3d_tc_wgrad2_partial_khm: real channel counts out of range");
  VXM_REQUIRE(*ndesc + 1 <= MAXRED, "conv3d_tc_wgrad2_partial_khm: too many pending reductions (flush first)");
  ReduceDesc* d = (ReduceDesc*)descs_host;
  size_t used = 0;
  int rc = wgrad2_launch(x, Cx, 0, gz, Cg, grad_w, grad_b, work, B, D, H, W, 1, Cout_real, Cin_real, 0, Cin_real, accumulate, as_stream(stream),
                         &d[*ndesc], &used, true);
  if (rc) return rc;
  VXM_REQUIRE(used <= work_bytes, "conv3d_tc_wgrad2_partial_khm: workspace too small");
  ++*ndesc;
  *work_used = used;
  return VXM_OK;
}

extern "C" int vxm_conv3d_tc_wgrad2_flush(const void* descs_host, int ndesc, void* stream) {
  VXM_REQUIRE(descs_host && ndesc > 0 && ndesc <= MAXRED, "conv3d_tc_wgrad2_flush: bad argument");
  ReduceBatch rb;
  const ReduceDesc* d = (const ReduceDesc*)descs_host;
  int blocks = 0;
  for (int i = 0; i < ndesc; ++i) {
    rb.d[i] = d[i];
    const int nb = d[i].blk_begin;      // wgrad2_launch stored the block count here
    rb.d[i].blk_begin = blocks;
    blocks += nb;
  }
  rb.n = ndesc; rb.total_blocks = blocks;
  wgrad2_reduce_multi_kernel<<<blocks, 256, 0, as_stream(stream)>>>(rb);
  return check_launch("conv3d_tc_wgrad2_reduce_multi");
}
