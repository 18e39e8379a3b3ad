// Conv3d k=3 on tcgen05, kw-stacked swizzled formulation of conv3d_tc_s.cu with TWO MMA-issuing warps that alternate whole
// slab steps (8-row tiles).  The default for every layer whose 8-row slab ring fits shared memory (tc.py:_use_s2; +5 % on the
// step over the single issuer, profiles/r2_conv_layers_ab.txt); the A operand is staged by TMA tensor copies, the epilogues
// are specialised per use (EPI), and the kernel carries the profiling switches / clock64 trace that produced
// profiles/r2_conv_ablation.md.
//
// Why two issuers: ncu on the single-issuer kernel (profiles/r1_final_conv_ncu.md) showed the issuer warp busy in its own
// per-step instruction stream (~130 serial instructions: barrier polls, descriptor construction, uniform datapath latencies)
// while the tensor pipe idled 48 % (thin layers) / 25 % (48->32).  Giving each of two issuers every other step doubles the time
// an issuer has per step.  (Splitting one step's two tile halves between two issuers bought nothing: both kept the per-step
// work; issuing the two halves interleaved is slower — round 2, measured.)
//
// Protocol (8-row tiles, 3 or 6 accumulators, 3 epilogue groups as in the single-issuer kernel):
//   * step t (global counter over the CTA's lifetime) belongs to issuer t & 1; it produces events 2t, 2t+1 (tile halves),
//     event e uses accumulator e % NACC (NACC = 3 or 6) and is drained by epilogue group e % 3;
//   * full[slot]  : both issuers walk ALL slabs in order with their own (slot, phase) cursor;
//   * empty[slot] : two arrivals, exactly ONE FROM EACH ISSUER, each made after that issuer has observed the slab's
//                   full phase and after its last MMA that reads the slab was issued: own step j arrives on slabs j and
//                   j+1; the issuer that does not own step 0 arrives on slab 0 at the start of the item; at the end of
//                   the item the would-be owner of step nd arrives on slabs nd and nd+1, the other one on slab nd+1.
//                   (A slot can therefore not be refilled before both cursors have passed it: no parity aliasing.)
//   * tfull[acc]  : one commit per event, consumed phase by phase by the owning epilogue group;
//   * tempty[issuer][acc] : the group that drains event e signals the issuer of event e+NACC (the next user of that
//                   accumulator), which waits for it before issuing it; every barrier has one waiter, consecutive phases.
// Entry point: vxm_conv3d_tcs2_fwd (same signature as vxm_conv3d_tcs_fwd; weights packed by vxm_conv3d_tcs_pack*).
#include <stdlib.h>

#include "tc_common.cuh"

namespace vxm {
namespace tcs2 {

using namespace vxm::tc;

constexpr int WT = 32, WUSE = 30;      // tile: HT (4 or 8) rows x 32 columns (30 written), slab = (HT + 2) x 32 voxel rows
constexpr int MAXSLOT = 16, MAXACC = 6;
constexpr int NLOADER = 64, NTHREADS = 512;   // warps 0-3 epilogue group 0, 4 / 5 MMA issuers (even / odd steps), 6-7 loader, 8-11 / 12-15 epilogue groups 1 / 2

struct ConvSArgs {
  const __nv_bfloat16* xa; const __nv_bfloat16* xb;
  const __nv_bfloat16* wpk; const float* bias;
  void* out; const __nv_bfloat16* mask;
  void* out2; int csplit;   // optional second bf16 output: channels [csplit, Cout) (single-pass dgrad of a concat layer)
  int B, D, H, W, Ca, Cb, up, upd, Cout, out_mode;
  float slope;
  int tiles_h, tiles_w, dchunk, nchunks, nitems, nslot;
  uint32_t wbytes;
  // TMA tile staging of the A operand (see conv3d_tc_s.cu): bit g of tma_mask = channel group g of every slab arrives as one
  // cp.async.bulk.tensor.5d of tensor map tm[g] starting at channel tc0[g]; zero padding = the copy's out-of-bounds fill
  int tma_mask, tc0[2];
  // ablation switches for profiling only (VXM_B200_TCS_DBG, never set in production): 1 = no MMAs issued, 2 = no TMEM
  // read-out, 4 = no global stores / mask loads, 8 = no slab copies
  int dbg;
  int nacc;       // TMEM accumulators in flight: 3 or 6
  // profiling aid (VXM_B200_TCS_TRACE = device address of a uint64 buffer, never set in production): CTA 0 records
  // clock64() stamps, trace[role * 4096 + n * 8 + k] for its n-th step / slab / tile (n < 512) at point k of role
  // 0 / 1 = MMA issuers, 2 = slab producer, 3 / 4 / 5 = epilogue groups (their first warp)
  unsigned long long* trace;
  alignas(64) CUtensorMap tm[2];
};

// byte offset inside a swizzled K-major tile whose rows are `width` bytes (32, 64 or 128): Swizzle<log2(width/16),4,3>
__host__ __device__ inline uint32_t swz(uint32_t off, uint32_t width) {
  return off ^ (((off >> 7) & (width / 16 - 1)) << 4);
}
__device__ __forceinline__ uint64_t make_desc_kmajor_swz(uint32_t saddr, uint32_t width) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                                   // LBO: unused for K inside one swizzle atom
  d |= (uint64_t)(((8u * width) >> 4) & 0x3FFF) << 32;      // SBO: 8 rows
  d |= (uint64_t)1 << 46;                                   // descriptor version 1
  d |= (uint64_t)(width == 128 ? 2 : (width == 64 ? 4 : 6)) << 61;   // SWIZZLE_128B / 64B / 32B
  return d;
}

// HT = 8: one slab step feeds TWO 4-row accumulators (one per epilogue group), halving the per-step issue / barrier
// overhead that bounds the thin layers and cutting the halo re-reads from 1.5x to 1.25x.
// EPI: epilogue specialisation.  The generic epilogue (0) decides bias / activation / mask / output layout / channel split
// per tile at run time and costs ~430 issued instructions per warp and tile: at one instruction per clock and scheduler the
// three epilogue warps of a scheduler then bound the thin layers (ablation: the barrier skeleton alone ran 92 us of the
// 180 us of the 16->16 layer, profiles/r2_conv_ablation.md).  1 = forward (bias + LeakyReLU, 0 <= slope <= 1, bf16
// channels-last, all COUT channels real), 2 = dgrad (LeakyReLU derivative from the saved activation, no bias): straight-line
// code with the pointers advanced incrementally.
template <int KD, int G0, int G1, int COUT, int HT, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tcs2_kernel(const __grid_constant__ ConvSArgs a) {
  static_assert(HT == 8, "alternating issuers are written for 8-row tiles");
  constexpr int SROWS = (HT + 2) * WT;
  constexpr int NH = HT / 4;
  constexpr int W0 = G0 * 2, W1 = G1 * 2;                 // row bytes of the two channel groups
  constexpr int NC8 = (G0 + G1) / 8;                      // 16-byte chunks per voxel
  constexpr uint32_t SLAB0 = SROWS * W0, SLAB1 = SROWS * W1;
  constexpr int NN = 3 * COUT;   // MMA N: (kw, co)
  // TMEM accumulators in flight = epilogue groups in use: group g owns accumulator g, so every mbarrier is waited on
  // phase by phase (a group that skipped ahead on a barrier would alias its parity)
  // Accumulators in flight: 3 (group g owns accumulator g) or 6 (group g owns g and g + 3, alternately).  The round trip
  // commit -> epilogue wake-up -> TMEM read -> release -> issuer wake-up measured ~1300 clk even with the MMAs and the
  // read-out switched off (profiles/r2_conv_ablation.md): with 3 accumulators that latency, not the tensor pipe, paces the
  // 16-channel layers.  A group still never skips a phase of a barrier it waits on (it owns its accumulators).
  const int NACC = a.nacc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const bool halfk = (a.Ca + a.Cb == 8);                  // 8 real channels in a 16-channel group: chunk 1 is zero-filled
  constexpr uint32_t slab_bytes = SLAB0 + SLAB1;
  const int NSLOT = a.nslot;
  uint8_t* s_w = smem;
  uint8_t* s_slab = smem + ((a.wbytes + 1023u) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_slab + NSLOT * slab_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAXSLOT;
  uint64_t* tfull = bars + 2 * MAXSLOT;
  uint64_t* tempty = tfull + MAXACC;
  uint64_t* wbar = tempty + 2 * MAXACC;   // tempty[issuer * MAXACC + accumulator]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = NACC * NN <= 128 ? 128u : (NACC * NN <= 256 ? 256u : 512u);

  const bool tma0 = a.tma_mask & 1, tma1 = (a.tma_mask & 2) != 0;
  const bool all_tma = tma0 && (G1 == 0 || tma1);        // no cp.async traffic at all: one producer thread
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], all_tma ? 1 : NLOADER); mbar_init(&empty[i], 2); }   // one arrival per issuer
    for (int i = 0; i < NACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); mbar_init(&tempty[MAXACC + i], 4); }   // one arrival per epilogue warp
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    mbar_expect_tx(wbar, a.wbytes);
    for (uint32_t off = 0; off < a.wbytes; off += 16384u) {
      uint32_t n = a.wbytes - off < 16384u ? a.wbytes - off : 16384u;
      bulk_g2s(s_w + off, reinterpret_cast<const uint8_t*>(a.wpk) + off, n, wbar);
    }
  }
  const int HW_tiles = a.tiles_h * a.tiles_w;
  unsigned long long* const trace = blockIdx.x == 0 ? a.trace : nullptr;
#define VXM_TR(role_, n_, k_)                                                                             \
  do {                                                                                                    \
    if (trace && (n_) < 512u && lane == 0) trace[(role_) * 4096 + (n_) * 8 + (k_)] = clock64();           \
  } while (0)

  if (warp == 6 || warp == 7) {
    // ================================ LOADER (64 threads) ================================
    const int lt = threadIdx.x - 6 * 32;
    uint32_t slot = 0, lphase = 1;   // producer side: the first lap passes on the fresh barriers
    uint32_t ltr = 0;
    const int Da = a.upd ? a.D >> 1 : a.D, Ha = a.up ? a.H >> 1 : a.H, Wa = a.up ? a.W >> 1 : a.W;
    const int nca8 = a.Ca >> 3;
    constexpr int nchunk = NC8 * SROWS;
    constexpr int KMAX = (nchunk + NLOADER - 1) / NLOADER;
    if (lt == 0) {
      if (tma0) tma_prefetch_desc(&a.tm[0]);
      if (tma1) tma_prefetch_desc(&a.tm[1]);
    }
    const uint32_t tma_bytes = (tma0 ? SLAB0 : 0u) + (tma1 ? SLAB1 : 0u);
    const bool nocopy = a.dbg & 8;
    for (int item = blockIdx.x; item < a.nitems && !(all_tma && lt != 0); item += gridDim.x) {
      const int wt = item % a.tiles_w, ht = (item / a.tiles_w) % a.tiles_h;
      const int ch = (item / HW_tiles) % a.nchunks, b = item / (HW_tiles * a.nchunks);
      const int h0 = ht * HT, w0 = wt * WUSE, d0 = ch * a.dchunk, d1 = min(d0 + a.dchunk, a.D);
      const int s_begin = KD == 3 ? d0 - 1 : d0, s_end = KD == 3 ? d1 + 1 : d1;
      int soff[KMAX];
      uint32_t doff[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int id = lt + k * NLOADER;
        soff[k] = -1;
        doff[k] = 0;
        if (id < nchunk) {
          const int c8 = id % NC8, row = id / NC8;
          const int r = row >> 5, c = row & 31;
          const int h = h0 - 1 + r, w = w0 - 1 + c;
          doff[k] = c8 < G0 / 8 ? swz((uint32_t)row * W0 + (uint32_t)c8 * 16u, W0)
                                : SLAB0 + swz((uint32_t)row * W1 + (uint32_t)(c8 - G0 / 8) * 16u, W1 ? W1 : 32);
          if (c8 < G0 / 8 ? tma0 : tma1) soff[k] = -2;      // this chunk's group arrives by tensor copy
          else if (h >= 0 && h < a.H && w >= 0 && w < a.W && !(halfk && c8 > 0)) {
            if (c8 < nca8) soff[k] = (((a.up ? h >> 1 : h) * Wa + (a.up ? w >> 1 : w)) * a.Ca + c8 * 8) << 1;
            else soff[k] = (((h * a.W + w) * a.Cb + (c8 - nca8) * 8) << 1) | 1;
          }
        }
      }
      for (int ds = s_begin; ds < s_end; ++ds) {
        if (warp == 6) VXM_TR(2, ltr, 0);
        mbar_wait(&empty[slot], lphase);
        if (warp == 6) VXM_TR(2, ltr, 1);
        uint8_t* slab = s_slab + (size_t)slot * slab_bytes;
        const bool dok = ds >= 0 && ds < a.D;
        const __nv_bfloat16* baseA = a.xa ? a.xa + (((size_t)b * Da + (dok ? (a.upd ? ds >> 1 : ds) : 0)) * Ha * Wa) * a.Ca : nullptr;
        const __nv_bfloat16* baseB = a.xb ? a.xb + (((size_t)b * a.D + (dok ? ds : 0)) * a.H * a.W) * a.Cb : nullptr;
        const __nv_bfloat16* dummy = a.xa ? a.xa : a.xb;
        if (lt == 0 && tma_bytes) {
          if (nocopy) { if (all_tma) mbar_arrive(&full[slot]); }
          else {
            if (all_tma) mbar_expect_tx(&full[slot], tma_bytes);
            else mbar_expect_tx_noarrive(&full[slot], tma_bytes);
            if (tma0) tma_load_5d(slab, &a.tm[0], a.tc0[0], w0 - 1, h0 - 1, ds, b, &full[slot]);
            if (tma1) tma_load_5d(slab + SLAB0, &a.tm[1], a.tc0[1], w0 - 1, h0 - 1, ds, b, &full[slot]);
          }
        }
        if (!all_tma) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (lt + k * NLOADER < nchunk && soff[k] != -2 && !nocopy) {
              const bool ok = dok && soff[k] >= 0;
              const __nv_bfloat16* src = ok ? ((soff[k] & 1) ? baseB : baseA) + (soff[k] >> 1) : dummy;
              cp_async16(slab + doff[k], src, ok ? 16u : 0u);
            }
          }
          cp_async_arrive_noinc(&full[slot]);
        }
        if (warp == 6) VXM_TR(2, ltr, 2);
        ++ltr;
        if (++slot == (uint32_t)NSLOT) { slot = 0; lphase ^= 1; }
      }
    }
  } else if (warp == 4 || warp == 5) {
    // ================================ MMA ISSUERS (whole warp each, one elected lane) ================================
    const uint32_t me = (uint32_t)(warp - 4);             // owns the slab steps with (global step & 1) == me
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t slab_u32 = smem_u32(s_slab), w_u32 = smem_u32(s_w);
    constexpr uint32_t WSTEP = (uint32_t)NN * (W0 + W1);          // bytes of packed weights per (kd, kh) step
    mbar_wait(wbar, 0);
    const uint64_t bdesc0 = make_desc_kmajor_swz(w_u32, W0);
    const uint64_t bdesc1 = make_desc_kmajor_swz(w_u32 + NN * W0, W1 ? W1 : 32);
    uint32_t wslot = 0, wphase = 0, wcur = 0;   // full-barrier cursor: ring slot / phase / global index of the next slab to observe
    uint32_t hslot = 0;                         // ring slot of the current step's first slab
    uint32_t gbase = 0;                         // global index of the current item's first slab
    uint32_t t = 0;                             // global step counter
    uint32_t acc0 = 0;                          // (2 t) % NACC: accumulator of the step's first tile half
    uint32_t phbits = 0;                        // per accumulator: phase of this issuer's tempty barrier
    auto observe = [&](uint32_t upto) {         // wait for every slab up to global index `upto`, in order
      while (wcur <= upto) {
        mbar_wait(&full[wslot], wphase);
        if (++wslot == (uint32_t)NSLOT) { wslot = 0; wphase ^= 1; }
        ++wcur;
      }
    };
    auto next_slot = [&](uint32_t s) { return s + 1 == (uint32_t)NSLOT ? 0u : s + 1; };
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int ch = (item / HW_tiles) % a.nchunks;
      const int d0 = ch * a.dchunk, d1 = min(d0 + a.dchunk, a.D);
      const int nd = d1 - d0;
      if (KD == 3 && (t & 1u) != me) {          // step 0 of this item is the other issuer's: release slab 0 on my behalf
        observe(gbase);
        if (elect_one()) umma_commit(&empty[hslot]);
        __syncwarp();
      }
      for (int j = 0; j < nd; ++j) {
        const bool mine = (t & 1u) == me;
        if (mine) {
          VXM_TR(me, t >> 1, 0);
          observe(gbase + (uint32_t)j + (KD == 3 ? 2u : 0u));
          VXM_TR(me, t >> 1, 1);
          tc_fence_after();
          uint64_t adesc0_kd[KD], adesc1_kd[KD];
          {
            uint32_t sl = hslot;
#pragma unroll
            for (int kd = 0; kd < KD; ++kd) {
              adesc0_kd[kd] = make_desc_kmajor_swz(slab_u32 + sl * slab_bytes, W0);
              adesc1_kd[kd] = make_desc_kmajor_swz(slab_u32 + sl * slab_bytes + SLAB0, W1 ? W1 : 32);
              sl = next_slot(sl);
            }
          }
#pragma unroll
          for (int hb = 0; hb < NH; ++hb) {
            const uint32_t acc = acc0 + hb >= (uint32_t)NACC ? acc0 + hb - NACC : acc0 + hb;
            if (2u * t + hb >= (uint32_t)NACC) {     // event 2t+hb-NACC used this accumulator: its drain is signalled to me
              mbar_wait(&tempty[me * MAXACC + acc], (phbits >> acc) & 1u);
              phbits ^= 1u << acc;
            }
            VXM_TR(me, t >> 1, 2 + 2 * hb);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * (uint32_t)NN;
            if (elect_one()) {
              if (!(a.dbg & 1)) {
#pragma unroll
              for (int kd = 0; kd < KD; ++kd) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                  const int st = kd * 3 + kh;
#pragma unroll
                  for (int k = 0; k < G0 / 16; ++k) {
                    const uint64_t adesc = adesc0_kd[kd] + (uint64_t)(((hb * 4 + kh) * WT * W0 + k * 32) >> 4);
                    const uint64_t bdesc = bdesc0 + (uint64_t)((st * WSTEP + k * 32) >> 4);
                    umma_f16(tmem_d, adesc, bdesc, idesc, (st | k) ? 1u : 0u);
                  }
#pragma unroll
                  for (int k = 0; k < G1 / 16; ++k) {
                    const uint64_t adesc = adesc1_kd[kd] + (uint64_t)(((hb * 4 + kh) * WT * W1 + k * 32) >> 4);
                    const uint64_t bdesc = bdesc1 + (uint64_t)((st * WSTEP + k * 32) >> 4);
                    umma_f16(tmem_d, adesc, bdesc, idesc, 1u);
                  }
                }
              }
              }
              umma_commit(&tfull[acc]);
              if (hb == NH - 1) {
                umma_commit(&empty[hslot]);                              // slab j: my last reader of it is this step
                if (KD == 3) umma_commit(&empty[next_slot(hslot)]);      // slab j+1: likewise
              }
            }
            __syncwarp();
            VXM_TR(me, t >> 1, 3 + 2 * hb);
          }
        } else if (KD == 1) {
          observe(gbase + (uint32_t)j);           // 2-D: a slab has one reader; the other issuer still observes and releases it
          if (elect_one()) umma_commit(&empty[hslot]);
          __syncwarp();
        }
        hslot = next_slot(hslot);
        ++t;
        acc0 += 2;
        if (acc0 >= (uint32_t)NACC) acc0 -= NACC;
      }
      if (KD == 3) {
        // hslot is slab nd's slot now.  The would-be owner of step nd releases slabs nd and nd+1, the other issuer slab nd+1.
        observe(gbase + (uint32_t)nd + 1u);
        const uint32_t h1 = next_slot(hslot);
        if (elect_one()) {
          if ((t & 1u) == me) umma_commit(&empty[hslot]);
          umma_commit(&empty[h1]);
        }
        __syncwarp();
        hslot = next_slot(h1);
        gbase += (uint32_t)nd + 2u;
      } else {
        gbase += (uint32_t)nd;
      }
    }
  } else {
    // ================================ EPILOGUE (3 groups x 4 warps; warp = tile row hh, lane = w') ==================
    // Group g drains the accumulators with (accumulator counter % 3) == g: the per-tile epilogue is a ~1300-cycle
    // dependent chain (TMEM load, 32 shuffles, bias / activation / mask, pack, store), so three tiles are kept in
    // flight; with two groups the epilogue, not the tensor pipe, bounds the thin layers (profiles/r1_*).
    const int grp = warp >= 12 ? 2 : (warp >= 8 ? 1 : 0);
    uint32_t turn = 0, tphase = 0;   // accumulator counter % NACC; phase of this group's tfull barrier
    uint32_t ev = 0;                 // global event counter (2 * step + tile half)
    const int wq = warp & 3;
    if constexpr (EPI != 0) {
      const float slope = a.slope;
      [[maybe_unused]] float bs[COUT];
      if constexpr (EPI == 1) {
#pragma unroll
        for (int c = 0; c < COUT; ++c) bs[c] = a.bias ? __ldg(a.bias + c) : 0.f;
      }
      const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
      uint32_t etr = 0;
      const size_t slice = (size_t)a.H * a.W * COUT;            // elements per output slice
      const int c1 = (EPI == 3 && a.out2) ? a.csplit : COUT;    // EPI 3: channels [0, c1) -> out, [c1, COUT) -> out2
      for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int wt = item % a.tiles_w, ht = (item / a.tiles_w) % a.tiles_h;
        const int ch = (item / HW_tiles) % a.nchunks, b = item / (HW_tiles * a.nchunks);
        const int w = wt * WUSE - 1 + lane, d0 = ch * a.dchunk, d1 = min(d0 + a.dchunk, a.D);
        const bool wok = lane >= 1 && lane <= WUSE && w < a.W;
        bool ok[NH];
        size_t off[NH];                                          // element offset of this lane's voxel in slice d0
#pragma unroll
        for (int hb = 0; hb < NH; ++hb) {
          const int h = ht * HT + hb * 4 + wq;
          ok[hb] = wok && h < a.H;
          off[hb] = ((((size_t)b * a.D + d0) * a.H + h) * a.W + w) * COUT;
        }
        for (int d = d0; d < d1; ++d) {
#pragma unroll
          for (int hb = 0; hb < NH; ++hb) {
            const uint32_t acc = turn;                           // accumulator of this event: (event counter) % NACC
            const bool mine = (int)turn == grp || (int)turn == grp + 3;
            if (++turn == (uint32_t)NACC) turn = 0;
            ++ev;
            const size_t o = off[hb];
            off[hb] += slice;
            if (!mine) continue;
            const uint32_t taddr = tlane + acc * (uint32_t)NN;
            const bool valid = ok[hb];
            [[maybe_unused]] uint32_t mreg[COUT / 16][8];
            if constexpr (EPI == 2) {
              if (valid) {
#pragma unroll
                for (int q = 0; q < COUT / 16; ++q) ld_global_nc_v8(a.mask + o + q * 16, mreg[q]);
              }
            }
            if (wq == 0) VXM_TR(3 + grp, etr, 0);
            mbar_wait(&tfull[acc], (tphase >> acc) & 1u);
            if (wq == 0) VXM_TR(3 + grp, etr, 1);
            tphase ^= 1u << acc;
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
              uint32_t r0[16], r1[16], r2[16];
              tmem_ld16(taddr + c0, r0);
              tmem_ld16(taddr + COUT + c0, r1);
              tmem_ld16(taddr + 2 * COUT + c0, r2);
              tmem_ld_wait();
              if (c0 + 16 >= COUT) {          // last TMEM read of this accumulator: hand it to the issuer of event e + NACC
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[((((ev - 1u) + (uint32_t)NACC) >> 1) & 1u) * MAXACC + acc]);
                if (wq == 0) VXM_TR(3 + grp, etr, 2);
              }
              float v[16];
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float p0 = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[c]), 1);
                const float p2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[c]), 1);
                v[c] = (p0 + __uint_as_float(r1[c])) + p2;      // out[w'] = P0[w'-1] + P1[w'] + P2[w'+1]
              }
              if constexpr (EPI == 1) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  const float x = v[c] + bs[c0 + c];
                  v[c] = fmaxf(x, x * slope);                    // LeakyReLU for 0 <= slope <= 1
                }
              } else if constexpr (EPI == 2) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {                    // sign bits of the saved bf16 activations
                  const uint32_t mw = mreg[c0 / 16][e];
                  if (mw & 0x8000u) v[2 * e] *= slope;
                  if (mw & 0x80000000u) v[2 * e + 1] *= slope;
                }
              }
              if (valid) {
                // o = voxel index * COUT; EPI 3 splits the channels between two tensors of c1 and COUT - c1 channels
                __nv_bfloat16* op = (EPI == 3 && c0 >= c1) ? reinterpret_cast<__nv_bfloat16*>(a.out2) + o / COUT * (COUT - c1) + (c0 - c1)
                                                           : reinterpret_cast<__nv_bfloat16*>(a.out) + (EPI == 3 ? o / COUT * c1 : o) + c0;
                st_global_v8(op, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]),
                             pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
              }
            }
            if (wq == 0) VXM_TR(3 + grp, etr, 3);
            ++etr;
          }
        }
      }
    } else {
    const size_t HWp = (size_t)a.H * a.W;
    constexpr int NBR = COUT <= 32 ? COUT : 1;     // bias kept in registers for the (forward) layer widths
    float biasr[NBR];
#pragma unroll
    for (int c = 0; c < NBR; ++c) biasr[c] = (a.bias && c < a.Cout) ? __ldg(a.bias + c) : 0.f;
    auto bias_at = [&](int c) -> float { return COUT <= 32 ? biasr[COUT <= 32 ? c : 0] : (a.bias ? __ldg(a.bias + c) : 0.f); };
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int wt = item % a.tiles_w, ht = (item / a.tiles_w) % a.tiles_h;
      const int ch = (item / HW_tiles) % a.nchunks, b = item / (HW_tiles * a.nchunks);
      const int w = wt * WUSE - 1 + lane, d0 = ch * a.dchunk, d1 = min(d0 + a.dchunk, a.D);
      for (int d = d0; d < d1; ++d) {
#pragma unroll
        for (int hb = 0; hb < NH; ++hb) {
        uint32_t eacc;
        {
          eacc = turn;
          const bool mine = (int)turn == grp || (int)turn == grp + 3;
          if (++turn == (uint32_t)NACC) turn = 0;
          ++ev;
          if (!mine) continue;
        }
        const int h = ht * HT + hb * 4 + wq;
        const bool valid = lane >= 1 && lane <= WUSE && h < a.H && w < a.W;
        const uint32_t acc = eacc;
        const size_t vox = (((size_t)b * a.D + d) * a.H + h) * a.W + w;
        // prefetch the LeakyReLU-derivative mask of this voxel before waiting for the tensor core
        uint4 mreg[COUT / 8];
        if (a.mask && valid && !(a.dbg & 4)) {
#pragma unroll
          for (int q = 0; q < COUT / 8; ++q)
            if (q * 8 < a.Cout) mreg[q] = __ldg(reinterpret_cast<const uint4*>(a.mask + vox * a.Cout) + q);
        }
        mbar_wait(&tfull[acc], (tphase >> acc) & 1u);
        tphase ^= 1u << acc;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * (uint32_t)NN;
        const int c1 = a.out2 ? a.csplit : a.Cout;          // channels [0,c1) -> out, [c1,Cout) -> out2
        // 16 output channels at a time: 3 x 16 TMEM columns (kw = 0,1,2), shuffle-combine across lanes, store
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16) {
          uint32_t r0[16], r1[16], r2[16];
          if (a.dbg & 2) {
#pragma unroll
            for (int c = 0; c < 16; ++c) r0[c] = r1[c] = r2[c] = 0u;
          } else {
            tmem_ld16(taddr + c0, r0);
            tmem_ld16(taddr + COUT + c0, r1);
            tmem_ld16(taddr + 2 * COUT + c0, r2);
            tmem_ld_wait();
          }
          if (c0 + 16 >= COUT) {          // last TMEM read of this accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[((((ev - 1u) + (uint32_t)NACC) >> 1) & 1u) * MAXACC + acc]);   // issuer of event e + NACC, e = ev - 1
          }
          float v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float p0 = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[c]), 1);
            const float p2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[c]), 1);
            v[c] = (p0 + __uint_as_float(r1[c])) + p2;      // out[w'] = P0[w'-1] + P1[w'] + P2[w'+1]
          }
          if (valid && c0 < a.Cout && !(a.dbg & 4)) {
            if (a.out_mode == 0) {
#pragma unroll
              for (int q = 0; q < 16; q += 8) {
                if (c0 + q < a.Cout) {
                  float x[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) x[e] = v[q + e] + bias_at(c0 + q + e);
                  if (a.mask) {
                    const uint4 m4 = mreg[(c0 + q) / 8];
                    const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&m4);
#pragma unroll
                    for (int e = 0; e < 8; ++e) if (__bfloat162float(mb[e]) < 0.f) x[e] *= a.slope;
                  } else if (a.slope >= 0.f) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = x[e] >= 0.f ? x[e] : x[e] * a.slope;
                  }
                  // a split never falls inside a group of 8 channels (csplit % 8 == 0)
                  const int cg = c0 + q;
                  __nv_bfloat16* oo = cg < c1 ? reinterpret_cast<__nv_bfloat16*>(a.out) + vox * c1 + cg
                                              : reinterpret_cast<__nv_bfloat16*>(a.out2) + vox * (a.Cout - c1) + (cg - c1);
                  *reinterpret_cast<uint4*>(oo) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                }
              }
            } else {
              float* o = reinterpret_cast<float*>(a.out);
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                if (c0 + c < a.Cout) {
                  float x = v[c] + bias_at(c0 + c);
                  if (a.slope >= 0.f) x = x >= 0.f ? x : x * a.slope;
                  o[(((size_t)b * a.Cout + c0 + c) * a.D + d) * HWp + (size_t)h * a.W + w] = x;
                }
              }
            }
          }
        }
        }
      }
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

}  // namespace tcs2
}  // namespace vxm

using namespace vxm;
using namespace vxm::tcs2;

// Same arguments as vxm_conv3d_tcs_fwd (conv3d_tc_s.cu); weights packed by vxm_conv3d_tcs_pack.  8-row tiles only:
// (Ca + Cb) in {8, 16, 32, 48 = 32 + 16}, padded Cout in {16, 32}.
extern "C" int vxm_conv3d_tcs2_fwd(const void* xa, const void* xb, const void* wpk, const float* bias, void* out, const void* mask,
                                   int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int coutp, int kd, int out_mode,
                                   float slope, void* out2, int csplit, void* stream) {
  VXM_REQUIRE(B > 0 && D > 0 && H > 4 && W > 0 && wpk && out, "conv3d_tcs2_fwd: bad argument");
  VXM_REQUIRE(kd == 1 || kd == 3, "conv3d_tcs2_fwd: kd must be 1 or 3");
  VXM_REQUIRE(coutp == 16 || coutp == 32 || (coutp == 48 && Ca + Cb == 32), "conv3d_tcs2_fwd: padded Cout must be 16 or 32 (48 for 32 input channels)");
  const int cin = Ca + Cb;
  VXM_REQUIRE(cin == 8 || cin == 16 || cin == 32 || (Ca == 32 && Cb == 16), "conv3d_tcs2_fwd: channel counts (%d,%d) unsupported", Ca, Cb);
  ConvSArgs a{};
  const int g0 = cin <= 16 ? 16 : 32, g1 = cin == 48 ? 16 : 0;
  a.xa = (const __nv_bfloat16*)xa; a.xb = (const __nv_bfloat16*)xb; a.wpk = (const __nv_bfloat16*)wpk; a.bias = bias;
  a.out = out; a.mask = (const __nv_bfloat16*)mask; a.out2 = out2; a.csplit = csplit;
  a.B = B; a.D = D; a.H = H; a.W = W; a.Ca = Ca; a.Cb = Cb; a.up = up; a.upd = (up && kd == 3) ? 1 : 0;
  a.Cout = Cout; a.out_mode = out_mode; a.slope = slope;
  a.wbytes = (uint32_t)((size_t)kd * 3 * (3 * coutp) * (g0 + g1) * sizeof(__nv_bfloat16));
  const size_t fixed = ((a.wbytes + 1023u) & ~1023u) + 1024 + 512;
  const size_t slab = (size_t)10 * WT * (g0 + g1) * 2;
  int nslot = (int)((227 * 1024 - fixed) / slab);
  {
    // ring depth: more slabs in flight hide the L2 / HBM latency of the tensor copies (VXM_B200_RING=8: A/B switch)
    const char* e = getenv("VXM_B200_RING");
    const int cap = e ? atoi(e) : MAXSLOT;
    if (nslot > cap) nslot = cap;
    if (nslot > MAXSLOT) nslot = MAXSLOT;
  }
  VXM_REQUIRE(nslot >= (kd == 3 ? 5 : 3), "conv3d_tcs2_fwd: not enough shared memory for 8-row slabs");
  a.nslot = nslot;
  a.tiles_h = (H + 7) / 8; a.tiles_w = (W + WUSE - 1) / WUSE;
  const int nsm = sm_count();
  const long long tiles = (long long)B * a.tiles_h * a.tiles_w;
  int best_nch = 1;
  double best_cost = 1e300;
  for (int nch = 1; nch <= 40 && nch <= D; ++nch) {
    const int dc = (D + nch - 1) / nch;
    const long long items = tiles * ((D + dc - 1) / dc);
    const double cost = (double)((items + nsm - 1) / nsm) * (dc + (kd == 3 ? 2.5 : 0.5));
    if (cost < best_cost - 1e-9) { best_cost = cost; best_nch = nch; }
  }
  a.dchunk = (D + best_nch - 1) / best_nch; a.nchunks = (D + a.dchunk - 1) / a.dchunk;
  a.nitems = (int)(tiles * a.nchunks);
  const size_t smem = fixed + (size_t)nslot * slab;
  const int grid = a.nitems < nsm ? a.nitems : nsm;
  cudaStream_t st = as_stream(stream);
  {
    const char* e = getenv("VXM_B200_TCS_DBG");
    a.dbg = e ? atoi(e) : 0;
    const char* tr = getenv("VXM_B200_TCS_TRACE");
    a.trace = tr ? (unsigned long long*)strtoull(tr, nullptr, 10) : nullptr;
    const char* na = getenv("VXM_B200_TCS_NACC");     // "3": A/B switch
    a.nacc = (6 * 3 * coutp <= 512 && !(na && na[0] == '3')) ? 6 : 3;
    // TMA plan: a channel group whose channels all come from one source tensor read at its own resolution (VXM_B200_TMA=0: none)
    a.tma_mask = 0; a.tc0[0] = a.tc0[1] = 0;
    const char* t = getenv("VXM_B200_TMA");
    if (!(t && t[0] == '0')) {
      const int lo[2] = {0, g0}, hi[2] = {g0, g0 + g1};
      for (int g = 0; g < 2; ++g) {
        if (hi[g] == lo[g]) continue;
        const void* base = nullptr;
        int C = 0, c0 = 0;
        if (a.xa && hi[g] <= Ca && !up) { base = xa; C = Ca; c0 = lo[g]; }
        else if (a.xb && lo[g] >= Ca && hi[g] <= Ca + Cb) { base = xb; C = Cb; c0 = lo[g] - Ca; }
        if (!base) continue;
        if (tc::make_act_tmap(&a.tm[g], base, B, D, H, W, C, hi[g] - lo[g], WT, 10) != 0) return VXM_ERR_CUDA;
        a.tma_mask |= 1 << g;
        a.tc0[g] = c0;
      }
    }
  }
  // epilogue specialisation (see the kernel): 3-D, bf16 channels-last output with all padded channels real, no channel split
  int epi = 0;
  {
    const char* e = getenv("VXM_B200_TCS_EPI");       // "0": generic epilogue everywhere (A/B switch)
    const bool plain = out_mode == 0 && Cout == coutp && !(e && e[0] == '0');
    if (plain && !out2 && !mask && slope >= 0.f && slope <= 1.f) epi = 1;
    else if (plain && !out2 && mask && !bias) epi = 2;
    else if (plain && !mask && !bias && slope < 0.f && (!out2 || csplit % 16 == 0)) epi = 3;
  }
#define VXM_TCS2_LAUNCH_E(KD_, G0_, G1_, CO_, E_)                                                                              \
  do {                                                                                                                        \
    VXM_CUDA(cudaFuncSetAttribute(conv_tcs2_kernel<KD_, G0_, G1_, CO_, 8, E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    conv_tcs2_kernel<KD_, G0_, G1_, CO_, 8, E_><<<grid, NTHREADS, smem, st>>>(a);                                               \
  } while (0)
#define VXM_TCS2_LAUNCH(KD_, G0_, G1_, CO_)                                                                                    \
  do {                                                                                                                        \
    if (epi == 1) VXM_TCS2_LAUNCH_E(KD_, G0_, G1_, CO_, 1);                                                                    \
    else if (epi == 2) VXM_TCS2_LAUNCH_E(KD_, G0_, G1_, CO_, 2);                                                               \
    else if (epi == 3) VXM_TCS2_LAUNCH_E(KD_, G0_, G1_, CO_, 3);                                                               \
    else VXM_TCS2_LAUNCH_E(KD_, G0_, G1_, CO_, 0);                                                                             \
  } while (0)
#define VXM_TCS2_G(KD_, CO_)                                                                                                  \
  do {                                                                                                                        \
    if (g0 == 16) VXM_TCS2_LAUNCH(KD_, 16, 0, CO_); else if (g1 == 0) VXM_TCS2_LAUNCH(KD_, 32, 0, CO_); else VXM_TCS2_LAUNCH(KD_, 32, 16, CO_); \
  } while (0)
  if (coutp == 48) { if (kd == 3) VXM_TCS2_LAUNCH(3, 32, 0, 48); else VXM_TCS2_LAUNCH(1, 32, 0, 48); }
  else if (kd == 3) { if (coutp == 16) VXM_TCS2_G(3, 16); else VXM_TCS2_G(3, 32); }
  else { if (coutp == 16) VXM_TCS2_G(1, 16); else VXM_TCS2_G(1, 32); }
  return check_launch("conv3d_tcs2_fwd");
}
