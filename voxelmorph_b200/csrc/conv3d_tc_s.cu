// Conv3d k=3 on tcgen05, "kw-stacked" formulation with 128/64/32-byte SWIZZLED K-major operands.
//
// Same algorithm as conv3d_tc_t.cu (N = 3*Cout stacks the kw taps, K loop over (kd, kh, Cin/16), kw shift by warp
// shuffle in the epilogue), but the shared-memory operands use the UMMA swizzled canonical layouts instead of
// SWIZZLE_NONE: every slab row is one voxel with all channels of a channel group contiguous (32 / 64 / 128 bytes)
// and the 16-byte chunks of a row are XOR-swizzled with the row index (Swizzle<B,4,3>), which the cp.async loader
// applies to its destination addresses and the weight packer applies on the host side of the operand.  The input
// channels are split into at most two groups of 16 / 32 / 64 channels (48 = 32 + 16), each with its own slab region.
// A (kd, kh) tap shifts the A operand by whole 32-voxel rows (a multiple of the 8-row swizzle period), so the start
// address stays pattern-aligned and the descriptor base offset is 0.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace vxm {
namespace tcs {

using namespace vxm::tc;

constexpr int WT = 32, WUSE = 30;      // tile: HT (4 or 8) rows x 32 columns (30 written), slab = (HT + 2) x 32 voxel rows
constexpr int MAXSLOT = 16, MAXACC = 4;
constexpr int NLOADER = 96, NTHREADS = 512, NGRP = 3;   // warps 0-3 epilogue group 0, 4 MMA issuer, 5-7 loader, 8-11 / 12-15 epilogue groups 1 / 2

struct ConvSArgs {
  const __nv_bfloat16* xa; const __nv_bfloat16* xb;
  const __nv_bfloat16* wpk; const float* bias;
  void* out; const __nv_bfloat16* mask;
  void* out2; int csplit;   // optional second bf16 output: channels [csplit, Cout) (single-pass dgrad of a concat layer)
  // split-precision (bf16x3) passes: `acc_in` (fp32 channels-last, COUT channels per voxel) is added to the tile before the
  // epilogue; out_mode 2 stores the raw fp32 sums back in that layout (no bias / activation), out_mode 3 applies bias +
  // activation and stores the result as a bf16 (hi, lo) pair: hi -> out, lo = bf16(x - hi) -> out_lo
  int B, D, H, W, Ca, Cb, up, upd, Cout, out_mode;
  float slope;
  int tiles_h, tiles_w, dchunk, nchunks, nitems, nslot;
  uint32_t wbytes;
  const float* acc_in; void* out_lo;
  // TMA tile staging of the A operand: bit g of tma_mask = channel group g of every slab arrives as ONE tiled tensor copy
  // (cp.async.bulk.tensor.5d, box {G channels, 32 columns, HT + 2 rows}) of tensor map tm[g], starting at channel tc0[g];
  // zero padding is the copy's out-of-bounds fill.  Groups read through the nearest-neighbour upsampler (decoder concat
  // layers), the 64-channel group that interleaves two sources and the 8-plane image input stay on the cp.async loader.
  int tma_mask, tc0[2];
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
// ACC: the split-precision epilogue (acc_in / out_mode 2, 3) is compiled in; the plain kernels (ACC = false) keep the
// round-1 epilogue — with the extra live registers the 32-channel variants spilled and lost up to 1.8x.
// EPI: epilogue specialisation (see conv3d_tc_s2.cu).  0 = generic run-time epilogue; 1 = forward (bias + LeakyReLU with
// 0 <= slope <= 1, bf16 channels-last, all COUT channels real); 2 = dgrad (LeakyReLU derivative from the saved activation);
// 3 = raw sums with an optional channel split at a multiple of 16 (single-pass dgrad of a concat layer).
template <int KD, int G0, int G1, int COUT, int HT, bool ACC, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tcs_kernel(const __grid_constant__ ConvSArgs a) {
  constexpr int SROWS = (HT + 2) * WT;
  constexpr int NH = HT / 4;
  constexpr int W0 = G0 * 2, W1 = G1 * 2;                 // row bytes of the two channel groups
  constexpr int NC8 = (G0 + G1) / 8;                      // 16-byte chunks per voxel
  constexpr uint32_t SLAB0 = SROWS * W0, SLAB1 = SROWS * W1;
  constexpr int NN = 3 * COUT;   // MMA N: (kw, co)
  // TMEM accumulators in flight = epilogue groups in use: group g owns accumulator g, so every mbarrier is waited on
  // phase by phase (a group that skipped ahead on a barrier would alias its parity)
  constexpr int NACC = (3 * NN <= 512) ? 3 : 2;
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
  uint64_t* wbar = tempty + MAXACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t tmem_cols = NACC * NN <= 128 ? 128u : (NACC * NN <= 256 ? 256u : 512u);

  const bool tma0 = a.tma_mask & 1, tma1 = (a.tma_mask & 2) != 0;
  const bool all_tma = tma0 && (G1 == 0 || tma1);        // no cp.async traffic at all: one producer thread
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], all_tma ? 1 : NLOADER); mbar_init(&empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }     // one arrival per epilogue warp
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

  if (warp >= 5 && warp < 8) {
    // ================================ LOADER (96 threads) ================================
    const int lt = threadIdx.x - 5 * 32;
    uint32_t slot = 0, lphase = 1;   // producer side: the first lap passes on the fresh barriers
    const int Da = a.upd ? a.D >> 1 : a.D, Ha = a.up ? a.H >> 1 : a.H, Wa = a.up ? a.W >> 1 : a.W;
    const int nca8 = a.Ca >> 3;
    constexpr int nchunk = NC8 * SROWS;
    constexpr int KMAX = (nchunk + NLOADER - 1) / NLOADER;
    if (lt == 0) {
      if (tma0) tma_prefetch_desc(&a.tm[0]);
      if (tma1) tma_prefetch_desc(&a.tm[1]);
    }
    const uint32_t tma_bytes = (tma0 ? SLAB0 : 0u) + (tma1 ? SLAB1 : 0u);
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
        mbar_wait(&empty[slot], lphase);
        uint8_t* slab = s_slab + (size_t)slot * slab_bytes;
        const bool dok = ds >= 0 && ds < a.D;
        const __nv_bfloat16* baseA = a.xa ? a.xa + (((size_t)b * Da + (dok ? (a.upd ? ds >> 1 : ds) : 0)) * Ha * Wa) * a.Ca : nullptr;
        const __nv_bfloat16* baseB = a.xb ? a.xb + (((size_t)b * a.D + (dok ? ds : 0)) * a.H * a.W) * a.Cb : nullptr;
        const __nv_bfloat16* dummy = a.xa ? a.xa : a.xb;
        if (lt == 0 && tma_bytes) {
          if (all_tma) mbar_expect_tx(&full[slot], tma_bytes);
          else mbar_expect_tx_noarrive(&full[slot], tma_bytes);
          if (tma0) tma_load_5d(slab, &a.tm[0], a.tc0[0], w0 - 1, h0 - 1, ds, b, &full[slot]);
          if (tma1) tma_load_5d(slab + SLAB0, &a.tm[1], a.tc0[1], w0 - 1, h0 - 1, ds, b, &full[slot]);
        }
        if (!all_tma) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (lt + k * NLOADER < nchunk && soff[k] != -2) {
              const bool ok = dok && soff[k] >= 0;
              const __nv_bfloat16* src = ok ? ((soff[k] & 1) ? baseB : baseA) + (soff[k] >> 1) : dummy;
              cp_async16(slab + doff[k], src, ok ? 16u : 0u);
            }
          }
          cp_async_arrive_noinc(&full[slot]);
        }
        if (++slot == (uint32_t)NSLOT) { slot = 0; lphase ^= 1; }
      }
    }
  } else if (warp == 4) {
    // ================================ MMA ISSUER (whole warp, one elected lane) ================================
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t slab_u32 = smem_u32(s_slab), w_u32 = smem_u32(s_w);
    constexpr uint32_t WSTEP = (uint32_t)NN * (W0 + W1);          // bytes of packed weights per (kd, kh) step
    mbar_wait(wbar, 0);
    const uint64_t bdesc0 = make_desc_kmajor_swz(w_u32, W0);
    const uint64_t bdesc1 = make_desc_kmajor_swz(w_u32 + NN * W0, W1 ? W1 : 32);
    // slot / phase bookkeeping is incremental (no runtime modulo: the issue loop is the critical path of the thin layers)
    uint32_t wslot = 0, wphase = 0;     // next slab to wait for
    uint32_t hslot = 0;                 // head of the kd window
    uint32_t acc = 0, aphase = 1;       // accumulator ring (consumer of tempty: first lap passes)
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int ch = (item / HW_tiles) % a.nchunks;
      const int d0 = ch * a.dchunk, d1 = min(d0 + a.dchunk, a.D);
      const int nd = d1 - d0;
      for (int j = 0; j < nd; ++j) {
        const int nwait = (KD == 3 && j == 0) ? 3 : 1;
        for (int q = 0; q < nwait; ++q) {
          mbar_wait(&full[wslot], wphase);
          if (++wslot == (uint32_t)NSLOT) { wslot = 0; wphase ^= 1; }
        }
        uint64_t adesc0_kd[KD], adesc1_kd[KD];
        {
          uint32_t sl = hslot;
#pragma unroll
          for (int kd = 0; kd < KD; ++kd) {
            adesc0_kd[kd] = make_desc_kmajor_swz(slab_u32 + sl * slab_bytes, W0);
            adesc1_kd[kd] = make_desc_kmajor_swz(slab_u32 + sl * slab_bytes + SLAB0, W1 ? W1 : 32);
            if (++sl == (uint32_t)NSLOT) sl = 0;
          }
        }
#pragma unroll
        for (int hb = 0; hb < NH; ++hb) {
        mbar_wait(&tempty[acc], aphase);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * (uint32_t)NN;
        if (elect_one()) {
#pragma unroll
          for (int kd = 0; kd < KD; ++kd) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              {
                const int st = kd * 3 + kh;
#pragma unroll
                for (int k = 0; k < G0 / 16; ++k) {      // start-address field is in 16-byte units: kh rows, 32 B per K step
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
          if (hb == NH - 1) umma_commit(&empty[hslot]);
        }
        __syncwarp();
        if (++acc == (uint32_t)NACC) { acc = 0; aphase ^= 1; }
        }
        if (++hslot == (uint32_t)NSLOT) hslot = 0;
      }
      if (KD == 3) {
        const uint32_t h1 = hslot + 1 == (uint32_t)NSLOT ? 0 : hslot + 1;
        if (elect_one()) {
          umma_commit(&empty[hslot]);
          umma_commit(&empty[h1]);
        }
        __syncwarp();
        hslot = h1 + 1 == (uint32_t)NSLOT ? 0 : h1 + 1;
      }
    }
  } else {
    // ================================ EPILOGUE (3 groups x 4 warps; warp = tile row hh, lane = w') ==================
    // Group g drains the accumulators with (accumulator counter % 3) == g: the per-tile epilogue is a ~1300-cycle
    // dependent chain (TMEM load, 32 shuffles, bias / activation / mask, pack, store), so three tiles are kept in
    // flight; with two groups the epilogue, not the tensor pipe, bounds the thin layers (profiles/r1_*).
    const int grp = warp >= 12 ? 2 : (warp >= 8 ? 1 : 0);
    uint32_t turn = 0, tphase = 0;   // accumulator counter % NACC; phase of this group's tfull barrier
    const int wq = warp & 3;
    if constexpr (EPI != 0) {
      const float slope = a.slope;
      [[maybe_unused]] float bs[EPI == 1 ? COUT : 1];
      if constexpr (EPI == 1) {
#pragma unroll
        for (int c = 0; c < COUT; ++c) bs[c] = a.bias ? __ldg(a.bias + c) : 0.f;
      }
      const uint32_t acc = (uint32_t)grp;
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * (uint32_t)NN;
      const int c1 = (EPI == 3 && a.out2) ? a.csplit : COUT;       // channels [0, c1) -> out, [c1, COUT) -> out2
      const size_t HWp = (size_t)a.H * a.W;
      for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        const int wt = item % a.tiles_w, ht = (item / a.tiles_w) % a.tiles_h;
        const int ch = (item / HW_tiles) % a.nchunks, b = item / (HW_tiles * a.nchunks);
        const int w = wt * WUSE - 1 + lane, d0 = ch * a.dchunk, d1 = min(d0 + a.dchunk, a.D);
        const bool wok = lane >= 1 && lane <= WUSE && w < a.W;
        bool ok[NH];
        size_t vx[NH];                                           // voxel index of this lane in slice d0
#pragma unroll
        for (int hb = 0; hb < NH; ++hb) {
          const int h = ht * HT + hb * 4 + wq;
          ok[hb] = wok && h < a.H;
          vx[hb] = (((size_t)b * a.D + d0) * a.H + h) * a.W + w;
        }
        for (int d = d0; d < d1; ++d) {
#pragma unroll
          for (int hb = 0; hb < NH; ++hb) {
            const bool mine = (int)turn == grp;
            if (++turn == (uint32_t)NACC) turn = 0;
            const size_t vox = vx[hb];
            vx[hb] += HWp;
            if (!mine) continue;
            const bool valid = ok[hb];
            [[maybe_unused]] uint32_t mreg[EPI == 2 ? COUT / 16 : 1][8];
            if constexpr (EPI == 2) {
              if (valid) {
#pragma unroll
                for (int q = 0; q < COUT / 16; ++q) ld_global_nc_v8(a.mask + vox * COUT + q * 16, mreg[q]);
              }
            }
            mbar_wait(&tfull[acc], tphase);
            tphase ^= 1;
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
              uint32_t r0[16], r1[16], r2[16];
              tmem_ld16(taddr + c0, r0);
              tmem_ld16(taddr + COUT + c0, r1);
              tmem_ld16(taddr + 2 * COUT + c0, r2);
              tmem_ld_wait();
              if (c0 + 16 >= COUT) {          // last TMEM read of this accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
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
                __nv_bfloat16* dst = (EPI == 3 && c0 >= c1) ? reinterpret_cast<__nv_bfloat16*>(a.out2) + vox * (COUT - c1) + (c0 - c1)
                                                            : reinterpret_cast<__nv_bfloat16*>(a.out) + vox * c1 + c0;
                uint4* op = reinterpret_cast<uint4*>(dst);
                st_global_v8(op, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]),
                             pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
              }
            }
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
        {
          const bool mine = (int)turn == grp;
          if (++turn == (uint32_t)NACC) turn = 0;
          if (!mine) continue;
        }
        const int h = ht * HT + hb * 4 + wq;
        const bool valid = lane >= 1 && lane <= WUSE && h < a.H && w < a.W;
        const uint32_t acc = (uint32_t)grp;
        const size_t vox = (((size_t)b * a.D + d) * a.H + h) * a.W + w;
        // prefetch the LeakyReLU-derivative mask of this voxel before waiting for the tensor core
        uint4 mreg[COUT / 8];
        if (a.mask && valid) {
#pragma unroll
          for (int q = 0; q < COUT / 8; ++q)
            if (q * 8 < a.Cout) mreg[q] = __ldg(reinterpret_cast<const uint4*>(a.mask + vox * a.Cout) + q);
        }
        mbar_wait(&tfull[acc], tphase);
        tphase ^= 1;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * (uint32_t)NN;
        const int c1 = a.out2 ? a.csplit : a.Cout;          // channels [0,c1) -> out, [c1,Cout) -> out2
        // 16 output channels at a time: 3 x 16 TMEM columns (kw = 0,1,2), shuffle-combine across lanes, store
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16) {
          uint32_t r0[16], r1[16], r2[16];
          [[maybe_unused]] float4 ain[ACC ? 4 : 1];
          if constexpr (ACC) {
            if (a.acc_in && valid) {      // partial sums of the earlier split-precision passes (issued before the TMEM wait)
              const float4* ap = reinterpret_cast<const float4*>(a.acc_in + vox * COUT + c0);
#pragma unroll
              for (int q = 0; q < 4; ++q) ain[q] = __ldg(ap + q);
            }
          }
          tmem_ld16(taddr + c0, r0);
          tmem_ld16(taddr + COUT + c0, r1);
          tmem_ld16(taddr + 2 * COUT + c0, r2);
          tmem_ld_wait();
          if (c0 + 16 >= COUT) {          // last TMEM read of this accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          float v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float p0 = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[c]), 1);
            const float p2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[c]), 1);
            v[c] = (p0 + __uint_as_float(r1[c])) + p2;      // out[w'] = P0[w'-1] + P1[w'] + P2[w'+1]
          }
          bool handled = false;
          if constexpr (ACC) {
            if (a.acc_in && valid) {
#pragma unroll
              for (int q = 0; q < 4; ++q) { v[4 * q] += ain[q].x; v[4 * q + 1] += ain[q].y; v[4 * q + 2] += ain[q].z; v[4 * q + 3] += ain[q].w; }
            }
            if (a.out_mode == 2) {
              handled = true;
              if (valid) {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + vox * COUT + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              }
            } else if (a.out_mode == 3) {
              handled = true;
              if (valid && c0 < a.Cout) {
#pragma unroll
                for (int q = 0; q < 16; q += 8) {
                  if (c0 + q < a.Cout) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                      float x0 = v[q + e] + bias_at(c0 + q + e), x1 = v[q + e + 1] + bias_at(c0 + q + e + 1);
                      if (a.slope >= 0.f) { x0 = x0 >= 0.f ? x0 : x0 * a.slope; x1 = x1 >= 0.f ? x1 : x1 * a.slope; }
                      const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                      hi[e >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                      lo[e >> 1] = pack_bf16x2(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
                    }
                    const size_t o = vox * a.Cout + c0 + q;
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out_lo) + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                  }
                }
              }
            }
          }
          if (!handled && valid && c0 < a.Cout) {
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

// fp32 (Cout, Cin, KD, 3, 3) -> bf16, per (kd,kh) step two swizzled K-major tiles [N rows = kw*COUT + co][G0 ch] and [N][G1 ch]
__global__ void pack_weights_s_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin, int KD, int COUT,
                                      int NN, int G0, int G1, int transposed) {
  const int T = KD * 9;
  const int CG = G0 + G1;
  const int total = KD * 3 * NN * CG;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % CG, n = (i / CG) % NN, st = i / (CG * NN);
    const int kd = st / 3, kh = st % 3, g = n / COUT, co = n % COUT;
    const int tap = (kd * 3 + kh) * 3 + g;
    float v = 0.f;
    if (!transposed) {
      if (co < Cout && ci < Cin) v = w[((size_t)co * Cin + ci) * T + tap];
    } else {
      if (co < Cin && ci < Cout) v = w[((size_t)ci * Cin + co) * T + (T - 1 - tap)];
    }
    // destination: step tile = [group 0: NN x G0][group 1: NN x G1], bytes swizzled inside each group tile
    const uint32_t W0 = G0 * 2, W1 = G1 * 2;
    uint32_t off;
    if (ci < G0) off = swz((uint32_t)n * W0 + (uint32_t)ci * 2u, W0);
    else off = (uint32_t)NN * W0 + swz((uint32_t)n * W1 + (uint32_t)(ci - G0) * 2u, W1);
    out[((size_t)st * NN * (W0 + W1) + off) / 2] = __float2bfloat16_rn(v);
  }
}

// All packed operands of a model in ONE launch (23 pack launches per training step otherwise: forward + transposed copy
// of every layer, a few kilobytes each).  `descs` lives in device memory; element ranges are consecutive.
struct PackDesc {
  const float* w;
  __nv_bfloat16* out;
  int Cout, Cin, KD, COUT, NN, G0, G1, transposed;
  int begin, count;      // global element range [begin, begin + count)
  // fold > 0: `w` is a 3-D (Cout, Cin, 3, 3, 3) weight whose kd taps are folded into the input channels of a 2-D operand
  // (KD = 1 here): operand input channel kd * fold + c <-> (tap kd, real channel c), fold = real channel count of the
  // operand's input side (Cin forward, Cout transposed)
  int fold;
};
__global__ void pack_weights_multi_kernel(const PackDesc* __restrict__ descs, int ndesc, int total) {
  for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += gridDim.x * blockDim.x) {
    int lo = 0, hi = ndesc - 1;
    while (lo < hi) {                         // last descriptor whose begin <= gi
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].begin <= gi) lo = mid; else hi = mid - 1;
    }
    const PackDesc d = descs[lo];
    const int i = gi - d.begin;
    const int T = d.KD * 9, CG = d.G0 + d.G1;
    const int ci = i % CG, n = (i / CG) % d.NN, st = i / (CG * d.NN);
    const int kd = st / 3, kh = st % 3, g = n / d.COUT, co = n % d.COUT;
    const int tap = (kd * 3 + kh) * 3 + g;
    float v = 0.f;
    if (d.fold) {
      const int fk = ci / d.fold, c = ci % d.fold;              // folded kd tap, real input channel of the operand
      const int tap3 = (fk * 3 + kh) * 3 + g;
      if (fk < 3) {
        if (!d.transposed) { if (co < d.Cout) v = d.w[((size_t)co * d.Cin + c) * 27 + tap3]; }
        else if (co < d.Cin) v = d.w[((size_t)c * d.Cin + co) * 27 + (26 - tap3)];
      }
    } else if (!d.transposed) {
      if (co < d.Cout && ci < d.Cin) v = d.w[((size_t)co * d.Cin + ci) * T + tap];
    } else {
      if (co < d.Cin && ci < d.Cout) v = d.w[((size_t)ci * d.Cin + co) * T + (T - 1 - tap)];
    }
    const uint32_t W0 = d.G0 * 2, W1 = d.G1 * 2;
    uint32_t off;
    if (ci < d.G0) off = swz((uint32_t)n * W0 + (uint32_t)ci * 2u, W0);
    else off = (uint32_t)d.NN * W0 + swz((uint32_t)n * W1 + (uint32_t)(ci - d.G0) * 2u, W1);
    d.out[((size_t)st * d.NN * (W0 + W1) + off) / 2] = __float2bfloat16_rn(v);
  }
}

// channel grouping of a convolution input of `cin` channels (8 counts as a zero-padded 16)
static void groups_of(int cin, int* g0, int* g1) {
  if (cin == 8 || cin == 16) { *g0 = 16; *g1 = 0; }
  else if (cin == 32) { *g0 = 32; *g1 = 0; }
  else if (cin == 48) { *g0 = 32; *g1 = 16; }
  else { *g0 = 64; *g1 = 0; }
}

// Which channel groups of the slab can be staged by tiled tensor copies: a group whose channels all come from one source
// tensor that is read at its own resolution.  VXM_B200_TMA=0 keeps every group on the cp.async loader (A/B switch).
static int plan_tma(ConvSArgs& a, int g0, int g1, int HTv) {
  a.tma_mask = 0;
  a.tc0[0] = a.tc0[1] = 0;
  const char* e = getenv("VXM_B200_TMA");
  if (e && e[0] == '0') return 0;
  const int lo[2] = {0, g0}, hi[2] = {g0, g0 + g1};
  for (int g = 0; g < 2; ++g) {
    if (hi[g] == lo[g]) continue;
    const void* base = nullptr;
    int C = 0, c0 = 0;
    if (a.xa && hi[g] <= a.Ca && !a.up) { base = a.xa; C = a.Ca; c0 = lo[g]; }
    else if (a.xb && lo[g] >= a.Ca && hi[g] <= a.Ca + a.Cb) { base = a.xb; C = a.Cb; c0 = lo[g] - a.Ca; }
    if (!base) continue;
    if (tc::make_act_tmap(&a.tm[g], base, a.B, a.D, a.H, a.W, C, hi[g] - lo[g], WT, HTv + 2) != 0) return -1;
    a.tma_mask |= 1 << g;
    a.tc0[g] = c0;
  }
  return 0;
}

}  // namespace tcs
}  // namespace vxm

using namespace vxm;
using namespace vxm::tcs;

extern "C" size_t vxm_conv3d_tcs_packed_bytes(int cin_eff, int coutp, int kd) {
  int g0, g1;
  groups_of(cin_eff <= 8 ? 8 : (cin_eff <= 16 ? 16 : (cin_eff <= 32 ? 32 : (cin_eff <= 48 ? 48 : 64))), &g0, &g1);
  return (size_t)kd * 3 * (3 * coutp) * (g0 + g1) * sizeof(__nv_bfloat16);
}

extern "C" int vxm_conv3d_tcs_pack(const float* w, void* wpk, int Cout, int Cin, int kd, int coutp, int transposed, void* stream) {
  VXM_REQUIRE(w && wpk && Cout > 0 && Cin > 0 && (kd == 1 || kd == 3) && (coutp == 16 || coutp == 32 || coutp == 48 || coutp == 64), "conv3d_tcs_pack: bad argument");
  int cin_eff = transposed ? Cout : Cin, nout = transposed ? Cin : Cout;
  VXM_REQUIRE(nout <= coutp, "conv3d_tcs_pack: %d output channels do not fit %d", nout, coutp);
  VXM_REQUIRE(cin_eff <= 64, "conv3d_tcs_pack: at most 64 input channels");
  int g0, g1;
  groups_of(cin_eff <= 8 ? 8 : (cin_eff <= 16 ? 16 : (cin_eff <= 32 ? 32 : (cin_eff <= 48 ? 48 : 64))), &g0, &g1);
  int NN = 3 * coutp;
  int total = kd * 3 * NN * (g0 + g1);
  pack_weights_s_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(w, (__nv_bfloat16*)wpk, Cout, Cin, kd, coutp, NN, g0, g1, transposed);
  return check_launch("conv3d_tcs_pack");
}

extern "C" size_t vxm_conv3d_tcs_pack_desc_bytes(void) { return sizeof(PackDesc); }

// Fill one host-side descriptor (the caller uploads the array once and keeps it while the pointers stay valid).
extern "C" int vxm_conv3d_tcs_pack_desc(void* desc_host, const float* w, void* wpk, int Cout, int Cin, int kd, int coutp, int transposed,
                                        int begin) {
  VXM_REQUIRE(desc_host && w && wpk && Cout > 0 && Cin > 0 && (kd == 1 || kd == 3) && (coutp == 16 || coutp == 32 || coutp == 48 || coutp == 64),
              "conv3d_tcs_pack_desc: bad argument");
  const int cin_eff = transposed ? Cout : Cin, nout = transposed ? Cin : Cout;
  VXM_REQUIRE(nout <= coutp && cin_eff <= 64, "conv3d_tcs_pack_desc: channel counts do not fit");
  int g0, g1;
  groups_of(cin_eff <= 8 ? 8 : (cin_eff <= 16 ? 16 : (cin_eff <= 32 ? 32 : (cin_eff <= 48 ? 48 : 64))), &g0, &g1);
  PackDesc d;
  d.w = w; d.out = (__nv_bfloat16*)wpk; d.Cout = Cout; d.Cin = Cin; d.KD = kd; d.COUT = coutp; d.NN = 3 * coutp; d.G0 = g0; d.G1 = g1;
  d.transposed = transposed; d.begin = begin; d.count = kd * 3 * d.NN * (g0 + g1); d.fold = 0;
  memcpy(desc_host, &d, sizeof(d));
  return d.count;      // elements of this operand (>= 0), so the caller can chain `begin`
}

// Descriptor of a kd-folded 2-D operand of the 3-D weight w (Cout, Cin, 3, 3, 3): the operand has 3 * (Cin | Cout if
// transposed) input channels (padded to 8 / 16) and kd = 1; packed size = vxm_conv3d_tcs_packed_bytes(3 * that, coutp, 1).
extern "C" int vxm_conv3d_tcs_pack_desc_fold(void* desc_host, const float* w, void* wpk, int Cout, int Cin, int coutp, int transposed, int begin) {
  VXM_REQUIRE(desc_host && w && wpk && Cout > 0 && Cin > 0 && (coutp == 16 || coutp == 32), "conv3d_tcs_pack_desc_fold: bad argument");
  const int real_in = transposed ? Cout : Cin, nout = transposed ? Cin : Cout;
  VXM_REQUIRE(3 * real_in <= 16 && nout <= coutp, "conv3d_tcs_pack_desc_fold: %d x 3 input channels / %d outputs do not fit", real_in, nout);
  int g0, g1;
  groups_of(3 * real_in <= 8 ? 8 : 16, &g0, &g1);
  PackDesc d;
  d.w = w; d.out = (__nv_bfloat16*)wpk; d.Cout = Cout; d.Cin = Cin; d.KD = 1; d.COUT = coutp; d.NN = 3 * coutp; d.G0 = g0; d.G1 = g1;
  d.transposed = transposed; d.begin = begin; d.count = 3 * d.NN * (g0 + g1); d.fold = real_in;
  memcpy(desc_host, &d, sizeof(d));
  return d.count;
}

extern "C" int vxm_conv3d_tcs_pack_multi(const void* descs_dev, int ndesc, int total, void* stream) {
  VXM_REQUIRE(descs_dev && ndesc > 0 && total > 0, "conv3d_tcs_pack_multi: bad argument");
  const int blocks = (total + 255) / 256, cap = sm_count() * 8;
  pack_weights_multi_kernel<<<blocks < cap ? blocks : cap, 256, 0, as_stream(stream)>>>((const PackDesc*)descs_dev, ndesc, total);
  return check_launch("conv3d_tcs_pack_multi");
}

extern "C" int vxm_conv3d_tcs_supported(int Ca, int Cb, int Cout) {
  int cin = Ca + Cb;
  bool split_ok = cin != 48 || (Ca == 32 && Cb == 16) || Cb == 0 || Ca == 0;   // 48 = 32 + 16 channel groups
  return (Cout <= 64) && (cin == 8 || cin == 16 || cin == 32 || cin == 48 || cin == 64) && Ca % 8 == 0 && Cb % 8 == 0 && split_ok;
}

static int conv_tcs_launch(const void* xa, const void* xb, const void* wpk, const float* bias, void* out, const void* mask,
                           int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int coutp, int kd, int out_mode,
                           float slope, void* out2, int csplit, const float* acc_in, void* out_lo, void* stream);

extern "C" int vxm_conv3d_tcs_fwd(const void* xa, const void* xb, const void* wpk, const float* bias, void* out, const void* mask,
                                  int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int coutp, int kd, int out_mode,
                                  float slope, void* out2, int csplit, void* stream) {
  VXM_REQUIRE(out_mode == 0 || out_mode == 1, "conv3d_tcs_fwd: out_mode must be 0 (bf16 channels-last) or 1 (fp32 planar)");
  return conv_tcs_launch(xa, xb, wpk, bias, out, mask, B, D, H, W, Ca, Cb, up, Cout, coutp, kd, out_mode, slope, out2, csplit,
                         nullptr, nullptr, stream);
}

extern "C" int vxm_conv3d_tcs_fwd_acc(const void* xa, const void* xb, const void* wpk, const float* bias, void* out, void* out_lo,
                                      const float* acc_in, int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int coutp,
                                      int kd, int out_mode, float slope, void* stream) {
  VXM_REQUIRE(out_mode >= 1 && out_mode <= 3, "conv3d_tcs_fwd_acc: out_mode must be 1 (fp32 planar), 2 (fp32 partial sums) or 3 (bf16 hi/lo pair)");
  VXM_REQUIRE(out_mode != 3 || out_lo, "conv3d_tcs_fwd_acc: out_mode 3 needs out_lo");
  return conv_tcs_launch(xa, xb, wpk, bias, out, nullptr, B, D, H, W, Ca, Cb, up, Cout, coutp, kd, out_mode, slope, nullptr, 0,
                         acc_in, out_lo, stream);
}

static int conv_tcs_launch(const void* xa, const void* xb, const void* wpk, const float* bias, void* out, const void* mask,
                           int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int coutp, int kd, int out_mode,
                           float slope, void* out2, int csplit, const float* acc_in, void* out_lo, void* stream) {
  VXM_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && wpk && out, "conv3d_tcs_fwd: bad argument");
  VXM_REQUIRE(kd == 1 || kd == 3, "conv3d_tcs_fwd: kd must be 1 or 3");
  VXM_REQUIRE(coutp == 16 || coutp == 32 || coutp == 48 || coutp == 64, "conv3d_tcs_fwd: padded Cout must be 16, 32, 48 or 64");
  VXM_REQUIRE(!out2 || (out_mode == 0 && csplit > 0 && csplit < Cout && csplit % 8 == 0 && !mask), "conv3d_tcs_fwd: bad output split");
  VXM_REQUIRE(Cout > 0 && Cout <= coutp && (out_mode == 1 || out_mode == 2 || Cout % 8 == 0), "conv3d_tcs_fwd: unsupported Cout %d", Cout);
  VXM_REQUIRE(vxm_conv3d_tcs_supported(Ca, Cb, Cout), "conv3d_tcs_fwd: channel counts (%d,%d)->%d unsupported", Ca, Cb, Cout);
  VXM_REQUIRE((Ca == 0 || xa) && (Cb == 0 || xb), "conv3d_tcs_fwd: missing source tensor");
  VXM_REQUIRE(!up || (H % 2 == 0 && W % 2 == 0 && (kd == 1 || D % 2 == 0)), "conv3d_tcs_fwd: upsampled source needs even sizes");
  ConvSArgs a{};
  const int cin = Ca + Cb;
  int g0, g1;
  groups_of(cin, &g0, &g1);
  a.xa = (const __nv_bfloat16*)xa; a.xb = (const __nv_bfloat16*)xb; a.wpk = (const __nv_bfloat16*)wpk; a.bias = bias;
  a.out = out; a.mask = (const __nv_bfloat16*)mask; a.out2 = out2; a.csplit = csplit;
  a.acc_in = acc_in; a.out_lo = out_lo;
  a.B = B; a.D = D; a.H = H; a.W = W; a.Ca = Ca; a.Cb = Cb; a.up = up; a.upd = (up && kd == 3) ? 1 : 0;
  a.Cout = Cout; a.out_mode = out_mode; a.slope = slope;
  a.wbytes = (uint32_t)vxm_conv3d_tcs_packed_bytes(cin, coutp, kd);
  const size_t fixed = ((a.wbytes + 1023u) & ~1023u) + 1024 + 512;
  // tile height: 8 rows (two accumulators per slab step) when the ring still holds >= 5 slabs and 4 accumulators fit TMEM
  int HTv = 4;
  {
    const size_t slab8 = (size_t)10 * WT * (g0 + g1) * 2;
    const int ns8 = (int)((227 * 1024 - fixed) / slab8);
    const char* e = getenv("VXM_B200_TCS_HT");
    if (g0 <= 32 && coutp <= 32 && ns8 >= (kd == 3 ? 5 : 3) && H > 4 && !(e && e[0] == '4')) HTv = 8;
  }
  a.tiles_h = (H + HTv - 1) / HTv; a.tiles_w = (W + WUSE - 1) / WUSE;
  int nsm = sm_count();
  // depth chunking: balance the persistent CTAs (waves of nsm items) against the 2 halo slabs every chunk re-loads
  const long long tiles = (long long)B * a.tiles_h * a.tiles_w;
  int best_nch = 1;
  double best_cost = 1e300;
  for (int nch = 1; nch <= 40 && nch <= D; ++nch) {
    const int dc = (D + nch - 1) / nch;
    const long long items = tiles * ((D + dc - 1) / dc);
    const long long waves = (items + nsm - 1) / nsm;
    const double cost = (double)waves * (dc + (kd == 3 ? 2.5 : 0.5));
    if (cost < best_cost - 1e-9) { best_cost = cost; best_nch = nch; }
  }
  a.dchunk = (D + best_nch - 1) / best_nch; a.nchunks = (D + a.dchunk - 1) / a.dchunk;
  a.nitems = (int)(tiles * a.nchunks);
  const size_t slab = (size_t)(HTv + 2) * WT * (g0 + g1) * 2;
  int nslot = (int)((227 * 1024 - fixed) / slab);
  {
    // ring depth: the slabs of up to 13 steps ahead (16-channel layers) hide the L2 / HBM latency of the tensor copies — with 8
    // slots the issuers waited ~550 clk per step for slab data (profiles/r2_conv_ablation.md).  VXM_B200_RING=8: A/B switch.
    const char* e = getenv("VXM_B200_RING");
    const int cap = e ? atoi(e) : MAXSLOT;
    if (nslot > cap) nslot = cap;
    if (nslot > MAXSLOT) nslot = MAXSLOT;
  }
  VXM_REQUIRE(nslot >= 4, "conv3d_tcs_fwd: not enough shared memory for the slab ring");
  a.nslot = nslot;
  size_t smem = fixed + (size_t)nslot * slab;
  int grid = a.nitems < nsm ? a.nitems : nsm;
  cudaStream_t st = as_stream(stream);
  if (plan_tma(a, g0, g1, HTv) != 0) return VXM_ERR_CUDA;
  const bool acc_epi = acc_in != nullptr || out_mode >= 2;
  // epilogue specialisation: 3-D, bf16 channels-last output, every padded channel real
  int epi = 0;
  {
    const char* e = getenv("VXM_B200_TCS_EPI");       // "0": generic epilogue everywhere (A/B switch)
    const bool plain = kd == 3 && out_mode == 0 && !acc_epi && Cout == coutp && !(e && e[0] == '0');
    if (plain && !out2 && !mask && slope >= 0.f && slope <= 1.f) epi = 1;
    else if (plain && !out2 && mask && !bias) epi = 2;
    else if (plain && !mask && !bias && slope < 0.f && (!out2 || csplit % 16 == 0)) epi = 3;
  }
#define VXM_TCS_LAUNCH_E(KD_, G0_, G1_, CO_, HT_, ACC_, E_)                                                                       \
  do {                                                                                                                            \
    VXM_CUDA(cudaFuncSetAttribute(conv_tcs_kernel<KD_, G0_, G1_, CO_, HT_, ACC_, E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    conv_tcs_kernel<KD_, G0_, G1_, CO_, HT_, ACC_, E_><<<grid, NTHREADS, smem, st>>>(a);                                          \
  } while (0)
#define VXM_TCS_LAUNCH(KD_, G0_, G1_, CO_, HT_)                                                                                   \
  do {                                                                                                                            \
    if (acc_epi) VXM_TCS_LAUNCH_E(KD_, G0_, G1_, CO_, HT_, true, 0);                                                              \
    else if (KD_ == 3 && epi == 1) VXM_TCS_LAUNCH_E(3, G0_, G1_, CO_, HT_, false, 1);                                             \
    else if (KD_ == 3 && epi == 2) VXM_TCS_LAUNCH_E(3, G0_, G1_, CO_, HT_, false, 2);                                             \
    else if (KD_ == 3 && epi == 3) VXM_TCS_LAUNCH_E(3, G0_, G1_, CO_, HT_, false, 3);                                             \
    else VXM_TCS_LAUNCH_E(KD_, G0_, G1_, CO_, HT_, false, 0);                                                                     \
  } while (0)
#define VXM_TCS_G8(KD_, CO_)                                                  \
  do {                                                                        \
    if (g0 == 16) VXM_TCS_LAUNCH(KD_, 16, 0, CO_, 8);                         \
    else if (g0 == 32 && g1 == 0) VXM_TCS_LAUNCH(KD_, 32, 0, CO_, 8);         \
    else VXM_TCS_LAUNCH(KD_, 32, 16, CO_, 8);                                 \
  } while (0)
#define VXM_TCS_G(KD_, CO_)                                                   \
  do {                                                                        \
    if (g0 == 16) VXM_TCS_LAUNCH(KD_, 16, 0, CO_, 4);                         \
    else if (g0 == 32 && g1 == 0) VXM_TCS_LAUNCH(KD_, 32, 0, CO_, 4);         \
    else if (g0 == 32) VXM_TCS_LAUNCH(KD_, 32, 16, CO_, 4);                   \
    else VXM_TCS_LAUNCH(KD_, 64, 0, CO_, 4);                                  \
  } while (0)
  if (HTv == 8) {
    if (kd == 3) { if (coutp == 16) VXM_TCS_G8(3, 16); else VXM_TCS_G8(3, 32); }
    else { if (coutp == 16) VXM_TCS_G8(1, 16); else VXM_TCS_G8(1, 32); }
  } else if (kd == 3) { if (coutp == 16) VXM_TCS_G(3, 16); else if (coutp == 32) VXM_TCS_G(3, 32); else if (coutp == 48) VXM_TCS_G(3, 48); else VXM_TCS_G(3, 64); }
  else { if (coutp == 16) VXM_TCS_G(1, 16); else if (coutp == 32) VXM_TCS_G(1, 32); else if (coutp == 48) VXM_TCS_G(1, 48); else VXM_TCS_G(1, 64); }
  return check_launch("conv3d_tcs_fwd");
}
