// Micro-benchmark (profiling aid, not part of the library): cycle cost of the synchronisation primitives the convolution
// pipeline is built from, measured with clock64() in one warp of one CTA.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I voxelmorph_b200/csrc tools/ubench/mbar_latency.cu -o gpurun_out/mbar_latency
#include <cstdio>
#include "tc_common.cuh"
using namespace vxm::tc;

__global__ void k(long long* out) {
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tslot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 32); mbar_init(&bar[3], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tslot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp != 0) return;
  long long t0, t1;
  // 1. clock64 overhead
  t0 = clock64();
  for (int i = 0; i < 64; ++i) asm volatile("" ::: "memory");
  t1 = clock64();
  if (lane == 0) out[0] = t1 - t0;
  // 2. arrive + try_wait on own barrier (count 1), 256 dependent round trips
  uint32_t ph = 0;
  t0 = clock64();
  for (int i = 0; i < 256; ++i) {
    if (lane == 0) mbar_arrive(&bar[0]);
    __syncwarp();
    while (!mbar_try_wait(&bar[0], ph)) {}
    ph ^= 1;
  }
  t1 = clock64();
  if (lane == 0) out[1] = (t1 - t0) / 256;
  // 3. try_wait alone on an already completed phase (parity of the previous phase), 256 times
  t0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 256; ++i) acc += mbar_try_wait(&bar[0], ph ^ 1) ? 1u : 0u;
  t1 = clock64();
  if (lane == 0) { out[2] = (t1 - t0) / 256; out[7] = acc; }
  // 4. 32-lane arrive (count 32) + wait
  ph = 0;
  t0 = clock64();
  for (int i = 0; i < 256; ++i) {
    mbar_arrive(&bar[2]);
    while (!mbar_try_wait(&bar[2], ph)) {}
    ph ^= 1;
  }
  t1 = clock64();
  if (lane == 0) out[3] = (t1 - t0) / 256;
  // 5. tcgen05.commit (nothing outstanding) -> barrier completion
  ph = 0;
  t0 = clock64();
  for (int i = 0; i < 256; ++i) {
    if (elect_one()) umma_commit(&bar[1]);
    __syncwarp();
    while (!mbar_try_wait(&bar[1], ph)) {}
    ph ^= 1;
  }
  t1 = clock64();
  if (lane == 0) out[4] = (t1 - t0) / 256;
  // 6. tcgen05.commit issue cost alone (4 commits back to back, then one wait for the 4 phases)
  t0 = clock64();
  for (int i = 0; i < 64; ++i) {
    if (elect_one()) umma_commit(&bar[3]);
    __syncwarp();
    while (!mbar_try_wait(&bar[3], (uint32_t)(i & 1))) {}
  }
  t1 = clock64();
  if (lane == 0) out[5] = (t1 - t0) / 64;
  // 7. elect + syncwarp + fence pair
  t0 = clock64();
  for (int i = 0; i < 256; ++i) { tc_fence_after(); if (elect_one()) asm volatile("" ::: "memory"); __syncwarp(); tc_fence_before(); }
  t1 = clock64();
  if (lane == 0) out[6] = (t1 - t0) / 256;
  __syncwarp();
  tmem_dealloc(tslot, 32);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  k<<<1, 64>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("status %s\nclock64 pair + empty loop: %lld clk\narrive(1 lane)+try_wait round trip: %lld clk\ntry_wait on a completed phase: %lld clk\n"
         "arrive(32 lanes)+try_wait: %lld clk\ntcgen05.commit -> barrier complete -> observed: %lld clk\nsame, alternate check: %lld clk\nfence/elect/syncwarp/fence: %lld clk\n(check %lld)\n",
         cudaGetErrorString(e), h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  return 0;
}
