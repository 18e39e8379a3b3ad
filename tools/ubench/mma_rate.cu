// Micro-benchmark (profiling aid): cycles per tcgen05.mma (kind::f16, bf16 -> fp32, cta_group::1) issued back to back by one thread,
// for the operand forms the convolution and weight-gradient kernels use.  Shared memory holds finite garbage; only timing matters.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I voxelmorph_b200/csrc -I include tools/ubench/mma_rate.cu voxelmorph_b200/csrc/build/capi.o -o tools/ubench/mma_rate.bin
#include <cstdio>
#include "tc_common.cuh"
using namespace vxm::tc;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int swz_code) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)swz_code << 61;       // 0 none, 2 = 128B, 4 = 64B, 6 = 32B
  return d;
}

struct Case { int M, N, a_mn, b_mn, wa, wb, nacc; const char* name; int boff = 0, kstep = 32; };

__global__ void __launch_bounds__(128, 1) k(long long* out_all, int ncase, const int* cfg) {
  long long* out = out_all + blockIdx.x * 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (uint32_t i = threadIdx.x * 16u; i < 160 * 1024; i += blockDim.x * 16u) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  if (threadIdx.x < 32) {
    uint32_t ph = 0;
    for (int c = 0; c < ncase; ++c) {
      const int M = cfg[c * 7], N = cfg[c * 7 + 1], a_mn = cfg[c * 7 + 2], b_mn = cfg[c * 7 + 3], wa = cfg[c * 7 + 4], wb = cfg[c * 7 + 5], nacc = cfg[c * 7 + 6] & 0xff, boff = (cfg[c * 7 + 6] >> 8) & 0xfff, kstep = cfg[c * 7 + 6] >> 20;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const uint32_t a_u = smem_u32(smem), b_u = smem_u32(smem + 96 * 1024) + (uint32_t)boff;
      auto code = [](int w) { return w == 128 ? 2 : (w == 64 ? 4 : (w == 32 ? 6 : 0)); };
      // K-major: rows of w bytes, 8-row groups contiguous (SBO = 8 w).  MN-major: rows (K index) of w bytes, atoms `lbo` apart.
      const uint64_t ad = a_mn ? desc(a_u, 6144u, 8u * wa, code(wa)) : desc(a_u, 16u, 8u * wa, code(wa));
      const uint64_t bd = b_mn ? desc(b_u, (uint32_t)wb, 8u * wb, code(wb)) : desc(b_u, 16u, 8u * wb, code(wb));
      long long t0 = 0, t1 = 0;
      for (int rep = 0; rep < 2; ++rep) {
        t0 = clock64();
        if (elect_one()) {
#pragma unroll 1
          for (int i = 0; i < 64; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) umma_f16(tm + (uint32_t)((j % nacc) * N), ad + (uint64_t)(j * (kstep >> 4)), bd + (uint64_t)(j * (kstep >> 4)), idesc, 1u);
          }
          umma_commit(&bar);
        }
        __syncwarp();
        while (!mbar_try_wait(&bar, ph)) {}
        ph ^= 1;
        t1 = clock64();
      }
      if ((threadIdx.x & 31) == 0) out[c] = (t1 - t0);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  // {M, N, a_mn_major, b_mn_major, row bytes of A, row bytes of B}
  const Case cases[] = {
      {128, 48, 0, 0, 32, 32, 2, "conv 16ch : A K-major SW32, B K-major SW32, N=48"},
      {128, 96, 0, 0, 64, 64, 2, "conv 32ch : A K-major SW64, B K-major SW64, N=96"},
      {128, 144, 0, 0, 64, 64, 2, "conv 32->48: A K-major SW64, B K-major SW64, N=144"},
      {128, 48, 0, 0, 128, 128, 2, "A K-major SW128, B K-major SW128, N=48"},
      {64, 48, 1, 1, 32, 32, 2, "wgrad 16/16: A MN-major SW32 (M=64), B MN-major SW32, N=48"},
      {128, 48, 1, 1, 64, 32, 2, "wgrad 32/16: A MN-major SW64 (M=128), B MN-major SW32, N=48"},
      {64, 96, 1, 1, 32, 64, 2, "wgrad 16/32: A MN-major SW32 (M=64), B MN-major SW64, N=96"},
      {128, 96, 1, 1, 64, 64, 2, "wgrad 32/32: A MN-major SW64 (M=128), B MN-major SW64, N=96"},
      {128, 96, 0, 1, 64, 64, 2, "A K-major SW64 (M=128), B MN-major SW64, N=96"},
      {128, 96, 1, 0, 64, 64, 2, "A MN-major SW64 (M=128), B K-major SW64, N=96"},
      {64, 48, 0, 0, 32, 32, 2, "A K-major SW32 (M=64), B K-major SW32, N=48"},
      {128, 256, 0, 0, 128, 128, 2, "A K-major SW128, B K-major SW128, N=256 (math-bound reference)"},
      {128, 48, 0, 0, 32, 32, 1, "conv 16ch, ALL MMAs INTO ONE ACCUMULATOR, N=48"},
      {128, 96, 0, 0, 64, 64, 1, "conv 32ch, one accumulator, N=96"},
      {64, 48, 1, 1, 32, 32, 1, "wgrad 16/16 (M=64), one accumulator, N=48"},
      {64, 48, 1, 1, 32, 32, 3, "wgrad 16/16 (M=64), three accumulators, N=48"},
      {128, 96, 1, 1, 64, 64, 1, "wgrad 32/32 (M=128), one accumulator, N=96"},
      {128, 96, 1, 1, 64, 64, 3, "wgrad 32/32 (M=128), three accumulators, N=96"},
      {64, 48, 1, 1, 32, 32, 3, "wgrad 16/16 as in the kernel: K chunks 512 B apart", 0, 512},
      {64, 48, 1, 1, 32, 32, 3, "wgrad 16/16 as in the kernel: 512 B K chunks, B (Toeplitz) starts at row 15 (+480 B)", 480, 512},
      {64, 48, 1, 1, 32, 32, 3, "wgrad 16/16: 512 B K chunks, B starts at row 16 (+512 B, aligned)", 512, 512},
      {128, 96, 1, 1, 64, 64, 3, "wgrad 32/32 as in the kernel: K chunks 1024 B apart, B starts at row 15 (+960 B)", 960, 1024},
      {128, 96, 1, 1, 64, 64, 3, "wgrad 32/32: 1024 B K chunks, B aligned (+1024 B)", 1024, 1024},
      {128, 48, 1, 1, 64, 32, 3, "wgrad 32/16 as in the kernel: B row 15 (+480 B), A chunks 1024 B apart", 480, 512},
  };
  const int n = sizeof(cases) / sizeof(cases[0]);
  int h[24 * 7];
  for (int i = 0; i < n; ++i) { h[i * 7] = cases[i].M; h[i * 7 + 1] = cases[i].N; h[i * 7 + 2] = cases[i].a_mn; h[i * 7 + 3] = cases[i].b_mn; h[i * 7 + 4] = cases[i].wa; h[i * 7 + 5] = cases[i].wb; h[i * 7 + 6] = cases[i].nacc | (cases[i].boff << 8) | (cases[i].kstep << 20); }
  int* dcfg; long long* dout;
  cudaMalloc(&dcfg, sizeof(h)); cudaMalloc(&dout, 148 * 32 * 8);
  cudaMemcpy(dcfg, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {1, 148}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<grid, 128, 200 * 1024>>>(dout, n, dcfg);          // warm
    cudaEventRecord(e0);
    for (int rep = 0; rep < 20; ++rep) k<<<grid, 128, 200 * 1024>>>(dout, n, dcfg);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long r[24];
    cudaMemcpy(r, dout, 24 * 8, cudaMemcpyDeviceToHost);
    long long tot = 0; for (int i = 0; i < n; ++i) tot += 2 * r[i];
    printf("status %s; grid %d: 512 MMAs (K = 16 each) per case issued back to back by one thread per CTA; 20 launches took %.3f ms -> %.0f MHz effective (cycles counted in CTA 0 / wall time)\n",
           cudaGetErrorString(e), grid, ms, tot * 20 / (ms * 1e3));
    for (int i = 0; i < n; ++i) printf("  %-72s %7.1f clk/MMA  (math floor %5.1f)\n", cases[i].name, r[i] / 512.0, 128 * cases[i].N / 256.0);
  }
  return 0;
}
