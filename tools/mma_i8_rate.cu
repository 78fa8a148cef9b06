// tcgen05.mma.kind::i8 (M = 128, K = 32) throughput against N on one SM / on all SMs: how long does a back-to-back
// stream of MMAs of a given N take per instruction?  (The stacked-digit scheme of k_project_q issues N = 224 ... 32.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I beta-cores_b200/csrc -o tools/mma_i8_rate tools/mma_i8_rate.cu
#include <cstdio>
#include <cstdlib>
#include "bc_common.cuh"
#include "bc_umma.cuh"
using namespace bc;

__global__ void __launch_bounds__(128, 1) k_rate(int N, int reps, int also_b_read, unsigned long long* cycles) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  unsigned char* As = smem;              // 128 rows x 128 B
  unsigned char* Bs = smem + 16384;      // 256 rows x 128 B
  for (int i = threadIdx.x; i < 16384 + 32768; i += blockDim.x) smem[i] = (unsigned char)(i * 7 + 1);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc(&tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_i8(N);
    const uint64_t ad = umma_desc_sw128(smem_u32(As)), bd = umma_desc_sw128(smem_u32(Bs));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    unsigned long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_i8(tbase + (uint32_t)((r & 1) * 256), ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (r | k) ? 1u : 0u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, pass & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

int main() {
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  unsigned long long* d;
  cudaMalloc(&d, 148 * sizeof(unsigned long long));
  const int reps = 2000;
  for (int grid : {1, 148}) {
    for (int N : {32, 64, 96, 128, 160, 192, 224, 256}) {
      k_rate<<<grid, 128, 60 * 1024>>>(N, reps, 0, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[148];
      cudaMemcpy(h, d, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      unsigned long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double per = (double)mx / (reps * 4.0);
      const double macs = 128.0 * N * 32.0;
      printf("{\"grid\": %d, \"N\": %d, \"cycles_per_mma\": %.1f, \"macs_per_cycle\": %.0f, \"ideal_cycles_at_8192\": %.1f}\n", grid, N, per, macs / per, macs / 8192.0);
    }
  }
  return 0;
}
