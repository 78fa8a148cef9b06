// Microbenchmark: FP64 DMMA (mma.sync f64) and DFMA peak on B200 (sm_100a).
// Establishes the fp64 roofline denominators for the projection kernel
// (MEASURED_PEAKS.json only carries HBM GB/s and bf16 TFLOP/s).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma_16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
    : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
      "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma_1688(double (&c)[4], const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
    : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma_1684(double (&c)[4], const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
    : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
    : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}
__device__ __forceinline__ void dmma_884(double (&c)[2], const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
    : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
}

// mode: 0 = m16n8k16, 1 = m16n8k8, 2 = m16n8k4, 3 = m8n8k4, 4 = DFMA only,
//       5 = mixed (even warps DMMA k16, odd warps DFMA)
template <int MODE>
__global__ void __launch_bounds__(512) k_peak(double* out, int iters, double seed) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = seed * (lane + i + 1) * 1e-3;
  for (int i = 0; i < 4; i++) b[i] = seed * (lane - i) * 1e-3;
  double acc = 0;
  bool do_mma = (MODE <= 3) || (MODE == 5 && (warp & 1) == 0);
  if (do_mma) {
    if (MODE == 3) {
      double c[8][2];
      for (int j = 0; j < 8; j++) { c[j][0] = 0; c[j][1] = 0; }
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma_884(c[j], a + (j & 7), b + (j & 3));
      }
      for (int j = 0; j < 8; j++) acc += c[j][0] + c[j][1];
    } else {
      double c[8][4];
      for (int j = 0; j < 8; j++) for (int q = 0; q < 4; q++) c[j][q] = 0;
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (MODE == 0 || MODE == 5) dmma_16816(c[j], a, b);
          else if (MODE == 1) dmma_1688(c[j], a + (j & 3), b + (j & 1));
          else dmma_1684(c[j], a + (j & 3), b + (j & 3));
        }
      }
      for (int j = 0; j < 8; j++) for (int q = 0; q < 4; q++) acc += c[j][q];
    }
  } else {
    double c[16];
    for (int j = 0; j < 16; j++) c[j] = seed * j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 8; r++) {
#pragma unroll
        for (int j = 0; j < 16; j++) c[j] = fma(c[j], a[j & 7], b[j & 3]);
      }
    }
    for (int j = 0; j < 16; j++) acc += c[j];
  }
  if (acc == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int threads, int blocks_per_sm, int iters, double flop_per_warp_iter_mma, double flop_per_warp_iter_fma) {
  int sms = 148;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * blocks_per_sm * threads));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) k_peak<MODE><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0 + w);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0));
    k_peak<MODE><<<sms * blocks_per_sm, threads>>>(out, iters, 2.0 + r);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double warps = (double)sms * blocks_per_sm * threads / 32;
  double mma_warps = warps, fma_warps = 0;
  if (MODE == 4) { mma_warps = 0; fma_warps = warps; }
  if (MODE == 5) { mma_warps = warps / 2; fma_warps = warps / 2; }
  double fl_mma = mma_warps * iters * flop_per_warp_iter_mma;
  double fl_fma = fma_warps * iters * flop_per_warp_iter_fma;
  printf("{\"bench\":\"%s\",\"threads\":%d,\"blocks_per_sm\":%d,\"ms\":%.4f,\"mma_tflops\":%.3f,\"fma_tflops\":%.3f,\"total_tflops\":%.3f}\n",
         name, threads, blocks_per_sm, best, fl_mma / best / 1e9, fl_fma / best / 1e9, (fl_mma + fl_fma) / best / 1e9);
  CK(cudaFree(out));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, p.multiProcessorCount, p.clockRate);
  const int it = 4000;
  const double fma_flop = 8.0 * 16 * 2 * 32;  // per warp-iteration of the DFMA loop
  for (int th : {128, 256, 512}) {
    for (int bps : {1, 2}) {
      if (th == 128) { run<0>("dmma_m16n8k16", 128, bps, it, 8 * 2.0 * 16 * 8 * 16, 0); }
      if (th == 256) { run<0>("dmma_m16n8k16", 256, bps, it, 8 * 2.0 * 16 * 8 * 16, 0); }
      if (th == 512) { run<0>("dmma_m16n8k16", 512, bps, it, 8 * 2.0 * 16 * 8 * 16, 0); }
    }
  }
  run<1>("dmma_m16n8k8", 256, 2, it, 8 * 2.0 * 16 * 8 * 8, 0);
  run<2>("dmma_m16n8k4", 256, 2, it, 8 * 2.0 * 16 * 8 * 4, 0);
  run<3>("dmma_m8n8k4", 256, 2, it, 8 * 2.0 * 8 * 8 * 4, 0);
  run<4>("dfma", 256, 2, it, 0, fma_flop);
  run<4>("dfma", 512, 2, it, 0, fma_flop);
  run<5>("mixed_dmma_dfma", 256, 2, it, 8 * 2.0 * 16 * 8 * 16, fma_flop);
  run<5>("mixed_dmma_dfma", 512, 2, it, 8 * 2.0 * 16 * 8 * 16, fma_flop);
  return 0;
}
