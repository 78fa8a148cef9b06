// Microbenchmark: FP64 pipe issue rate on B200 in warp-instructions per cycle per SM sub-partition, as a function of
// resident warps per sub-partition and independent chains per thread (clock64-based: independent of the SM clock).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ipc fp64_ipc.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double* out, long long* cyc, int iters, double s) {
  double c[ILP];
  for (int j = 0; j < ILP; ++j) c[j] = s * (threadIdx.x + j);
  const double a = s * 1.0000001, b = s * 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        if (OP == 0) c[j] = fma(c[j], a, b);
        if (OP == 1) c[j] = c[j] + b;
        if (OP == 2) c[j] = c[j] * a;
      }
    }
  }
  long long t1 = clock64();
  double acc = 0;
  for (int j = 0; j < ILP; ++j) acc += c[j];
  if (acc == 1.2345) out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP, int OP>
void run(int warps_per_smsp) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 148);
  const int iters = 2000, threads = warps_per_smsp * 4 * 32;
  k<ILP, OP><<<148, threads>>>(out, cyc, iters, 1.0);
  k<ILP, OP><<<148, threads>>>(out, cyc, iters, 2.0);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double inst_per_warp = (double)iters * 16 * ILP;
  printf("{\"op\":\"%s\",\"ilp\":%d,\"warps_per_smsp\":%d,\"cycles\":%.0f,\"warp_inst_per_cycle_per_smsp\":%.4f,\"cycles_per_dependent_inst\":%.2f}\n",
         OP == 0 ? "dfma" : OP == 1 ? "dadd" : "dmul", ILP, warps_per_smsp, avg, inst_per_warp * warps_per_smsp / avg, avg / (iters * 16.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, 0>(1); run<2, 0>(1); run<4, 0>(1); run<8, 0>(1); run<16, 0>(1);
  run<1, 0>(2); run<4, 0>(2); run<8, 0>(2); run<8, 0>(4); run<8, 0>(8);
  run<8, 1>(2); run<8, 2>(2); run<8, 1>(4);
  return 0;
}
