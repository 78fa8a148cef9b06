// Microbenchmark: FP64-pipe utilisation of the potential evaluation alone (LogisticF<BETALIK, 20>::evalv<W>), 4 warps per
// sub-partition, no TMEM / conversion / reductions around it.  Tells how much of k_project_q's FP64 idle time is the
// evaluation's own instruction mix and how much comes from the rest of the epilogue.
#include <cstdio>
#include <cuda_runtime.h>
#include "../beta-cores_b200/csrc/bc_models.cuh"
using namespace bc;
template <int W>
__global__ void __launch_bounds__(512, 1) k(double* out, long long* cyc, int iters, const ModelParams mp) {
  double c[W], ca[W], f[W], acc = 0.0;
  for (int j = 0; j < W; ++j) { c[j] = 0.01 * (threadIdx.x % 97) - 0.3 * j; ca[j] = 0.0; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    LogisticF<KIND_BETALIK, 20>::evalv<W>(c, 0.0, ca, mp, f);
#pragma unroll
    for (int j = 0; j < W; ++j) { acc += f[j]; c[j] = c[j] * 0.999 + 1e-3; }
  }
  long long t1 = clock64();
  if (acc == 1.2345) out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int W>
void run(int wps) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 148);
  ModelParams mp; for (int i = 0; i < 8; ++i) mp.p[i] = 0; mp.p[0] = 0.1; mp.p[1] = 11.0; mp.p[2] = 700.0;
  for (int i = 0; i <= kPowPolyMax; ++i) mp.q[i] = 1.0 / (1 + i);
  const int iters = 4000, threads = wps * 128;
  k<W><<<148, threads>>>(out, cyc, iters, mp);
  k<W><<<148, threads>>>(out, cyc, iters, mp);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("{\"W\":%d,\"warps_per_smsp\":%d,\"cycles_per_eval_per_warp\":%.1f,\"cycles_per_eval_per_smsp\":%.1f}\n", W, wps, avg / (iters * (double)W),
         avg / (iters * (double)W * wps));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<4>(4); run<4>(2); run<4>(1); run<8>(2); run<2>(4); run<1>(4);
  return 0;
}
