// Microbenchmark: does the operand kind of a DFMA (register / constant bank / uniform register) change its issue rate?
#include <cstdio>
#include <cuda_runtime.h>
struct Coef { double q[16]; };
__constant__ double kC[16];
template <int ILP, int OP>
__global__ void k(double* out, long long* cyc, int iters, double s, const Coef P) {
  double c[ILP], x[ILP];
  for (int j = 0; j < ILP; ++j) { c[j] = s * (threadIdx.x + j); x[j] = s * 1e-3 * (j + 1); }
  double r[16];
  for (int j = 0; j < 16; ++j) r[j] = s * j;   // register-resident coefficients (OP 0)
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int st = 0; st < 16; ++st) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        if (OP == 0) c[j] = fma(c[j], x[j], r[st]);
        if (OP == 1) c[j] = fma(c[j], x[j], P.q[st]);   // kernel-parameter constant bank
        if (OP == 2) c[j] = fma(c[j], x[j], kC[st]);    // __constant__
        if (OP == 3) c[j] = fma(c[j], x[j], 1.0 + st);  // immediate
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
void run(int wps, const char* name) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 148);
  Coef P; for (int i = 0; i < 16; ++i) P.q[i] = 0.5 + i;
  const int iters = 2000, threads = wps * 4 * 32;
  k<ILP, OP><<<148, threads>>>(out, cyc, iters, 1.0, P);
  k<ILP, OP><<<148, threads>>>(out, cyc, iters, 2.0, P);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("{\"operand\":\"%s\",\"ilp\":%d,\"warps_per_smsp\":%d,\"warp_inst_per_cycle_per_smsp\":%.4f}\n", name, ILP, wps,
         (double)iters * 16 * ILP * wps / avg);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  double hc[16]; for (int i = 0; i < 16; ++i) hc[i] = 0.25 + i;
  cudaMemcpyToSymbol(kC, hc, sizeof(hc));
  run<4, 0>(2, "register"); run<4, 1>(2, "param_cbank"); run<4, 2>(2, "constant"); run<4, 3>(2, "immediate");
  run<4, 0>(4, "register"); run<4, 1>(4, "param_cbank"); run<4, 2>(4, "constant"); run<4, 3>(4, "immediate");
  return 0;
}
