// Modular-multiplication throughput: IMAD.WIDE path (field.cuh) vs DFMA path (field52.cuh), BN254 Fp
// and BLS12-381 Fp.  Each thread runs a dependent chain x <- x*y; many warps per SM saturate the pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I gnark_b200/csrc -o tools/mulbench tools/mulbench.cu
#include <cfenv>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "field.cuh"
#include "field52.cuh"
using namespace gb200;

#define ITER 2000

template <class F>
__global__ void k_imad(F* io, F y) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  F x = io[t];
  for (int i = 0; i < ITER; i++) x = x * y;
  io[t] = x;
}
template <class P52>
__global__ void k_dfma(F52<P52>* io, F52<P52> y) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  F52<P52> x = io[t];
  const D52<P52> dy(y);
  for (int i = 0; i < ITER; i++) x = mul52<P52>(D52<P52>(x), dy);
  io[t] = x;
}

template <class F, class P52>
void run(const char* name, int sms) {
  const int threads = 128, blocks = sms * 8;
  const int n = threads * blocks;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // IMAD
  std::vector<F> h(n);
  for (int t = 0; t < n; t++) { h[t] = F::one(); h[t].l[0] += t; }
  F y = F::one(); y.l[1] = 12345;
  F* d; cudaMalloc(&d, n * sizeof(F));
  cudaMemcpy(d, h.data(), n * sizeof(F), cudaMemcpyHostToDevice);
  k_imad<F><<<blocks, threads>>>(d, y); cudaDeviceSynchronize();
  cudaEventRecord(e0); k_imad<F><<<blocks, threads>>>(d, y); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double rate_i = (double)n * ITER / (ms * 1e-3);
  printf("%-14s IMAD.WIDE path: %8.3f ms  %7.2f G modmul/s\n", name, ms, rate_i / 1e9);
  // DFMA
  using G = F52<P52>;
  std::vector<G> g(n);
  for (int t = 0; t < n; t++) { for (int i = 0; i < P52::L; i++) g[t].l[i] = (int64_t)P52::one52(i); g[t].l[0] += (t % 1000); }
  G y2; for (int i = 0; i < P52::L; i++) y2.l[i] = (int64_t)P52::r2_52(i);
  G* d2; cudaMalloc(&d2, n * sizeof(G));
  cudaMemcpy(d2, g.data(), n * sizeof(G), cudaMemcpyHostToDevice);
  k_dfma<P52><<<blocks, threads>>>(d2, y2); cudaDeviceSynchronize();
  cudaMemcpy(d2, g.data(), n * sizeof(G), cudaMemcpyHostToDevice);
  cudaEventRecord(e0); k_dfma<P52><<<blocks, threads>>>(d2, y2); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  double rate_d = (double)n * ITER / (ms * 1e-3);
  printf("%-14s DFMA path     : %8.3f ms  %7.2f G modmul/s   (x%.2f)\n", name, ms, rate_d / 1e9, rate_d / rate_i);
  // check thread 7 against the host (round-toward-zero emulation)
  G got; cudaMemcpy(&got, d2 + 7, sizeof(G), cudaMemcpyDeviceToHost);
  fesetround(FE_TOWARDZERO);
  G x = g[7];
  for (int i = 0; i < ITER; i++) x = mul52<P52>(x, y2);
  fesetround(FE_TONEAREST);
  bool ok = true;
  for (int i = 0; i < P52::L; i++) ok &= (x.l[i] == got.l[i]);
  printf("%-14s DFMA device result %s the host emulation\n", name, ok ? "MATCHES" : "DIFFERS FROM");
  cudaFree(d); cudaFree(d2);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  run<bn254_fp, bn254_fp_params52>("bn254 Fp", p.multiProcessorCount);
  run<bls12_381_fp, bls12_381_fp_params52>("bls12-381 Fp", p.multiProcessorCount);
  return 0;
}
