// Developer probe: effective SM clock (clock64 / globaltimer) over the life of one long fp64-heavy kernel.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void burn(double* out, unsigned long long* samples, int nsamp, long long iters_per_sample) {
  double a = threadIdx.x*1e-3 + 1.0, b = 1.0000001, c = 1e-9;
  unsigned long long t_prev, c_prev;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_prev));
  c_prev = clock64();
  for (int s = 0; s < nsamp; ++s) {
    for (long long k = 0; k < iters_per_sample; ++k) { a = fma(a, b, c); a = fma(a, b, -c); a = fma(a, b, c); a = fma(a, b, -c); }
    unsigned long long t, cc;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    cc = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { samples[2*s] = t - t_prev; samples[2*s+1] = cc - c_prev; }
    t_prev = t; c_prev = cc;
  }
  out[blockIdx.x*blockDim.x + threadIdx.x] = a;
}
int main(int argc, char** argv) {
  int nsamp = 60; long long ips = argc > 1 ? atoll(argv[1]) : 200000;
  int blocks = argc > 2 ? atoi(argv[2]) : 296*2;
  double* out; unsigned long long* samples;
  cudaMalloc(&out, blocks*256*sizeof(double)); cudaMallocManaged(&samples, 2*nsamp*sizeof(unsigned long long));
  burn<<<blocks, 256>>>(out, samples, 4, 1000); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); burn<<<blocks, 256>>>(out, samples, nsamp, ips); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("kernel %.1f ms, blocks %d\n", ms, blocks);
  double tacc = 0;
  for (int s = 0; s < nsamp; ++s) { tacc += samples[2*s]*1e-6; printf("t=%7.1f ms  eff SM clock %.0f MHz\n", tacc, samples[2*s+1]/(samples[2*s]*1e-3)); }
  return 0;
}
