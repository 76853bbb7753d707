// microbench.cu -- FP64 pipe peak (the speed-of-light gauge of the fused sweep) and
// accuracy of the reciprocal used in the WENO weights.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

template <int OP>   // 0 DFMA, 1 DMUL, 2 DADD
__global__ void k_fp64(double* out, int iters, double a, double b)
{
  double x0 = threadIdx.x * 1e-9 + 1.0, x1 = x0 + 0.1, x2 = x0 + 0.2, x3 = x0 + 0.3, x4 = x0 + 0.4, x5 = x0 + 0.5, x6 = x0 + 0.6, x7 = x0 + 0.7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (OP == 0) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b); }
      if (OP == 1) { x0 *= a; x1 *= a; x2 *= a; x3 *= a; x4 *= a; x5 *= a; x6 *= a; x7 *= a; }
      if (OP == 2) { x0 += b; x1 += b; x2 += b; x3 += b; x4 += b; x5 += b; x6 += b; x7 += b; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ double rcp3(double x)
{
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0); e = fma(e, e, e); return fma(r, e, r);
}
__device__ __forceinline__ double rcp5(double x)
{
  double r = rcp3(x); double e = fma(-x, r, 1.0); return fma(r, e, r);
}
__global__ void k_rcp(const double* x, double* e0, double* e3, double* e5, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double ex = 1.0 / x[i], r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i]));
  e0[i] = fabs(r - ex) / ex; e3[i] = fabs(rcp3(x[i]) - ex) / ex; e5[i] = fabs(rcp5(x[i]) - ex) / ex;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, clock attr %d kHz\n", p.name, p.multiProcessorCount, clk);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const char* names[3] = { "DFMA", "DMUL", "DADD" };
  for (int op = 0; op < 3; op++) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(a);
      if (op == 0) k_fp64<0><<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
      if (op == 1) k_fp64<1><<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
      if (op == 2) k_fp64<2><<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (rep > 0 && ms < best) best = ms;
    }
    double ninst = (double)blocks * threads * iters * 64.0;
    printf("%s: %.3f ms, %.3f T thread-instr/s (%.2f per SM per clk at 1.965 GHz)%s\n", names[op], best,
           ninst / (best * 1e-3) / 1e12, ninst / (best * 1e-3) / p.multiProcessorCount / 1.965e9,
           op == 0 ? "  [x2 = FLOP/s]" : "");
  }
  const int n = 1 << 20;
  double* hx = (double*)malloc(n * sizeof(double));
  srand(7);
  for (int i = 0; i < n; i++) { double m = 1.0 + rand() / (double)RAND_MAX; int e = rand() % 400 - 200; hx[i] = ldexp(m, e); }
  double *dx, *e0, *e3, *e5; cudaMalloc(&dx, n * 8); cudaMalloc(&e0, n * 8); cudaMalloc(&e3, n * 8); cudaMalloc(&e5, n * 8);
  cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
  k_rcp<<<n / 256, 256>>>(dx, e0, e3, e5, n);
  double* h = (double*)malloc(n * 8);
  double* ptr[3] = { e0, e3, e5 };
  const char* nm[3] = { "seed (MUFU.RCP64H)", "seed + cubic step", "seed + cubic + quadratic" };
  for (int k = 0; k < 3; k++) {
    cudaMemcpy(h, ptr[k], n * 8, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < n; i++) if (h[i] > m) m = h[i];
    printf("rcp %-28s max rel err vs IEEE 1/x: %.3e\n", nm[k], m);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
