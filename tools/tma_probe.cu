// tma_probe.cu -- checks, on the B200, the three TMA behaviours the fused sweep (sweep_tma.cuh) builds on:
//   1. 4-D tiled loads of FP64 boxes with negative / out-of-range coordinates (zero fill), with and
//      without SWIZZLE_64B (inner box extent 8 doubles = 64 B), and the address permutation of the swizzle;
//   2. tiled stores clipped at the tensor bounds;
//   3. cp.reduce.async.bulk.tensor ... .add on FP64 (read-modify-write of the right-hand side in L2).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode()
{
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return (EncodeFn)fn;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// box: (b0, b1, b2, nf) doubles; one CTA, thread 0 drives the TMA
__global__ void probe(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout,
                      int c0, int c1, int c2, int nbox, double* dump, int s0, int s1, int s2, int mode, int stage)
{
  extern __shared__ __align__(1024) unsigned char raw[];
  double* tile = (double*)raw;
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (stage & 1) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(nbox * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(tile)), "l"(&tin), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(0) : "memory");
  }
  // wait phase 0
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  for (int i = threadIdx.x; i < nbox; i += blockDim.x) dump[i] = tile[i];
  }
  else { for (int i = threadIdx.x; i < nbox; i += blockDim.x) tile[i] = 0.0; }
  __syncthreads();
  if (!(stage & 2)) return;
  // modify: tile += 1000 (generic proxy), then store / reduce through the second map
  for (int i = threadIdx.x; i < nbox; i += blockDim.x) tile[i] = (mode == 2) ? 0.5 : tile[i] + 1000.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    if (mode == 1)
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                   :: "l"(&tout), "r"(smem_u32(tile)), "r"(s0), "r"(s1), "r"(s2), "r"(0) : "memory");
    else
      asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                   :: "l"(&tout), "r"(smem_u32(tile)), "r"(s0), "r"(s1), "r"(s2), "r"(0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

static int make_map(EncodeFn enc, CUtensorMap* m, double* base, int P0, int P1, int P2, int NF, int b0, int b1, int b2, int bf,
                    CUtensorMapSwizzle sw)
{
  cuuint64_t dims[4] = { (cuuint64_t)P0, (cuuint64_t)P1, (cuuint64_t)P2, (cuuint64_t)NF };
  cuuint64_t strides[3] = { (cuuint64_t)P0 * 8, (cuuint64_t)P0 * P1 * 8, (cuuint64_t)P0 * P1 * P2 * 8 };
  cuuint32_t box[4] = { (cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, (cuuint32_t)bf };
  cuuint32_t es[4] = { 1, 1, 1, 1 };
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed: %d\n", (int)r); return 1; }
  return 0;
}

int main(int argc, char** argv)
{
  const int stage = argc > 1 ? atoi(argv[1]) : 3;
  const int only = argc > 2 ? atoi(argv[2]) : -1;
  EncodeFn enc = get_encode();
  if (!enc) { printf("FAIL: no cuTensorMapEncodeTiled\n"); return 1; }
  const int P0 = 38, P1 = 22, P2 = 14, NF = 5;
  const long long npg = (long long)P0 * P1 * P2;
  std::vector<double> h(npg * NF);
  for (long long i = 0; i < npg * NF; i++) h[i] = (double)i + 0.25;
  double *d_in, *d_out, *d_dump;
  cudaMalloc(&d_in, h.size() * 8); cudaMalloc(&d_out, h.size() * 8); cudaMalloc(&d_dump, 65536);
  cudaMemcpy(d_in, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  int fails = 0;
  struct Case { const char* name; int b0, b1, b2; CUtensorMapSwizzle sw; int c0, c1, c2; } cases[] = {
    { "x-sweep box (32,8,1) no swizzle, interior", 32, 8, 1, CU_TENSOR_MAP_SWIZZLE_NONE, 2, 3, 4 },
    { "y box (8,32,1) NO swizzle, interior", 8, 32, 1, CU_TENSOR_MAP_SWIZZLE_NONE, 10, 0, 5 },
    { "y box (8,32,1) NO swizzle, negative y", 8, 32, 1, CU_TENSOR_MAP_SWIZZLE_NONE, 2, -26, 5 },
    { "y box (8,32,1) swizzle 32B, interior", 8, 32, 1, CU_TENSOR_MAP_SWIZZLE_32B, 11, 0, 5 },
    { "y box (8,32,1) swizzle 128B, interior", 8, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B, 10, 0, 5 },
    { "y box (16,16,1) swizzle 128B, interior", 16, 16, 1, CU_TENSOR_MAP_SWIZZLE_128B, 10, 0, 5 },
    { "x-sweep box (32,8,1) no swizzle, negative x", 32, 8, 1, CU_TENSOR_MAP_SWIZZLE_NONE, -26, 3, 4 },
    { "x-sweep box (32,8,1) past the end", 32, 8, 1, CU_TENSOR_MAP_SWIZZLE_NONE, 22, 18, 13 },
    { "y-sweep box (8,32,1) swizzle 64B", 8, 32, 1, CU_TENSOR_MAP_SWIZZLE_64B, 2, -26, 5 },
    { "y-sweep box (8,32,1) swizzle 64B, interior", 8, 32, 1, CU_TENSOR_MAP_SWIZZLE_64B, 10, 3, 5 },
    { "z-sweep box (8,1,32) swizzle 64B", 8, 1, 32, CU_TENSOR_MAP_SWIZZLE_64B, 34, 7, -20 },
    { "z-sweep box (8,1,32) swizzle 64B, store-able", 8, 1, 32, CU_TENSOR_MAP_SWIZZLE_64B, 34, 7, 3 },
    { "y-sweep box (8,32,1) swizzle 64B, past the end", 8, 32, 1, CU_TENSOR_MAP_SWIZZLE_64B, 34, 3, 13 },
  };
  int ci = -1;
  for (auto& c : cases) {
    ci++; if (only >= 0 && ci != only) continue;
    for (int mode = 1; mode <= 2; mode++) {
      CUtensorMap tin, tout;
      if (make_map(enc, &tin, d_in, P0, P1, P2, NF, c.b0, c.b1, c.b2, NF, c.sw)) return 1;
      if (make_map(enc, &tout, d_out, P0, P1, P2, NF, c.b0, c.b1, c.b2, NF, c.sw)) return 1;
      const int nbox = c.b0 * c.b1 * c.b2 * NF;
      std::vector<double> o0(h.size());
      for (size_t i = 0; i < o0.size(); i++) o0[i] = -(double)i;
      cudaMemcpy(d_out, o0.data(), o0.size() * 8, cudaMemcpyHostToDevice);
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, nbox * 8 + 1024);
      probe<<<1, 128, nbox * 8 + 1024>>>(tin, tout, c.c0, c.c1, c.c2, nbox, d_dump, c.c0, c.c1, c.c2, mode, stage);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("FAIL %s mode %d stage %d: %s\n", c.name, mode, stage, cudaGetErrorString(e)); return 1; }
      std::vector<double> dump(nbox), o1(h.size());
      cudaMemcpy(dump.data(), d_dump, nbox * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(o1.data(), d_out, o1.size() * 8, cudaMemcpyDeviceToHost);
      // expected tile contents: element (i0,i1,i2,f) of the box at dense index ((f*b2+i2)*b1+i1)*b0+i0, swizzled
      int bad_load = 0, bad_store = 0;
      std::vector<char> touched(h.size(), 0);
      for (int f = 0; f < NF; f++) for (int i2 = 0; i2 < c.b2; i2++) for (int i1 = 0; i1 < c.b1; i1++) for (int i0 = 0; i0 < c.b0; i0++) {
        long long dense = (((long long)f * c.b2 + i2) * c.b1 + i1) * c.b0 + i0;
        long long byte = dense * 8;
        if (c.sw == CU_TENSOR_MAP_SWIZZLE_64B) byte ^= ((byte >> 7) & 3) << 4;
        if (c.sw == CU_TENSOR_MAP_SWIZZLE_32B) byte ^= ((byte >> 7) & 1) << 4;
        if (c.sw == CU_TENSOR_MAP_SWIZZLE_128B) byte ^= ((byte >> 7) & 7) << 4;
        const int g0 = c.c0 + i0, g1 = c.c1 + i1, g2 = c.c2 + i2;
        const bool in = g0 >= 0 && g0 < P0 && g1 >= 0 && g1 < P1 && g2 >= 0 && g2 < P2;
        const long long gi = f * npg + g0 + (long long)P0 * (g1 + (long long)P1 * g2);
        const double want = in ? h[gi] : 0.0;
        if ((stage & 1) && dump[byte / 8] != want) bad_load++;
        if (in) {
          touched[gi] = 1;
          const double ws = (mode == 1) ? ((stage & 1) ? want : 0.0) + 1000.0 : o0[gi] + 0.5;
          if ((stage & 2) && o1[gi] != ws) bad_store++;
        }
      }
      for (size_t i = 0; i < o1.size(); i++) if (!touched[i] && o1[i] != o0[i]) bad_store++;
      printf("%s %-52s mode %s: load mismatches %d, store mismatches %d\n", (bad_load || bad_store) ? "FAIL" : "ok  ", c.name,
             mode == 1 ? "store " : "reduce", bad_load, bad_store);
      fails += (bad_load || bad_store) ? 1 : 0;
    }
  }
  printf(fails ? "TMA PROBE: FAIL\n" : "TMA PROBE: all ok\n");
  return fails ? 2 : 0;
}
