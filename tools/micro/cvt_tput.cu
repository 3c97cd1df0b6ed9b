// Throughput of the fp32 -> packed fp16 conversion (cvt.rn.f16x2.f32 = F2FP) vs FFMA vs MUFU.EX2, one SM.
#include <cstdio>
#include <cuda_fp16.h>
template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float a = threadIdx.x * 1e-3f + 0.1f, b = a + 1.f, c = a + 2.f, d = a + 3.f;
  unsigned u0 = 0, u1 = 0, u2 = 0, u3 = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) {  // F2FP: 4 packs (8 elements)
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u0) : "f"(a), "f"(b));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u1) : "f"(c), "f"(d));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u2) : "f"(b), "f"(c));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u3) : "f"(d), "f"(a));
      a += __uint_as_float(u0 & 1); 
    } else if (OP == 1) {  // FFMA x4
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b) : "f"(c), "f"(d));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(c) : "f"(d), "f"(a));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(d) : "f"(a), "f"(b));
    } else {  // MUFU x4
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = a + b + c + d + __uint_as_float(u0 ^ u1 ^ u2 ^ u3);
}
int main() {
  float* o; long long* c; long long h;
  cudaMalloc(&o, 8192); cudaMalloc(&c, 8);
  const int iters = 8192, threads = 512;
  k<0><<<1, threads>>>(o, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("F2FP  : %.2f warp-instr-lanes / cycle / SM  (%.2f packs/clk)\n", 4.0 * iters * threads / h, 4.0 * iters * threads / h);
  k<1><<<1, threads>>>(o, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("FFMA  : %.2f lanes / cycle / SM\n", 4.0 * iters * threads / h);
  k<2><<<1, threads>>>(o, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("MUFU  : %.2f lanes / cycle / SM\n", 4.0 * iters * threads / h);
  return 0;
}
