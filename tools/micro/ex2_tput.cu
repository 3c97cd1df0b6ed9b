// MUFU.EX2 throughput: fp32 ex2.approx vs packed ex2.approx.f16x2 (standalone microbenchmark, nvcc -arch=sm_100a)
#include <cstdio>
#include <cuda_fp16.h>
__global__ void k32(float* out, int iters, long long* cyc) {
  float a = threadIdx.x * 1e-3f, b = a + 1.f, c = a + 2.f, d = a + 3.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = a + b + c + d;
}
__global__ void k16(unsigned* out, int iters, long long* cyc) {
  unsigned a = 0x38003400u + threadIdx.x, b = a + 7, c = a + 11, d = a + 13;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(c));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(d));
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = a ^ b ^ c ^ d;
}
int main() {
  float* o; unsigned* o2; long long* c; long long h;
  cudaMalloc(&o, 4096); cudaMalloc(&o2, 4096); cudaMalloc(&c, 8);
  for (int threads : {128, 256, 512}) {
    int iters = 4096;
    k32<<<1, threads>>>(o, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("f32   threads=%d: %.2f ex2 elements / cycle / SM\n", threads, 4.0 * iters * threads / h);
    k16<<<1, threads>>>(o2, iters, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("f16x2 threads=%d: %.2f ex2 elements / cycle / SM\n", threads, 8.0 * iters * threads / h);
  }
  return 0;
}
