// Throughput version: 16 warps on one SM issue independent LDS.128 with a fixed per-lane row pattern;
// prints SM cycles per warp-instruction (= wavefronts if the pipe retires one wavefront per cycle).
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
__global__ void k(const int* rows, int pat, float* out, long long* cyc) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(0, 0, 0, 0);
  __syncthreads();
  const int r = rows[pat * 32 + (threadIdx.x & 31)];
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const unsigned base0 = (unsigned)__cvta_generic_to_shared(sm) + r * 16;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 2048; ++it) {
    float4 v0, v1, v2, v3;
    const unsigned base = base0 + ((it & 7) << 11);   // + multiples of 2 KB: same bank pattern, new address
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(base) : "memory");
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+128];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(base) : "memory");
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+256];" : "=f"(v2.x), "=f"(v2.y), "=f"(v2.z), "=f"(v2.w) : "r"(base) : "memory");
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+384];" : "=f"(v3.x), "=f"(v3.y), "=f"(v3.z), "=f"(v3.w) : "r"(base) : "memory");
    a0 += v0.x + v0.y + v0.z + v0.w; a1 += v1.x + v1.y + v1.z + v1.w; a2 += v2.x + v2.y + v2.z + v2.w; a3 += v3.x + v3.y + v3.z + v3.w;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = a0 + a1 + a2 + a3;
}
int main() {
  const int NP = 20;
  static int h[NP][32];
  srand(7);
  for (int l = 0; l < 32; ++l) {
    h[0][l] = l;
    h[1][l] = 8 * (l % 8) + l / 8;
    h[2][l] = ((l / 8) * 2 + (l % 8) / 4) + 8 * (l % 4) + 64 * (l / 8);
    h[3][l] = (l % 4) * 8 + (l / 4);
  }
  for (int p = 4; p < NP; ++p)
    for (int l = 0; l < 32; ++l) {
      bool dup;
      do {
        h[p][l] = rand() % 2000;
        dup = false;
        for (int j = 0; j < l; ++j) dup |= h[p][j] == h[p][l];
      } while (dup);
    }
  int* d; float* o; long long* c;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, 512 * 4); cudaMalloc(&c, 8);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  for (int p = 0; p < NP; ++p) {
    k<<<1, 512, 65536 + 1024>>>(d, p, o, c);
    long long hc;
    cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
    int cnt[8] = {0}, wq = 0;
    for (int l = 0; l < 32; ++l) cnt[h[p][l] % 8]++;
    int wg = std::max(4, *std::max_element(cnt, cnt + 8));
    for (int q = 0; q < 4; ++q) { int c8[8] = {0}; for (int l = 8 * q; l < 8 * q + 8; ++l) c8[h[p][l] % 8]++; wq += *std::max_element(c8, c8 + 8); }
    printf("pattern %2d: %.2f cycles/warp-instr  global=%d quarter=%d\n", p, hc / (2048.0 * 4 * 16), wg, wq);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
