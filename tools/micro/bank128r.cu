// Random-row version of bank128.cu: prints measured cycles next to the wavefront counts two models predict.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
__global__ void k(const int* rows, int npat, float* out, long long* cyc) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 0, 0, 0);
  __syncthreads();
  for (int p = 0; p < npat; ++p) {
    int r = rows[p * 32 + threadIdx.x];
    float acc = 0;
    __syncwarp();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 4096; ++it) {
      float4 v = sm[r];
      acc += v.y;
      r += __float_as_int(v.z);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[p] = t1 - t0;
    out[p * 32 + threadIdx.x] = acc + r;
  }
}
int main() {
  const int NP = 24;
  static int h[NP][32];
  srand(7);
  for (int p = 0; p < NP; ++p)
    for (int l = 0; l < 32; ++l) {
      bool dup;
      do {
        h[p][l] = rand() % 2600;
        dup = false;
        for (int j = 0; j < l; ++j) dup |= h[p][j] == h[p][l];
      } while (dup);
    }
  int* d; float* o; long long* c;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, NP * 32 * 4); cudaMalloc(&c, NP * 8);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<<<1, 32, 65536>>>(d, NP, o, c);
  static long long hc[NP];
  cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  for (int p = 0; p < NP; ++p) {
    int cnt[8] = {0}, wq = 0, wh = 0;
    for (int l = 0; l < 32; ++l) cnt[h[p][l] % 8]++;
    int wg = std::max(4, *std::max_element(cnt, cnt + 8));
    for (int q = 0; q < 4; ++q) { int c8[8] = {0}; for (int l = 8 * q; l < 8 * q + 8; ++l) c8[h[p][l] % 8]++; wq += *std::max_element(c8, c8 + 8); }
    for (int q = 0; q < 2; ++q) { int c8[8] = {0}; for (int l = 16 * q; l < 16 * q + 16; ++l) c8[h[p][l] % 8]++; wh += std::max(2, *std::max_element(c8, c8 + 8)); }
    printf("pattern %2d: %.2f cycles  global=%d quarter=%d half=%d  counts", p, hc[p] / 4096.0, wg, wq, wh);
    for (int i = 0; i < 8; ++i) printf(" %d", cnt[i]);
    printf("\n");
  }
  return 0;
}
