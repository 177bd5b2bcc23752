// Micro-benchmark: how does the shared-memory pipe split a warp-wide 128-bit access into wavefronts?
// Each pattern gives every lane a 16-byte row index; the kernel times a dependent chain of LDS.128
// with clock64 and prints cycles per access (4 wavefronts = conflict-free).
#include <cstdio>
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
      acc += v.y;                      // v.y == 0: keeps the chain dependent without changing r
      r += __float_as_int(v.z);        // + 0
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[p] = t1 - t0;
    out[p * 32 + threadIdx.x] = acc + r;
  }
}
int main() {
  const int NP = 8;
  int h[NP][32];
  for (int l = 0; l < 32; ++l) {
    h[0][l] = l;                                   // A: conflict-free
    h[1][l] = 8 * (l % 8) + l / 8;                 // B: quarter q all == q mod 8 (8-way inside consecutive-8 quarters; 4 residues overall)
    h[2][l] = 8 * l;                               // C: all same bank group (32-way)
    h[3][l] = (l / 8) * 8 + ((l * 5) % 8);         // D: distinct residues per consecutive quarter, shuffled
    h[4][l] = ((l / 8) * 2 + (l % 8) / 4) + 8 * (l % 4) + 64 * (l / 8);  // E: consecutive quarter q holds residues {2q,2q+1} 4x each; globally each residue 4x
    h[5][l] = (l % 4) * 8 + (l / 4);               // F: lanes l, l+4, l+8.. : stride-4 groups {l%4 fixed} share residue? rows = (l%4)*8 + l/4: residue = l/4 (0..7), 4 lanes each: consecutive lanes 4j..4j+3 same residue
    h[6][l] = (l % 16 < 8) ? l % 8 : 8 + (l % 8) * 8;  // G: first half of each half-warp conflict-free, second half 8-way
    h[7][l] = l % 8;                               // H: 4 lanes per row, same address (broadcast)
  }
  int* d; float* o; long long* c;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, NP * 32 * 4); cudaMalloc(&c, NP * 8);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<<<1, 32, 65536>>>(d, NP, o, c);
  long long hc[NP];
  cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  const char* names = "ABCDEFGH";
  for (int p = 0; p < NP; ++p) printf("pattern %c: %.2f cycles/access\n", names[p], hc[p] / 4096.0);
  return 0;
}
