// C-ABI plumbing of libqmc_b200.so: version, thread-local error string, launch counter, the
// host-buffer end-to-end entry point, and the small dense helpers (get_tensor, NMSE terms).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "qmc_common.cuh"

namespace qmc {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// X[b][k][p] = sum_r S[b][r][p] * C[b][r][k]   (get_tensor, quantization_model.py:79-86; same
// left-to-right order over r as the reference, but fused multiply-adds)
template <int RP>
__global__ void get_tensor_kernel(const float* __restrict__ S, const float* __restrict__ C, int IJ, int K,
                                  int R, float* __restrict__ X) {
  const int b = blockIdx.z, k = blockIdx.y;
  const float* Sb = S + (int64_t)b * R * IJ;
  const float* Cb = C + (int64_t)b * R * K;
  float c[RP];
#pragma unroll
  for (int r = 0; r < RP; ++r) c[r] = r < R ? Cb[r * K + k] : 0.0f;
  float* Xr = X + ((int64_t)b * K + k) * IJ;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < IJ; p += gridDim.x * blockDim.x) {
    float t = 0.0f;
#pragma unroll
    for (int r = 0; r < RP; ++r)
      if (r < R) t = fmaf(Sb[(int64_t)r * IJ + p], c[r], t);
    Xr[p] = t;
  }
}

// per map: sum (xhat - xref)^2 and sum xref^2 with xhat formed on the fly
template <int RP>
__global__ void nmse_terms_kernel(const float* __restrict__ S, const float* __restrict__ C,
                                  const float* __restrict__ Xref, int IJ, int K, int R, int log_domain,
                                  float offset, double* __restrict__ out) {
  const int b = blockIdx.z, k = blockIdx.y;
  const float* Sb = S + (int64_t)b * R * IJ;
  const float* Cb = C + (int64_t)b * R * K;
  float c[RP];
#pragma unroll
  for (int r = 0; r < RP; ++r) c[r] = r < R ? Cb[r * K + k] : 0.0f;
  const float* Xr = Xref + ((int64_t)b * K + k) * IJ;
  double num = 0.0, den = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < IJ; p += gridDim.x * blockDim.x) {
    float t = 0.0f;
#pragma unroll
    for (int r = 0; r < RP; ++r)
      if (r < R) t = fmaf(Sb[(int64_t)r * IJ + p], c[r], t);
    float ref = Xr[p];
    if (log_domain) {
      t = logf(t + offset);
      ref = logf(ref + offset);
    }
    const float d = t - ref;
    num += (double)d * d;
    den += (double)ref * ref;
  }
  num = warp_sum(num);
  den = warp_sum(den);
  __shared__ double sn[8], sd[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sn[warp] = num; sd[warp] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, d = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sn[w]; d += sd[w]; }
    atomicAdd(out + 2 * b, a);
    atomicAdd(out + 2 * b + 1, d);
  }
}

}  // namespace qmc

using namespace qmc;

extern "C" int qmc_abi_version(void) { return QMC_ABI_VERSION; }
extern "C" const char* qmc_last_error(void) { return g_err; }
extern "C" int64_t qmc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

#define QMC_DISPATCH_RP(R, CALL)                                      \
  do {                                                                \
    if ((R) <= 1) { constexpr int RP = 1; CALL; }                     \
    else if ((R) <= 2) { constexpr int RP = 2; CALL; }                \
    else if ((R) <= 4) { constexpr int RP = 4; CALL; }                \
    else if ((R) <= 8) { constexpr int RP = 8; CALL; }                \
    else if ((R) <= 16) { constexpr int RP = 16; CALL; }              \
    else { constexpr int RP = 32; CALL; }                             \
  } while (0)

extern "C" int qmc_get_tensor(const float* S_dev, const float* C_dev, int B, int IJ, int K, int R,
                              float* X_out_dev, void* stream) {
  QMC_REQUIRE(S_dev && C_dev && X_out_dev, "null argument");
  QMC_REQUIRE(B > 0 && IJ > 0 && K > 0 && R > 0 && R <= QMC_MAX_RANK && K <= 65535 && B <= 65535, "bad sizes");
  int gx = (IJ + 255) / 256;
  if (gx > 1024) gx = 1024;
  dim3 grid(gx, K, B);
  QMC_DISPATCH_RP(R, (get_tensor_kernel<RP><<<grid, 256, 0, (cudaStream_t)stream>>>(S_dev, C_dev, IJ, K, R, X_out_dev)));
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

extern "C" int qmc_nmse_terms(const float* S_dev, const float* C_dev, const float* X_ref_dev, int B, int IJ,
                              int K, int R, int log_domain, float offset, double* out_dev, void* stream) {
  QMC_REQUIRE(S_dev && C_dev && X_ref_dev && out_dev, "null argument");
  QMC_REQUIRE(B > 0 && IJ > 0 && K > 0 && R > 0 && R <= QMC_MAX_RANK && K <= 65535 && B <= 65535, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  QMC_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, sizeof(double) * 2 * B, st));
  int gx = (IJ + 255) / 256;
  if (gx > 64) gx = 64;
  dim3 grid(gx, K, B);
  QMC_DISPATCH_RP(R, (nmse_terms_kernel<RP><<<grid, 256, 0, st>>>(S_dev, C_dev, X_ref_dev, IJ, K, R, log_domain,
                                                                    offset, out_dev)));
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

extern "C" int qmc_nll_fwd_bwd_gather_host(const float* S_host, const float* C_host, float* S_scratch_dev,
                                           float* C_scratch_dev, const qmc_obs_view_t* obs,
                                           const qmc_likelihood_t* lik, int B, int IJ, int K, int R, int algo,
                                           int tile_warps, double* nll_scratch_dev, float* gS_scratch_dev,
                                           float* gC_scratch_dev, double* nll_host, float* gS_host,
                                           float* gC_host, void* stream) {
  QMC_REQUIRE(S_host && C_host && S_scratch_dev && C_scratch_dev && nll_scratch_dev && nll_host && lik, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nS = sizeof(float) * (size_t)B * R * IJ, nC = sizeof(float) * (size_t)B * R * K;
  QMC_CUDA_CHECK(cudaMemcpyAsync(S_scratch_dev, S_host, nS, cudaMemcpyHostToDevice, st));
  QMC_CUDA_CHECK(cudaMemcpyAsync(C_scratch_dev, C_host, nC, cudaMemcpyHostToDevice, st));
  const int rc = qmc_nll_fwd_bwd_gather(S_scratch_dev, (int64_t)R * IJ, IJ, 1, C_scratch_dev, obs, lik, B, IJ, K, R,
                                        algo, tile_warps, nll_scratch_dev, gS_scratch_dev, gC_scratch_dev, stream);
  if (rc != QMC_OK) return rc;
  QMC_CUDA_CHECK(cudaMemcpyAsync(nll_host, nll_scratch_dev, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
  if (!(lik->flags & QMC_FORWARD_ONLY)) {
    QMC_REQUIRE(gS_host && gC_host, "null gradient host buffers");
    QMC_CUDA_CHECK(cudaMemcpyAsync(gS_host, gS_scratch_dev, nS, cudaMemcpyDeviceToHost, st));
    QMC_CUDA_CHECK(cudaMemcpyAsync(gC_host, gC_scratch_dev, nC, cudaMemcpyDeviceToHost, st));
  }
  QMC_CUDA_CHECK(cudaStreamSynchronize(st));
  return QMC_OK;
}
