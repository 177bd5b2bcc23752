// C-ABI plumbing of libqmc_b200.so: version, thread-local error string, launch counter, the
// host-buffer end-to-end entry point, and the small dense helpers (get_tensor, NMSE terms).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include <mutex>

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

// One-bit BCE form of the likelihood (NegLikelihood, quantization_model.py:97-113): per element
//   p = F_probit(x - mean, std) = 0.5*(1 + erf((x - mean)/(std*1.414213)))   or   F_sigmoid(x - mean),
//   loss = -(t*max(log p, -100) + (1 - t)*max(log(1 - p), -100)),  mean over all elements (nn.BCELoss),
// with the reference's own fp32 arithmetic (p formed first, then the logarithms -- including its saturation to
// -100 in the tails), and the gradient torch's BCELoss backward produces: (p - t)/max(p(1 - p), 1e-12) * dp/dx / n.
__global__ void bce_one_bit_kernel(const float* __restrict__ x, const float* __restrict__ target, int64_t n, float mean,
                                   float inv_a, int probit, double inv_n, double* __restrict__ loss, float* __restrict__ gx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float y = x[i] - mean, t = target[i];
    float p, dpdx;
    if (probit) {
      const float z = y * inv_a;
      p = 0.5f * (1.0f + erff(z));
      dpdx = kInvSqrtPi * inv_a * expf(-z * z);
    } else {
      p = 1.0f / (1.0f + expf(-y));
      dpdx = p * (1.0f - p);
    }
    const float lp = fmaxf(logf(p), -100.0f), lq = fmaxf(logf(1.0f - p), -100.0f);
    acc -= (double)(t * lp + (1.0f - t) * lq);
    if (gx) gx[i] = (float)((double)((p - t) / fmaxf((1.0f - p) * p, 1e-12f) * dpdx) * inv_n);
  }
  acc = warp_sum(acc);
  __shared__ double sacc[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sacc[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sacc[w];
    atomicAdd(loss, a * inv_n);
  }
}

}  // namespace qmc

using namespace qmc;

extern "C" int qmc_bce_one_bit(const float* x_dev, const float* target_dev, int64_t n, float mean, float noise_std,
                               int probit, double* loss_out_dev, float* gx_out_dev, void* stream) {
  QMC_REQUIRE(x_dev && target_dev && loss_out_dev && n > 0, "bad arguments");
  QMC_REQUIRE(!probit || noise_std > 0.0f, "the probit form needs a positive std");
  cudaStream_t st = (cudaStream_t)stream;
  QMC_CUDA_CHECK(cudaMemsetAsync(loss_out_dev, 0, sizeof(double), st));
  int64_t g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  bce_one_bit_kernel<<<(unsigned)g, 256, 0, st>>>(x_dev, target_dev, n, mean, probit ? 1.0f / probit_scale(noise_std) : 1.0f,
                                                   probit, 1.0 / (double)n, loss_out_dev, gx_out_dev);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

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

// Host-buffer entry: the batch is cut into chunks of maps that flow through a few internal streams, so
// that the host->device copy of one chunk, the kernel of the previous one and the device->host copy of
// the one before overlap (PCIe is full duplex; the kernel is a small fraction of either copy).
namespace {
constexpr int kPipeStreams = 3;
struct HostPipe {  // per device: copy/compute streams and the event that orders them after the caller's
  std::mutex mu;   // calls on one device are serialised; different devices proceed independently
  bool ready = false;
  cudaStream_t s[kPipeStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t start = nullptr;
};
constexpr int kMaxDevices = 64;
HostPipe g_pipes[kMaxDevices];
}  // namespace

extern "C" int qmc_nll_fwd_bwd_gather_host(const float* S_host, const float* C_host, float* S_scratch_dev,
                                           float* C_scratch_dev, const qmc_obs_view_t* obs,
                                           const qmc_likelihood_t* lik, int B, int IJ, int K, int R, int algo,
                                           int tile_warps, double* nll_scratch_dev, float* gS_scratch_dev,
                                           float* gC_scratch_dev, double* nll_host, float* gS_host,
                                           float* gC_host, void* stream) {
  QMC_REQUIRE(S_host && C_host && S_scratch_dev && C_scratch_dev && nll_scratch_dev && nll_host && lik && obs, "null argument");
  const bool grad = !(lik->flags & QMC_FORWARD_ONLY);
  QMC_REQUIRE(!grad || (gS_host && gC_host && gS_scratch_dev && gC_scratch_dev), "null gradient buffers");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  QMC_CUDA_CHECK(cudaGetDevice(&dev));
  QMC_REQUIRE(dev >= 0 && dev < kMaxDevices, "device ordinal %d out of range", dev);
  HostPipe& g_pipe = g_pipes[dev];
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  if (!g_pipe.ready) {
    for (auto& x : g_pipe.s) QMC_CUDA_CHECK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    QMC_CUDA_CHECK(cudaEventCreateWithFlags(&g_pipe.start, cudaEventDisableTiming));
    g_pipe.ready = true;
  }
  QMC_CUDA_CHECK(cudaEventRecord(g_pipe.start, st));
  for (auto& x : g_pipe.s) QMC_CUDA_CHECK(cudaStreamWaitEvent(x, g_pipe.start, 0));

  // enough chunks that the fill and drain of the pipeline (one chunk's copy in, one chunk's copy out) are a small
  // part of the whole, few enough that a chunk's copies stay long (measured on the B200 host: 8 chunks on 3 streams
  // 4.46 ms per cfg3 step, 16 chunks on 4 streams 4.84 ms)
  const int n_chunks = B >= 512 ? 8 : (B >= 64 ? 4 : 1);
  const int per = (B + n_chunks - 1) / n_chunks;
  int rc_all = QMC_OK;
  const size_t sS = (size_t)R * IJ, sC = (size_t)R * K;
  const int64_t streams_per_map = obs->n_sub, rows_per_map = (int64_t)obs->n_sub * K;
  for (int c = 0, b0 = 0; b0 < B; ++c, b0 += per) {
    const int nb = (B - b0) < per ? (B - b0) : per;
    cudaStream_t cs = g_pipe.s[c % kPipeStreams];
    // a failure must not return while earlier chunks' copies into the caller's buffers are still in flight
#define QMC_PIPE_CHECK(expr)                                                                      \
    do {                                                                                          \
      cudaError_t _e = (expr);                                                                    \
      if (_e != cudaSuccess) {                                                                    \
        rc_all = set_error(QMC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));         \
        goto drain;                                                                               \
      }                                                                                           \
    } while (0)
    QMC_PIPE_CHECK(cudaMemcpyAsync(S_scratch_dev + b0 * sS, S_host + b0 * sS, sizeof(float) * nb * sS, cudaMemcpyHostToDevice, cs));
    QMC_PIPE_CHECK(cudaMemcpyAsync(C_scratch_dev + b0 * sC, C_host + b0 * sC, sizeof(float) * nb * sC, cudaMemcpyHostToDevice, cs));
    qmc_obs_view_t v = *obs;  // the observation arrays of maps b0.. (offsets stored in them are absolute)
    if (v.row_off_dev) v.row_off_dev += b0 * rows_per_map;
    if (v.nrows_dev) v.nrows_dev += b0 * streams_per_map;
    if (v.words_dev && v.stream_stride > 0) v.words_dev += b0 * streams_per_map * v.stream_stride;
    else if (v.stream_off_dev) v.stream_off_dev += b0 * streams_per_map;
    const int rc = qmc_nll_fwd_bwd_gather(S_scratch_dev + b0 * sS, (int64_t)sS, IJ, 1, C_scratch_dev + b0 * sC, &v, lik, nb,
                                          IJ, K, R, algo, tile_warps, nll_scratch_dev + b0,
                                          grad ? gS_scratch_dev + b0 * sS : nullptr, grad ? gC_scratch_dev + b0 * sC : nullptr, cs);
    if (rc != QMC_OK) { rc_all = rc; goto drain; }
    QMC_PIPE_CHECK(cudaMemcpyAsync(nll_host + b0, nll_scratch_dev + b0, sizeof(double) * nb, cudaMemcpyDeviceToHost, cs));
    if (grad) {
      QMC_PIPE_CHECK(cudaMemcpyAsync(gS_host + b0 * sS, gS_scratch_dev + b0 * sS, sizeof(float) * nb * sS, cudaMemcpyDeviceToHost, cs));
      QMC_PIPE_CHECK(cudaMemcpyAsync(gC_host + b0 * sC, gC_scratch_dev + b0 * sC, sizeof(float) * nb * sC, cudaMemcpyDeviceToHost, cs));
    }
  }
#undef QMC_PIPE_CHECK
drain:
  for (auto& x : g_pipe.s) {
    const cudaError_t e = cudaStreamSynchronize(x);
    if (e != cudaSuccess && rc_all == QMC_OK) rc_all = set_error(QMC_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
  }
  return rc_all;
}
