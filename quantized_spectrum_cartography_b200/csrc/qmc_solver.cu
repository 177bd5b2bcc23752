// Fused factor update of the alternating solver: gradient of the per-map Frobenius regulariser, Adam
// moments and step, projection onto the non-negative orthant, and the squared norm of the updated
// factor for the next iteration -- one pass over the factor instead of the ~25 elementwise launches the
// reference's torch.optim.Adam + torch.norm + masked assignment amount to
// (qmc/qmc.ipynb c1:126-128 optimisers, :151/:209 cost = nll + lambda*||.||_F, :153-157/:211-212 step and
// `C[C<0] = 0`).  Arithmetic follows torch.optim.Adam (no amsgrad, no weight decay) operation by
// operation in fp32.
#include "qmc_common.cuh"

namespace qmc {

constexpr int UPD_THREADS = 256;
constexpr int UPD_ITEMS = 4;  // float4 per thread

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double ws[UPD_THREADS / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < UPD_THREADS / 32) t = ws[threadIdx.x];
  if (threadIdx.x < 32) t = warp_sum(t);
  return t;  // valid in thread 0
}

struct AdamParams {
  float lr, beta1, beta2, eps, lam;
  int step;
  const int32_t* step_dev;
  int project;
};

template <bool VEC>
__global__ void __launch_bounds__(UPD_THREADS) adam_frob_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                                float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                                const double* __restrict__ sumsq_in,
                                                                double* __restrict__ sumsq_out, const AdamParams a) {
  const int b = blockIdx.y;
  const int t = a.step + (a.step_dev ? *a.step_dev : 0);
  const float bc1 = 1.0f - powf(a.beta1, (float)t), bc2 = 1.0f - powf(a.beta2, (float)t);
  const float step_size = a.lr / bc1, bc2_sqrt = 1.0f / sqrtf(bc2);  // reciprocal: see adam_one
  const double ss = sumsq_in ? sumsq_in[b] : 0.0;
  const float nrm = (float)sqrt(ss);
  const float coef = (a.lam != 0.0f && nrm > 0.0f) ? a.lam / nrm : 0.0f;
  const float one_m_b1 = 1.0f - a.beta1, one_m_b2 = 1.0f - a.beta2;
  const bool project = a.project != 0;
  const int64_t base = (int64_t)b * n;
  double acc = 0.0;
  if (VEC) {
    const int64_t nv = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p + base);
    const float4* g4 = reinterpret_cast<const float4*>(g + base);
    float4* m4 = reinterpret_cast<float4*>(m + base);
    float4* v4 = reinterpret_cast<float4*>(v + base);
#pragma unroll
    for (int j = 0; j < UPD_ITEMS; ++j) {
      const int64_t i = ((int64_t)blockIdx.x * UPD_ITEMS + j) * UPD_THREADS + threadIdx.x;
      if (i < nv) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = g4[i];
        pp.x = adam_one(pp.x, gg.x, mm.x, vv.x, coef, one_m_b1, a.beta2, one_m_b2, step_size, bc2_sqrt, a.eps, project);
        pp.y = adam_one(pp.y, gg.y, mm.y, vv.y, coef, one_m_b1, a.beta2, one_m_b2, step_size, bc2_sqrt, a.eps, project);
        pp.z = adam_one(pp.z, gg.z, mm.z, vv.z, coef, one_m_b1, a.beta2, one_m_b2, step_size, bc2_sqrt, a.eps, project);
        pp.w = adam_one(pp.w, gg.w, mm.w, vv.w, coef, one_m_b1, a.beta2, one_m_b2, step_size, bc2_sqrt, a.eps, project);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
        acc += (double)pp.x * pp.x + (double)pp.y * pp.y + (double)pp.z * pp.z + (double)pp.w * pp.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < UPD_ITEMS * 4; ++j) {
      const int64_t i = ((int64_t)blockIdx.x * UPD_ITEMS * 4 + j) * UPD_THREADS + threadIdx.x;
      if (i < n) {
        float pp = p[base + i], mm = m[base + i], vv = v[base + i];
        pp = adam_one(pp, g[base + i], mm, vv, coef, one_m_b1, a.beta2, one_m_b2, step_size, bc2_sqrt, a.eps, project);
        p[base + i] = pp; m[base + i] = mm; v[base + i] = vv;
        acc += (double)pp * pp;
      }
    }
  }
  if (sumsq_out) {
    const double tot = block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(sumsq_out + b, tot);
  }
}

__global__ void __launch_bounds__(UPD_THREADS) sumsq_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t base = (int64_t)b * n;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * UPD_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * UPD_THREADS) {
    const float t = x[base + i];
    acc += (double)t * t;
  }
  const double tot = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out + b, tot);
}

__global__ void counter_add_kernel(int32_t* ctr, int add) { *ctr += add; }

}  // namespace qmc

using namespace qmc;

extern "C" int qmc_sumsq_per_map(const float* x_dev, int B, int64_t n_per_map, double* out_dev, void* stream) {
  QMC_REQUIRE(x_dev && out_dev, "null argument");
  QMC_REQUIRE(B > 0 && B <= 65535 * 1 && n_per_map > 0, "bad sizes B=%d n=%lld", B, (long long)n_per_map);
  cudaStream_t st = (cudaStream_t)stream;
  QMC_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, sizeof(double) * B, st));
  int64_t bx = (n_per_map + UPD_THREADS * 8 - 1) / (UPD_THREADS * 8);
  if (bx > 64) bx = 64;
  sumsq_kernel<<<dim3((unsigned)bx, (unsigned)B), UPD_THREADS, 0, st>>>(x_dev, n_per_map, out_dev);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

extern "C" int qmc_adam_frob_project(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, int B,
                                     int64_t n_per_map, const double* sumsq_in_dev, double* sumsq_out_dev, float lr,
                                     float beta1, float beta2, float eps, float lam, int project, int step,
                                     const int32_t* step_dev, void* stream) {
  QMC_REQUIRE(p_dev && g_dev && m_dev && v_dev, "null argument");
  QMC_REQUIRE(B > 0 && B <= 65535 && n_per_map > 0, "bad sizes B=%d n=%lld", B, (long long)n_per_map);
  QMC_REQUIRE(lam == 0.0f || sumsq_in_dev, "the Frobenius regulariser needs the squared norms of the factor");
  QMC_REQUIRE(step >= 0 && (step > 0 || step_dev), "Adam steps count from 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (sumsq_out_dev) QMC_CUDA_CHECK(cudaMemsetAsync(sumsq_out_dev, 0, sizeof(double) * B, st));
  AdamParams a{lr, beta1, beta2, eps, lam, step, step_dev, project};
  const bool vec = (n_per_map % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(p_dev) | reinterpret_cast<uintptr_t>(g_dev) |
                     reinterpret_cast<uintptr_t>(m_dev) | reinterpret_cast<uintptr_t>(v_dev)) & 15) == 0;
  const int64_t per_block = (int64_t)UPD_THREADS * UPD_ITEMS * 4;  // floats per block
  const int64_t bx = (n_per_map + per_block - 1) / per_block;
  QMC_REQUIRE(bx <= 0x7fffffff, "map too large");
  dim3 grid((unsigned)bx, (unsigned)B);
  if (vec) adam_frob_kernel<true><<<grid, UPD_THREADS, 0, st>>>(p_dev, g_dev, m_dev, v_dev, n_per_map, sumsq_in_dev, sumsq_out_dev, a);
  else adam_frob_kernel<false><<<grid, UPD_THREADS, 0, st>>>(p_dev, g_dev, m_dev, v_dev, n_per_map, sumsq_in_dev, sumsq_out_dev, a);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

extern "C" int qmc_counter_add(int32_t* counter_dev, int add, void* stream) {
  QMC_REQUIRE(counter_dev, "null argument");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, add);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}
