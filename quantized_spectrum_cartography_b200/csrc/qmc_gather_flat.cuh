// Flat observed-entry kernel: one thread per entry, factors through L2, warp-aggregated global atomics.
#pragma once
#include "qmc_gather_common.cuh"

namespace qmc {

// ------------------------------------------------------------------------------------------------
// flat kernel
// ------------------------------------------------------------------------------------------------
template <int RP, int EPI, bool LOGD, bool GRAD>
__global__ void __launch_bounds__(256) gather_flat_kernel(const GatherParams prm) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int64_t rows_per_map = (int64_t)prm.n_sub * prm.K;
  const int64_t beg = prm.row_off[b * rows_per_map];
  const int64_t end = prm.row_off[(b + 1) * rows_per_map];
  const float* __restrict__ Sb = prm.S + b * prm.sB;
  const float* __restrict__ Cb = prm.C + (int64_t)b * prm.R * prm.K;
  float* gSb = GRAD ? prm.gS + b * prm.sB : nullptr;
  float* gCb = GRAD ? prm.gC + (int64_t)b * prm.R * prm.K : nullptr;

  float nll_part = 0.0f;
  // whole warps iterate together so the shuffles below always see 32 lanes
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = beg + (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < end; base += stride) {
    const int64_t i = base + lane;
    const bool valid = i < end;
    int k = 0, p = 0, lv = 0;
    if (valid) {
      const int id = prm.idx[i];
      lv = prm.lvl[i];
      k = fast_div((uint32_t)id, prm.div_magic, prm.div_shift);
      p = id - k * prm.IJ;
    }
    float s[RP], c[RP];
    float t = 0.0f;
#pragma unroll
    for (int r = 0; r < RP; ++r) {
      const bool on = valid && r < prm.R;
      s[r] = on ? __ldg(Sb + r * prm.sR + p * prm.sP) : 0.0f;
      c[r] = on ? __ldg(Cb + r * prm.K + k) : 0.0f;
      t = fmaf(s[r], c[r], t);
    }
    float g = 0.0f;
    if (valid) {
      float dxdt;
      const BinEval ev = eval_entry<EPI, LOGD>(prm, t, lv, dxdt);
      nll_part -= ev.logp;
      g = ev.gx * dxdt;
    }
    if (GRAD) {
      // gS: scattered pixels, one atomic per (entry, r)
      if (valid) {
#pragma unroll
        for (int r = 0; r < RP; ++r)
          if (r < prm.R) atomicAdd(gSb + r * prm.sR + p * prm.sP, g * c[r]);
      }
      // gC: entries are band-sorted, so a warp usually sees one band: aggregate, one atomic per r
      const int k0 = __shfl_sync(0xffffffffu, k, 0);
      const bool uniform = __all_sync(0xffffffffu, !valid || k == k0);
      if (uniform) {
        float v[RP];
#pragma unroll
        for (int r = 0; r < RP; ++r) v[r] = g * s[r];
        const float tot = warp_transpose_sum<RP>(v, lane);
        const int r_own = warp_transpose_owner<RP>(lane);
        if ((lane & (32 / RP - 1)) == 0 && r_own < prm.R) atomicAdd(gCb + r_own * prm.K + k0, tot);
      } else if (valid) {
#pragma unroll
        for (int r = 0; r < RP; ++r)
          if (r < prm.R) atomicAdd(gCb + r * prm.K + k, g * s[r]);
      }
    }
  }
  // NLL: fp32 per thread (a handful of terms), fp64 from the warp level up
  double w = warp_sum((double)nll_part);
  __shared__ double wsum[8];
  if (lane == 0) wsum[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += wsum[i];
    if (tot != 0.0 || (blockIdx.x == 0)) atomicAdd(prm.nll + b, tot);
  }
}

template <int RP>
int launch_flat_rp(const GatherParams& prm, int epi, bool logd, bool grad, cudaStream_t st) {
  // size the grid from the average entries per map; the kernel is grid-stride
  const int threads = 256;
  int64_t per_map_guess = (int64_t)prm.K * prm.IJ;  // upper bound; the loop exits early
  int64_t want = (per_map_guess + threads - 1) / threads;
  int bx = (int)(want < 148 * 8 ? want : 148 * 8);
  if (bx < 1) bx = 1;
  dim3 grid(bx, prm.B);
#define QMC_GO(E, L, G) gather_flat_kernel<RP, E, L, G><<<grid, threads, 0, st>>>(prm)
  QMC_GATHER_SWITCH(QMC_GO);
#undef QMC_GO
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

}  // namespace qmc
