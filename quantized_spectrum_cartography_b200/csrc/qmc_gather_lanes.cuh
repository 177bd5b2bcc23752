// Lane-stream observed-entry kernel: per-lane band walks, C row and gC accumulator in registers.
#pragma once
#include "qmc_gather_common.cuh"

namespace qmc {

// ------------------------------------------------------------------------------------------------
// lanes kernel
// ------------------------------------------------------------------------------------------------
// Same tiling and ownership as the tiled kernel (one CTA per (map, pixel tile), each warp owns a pixel
// sub-tile exclusively), on the lane-stream layout of qmc_obs_build_lanes: every lane of a warp walks
// the entries of ONE band at a time, so C[band] and the band's gC accumulator stay in registers (no
// shared-memory traffic, no warp reduction, one plain store per (warp, band)); the builder guarantees
// that the 32 entries of a step hit 32 different pixels, so the gS update is a plain shared-memory
// read-modify-write.  Per entry the shared-memory pipe sees one S row read and one gS row
// read-modify-write -- half of what the tiled/matched kernels need -- and the bank groups inside a
// quarter-warp are distinct wherever the builder had a choice.
// Shared memory (floats): Ssm[TP+32][RP] | gSsm[TP+32][RP] | Csm[K+1][RP] | gCw[W][K+1][RP]; the 32 extra
// pixel rows and the extra band row absorb padding words.
__device__ __forceinline__ uint4 ldg_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <int RP, int EPI, bool LOGD, bool GRAD>
__global__ void __launch_bounds__(256, 2) gather_lanes_kernel(const GatherParams prm) {
  extern __shared__ __align__(16) float smem[];
  const int W = prm.tile_warps, K = prm.K;
  const int TP = prm.sub_pixels * W;
  const int TPD = TP + 32;  // + dummy pixel rows
  float* Ssm = smem;
  float* gSsm = Ssm + (size_t)TPD * RP;
  float* Csm = gSsm + (GRAD ? (size_t)TPD * RP : 0);
  float* gCw = Csm + (size_t)(K + 1) * RP;
  __shared__ uint64_t mbar;
  __shared__ double wsum[8];

  const int b = blockIdx.x / prm.tiles_per_map;
  const int tile = blockIdx.x - b * prm.tiles_per_map;
  const int p0 = tile * TP;
  const int np = min(TP, prm.IJ - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nthr = blockDim.x;
  const float* __restrict__ Sb = prm.S + b * prm.sB;
  const float* __restrict__ Cb = prm.C + (int64_t)b * prm.R * K;
  const bool bulk = (prm.sR == 1 && prm.sP == RP && prm.R == RP && (RP % 4) == 0 &&
                     ((reinterpret_cast<uintptr_t>(Sb) | (GRAD ? reinterpret_cast<uintptr_t>(prm.gS + b * prm.sB) : 0)) & 15) == 0);
  if (bulk) {
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)np * RP * sizeof(float);
      mbar_expect_tx(&mbar, bytes);
      bulk_g2s(Ssm, Sb + (int64_t)p0 * RP, bytes, &mbar);
    }
  } else {
#pragma unroll
    for (int r = 0; r < RP; ++r)
      for (int pl = threadIdx.x; pl < np; pl += nthr)
        Ssm[pl * RP + r] = (r < prm.R) ? __ldg(Sb + r * prm.sR + (int64_t)(p0 + pl) * prm.sP) : 0.0f;
  }
  for (int i = np * RP + threadIdx.x; i < TPD * RP; i += nthr) Ssm[i] = 0.0f;  // tail + dummy rows
#pragma unroll
  for (int r = 0; r < RP; ++r)
    for (int k = threadIdx.x; k <= K; k += nthr) Csm[k * RP + r] = (r < prm.R && k < K) ? __ldg(Cb + r * K + k) : 0.0f;
  if (GRAD) {
    for (int i = threadIdx.x; i < TPD * RP; i += nthr) gSsm[i] = 0.0f;
    for (int i = threadIdx.x; i < W * (K + 1) * RP; i += nthr) gCw[i] = 0.0f;
  }
  const int64_t stream = (int64_t)b * prm.n_sub + (int64_t)tile * W + warp;
  const uint4* gp = reinterpret_cast<const uint4*>(prm.words + prm.stream_off[stream]) + lane;
  const int ngroups = prm.nrows[stream] >> 2;
  // two groups of look-ahead
  const uint32_t padw = (0xFFu << 24) | ((uint32_t)K << LW_BAND_SHIFT) | (uint32_t)(TP + lane);
  const uint4 padg = make_uint4(padw, padw, padw, padw);
  uint4 wa = ngroups > 0 ? ldg_u4(gp) : padg;
  uint4 wb = ngroups > 1 ? ldg_u4(gp + 32) : padg;
  __syncthreads();
  if (bulk) mbar_wait(&mbar, 0);

  constexpr uint32_t ROWB = RP * sizeof(float);
  const uint32_t S_a = smem_u32(Ssm), C_a = smem_u32(Csm);
  const uint32_t gS_delta = smem_u32(gSsm) - S_a;
  const uint32_t gC_delta = smem_u32(gCw + (size_t)warp * (K + 1) * RP) - C_a;

  float nll_part = 0.0f;
  float c[RP], acc[RP];
  uint32_t cur_key = wa.x & LW_BAND_MASK;  // band bits of the lane's current band
  auto load_band = [&](uint32_t key) {
    const uint32_t crow = C_a + (key >> LW_BAND_SHIFT) * ROWB;
    if (RP % 4 == 0) {
#pragma unroll
      for (int r = 0; r < RP; r += 4) {
        const float4 c4 = lds128_ro(crow + r * 4);
        c[r] = c4.x; c[r + 1] = c4.y; c[r + 2] = c4.z; c[r + 3] = c4.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < RP; ++r) c[r] = lds32_ro(crow + r * 4);
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) acc[r] = 0.0f;
  };
  auto flush_band = [&](uint32_t key) {  // one plain store per (warp, band): the band is this lane's alone
    if (!GRAD) return;
    const uint32_t grow = C_a + gC_delta + (key >> LW_BAND_SHIFT) * ROWB;
    if (RP % 4 == 0) {
#pragma unroll
      for (int r = 0; r < RP; r += 4) sts128(grow + r * 4, make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]));
    } else {
#pragma unroll
      for (int r = 0; r < RP; ++r) sts32(grow + r * 4, acc[r]);
    }
  };
  load_band(cur_key);

  // likelihood + gradient scale of NS steps (independent chains, interleaved by the compiler)
  auto steps = [&](const uint32_t (&w)[4], auto ns_tag, auto switch_tag) {
    constexpr int NS = decltype(ns_tag)::value;
    constexpr bool SWITCH = decltype(switch_tag)::value;
    float g[NS], sv[NS][RP];
    uint32_t srow[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      if (SWITCH) {
        const uint32_t key = w[j] & LW_BAND_MASK;
        if (key != cur_key) {
          flush_band(cur_key);
          cur_key = key;
          load_band(key);
        }
      }
      srow[j] = S_a + (w[j] & LW_PIX_MASK) * ROWB;
      if (RP % 4 == 0) {
#pragma unroll
        for (int r = 0; r < RP; r += 4) {
          const float4 s4 = lds128_ro(srow[j] + r * 4);
          sv[j][r] = s4.x; sv[j][r + 1] = s4.y; sv[j][r + 2] = s4.z; sv[j][r + 3] = s4.w;
        }
      } else {
#pragma unroll
        for (int r = 0; r < RP; ++r) sv[j][r] = lds32_ro(srow[j] + r * 4);
      }
      float t = 0.0f;
#pragma unroll
      for (int r = 0; r < RP; ++r) t = fmaf(sv[j][r], c[r], t);
      const int lv = (int)(w[j] >> 24);
      const bool ok = lv != 0xFF;
      float dxdt;
      const BinEval ev = eval_entry<EPI, LOGD>(prm, t, (EPI == EPI_ONEBIT) ? (lv & 1) : lv, dxdt);
      nll_part -= ok ? ev.logp : 0.0f;
      g[j] = ok ? ev.gx * dxdt : 0.0f;
      if (GRAD && SWITCH) {
        // sequential form: finish this step's updates before a later step may change band
        const uint32_t rs = srow[j] + gS_delta;
        if (RP % 4 == 0) {
#pragma unroll
          for (int r = 0; r < RP; r += 4) {
            float4 v = lds128(rs + r * 4);
            v.x = fmaf(g[j], c[r], v.x); v.y = fmaf(g[j], c[r + 1], v.y);
            v.z = fmaf(g[j], c[r + 2], v.z); v.w = fmaf(g[j], c[r + 3], v.w);
            sts128(rs + r * 4, v);
          }
        } else {
#pragma unroll
          for (int r = 0; r < RP; ++r) sts32(rs + r * 4, fmaf(g[j], c[r], lds32(rs + r * 4)));
        }
#pragma unroll
        for (int r = 0; r < RP; ++r) acc[r] = fmaf(g[j], sv[j][r], acc[r]);
        __syncwarp();
      }
    }
    if (GRAD && !SWITCH) {
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const uint32_t rs = srow[j] + gS_delta;
        if (RP % 4 == 0) {
#pragma unroll
          for (int r = 0; r < RP; r += 4) {
            float4 v = lds128(rs + r * 4);
            v.x = fmaf(g[j], c[r], v.x); v.y = fmaf(g[j], c[r + 1], v.y);
            v.z = fmaf(g[j], c[r + 2], v.z); v.w = fmaf(g[j], c[r + 3], v.w);
            sts128(rs + r * 4, v);
          }
        } else {
#pragma unroll
          for (int r = 0; r < RP; ++r) sts32(rs + r * 4, fmaf(g[j], c[r], lds32(rs + r * 4)));
        }
#pragma unroll
        for (int r = 0; r < RP; ++r) acc[r] = fmaf(g[j], sv[j][r], acc[r]);
        __syncwarp();  // the next step may touch the same pixel from another lane
      }
    }
  };

  for (int grp = 0; grp < ngroups; ++grp) {
    const uint4 wv = wa;
    wa = wb;
    wb = (grp + 2 < ngroups) ? ldg_u4(gp + (size_t)(grp + 2) * 32) : padg;
    const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
    // bands are contiguous runs in a lane's stream: the last word tells whether the group changes band
    const bool sw = ((wv.w ^ cur_key) & LW_BAND_MASK) != 0;
    if (__any_sync(0xffffffffu, sw)) {
      steps(w, std::integral_constant<int, 4>{}, std::true_type{});
    } else {
      steps(w, std::integral_constant<int, 4>{}, std::false_type{});
    }
  }
  flush_band(cur_key);

  double wsumv = warp_sum((double)nll_part);
  if (lane == 0) wsum[warp] = wsumv;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < W; ++i) tot += wsum[i];
    if (prm.tiles_per_map == 1) prm.nll[b] = tot;
    else atomicAdd(prm.nll + b, tot);
  }
  if (!GRAD) return;
  float* gSb = prm.gS + b * prm.sB;
  if (bulk) {
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) bulk_s2g(gSb + (int64_t)p0 * RP, gSsm, (uint32_t)np * RP * sizeof(float));
  } else {
#pragma unroll
    for (int r = 0; r < RP; ++r)
      if (r < prm.R)
        for (int pl = threadIdx.x; pl < np; pl += nthr) gSb[r * prm.sR + (int64_t)(p0 + pl) * prm.sP] = gSsm[pl * RP + r];
  }
  float* gCb = prm.gC + (int64_t)b * prm.R * K;
#pragma unroll
  for (int r = 0; r < RP; ++r) {
    if (r >= prm.R) break;
    for (int k = threadIdx.x; k < K; k += nthr) {
      float v = 0.0f;
      for (int w2 = 0; w2 < W; ++w2) v += gCw[((size_t)w2 * (K + 1) + k) * RP + r];
      if (prm.tiles_per_map == 1) gCb[r * K + k] = v;
      else atomicAdd(gCb + r * K + k, v);
    }
  }
  if (bulk && threadIdx.x == 0) bulk_wait_all();
}

template <int RP, int EPI, bool LOGD, bool GRAD>
static int launch_lanes_one(const GatherParams& prm, cudaStream_t st) {
  const size_t smem = lanes_smem_bytes(prm.K, RP, prm.sub_pixels, prm.tile_warps, GRAD);
  auto kern = gather_lanes_kernel<RP, EPI, LOGD, GRAD>;
  QMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ctas = (int64_t)prm.B * prm.tiles_per_map;
  QMC_REQUIRE(ctas <= 0x7fffffff, "too many CTAs (%lld)", (long long)ctas);
  kern<<<(unsigned)ctas, prm.tile_warps * 32, smem, st>>>(prm);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

template <int RP>
int launch_lanes_rp(const GatherParams& prm, int epi, bool logd, bool grad, cudaStream_t st) {
#define QMC_GO(E, L, G) return launch_lanes_one<RP, E, L, G>(prm, st)
  QMC_GATHER_SWITCH(QMC_GO);
#undef QMC_GO
  return QMC_OK;
}

}  // namespace qmc
