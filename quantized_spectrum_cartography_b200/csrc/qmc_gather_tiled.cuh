// Tiled observed-entry kernel: one CTA per (map, pixel tile), band rows per warp.
#pragma once
#include "qmc_gather_common.cuh"

namespace qmc {

// ------------------------------------------------------------------------------------------------
// tiled kernel
// ------------------------------------------------------------------------------------------------
// One CTA per (map, pixel tile), tile_warps warps; warp w owns pixel sub-tile (tile*W + w) exclusively
// and walks its entries -- one contiguous stream, rows (bands) in increasing order -- 32*UNR at a
// time.  Shared memory (floats):
//   Ssm[TP][RP] | Csm[K][RP] | gSsm[TP][RP] | gCw[Wc][K][RP] | scratch[W][32][RP] | offs[W][K+2] (int)
// TP = tile pixels.  gCw holds one private copy of gC per warp (Wc = W) when that fits, otherwise a
// single copy updated with shared-memory atomics (Wc = 1).
//
// Per 32-entry chunk: phase A (pure math, UNR chunks interleaved for ILP) computes x, log P and
// g = dNLL/dt for every entry; phase B applies the gradient updates band segment by band segment:
// inside one band the pixels of a sub-tile are distinct, so gS is a plain shared-memory
// read-modify-write; gC accumulates in registers and is reduced across the warp once per band.

template <int RP, int EPI, bool LOGD, bool GRAD, int UNR, bool PRIV>
__global__ void __launch_bounds__(256, 2) gather_tiled_kernel(const GatherParams prm) {
  extern __shared__ __align__(16) float smem[];
  const int W = prm.tile_warps, K = prm.K;
  const int TP = prm.sub_pixels * W;
  constexpr bool priv = PRIV;
  float* Ssm = smem;
  float* Csm = Ssm + (size_t)TP * RP;
  float* gSsm = Csm + (size_t)K * RP;
  float* gCw = gSsm + (GRAD ? (size_t)TP * RP : 0);
  float* scratch = gCw + (GRAD ? (size_t)(priv ? W : 1) * K * RP : 0);
  int* offs = reinterpret_cast<int*>(scratch + (GRAD ? (size_t)W * 32 * RP : 0));
  __shared__ uint64_t mbar;
  __shared__ double wsum[16];

  const int b = blockIdx.x / prm.tiles_per_map;
  const int tile = blockIdx.x - b * prm.tiles_per_map;
  const int p0 = tile * TP;
  const int np = min(TP, prm.IJ - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nthr = blockDim.x;

  const float* __restrict__ Sb = prm.S + b * prm.sB;
  const float* __restrict__ Cb = prm.C + (int64_t)b * prm.R * K;
  // pixel-major storage ([IJ][R], R == RP a multiple of 4): the tile is one contiguous, 16-byte
  // aligned run -> one TMA bulk copy in, one out
  const bool bulk = (prm.sR == 1 && prm.sP == RP && prm.R == RP && (RP % 4) == 0 &&
                     ((reinterpret_cast<uintptr_t>(Sb) | (GRAD ? reinterpret_cast<uintptr_t>(prm.gS + b * prm.sB) : 0)) & 15) == 0);

  // ---- stage the factor tiles ---------------------------------------------------------------
  if (bulk) {
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)np * RP * sizeof(float);
      mbar_expect_tx(&mbar, bytes);
      bulk_g2s(Ssm, Sb + (int64_t)p0 * RP, bytes, &mbar);
    }
  } else {
    // emitter-major storage (the reference's): coalesced row reads, transposed into [p][r]
#pragma unroll
    for (int r = 0; r < RP; ++r)
      for (int pl = threadIdx.x; pl < np; pl += nthr)
        Ssm[pl * RP + r] = (r < prm.R) ? __ldg(Sb + r * prm.sR + (int64_t)(p0 + pl) * prm.sP) : 0.0f;
  }
#pragma unroll
  for (int r = 0; r < RP; ++r)
    for (int k = threadIdx.x; k < K; k += nthr) Csm[k * RP + r] = (r < prm.R) ? __ldg(Cb + r * K + k) : 0.0f;
  if (GRAD) {
    for (int i = threadIdx.x; i < np * RP; i += nthr) gSsm[i] = 0.0f;
    for (int i = threadIdx.x; i < (priv ? W : 1) * K * RP; i += nthr) gCw[i] = 0.0f;
  }
  // this warp's entries: rows (b, tile*W + warp, 0..K-1), one contiguous stream
  const int64_t row0 = ((int64_t)b * prm.n_sub + (int64_t)tile * W + warp) * K;
  const int64_t beg = prm.row_off[row0];
  const int n = (int)(prm.row_off[row0 + K] - beg);
  int* offs_w = offs + warp * (K + 2);
  for (int i = lane; i <= K; i += 32) offs_w[i] = (int)(prm.row_off[row0 + i] - beg);
  if (lane == 0) offs_w[K + 1] = 0x7fffffff;
  __syncthreads();
  if (bulk) mbar_wait(&mbar, 0);

  const int32_t* __restrict__ idxw = prm.idx + beg;
  const uint8_t* __restrict__ lvlw = prm.lvl + beg;
  float* gCmine = gCw + (PRIV ? (size_t)warp * K * RP : 0);
  const int IJ = prm.IJ, dshift = prm.div_shift;
  const uint32_t dmagic = prm.div_magic;

  float nll_part = 0.0f;
  float acc[RP];
#pragma unroll
  for (int r = 0; r < RP; ++r) acc[r] = 0.0f;
  int cur_off = 0, next_off = offs_w[1];  // rows [cur_off, next_off) = current band

  // shared addresses used in the loop
  constexpr uint32_t ROWB = RP * sizeof(float);  // bytes per pixel row / band row
  const uint32_t S_a = smem_u32(Ssm), C_a = smem_u32(Csm);
  const uint32_t gS_delta = smem_u32(gSsm) - S_a;  // gS row address = S row address + delta
  const uint32_t scr_a = smem_u32(scratch) + (uint32_t)(warp * 32 + lane) * ROWB;

  // gS read-modify-write + gC register accumulation for the lanes selected by `on`.  Straight-line:
  // masked-off lanes update a private scratch row with g = 0 instead of branching around the code.
  auto update = [&](bool on, uint32_t s_row, float g, const float (&sv)[RP], const float (&cv)[RP]) {
    const uint32_t row = on ? s_row + gS_delta : scr_a;
    const float ge = on ? g : 0.0f;
    if (RP % 4 == 0) {
#pragma unroll
      for (int r = 0; r < RP; r += 4) {
        float4 v = lds128(row + r * 4);
        v.x = fmaf(ge, cv[r], v.x); v.y = fmaf(ge, cv[r + 1], v.y);
        v.z = fmaf(ge, cv[r + 2], v.z); v.w = fmaf(ge, cv[r + 3], v.w);
        sts128(row + r * 4, v);
      }
    } else {
#pragma unroll
      for (int r = 0; r < RP; ++r) sts32(row + r * 4, fmaf(ge, cv[r], lds32(row + r * 4)));
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) acc[r] = fmaf(ge, sv[r], acc[r]);
  };
  // band kcur is complete: reduce its gC contribution across the warp, move to the next band
  const bool writer = (lane & (32 / RP - 1)) == 0;
  uint32_t gc_a = smem_u32(gCmine + warp_transpose_owner<RP>(lane));  // advances by one band row per band
  uint32_t off_a = smem_u32(offs_w + 2);                              // &offs_w[kcur + 2]
  auto end_band = [&]() {
    if (next_off > cur_off) {  // the band had entries in this sub-tile
      const float tot = warp_transpose_sum<RP>(acc, lane);
      if (writer) {
        if (PRIV) sts32(gc_a, tot);
        else atomicAdd(gCmine + (gc_a - smem_u32(gCmine)) / 4, tot);
      }
#pragma unroll
      for (int r = 0; r < RP; ++r) acc[r] = 0.0f;
    }
    cur_off = next_off;
    gc_a += ROWB;
    next_off = lds32i(off_a);
    off_a += 4;
  };

  // two super-chunks of look-ahead: the loads issued in iteration i are consumed in iteration i+2
  constexpr int SUPER = 32 * UNR;
  int id_a[UNR], lv_a[UNR], id_b[UNR], lv_b[UNR];
  const int32_t* ip = idxw + lane;   // running per-lane pointers: loads use immediate offsets
  const uint8_t* lp = lvlw + lane;
  int rem = n - lane;                 // entries left from this lane's position
#pragma unroll
  for (int j = 0; j < UNR; ++j) {
    id_a[j] = 32 * j < rem ? __ldg(ip + 32 * j) : -1;
    lv_a[j] = 32 * j < rem ? (int)__ldg(lp + 32 * j) : 0;
    id_b[j] = SUPER + 32 * j < rem ? __ldg(ip + SUPER + 32 * j) : -1;
    lv_b[j] = SUPER + 32 * j < rem ? (int)__ldg(lp + SUPER + 32 * j) : 0;
  }

  for (int pos0 = 0; pos0 < n; pos0 += SUPER, ip += SUPER, lp += SUPER, rem -= SUPER) {
    int id_c[UNR], lv_c[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      id_c[j] = id_a[j];
      lv_c[j] = lv_a[j];
      id_a[j] = id_b[j];
      lv_a[j] = lv_b[j];
      const bool more = 2 * SUPER + 32 * j < rem;
      id_b[j] = more ? __ldg(ip + 2 * SUPER + 32 * j) : -1;
      lv_b[j] = more ? (int)__ldg(lp + 2 * SUPER + 32 * j) : 0;
    }
    // ---- phase A: likelihood of UNR independent chunks ------------------------------------------
    float g[UNR], sv[UNR][RP], cv[UNR][RP];
    uint32_t srow[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      const bool valid = id_c[j] >= 0;
      const int id = valid ? id_c[j] : p0;  // (band 0, local pixel 0): harmless stand-in
      const int k = fast_div((uint32_t)id, dmagic, dshift);
      srow[j] = S_a + (uint32_t)(id - k * IJ - p0) * ROWB;
      const uint32_t crow = C_a + (uint32_t)k * ROWB;
      if (RP % 4 == 0) {
#pragma unroll
        for (int r = 0; r < RP; r += 4) {
          const float4 s4 = lds128_ro(srow[j] + r * 4);
          const float4 c4 = lds128_ro(crow + r * 4);
          sv[j][r] = s4.x; sv[j][r + 1] = s4.y; sv[j][r + 2] = s4.z; sv[j][r + 3] = s4.w;
          cv[j][r] = c4.x; cv[j][r + 1] = c4.y; cv[j][r + 2] = c4.z; cv[j][r + 3] = c4.w;
        }
      } else {
#pragma unroll
        for (int r = 0; r < RP; ++r) {
          sv[j][r] = lds32_ro(srow[j] + r * 4);
          cv[j][r] = lds32_ro(crow + r * 4);
        }
      }
      float t = 0.0f;
#pragma unroll
      for (int r = 0; r < RP; ++r) t = fmaf(sv[j][r], cv[j][r], t);
      float dxdt;
      const BinEval ev = eval_entry<EPI, LOGD>(prm, t, lv_c[j], dxdt);
      nll_part -= valid ? ev.logp : 0.0f;
      g[j] = valid ? ev.gx * dxdt : 0.0f;
    }
    if (!GRAD) continue;
    // ---- phase B: gradient updates, band segment by band segment ---------------------------------
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      const int cstart = pos0 + 32 * j;
      if (cstart >= n) break;
      const int cend = min(cstart + 32, n);
      const int pos = cstart + lane;
      const bool valid = pos < cend;
      if (next_off >= cend) {
        // the whole chunk lies in the current band (distinct pixels, exclusive to this warp)
        update(valid, srow[j], g[j], sv[j], cv[j]);
        if (next_off == cend) end_band();
      } else {
        // a band ends inside the chunk: its lanes first, then the rest
        int bnd = next_off;
        update(pos < bnd, srow[j], g[j], sv[j], cv[j]);
        __syncwarp();
        end_band();
        while (next_off < cend) {  // (rare) further whole bands inside this chunk
          update(pos >= bnd && pos < next_off, srow[j], g[j], sv[j], cv[j]);
          __syncwarp();
          bnd = next_off;
          end_band();
        }
        update(pos >= bnd && valid, srow[j], g[j], sv[j], cv[j]);
        if (next_off == cend) end_band();
      }
      __syncwarp();
    }
  }

  // ---- NLL ---------------------------------------------------------------------------------------
  double w = warp_sum((double)nll_part);
  if (lane == 0) wsum[warp] = w;
  __syncthreads();  // also orders all gS/gC shared-memory updates before the write-back
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < W; ++i) tot += wsum[i];
    if (prm.tiles_per_map == 1) prm.nll[b] = tot;
    else atomicAdd(prm.nll + b, tot);
  }
  if (!GRAD) return;

  // ---- write the gradient tiles back -------------------------------------------------------------
  float* gSb = prm.gS + b * prm.sB;
  if (bulk) {
    fence_async_smem();  // generic-proxy writes to gSsm -> visible to the bulk-copy engine
    __syncthreads();
    if (threadIdx.x == 0) bulk_s2g(gSb + (int64_t)p0 * RP, gSsm, (uint32_t)np * RP * sizeof(float));
  } else {
#pragma unroll
    for (int r = 0; r < RP; ++r)
      if (r < prm.R)
        for (int pl = threadIdx.x; pl < np; pl += nthr) gSb[r * prm.sR + (int64_t)(p0 + pl) * prm.sP] = gSsm[pl * RP + r];
  }
  float* gCb = prm.gC + (int64_t)b * prm.R * K;
  const int wc = priv ? W : 1;
#pragma unroll
  for (int r = 0; r < RP; ++r) {
    if (r >= prm.R) break;
    for (int k = threadIdx.x; k < K; k += nthr) {
      float v = 0.0f;
      for (int w2 = 0; w2 < wc; ++w2) v += gCw[((size_t)w2 * K + k) * RP + r];
      if (prm.tiles_per_map == 1) gCb[r * K + k] = v;
      else atomicAdd(gCb + r * K + k, v);
    }
  }
  if (bulk && threadIdx.x == 0) bulk_wait_all();  // smem must stay alive until the engine has read it
}

template <int RP, int EPI, bool LOGD, bool GRAD>
static int launch_tiled_one(const GatherParams& prm, cudaStream_t st) {
  const size_t smem = tiled_smem_bytes(prm.K, RP, prm.sub_pixels, prm.tile_warps, GRAD);
  constexpr int UNR = RP <= 4 ? 2 : 1;
  auto kern = gc_private(prm.K, RP, prm.tile_warps) ? gather_tiled_kernel<RP, EPI, LOGD, GRAD, UNR, true>
                                                     : gather_tiled_kernel<RP, EPI, LOGD, GRAD, UNR, false>;
  if (RP == 4 && EPI == EPI_ONEBIT && GRAD && !LOGD && gc_private(prm.K, RP, prm.tile_warps)) {
    // deeper interleave for few-warp tiles (long rows, low occupancy): 4 chunks in flight per warp
    const char* e = getenv("QMC_TILED_UNR");
    const int want = e ? atoi(e) : (prm.tile_warps <= 4 ? 4 : 2);
    if (want == 4) kern = gather_tiled_kernel<RP, EPI, LOGD, GRAD, (RP == 4 ? 4 : UNR), true>;
    if (want == 1) kern = gather_tiled_kernel<RP, EPI, LOGD, GRAD, (RP == 4 ? 1 : UNR), true>;
  }
  QMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ctas = (int64_t)prm.B * prm.tiles_per_map;
  QMC_REQUIRE(ctas <= 0x7fffffff, "too many CTAs (%lld)", (long long)ctas);
  kern<<<(unsigned)ctas, prm.tile_warps * 32, smem, st>>>(prm);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

template <int RP>
int launch_tiled_rp(const GatherParams& prm, int epi, bool logd, bool grad, cudaStream_t st) {
#define QMC_GO(E, L, G) return launch_tiled_one<RP, E, L, G>(prm, st)
  QMC_GATHER_SWITCH(QMC_GO);
#undef QMC_GO
  return QMC_OK;
}

}  // namespace qmc
