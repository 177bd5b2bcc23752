// Observed-entry ("gather") kernels: fused low-rank reconstruction x = <S[:,p], C[:,k]>, optional
// log link, quantized probit NLL and the gradients w.r.t. both factors, visiting only observed
// entries.  Replaces get_tensor -> prob_probit -> -sum(Wx*log P) -> backward of the reference
// (qmc/quantization_model.py:22-39,57-61,70-86; caller qmc/qmc.ipynb c1:145-153).
//
//   flat  : one thread per observed entry, factors through L2, warp-aggregated global atomics.
//   tiled : one CTA per (map, pixel tile); S tile, C, and both gradient tiles live in shared
//           memory; every warp owns a pixel sub-tile exclusively, so gS updates are plain
//           shared-memory read-modify-writes (no atomics); entries arrive sorted by band inside a
//           sub-tile, so gC accumulates in registers and is reduced across the warp once per band.
#include <cstdlib>

#include "qmc_common.cuh"

namespace qmc {

struct GatherParams {
  const float* S;
  int64_t sB, sR, sP;
  const float* C;
  const int32_t* idx;
  const uint8_t* lvl;
  const int64_t* row_off;
  double* nll;
  float* gS;
  float* gC;
  int n_sub, sub_pixels;
  int B, IJ, K, R;
  uint32_t div_magic;  // k = umulhi(idx, div_magic) >> div_shift  (idx < 2^31)
  int div_shift;
  int tiles_per_map, tile_warps;
  float inv_a, offset;
  float thr;  // one-bit fast path threshold
  float bounds[QMC_MAX_BOUNDS];
};

enum : int { EPI_STABLE = 0, EPI_REFERENCE = 1, EPI_ONEBIT = 2 };

__device__ __forceinline__ int fast_div(uint32_t n, uint32_t magic, int shift) {
  return (int)(__umulhi(n, magic) >> shift);
}

template <int EPI, bool LOGD>
__device__ __forceinline__ BinEval eval_entry(const GatherParams& prm, float t, int lvl, float& dxdt) {
  float x = t;
  dxdt = 1.0f;
  if (LOGD) {
    const float u = t + prm.offset;
    x = logf(u);
    dxdt = 1.0f / u;
  }
  if (EPI == EPI_ONEBIT) {
    return probit_one_sided_fast(prm.thr, lvl ? -prm.inv_a : prm.inv_a, x);
  } else if (EPI == EPI_REFERENCE) {
    return probit_bin_reference(prm.bounds[lvl], prm.bounds[lvl + 1], x, prm.inv_a);
  } else {
    return probit_bin_stable<true>(prm.bounds[lvl], prm.bounds[lvl + 1], x, prm.inv_a);
  }
}

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

constexpr size_t kPrivateGcBytes = 32 * 1024;

__host__ __device__ inline bool gc_private(int K, int RP, int W) {
  return (size_t)W * K * RP * sizeof(float) <= kPrivateGcBytes;
}

static size_t tiled_smem_bytes(int K, int RP, int sub_pixels, int W, bool grad) {
  const size_t TP = (size_t)sub_pixels * W;
  const size_t wc = gc_private(K, RP, W) ? W : 1;
  size_t fl = TP * RP + (size_t)K * RP;
  if (grad) fl += TP * RP + wc * K * RP;
  if (grad) fl += (size_t)W * 32 * RP;  // per-lane scratch rows for masked-off updates
  return fl * sizeof(float) + (size_t)W * (K + 2) * sizeof(int) + 16;
}

// ---- bulk (TMA) copies of a contiguous tile ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- shared-memory accesses by 32-bit shared address (no generic-pointer arithmetic in the loop) --
// read-only data of the main loop (S and C tiles): plain asm, free to be scheduled
__device__ __forceinline__ float4 lds128_ro(uint32_t a) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32_ro(uint32_t a) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
// read-modify-write data (gradient tiles): ordered
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ int lds32i(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

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

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
template <int RP, int EPI, bool LOGD, bool GRAD>
static int launch_one(const GatherParams& prm, int algo, cudaStream_t st) {
  if (algo == QMC_ALGO_FLAT) {
    // size the grid from the average entries per map; the kernel is grid-stride
    const int threads = 256;
    int64_t per_map_guess = (int64_t)prm.K * prm.IJ;  // upper bound; the loop exits early
    int64_t want = (per_map_guess + threads - 1) / threads;
    int bx = (int)(want < 148 * 8 ? want : 148 * 8);
    if (bx < 1) bx = 1;
    dim3 grid(bx, prm.B);
    gather_flat_kernel<RP, EPI, LOGD, GRAD><<<grid, threads, 0, st>>>(prm);
  } else {
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
  }
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

template <int RP, int EPI>
static int launch_rp_epi(const GatherParams& prm, int algo, bool logd, bool grad, cudaStream_t st) {
  if (logd) return grad ? launch_one<RP, EPI, true, true>(prm, algo, st) : launch_one<RP, EPI, true, false>(prm, algo, st);
  return grad ? launch_one<RP, EPI, false, true>(prm, algo, st) : launch_one<RP, EPI, false, false>(prm, algo, st);
}

template <int RP>
static int launch_rp(const GatherParams& prm, int algo, int epi, bool logd, bool grad, cudaStream_t st) {
  switch (epi) {
    case EPI_ONEBIT: return launch_rp_epi<RP, EPI_ONEBIT>(prm, algo, logd, grad, st);
    case EPI_REFERENCE: return launch_rp_epi<RP, EPI_REFERENCE>(prm, algo, logd, grad, st);
    default: return launch_rp_epi<RP, EPI_STABLE>(prm, algo, logd, grad, st);
  }
}

}  // namespace qmc

using namespace qmc;

extern "C" int64_t qmc_tiled_smem_bytes(int K, int R, int sub_pixels, int tile_warps) {
  if (K <= 0 || R <= 0 || R > QMC_MAX_RANK || sub_pixels <= 0 || tile_warps <= 0 || tile_warps > 8) return 0;
  int RP = 1;
  while (RP < R) RP <<= 1;
  const size_t b = tiled_smem_bytes(K, RP, sub_pixels, tile_warps, true);
  return b <= 227 * 1024 ? (int64_t)b : 0;
}

extern "C" int qmc_nll_fwd_bwd_gather(const float* S_dev, int64_t s_stride_b, int64_t s_stride_r,
                                      int64_t s_stride_p, const float* C_dev, const qmc_obs_view_t* obs,
                                      const qmc_likelihood_t* lik, int B, int IJ, int K, int R, int algo,
                                      int tile_warps, double* nll_out_dev, float* gS_out_dev,
                                      float* gC_out_dev, void* stream) {
  QMC_REQUIRE(S_dev && C_dev && obs && lik && nll_out_dev, "null argument");
  QMC_REQUIRE(obs->idx_dev && obs->lvl_dev && obs->row_off_dev, "null observation arrays");
  QMC_REQUIRE(B > 0 && IJ > 0 && K > 0 && R > 0, "bad sizes B=%d IJ=%d K=%d R=%d", B, IJ, K, R);
  QMC_REQUIRE(R <= QMC_MAX_RANK, "rank %d > %d", R, QMC_MAX_RANK);
  QMC_REQUIRE((int64_t)K * IJ < (1LL << 31), "K*IJ = %lld does not fit the int32 linear index", (long long)K * IJ);
  QMC_REQUIRE(lik->n_bounds >= 2 && lik->n_bounds <= QMC_MAX_BOUNDS, "n_bounds %d out of range", lik->n_bounds);
  QMC_REQUIRE(lik->noise_std > 0.0f, "noise_std must be positive");
  QMC_REQUIRE(obs->n_sub > 0 && obs->sub_pixels > 0 && (int64_t)obs->n_sub * obs->sub_pixels >= IJ,
              "sub-tiles (%d x %d) do not cover IJ=%d", obs->n_sub, obs->sub_pixels, IJ);
  const bool grad = !(lik->flags & QMC_FORWARD_ONLY);
  QMC_REQUIRE(!grad || (gS_out_dev && gC_out_dev), "gradient outputs are NULL without QMC_FORWARD_ONLY");
  cudaStream_t st = (cudaStream_t)stream;

  GatherParams prm;
  prm.S = S_dev; prm.sB = s_stride_b; prm.sR = s_stride_r; prm.sP = s_stride_p;
  prm.C = C_dev; prm.idx = obs->idx_dev; prm.lvl = obs->lvl_dev; prm.row_off = obs->row_off_dev;
  prm.nll = nll_out_dev; prm.gS = gS_out_dev; prm.gC = gC_out_dev;
  prm.n_sub = obs->n_sub; prm.sub_pixels = obs->sub_pixels;
  prm.B = B; prm.IJ = IJ; prm.K = K; prm.R = R;
  // exact division of any idx < 2^31 by IJ (Granlund-Montgomery with N = 31: l = ceil(log2 IJ),
  // m = floor(2^(31+l)/IJ) + 1 fits 32 bits, q = (m*n) >> (31+l) = umulhi(n, m) >> (l-1))
  QMC_REQUIRE(IJ > 1, "IJ must be > 1");
  int l = 0;
  while ((1LL << l) < IJ) ++l;
  prm.div_magic = (uint32_t)(((1ULL << (31 + l)) / (uint64_t)IJ) + 1ULL);
  prm.div_shift = l - 1;
  const float a = probit_scale(lik->noise_std);
  prm.inv_a = 1.0f / a;
  prm.offset = lik->offset;
  for (int i = 0; i < lik->n_bounds; ++i) prm.bounds[i] = lik->bounds[i];
  prm.thr = lik->n_bounds >= 3 ? lik->bounds[1] : 0.0f;

  // epilogue: reference-literal, one-bit fast path (both outer bounds numerically infinite), or general
  int epi = EPI_STABLE;
  if (lik->flags & QMC_EPI_REFERENCE) epi = EPI_REFERENCE;
  else if (lik->n_bounds == 3) {
    // erfc(z) == 0 exactly (even in double) for z > 27; require the sentinel to sit that far out for
    // any |x| < half its magnitude
    const float lo = lik->bounds[0], hi = lik->bounds[2];
    const bool inf_lo = lo < 0 && (-lo * 0.5f) * prm.inv_a > 30.0f;
    const bool inf_hi = hi > 0 && (hi * 0.5f) * prm.inv_a > 30.0f;
    if (inf_lo && inf_hi && fabsf(lo) >= 1e4f && fabsf(hi) >= 1e4f) epi = EPI_ONEBIT;
  }
  const bool logd = (lik->flags & QMC_LOG_DOMAIN) != 0;

  int RP = 1;
  while (RP < R) RP <<= 1;

  if (algo == QMC_ALGO_AUTO) {
    const int64_t smem = qmc_tiled_smem_bytes(K, R, obs->sub_pixels, tile_warps > 0 ? tile_warps : 1);
    algo = (tile_warps > 0 && smem > 0 && obs->n_sub % tile_warps == 0 && (int64_t)B * (obs->n_sub / tile_warps) >= 64)
               ? QMC_ALGO_TILED : QMC_ALGO_FLAT;
  }
  if (algo == QMC_ALGO_TILED) {
    QMC_REQUIRE(tile_warps > 0 && tile_warps <= 8, "tile_warps %d out of range [1, 8]", tile_warps);
    QMC_REQUIRE(obs->n_sub % tile_warps == 0, "n_sub %d is not a multiple of tile_warps %d", obs->n_sub, tile_warps);
    QMC_REQUIRE(tiled_smem_bytes(K, RP, obs->sub_pixels, tile_warps, grad) <= 227 * 1024,
                "tile of %d pixels x rank %d does not fit shared memory", obs->sub_pixels * tile_warps, RP);
    prm.tile_warps = tile_warps;
    prm.tiles_per_map = obs->n_sub / tile_warps;
  } else if (algo == QMC_ALGO_FLAT) {
    prm.tile_warps = 0; prm.tiles_per_map = 0;
  } else {
    return set_error(QMC_ERR_INVALID, "unknown algo %d", algo);
  }

  // zero what the kernels accumulate into
  const bool atomics_on_out = (algo == QMC_ALGO_FLAT) || prm.tiles_per_map > 1;
  if (atomics_on_out) {
    QMC_CUDA_CHECK(cudaMemsetAsync(nll_out_dev, 0, sizeof(double) * B, st));
    if (grad) QMC_CUDA_CHECK(cudaMemsetAsync(gC_out_dev, 0, sizeof(float) * (size_t)B * R * K, st));
  }
  if (grad && algo == QMC_ALGO_FLAT) {
    // gS may be strided: zero the dense extent it spans only when it is compact
    const bool dense = (s_stride_p == 1 && s_stride_r == IJ) || (s_stride_r == 1 && s_stride_p == R);
    QMC_REQUIRE(dense && s_stride_b == (int64_t)R * IJ, "flat kernel needs a compact gS layout");
    QMC_CUDA_CHECK(cudaMemsetAsync(gS_out_dev, 0, sizeof(float) * (size_t)B * R * IJ, st));
  }

  switch (RP) {
    case 1: return launch_rp<1>(prm, algo, epi, logd, grad, st);
    case 2: return launch_rp<2>(prm, algo, epi, logd, grad, st);
    case 4: return launch_rp<4>(prm, algo, epi, logd, grad, st);
    case 8: return launch_rp<8>(prm, algo, epi, logd, grad, st);
    case 16: return launch_rp<16>(prm, algo, epi, logd, grad, st);
    case 32: return launch_rp<32>(prm, algo, epi, logd, grad, st);
  }
  return set_error(QMC_ERR_UNSUPPORTED, "rank %d", R);
}
