// Shared pieces of the observed-entry ("gather") kernels: launch parameters, the per-entry likelihood
// dispatch, TMA bulk copies and shared-memory access helpers.  See qmc_gather.cu for the overview.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "qmc_common.cuh"

namespace qmc {

struct GatherParams {
  const float* S;
  int64_t sB, sR, sP;
  const float* C;
  const int32_t* idx;
  const uint8_t* lvl;
  const int64_t* row_off;
  const uint32_t* words;      // lane-stream layout (qmc_obs_build_lanes)
  const int64_t* stream_off;  // lane-stream layout: first word of every (map, sub-tile) stream
  const int32_t* nrows;       // lane-stream layout: steps per stream (multiple of 4)
  int64_t stream_stride;      // lane-stream layout: > 0 = uniform stream capacity in words (no table look-up)
  int n_runs;                 // lane-stream layout: run-table entries per lane (the stream starts with n_runs*32 words)
  int word16;                 // lane-stream layout: 16-bit words (two groups per 16 bytes) instead of 32-bit ones
  int lvl_bits;               // 16-bit words: bits of the level field
  int has_cont;               // lane-stream layout: some band is split over several lanes (continuation rows in use)
  int map_mod;                // lanes kernel, forward only: map b uses the observations and C of map b % map_mod (0 = off)
  int lookahead;              // lanes kernel: CTAs resident on the device (prefetch distance in CTAs)
  int want_gs, want_gc;       // lanes kernel: which gradients the caller needs (QMC_SKIP_GS / QMC_SKIP_GC)
  // fused S-step (qmc_solver_s_step_fused): S is updated in place from the gS tile in shared memory
  int fuse_update;
  float* S_rw;                // = S
  float* adam_m;
  float* adam_v;
  const double* ss_in;
  double* ss_out;
  float lr, beta1, beta2, eps, lam;
  int project, step;
  const int32_t* step_dev;
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
  int one_sided;  // logistic model: three boundaries whose outer two are numerically infinite
  int n_bounds;   // entries of bounds[] in use
  float bounds[QMC_MAX_BOUNDS];
};

enum : int { EPI_STABLE = 0, EPI_REFERENCE = 1, EPI_ONEBIT = 2, EPI_LSQ = 3, EPI_LOGISTIC = 4 };

__device__ __forceinline__ int fast_div(uint32_t n, uint32_t magic, int shift) {
  return (int)(__umulhi(n, magic) >> shift);
}

// bnd: the boundary table -- prm.bounds (kernel-parameter bank: a per-lane level makes it a divergent constant
// access, replayed once per distinct level of the warp) or a copy in shared memory (lanes kernel: a plain gather).
// FASTLOG: SFU-grade log link (lg2.approx is good to 2^-22 absolute in log2 units, far below what the bin widths can
// see; a non-positive argument gives NaN / -inf exactly as logf does); never for the reference's literal epilogue.
template <int EPI, bool LOGD, bool FASTLOG = false>
__device__ __forceinline__ BinEval eval_entry(const GatherParams& prm, const float* __restrict__ bnd, float t, int lvl, float& dxdt) {
  float x = t;
  dxdt = 1.0f;
  if (LOGD) {
    const float u = t + prm.offset;
    if (FASTLOG && EPI != EPI_REFERENCE) {
      x = kLn2 * lg2_approx(u);
      dxdt = rcp_approx(u);
    } else {
      x = logf(u);
      dxdt = 1.0f / u;
    }
  }
  if (EPI == EPI_LSQ) {
    // masked least squares on the bin mid-point (quantization_model_log.py:43-51, qmc_dowjons.ipynb c1:112):
    // "logp" = -(x - mid)^2 so that the callers' nll -= logp accumulates the squared residual
    const float d = x - 0.5f * (bnd[lvl] + bnd[lvl + 1]);
    BinEval o;
    o.logp = -d * d;
    o.gx = 2.0f * d;
    return o;
  } else if (EPI == EPI_LOGISTIC) {
    if (prm.one_sided) return logistic_one_sided_fast(prm.thr, lvl ? prm.inv_a : -prm.inv_a, x);  // uniform branch
    return logistic_bin(bnd[lvl], bnd[lvl + 1], x, prm.inv_a);
  } else if (EPI == EPI_ONEBIT) {
    return probit_one_sided_fast(prm.thr, lvl ? -prm.inv_a : prm.inv_a, x);
  } else if (EPI == EPI_REFERENCE) {
    return probit_bin_reference(bnd[lvl], bnd[lvl + 1], x, prm.inv_a);
  } else {
    return probit_bin_stable<true>(bnd[lvl], bnd[lvl + 1], x, prm.inv_a);
  }
}
template <int EPI, bool LOGD>
__device__ __forceinline__ BinEval eval_entry(const GatherParams& prm, float t, int lvl, float& dxdt) {
  return eval_entry<EPI, LOGD, false>(prm, prm.bounds, t, lvl, dxdt);
}

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

// ---- lane-stream layout (qmc_obs_build_lanes, include/qmc_b200.h) ---------------------------------
// 32-bit words: bit 31 = level & 1, bits 24..30 = level >> 1 (0xFF in bits 24..31 = padding), bits 0..14 = pixel.
// 16-bit words: the top lvl_bits bits = level, the next bit = padding flag, the rest = pixel.
// Run-table entry: bits 0..8 = gC row, bits 9..17 = band, bits 18..31 = groups.
constexpr uint32_t LW_PIX_MASK = 0x7FFFu;
constexpr uint32_t LW_RUN_ROW_MASK = 0x1FFu;
constexpr int LW_RUN_BAND_SHIFT = 9;
constexpr int LW_RUN_LEN_SHIFT = 18;
constexpr int LANES_CONT_ROWS = 32;  // gC rows K+1 .. K+32 of a warp: continuation pieces, one per lane

constexpr int LANES_LOOKAHEAD = 4;   // groups of stream look-ahead per warp
__host__ __device__ inline int lanes_ring_slots(bool w16) { return w16 ? LANES_LOOKAHEAD / 2 : LANES_LOOKAHEAD; }

// Shared memory of the lanes kernel (offsets in floats, every section 16-byte aligned):
//   Ssm[TP][RP] | gSsm[TP][RP] | Csm[K+1][RP] | gCw[W][K+1+32][RP] | hdr[W][n_runs][32] | ring[W][slots][32][4]
// Ssm comes first so that the address of an S row is the row offset plus a compile-time constant.
struct LanesLayout {
  uint32_t gS, C, gC, hdr, ring, total;
};
__host__ __device__ inline LanesLayout lanes_layout(int K, int RP, int TP, int W, bool grad, int n_runs, bool w16) {
  LanesLayout L;
  const uint32_t tile = ((uint32_t)TP * RP + 3u) & ~3u;
  L.gS = tile;
  L.C = L.gS + (grad ? tile : 0u);
  L.gC = L.C + (((uint32_t)(K + 1) * RP + 3u) & ~3u);
  L.hdr = L.gC + (grad ? (((uint32_t)W * (K + 1 + LANES_CONT_ROWS) * RP + 3u) & ~3u) : 0u);
  L.ring = L.hdr + (uint32_t)W * n_runs * 32u;
  L.total = L.ring + (uint32_t)W * lanes_ring_slots(w16) * 128u;
  return L;
}
static size_t lanes_smem_bytes(int K, int RP, int sub_pixels, int W, bool grad, int n_runs, bool w16) {
  return (size_t)lanes_layout(K, RP, sub_pixels * W, W, grad, n_runs, w16).total * sizeof(float) + 16;
}
// per-family launchers, one explicit instantiation per padded rank (qmc_gather_inst.cu)
template <int RP> int launch_flat_rp(const GatherParams& prm, int epi, bool logd, bool grad, cudaStream_t st);
template <int RP> int launch_tiled_rp(const GatherParams& prm, int epi, bool logd, bool grad, cudaStream_t st);
template <int RP> int launch_lanes_rp(const GatherParams& prm, int epi, bool logd, bool grad, cudaStream_t st);

// expands to the (EPI, LOGD, GRAD) switch around GO(EPI, LOGD, GRAD)
#define QMC_GATHER_SWITCH(GO)                                                        \
  do {                                                                               \
    switch (epi) {                                                                   \
      case EPI_ONEBIT:                                                               \
        if (logd) { if (grad) GO(EPI_ONEBIT, true, true); else GO(EPI_ONEBIT, true, false); }          \
        else { if (grad) GO(EPI_ONEBIT, false, true); else GO(EPI_ONEBIT, false, false); }             \
        break;                                                                       \
      case EPI_REFERENCE:                                                            \
        if (logd) { if (grad) GO(EPI_REFERENCE, true, true); else GO(EPI_REFERENCE, true, false); }    \
        else { if (grad) GO(EPI_REFERENCE, false, true); else GO(EPI_REFERENCE, false, false); }       \
        break;                                                                       \
      case EPI_LOGISTIC:                                                             \
        if (logd) { if (grad) GO(EPI_LOGISTIC, true, true); else GO(EPI_LOGISTIC, true, false); }      \
        else { if (grad) GO(EPI_LOGISTIC, false, true); else GO(EPI_LOGISTIC, false, false); }         \
        break;                                                                       \
      case EPI_LSQ:                                                                  \
        if (logd) { if (grad) GO(EPI_LSQ, true, true); else GO(EPI_LSQ, true, false); }                \
        else { if (grad) GO(EPI_LSQ, false, true); else GO(EPI_LSQ, false, false); }                   \
        break;                                                                       \
      default:                                                                       \
        if (logd) { if (grad) GO(EPI_STABLE, true, true); else GO(EPI_STABLE, true, false); }          \
        else { if (grad) GO(EPI_STABLE, false, true); else GO(EPI_STABLE, false, false); }             \
    }                                                                                \
  } while (0)

}  // namespace qmc
