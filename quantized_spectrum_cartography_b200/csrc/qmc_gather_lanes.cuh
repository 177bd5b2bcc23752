// Lane-stream observed-entry kernel: per-lane band walks, C row and gC accumulator in registers.
#pragma once
#include "qmc_gather_common.cuh"

namespace qmc {

// ------------------------------------------------------------------------------------------------
// lanes kernel
// ------------------------------------------------------------------------------------------------
// Same tiling and ownership as the tiled kernel (one CTA per (map, pixel tile), each warp owns a pixel
// sub-tile exclusively), on the lane-stream layout of qmc_obs_build_lanes: every lane of a warp walks
// a list of runs -- (band, number of 4-step groups) pairs from the stream's run table -- so C[band] and
// the band's gC accumulator stay in registers (no shared-memory traffic, no warp reduction, one plain
// store per run); the builder guarantees that the 32 entries of a step hit 32 different pixels, so the
// gS update is a plain shared-memory read-modify-write, and that a lane changes band only at a group
// (4-step) boundary.  Per entry the shared-memory pipe sees one S row read and one gS row
// read-modify-write; a padding lane re-reads a row that a real lane of its quarter-warp reads anyway
// (a broadcast) and its update is predicated off.
//
// Stream words carry only (level, padding flag, tile-local pixel): 16 bits when that fits (W16: one
// 16-byte load per lane brings two groups), 32 bits otherwise.  The band comes from the run table.
//
// A band may be split over several lanes (runs cut at group boundaries so that all lanes get the same
// number of groups).  The piece that starts a band stores its gC partial to the band's row of the warp's
// private gC copy; a continuation piece is always the first run of its lane and stores to a row of its
// own (K+1+lane), merged into the band's row when the warp is done -- no atomics on the way.
//
// The warps of a CTA are autonomous: each stages its own S slice (its own TMA bulk copy and
// mbarrier), zeroes and later writes its own gS slice, and keeps a private gC copy; the only CTA-wide
// synchronisation is one early barrier after C is staged.  The last warp to finish folds the gC
// copies and the NLL partials and writes them out.
// Shared memory: see lanes_layout (qmc_gather_common.cuh); band row K of Csm / gCw is a dummy (zeros /
// absorbs the flush of a lane that owns nothing).
//
// Convergence: the gS read-modify-writes of a step may touch a pixel that another lane touched in an
// earlier step, so the warp must execute them in step order.  The only divergent code of the main loop is
// the run switch; it is followed by __syncwarp(), and everything after it up to the next switch is
// straight-line predicated code, which keeps the warp converged.

// ---- packed fp32x2 arithmetic (FFMA2/FMUL2/FADD2, sm_100) --------------------------------------
struct f2 {
  unsigned long long v;
};
__device__ __forceinline__ f2 mk2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void un2(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 bc2(float x) { return mk2(x, x); }

// ---- predicated shared-memory stores (padding lanes) ----------------------------------------------
__device__ __forceinline__ void sts128_if(uint32_t a, float4 v, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
               ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)p) : "memory");
}
__device__ __forceinline__ void sts32_if(uint32_t a, float v, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.shared.f32 [%0], %1;\n\t}" ::"r"(a), "f"(v), "r"((int)p) : "memory");
}

// 16-byte / 4-byte asynchronous global -> shared copies (LDGSTS) and their in-order group accounting
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// GMODE: 0 = NLL only, 1 = both gradients, 2 = gC only (QMC_SKIP_GS), 3 = gS only (QMC_SKIP_GC),
//        4 = gS only, consumed in place by the fused Adam update of S (qmc_solver_s_step_fused)
// 64 KB of zeros in global memory (L2-resident in practice): source of the bulk copy that clears a warp's gS
// slice, so the clearing costs neither issue slots nor shared-memory-pipe wavefronts
constexpr uint32_t LANES_ZERO_BYTES = 64 * 1024;
__device__ __align__(128) unsigned char g_lanes_zero[LANES_ZERO_BYTES];

template <int RP, int EPI, bool LOGD, int GMODE, bool W16>
__global__ void __launch_bounds__(256, 2) gather_lanes_kernel(const GatherParams prm) {
  constexpr bool GRAD = GMODE != 0;
  constexpr int SLOTS = W16 ? LANES_LOOKAHEAD / 2 : LANES_LOOKAHEAD;  // ring slots of 16 bytes per lane
  extern __shared__ __align__(16) float smem[];
  const int W = prm.tile_warps, K = prm.K;
  const int TP = prm.sub_pixels * W;
  const LanesLayout L = lanes_layout(K, RP, TP, W, GRAD, prm.n_runs, W16);
  float* Ssm = smem;
  float* gSsm = smem + L.gS;
  float* Csm = smem + L.C;
  float* gCw = smem + L.gC;
  constexpr int XROWS = 1 + LANES_CONT_ROWS;  // dummy row + continuation rows behind the K band rows
  __shared__ uint64_t mbar[8];
  __shared__ double wsum[8];
  __shared__ int done;
  // multi-level epilogues: the boundary table in shared memory (per-lane levels: a gather instead of a divergent
  // constant-bank index); not needed, and not allocated, by the packed one-bit epilogue
  constexpr bool NEED_BND = !(EPI == EPI_ONEBIT);
  __shared__ float bnd_sm[NEED_BND ? QMC_MAX_BOUNDS + 1 : 1];
  if (NEED_BND) {
    for (int i = threadIdx.x; i < prm.n_bounds; i += blockDim.x) bnd_sm[i] = prm.bounds[i];   // before the CTA barrier below
  }

  const int b = blockIdx.x / prm.tiles_per_map;
  const int tile = blockIdx.x - b * prm.tiles_per_map;
  const int p0 = tile * TP;
  const int np = min(TP, prm.IJ - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nthr = blockDim.x;
  // candidate batches (random-restart latent search): several maps of the launch share one observation set and C
  const int bo = prm.map_mod > 0 ? b % prm.map_mod : b;
  const float* __restrict__ Sb = prm.S + b * prm.sB;
  const float* __restrict__ Cb = prm.C + (int64_t)bo * prm.R * K;
  constexpr bool do_gs = GMODE == 1 || GMODE == 3 || GMODE == 4, do_gc = GMODE == 1 || GMODE == 2;
  constexpr bool FUSE = GMODE == 4;  // bulk layout guaranteed by the host
  float* gSb = (do_gs && !FUSE) ? prm.gS + b * prm.sB : nullptr;
  // pixel-major storage ([IJ][R], R == RP a multiple of 4): a slice is one contiguous, 16-byte aligned run
  const bool bulk = (prm.sR == 1 && prm.sP == RP && prm.R == RP && (RP % 4) == 0 &&
                     ((reinterpret_cast<uintptr_t>(Sb) | ((do_gs && !FUSE) ? reinterpret_cast<uintptr_t>(gSb) : 0)) & 15) == 0);
  // this warp's pixel slice of the tile
  const int sl0 = min(warp * prm.sub_pixels, np);
  const int sln = min((warp + 1) * prm.sub_pixels, np) - sl0;
  float* Sw = Ssm + (size_t)sl0 * RP;
  float* gSw = gSsm + (size_t)sl0 * RP;

  // ---- prologue: every warp stages its own slice; one CTA barrier for C -------------------------
  const uint32_t slice_bytes = (uint32_t)sln * RP * sizeof(float);
  const bool zero_by_copy = do_gs && (RP % 4) == 0 && slice_bytes <= LANES_ZERO_BYTES;  // the shared-memory side is always aligned
  const bool use_bar = sln > 0 && (bulk || zero_by_copy);
  if (use_bar && lane == 0) {
    mbar_init(&mbar[warp], 1);
    mbar_expect_tx(&mbar[warp], (bulk ? slice_bytes : 0u) + (zero_by_copy ? slice_bytes : 0u));
    if (bulk) bulk_g2s(Sw, Sb + (int64_t)(p0 + sl0) * RP, slice_bytes, &mbar[warp]);
    if (zero_by_copy) bulk_g2s(gSw, g_lanes_zero, slice_bytes, &mbar[warp]);
  }
  if (threadIdx.x == 0) done = 0;
  // every global load of the prologue is issued before anything waits on one of them (the warp issues in
  // order: a store of loaded data would hold back the loads behind it for a full memory round trip)
  const int64_t stream = (int64_t)bo * prm.n_sub + (int64_t)tile * W + warp;
  const uint32_t* sbase = prm.words + (prm.stream_stride > 0 ? stream * prm.stream_stride : prm.stream_off[stream]);
  const uint4* gp = reinterpret_cast<const uint4*>(sbase + (size_t)prm.n_runs * 32) + lane;
  // the run table of the stream (n_runs entries per lane) and SLOTS ring slots of look-ahead, filled by
  // asynchronous copies (the stream is read once, straight from DRAM; cp.async groups complete in order, which
  // plain loads sharing a scoreboard do not guarantee).  Every stream has room for at least SLOTS slots, so
  // the first copies need not know its length.
  const uint32_t hdr_a = smem_u32(smem + L.hdr) + (uint32_t)(warp * prm.n_runs * 32 + lane) * 4u;
  const uint32_t ring_a = smem_u32(smem + L.ring) + (uint32_t)((warp * SLOTS) * 32 + lane) * 16u;
  // emitter-major storage with contiguous pixels (the reference's [R][IJ]; rows are not 16-byte aligned at IJ = 2601, so
  // no bulk copy): lane = pixel.  Each lane loads the R emitters of its pixels with coalesced 4-byte loads (one
  // 128-byte run per emitter and instruction), NU pixels per lane in flight at once, and stores whole [p][r] rows
  // with one 16-byte shared-memory store each -- a transposition through registers, ~5 instructions per pixel row
  // instead of ~15 for element-wise copies.  The loads are issued here, with the other loads of the prologue; the
  // stores follow further down, once everything else is on its way.
  constexpr int NU = RP <= 4 ? 11 : RP == 8 ? 5 : 2;   // pixels per lane in flight (registers: NU * RP)
  const bool em_rows = !bulk && prm.sP == 1 && (RP % 4) == 0;
  float emv[(RP % 4) == 0 ? NU : 1][RP];
  const int em_rem = sln - lane;   // pixel u * 32 + lane is inside the slice iff u * 32 < em_rem
  if (em_rows) {
    // one base pointer per emitter row; the loads then differ by immediate offsets only, and one comparison per
    // pixel is the whole predicate when the rank is its own padded rank (the usual case)
    const float* erow[RP];
#pragma unroll
    for (int r = 0; r < RP; ++r) erow[r] = Sb + (int64_t)(r < prm.R ? r : 0) * prm.sR + (p0 + sl0) + lane;
    if (prm.R == RP) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const bool in = u * 32 < em_rem;
#pragma unroll
        for (int r = 0; r < RP; ++r) emv[u][r] = in ? __ldg(erow[r] + u * 32) : 0.0f;
      }
    } else {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
#pragma unroll
        for (int r = 0; r < RP; ++r) emv[u][r] = (u * 32 < em_rem && r < prm.R) ? __ldg(erow[r] + u * 32) : 0.0f;
      }
    }
  } else if (!bulk) {
    // any other strides: the slice is transposed into [p][r] rows by 4-byte asynchronous copies, all in flight at
    // once and in a cp.async group of their own, older than the ring's.  RP consecutive lanes take the RP emitters
    // of one pixel: the shared-memory side of a copy is 32 consecutive words (no bank conflict).
    constexpr int PPI = RP <= 32 ? 32 / RP : 1;  // pixels per copy instruction
    const int r = lane % RP, pq = lane / RP;
    const uint32_t Sw_a = smem_u32(Sw);
    const float* src = Sb + (int64_t)r * prm.sR + (int64_t)(p0 + sl0) * prm.sP;
    if (r < prm.R) {
      for (int pl = pq; pl < sln; pl += PPI) cp_async4(Sw_a + (uint32_t)(pl * RP + r) * 4u, src + (int64_t)pl * prm.sP);
    } else {
      for (int pl = pq; pl < sln; pl += PPI) Sw[pl * RP + r] = 0.0f;  // padded rank
    }
    cp_async_commit();
  }
  for (int i = 0; i < prm.n_runs; ++i) cp_async4(hdr_a + i * 128, sbase + i * 32 + lane);
#pragma unroll
  for (int d = 0; d < SLOTS; ++d) {
    cp_async16(ring_a + d * 512, gp + d * 32);
    cp_async_commit();  // the first group also holds the run table
  }
  const int nrows_v = __ldg(prm.nrows + stream);
  const int cn = RP * (K + 1);
  float cv0 = 0.0f, cv1 = 0.0f;  // C staged as [k][r]: element i -> (r = i / (K+1), k = i % (K+1))
  {
    const int i0 = threadIdx.x, i1 = threadIdx.x + nthr;
    const int r0 = i0 / (K + 1), k0 = i0 - r0 * (K + 1), r1 = i1 / (K + 1), k1 = i1 - r1 * (K + 1);
    if (i0 < cn && r0 < prm.R && k0 < K) cv0 = __ldg(Cb + r0 * K + k0);
    if (i1 < cn && r1 < prm.R && k1 < K) cv1 = __ldg(Cb + r1 * K + k1);
  }
  // pull the first bytes the CTA that will follow this one on its SM needs (S slice, C, head of the
  // stream) into L2 now, so that its prologue does not wait on DRAM
  if (prm.stream_stride > 0 && prm.lookahead > 0 && (int64_t)blockIdx.x + prm.lookahead < (int64_t)gridDim.x) {
    const int nb = blockIdx.x + prm.lookahead;
    const int b2 = nb / prm.tiles_per_map, tile2 = nb - b2 * prm.tiles_per_map;
    const int b2o = prm.map_mod > 0 ? b2 % prm.map_mod : b2;
    const int64_t stream2 = (int64_t)b2o * prm.n_sub + (int64_t)tile2 * W + warp;
    if (lane < 16) prefetch_l2(reinterpret_cast<const char*>(prm.words + stream2 * prm.stream_stride) + lane * 128);
    if (lane == 16) prefetch_l2(prm.nrows + stream2);
    if (warp == 0 && lane >= 24) {
      const char* c2 = reinterpret_cast<const char*>(prm.C + (int64_t)b2o * prm.R * K);
      for (int o = (lane - 24) * 128; o < prm.R * K * 4; o += 8 * 128) prefetch_l2(c2 + o);
    }
    const int np2 = min(TP, prm.IJ - tile2 * TP);
    const int s20 = min(warp * prm.sub_pixels, np2), s2n = min((warp + 1) * prm.sub_pixels, np2) - s20;
    if (bulk) {
      const char* s2 = reinterpret_cast<const char*>(prm.S + b2 * prm.sB + (int64_t)(tile2 * TP + s20) * RP);
      for (int o = lane * 128; o < s2n * RP * 4; o += 32 * 128) prefetch_l2(s2 + o);
    } else if (prm.sP == 1) {  // emitter-major: R contiguous runs of the slice's pixels
      for (int r = 0; r < prm.R; ++r) {
        const char* s2 = reinterpret_cast<const char*>(prm.S + b2 * prm.sB + r * prm.sR + (tile2 * TP + s20));
        for (int o = lane * 128; o < s2n * 4 + 128; o += 32 * 128) prefetch_l2(s2 + o);
      }
    }
  }
  if (FUSE && sln > 0) {  // the Adam moments of this warp's slice are needed at the very end: pull them into L2 now
    const char* m2 = reinterpret_cast<const char*>(prm.adam_m + b * prm.sB + (int64_t)(p0 + sl0) * RP);
    const char* v2 = reinterpret_cast<const char*>(prm.adam_v + b * prm.sB + (int64_t)(p0 + sl0) * RP);
    for (int o = lane * 128; o < (int)slice_bytes; o += 32 * 128) {
      prefetch_l2(m2 + o);
      prefetch_l2(v2 + o);
    }
  }
  if (em_rows) {
    if constexpr ((RP % 4) == 0) {
      float* srow = Sw + lane * RP;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        if (u * 32 < em_rem) {
#pragma unroll
          for (int r = 0; r < RP; r += 4)
            *reinterpret_cast<float4*>(srow + u * 32 * RP + r) = make_float4(emv[u][r], emv[u][r + 1], emv[u][r + 2], emv[u][r + 3]);
        }
      }
      // slices longer than 32 * NU pixels: the rest in further rounds (each one a memory round trip)
      for (int base = 32 * NU; base < sln; base += 32 * NU) {
        const float* src = Sb + (int64_t)(p0 + sl0) + base + lane;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
#pragma unroll
          for (int r = 0; r < RP; ++r)
            emv[u][r] = (base + u * 32 + lane < sln && r < prm.R) ? __ldg(src + (int64_t)r * prm.sR + u * 32) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int pl = base + u * 32 + lane;
          if (pl < sln) {
#pragma unroll
            for (int r = 0; r < RP; r += 4)
              *reinterpret_cast<float4*>(Sw + pl * RP + r) = make_float4(emv[u][r], emv[u][r + 1], emv[u][r + 2], emv[u][r + 3]);
          }
        }
      }
    }
  }
  if (GRAD) {
    float* zc = gCw + (size_t)warp * (K + XROWS) * RP;
    if (RP % 4 == 0) {
      float4* z = reinterpret_cast<float4*>(gSw);
      for (int i = lane; i < ((do_gs && !zero_by_copy) ? sln * (RP / 4) : 0); i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      float4* zc4 = reinterpret_cast<float4*>(zc);
      for (int i = lane; i < (K + XROWS) * (RP / 4); i += 32) zc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int i = lane; i < (do_gs ? sln * RP : 0); i += 32) gSw[i] = 0.0f;
      for (int i = lane; i < (K + XROWS) * RP; i += 32) zc[i] = 0.0f;
    }
  }
  {
    const int i0 = threadIdx.x, i1 = threadIdx.x + nthr;
    if (i0 < cn) Csm[(i0 % (K + 1)) * RP + i0 / (K + 1)] = cv0;
    if (i1 < cn) Csm[(i1 % (K + 1)) * RP + i1 / (K + 1)] = cv1;
    for (int i = threadIdx.x + 2 * nthr; i < cn; i += nthr) {
      const int r = i / (K + 1), k = i - r * (K + 1);
      Csm[k * RP + r] = (r < prm.R && k < K) ? __ldg(Cb + r * K + k) : 0.0f;
    }
  }
  const int ngroups = nrows_v >> 2;
  __syncthreads();  // Csm and `done` are ready; the slices are private to their warps
  if (use_bar) mbar_wait(&mbar[warp], 0);
  if (!bulk && !em_rows) cp_async_wait<SLOTS>();  // the S slice (every group but the ring's SLOTS newest)
  __syncwarp();

  constexpr uint32_t ROWB = RP * sizeof(float);
  const uint32_t S_a = smem_u32(Ssm), C_a = smem_u32(Csm);
  const uint32_t gS_delta = smem_u32(gSsm) - S_a;
  const uint32_t gC_a = smem_u32(gCw + (size_t)warp * (K + XROWS) * RP);

  float nll_part = 0.0f;               // generic epilogues: sum of -log P
  f2 nl2 = mk2(0.0f, 0.0f);            // packed one-bit epilogue: sum of log2 P, two partial sums
  float c[RP], acc[RP];
  uint32_t cur_row = (uint32_t)K;      // gC row of the current run (dummy until the first run)
  int run_left = 0;                    // groups left in the current run
  uint32_t run_a = hdr_a;              // next run-table entry of this lane
#pragma unroll
  for (int r = 0; r < RP; ++r) { c[r] = 0.0f; acc[r] = 0.0f; }

  auto flush_run = [&]() {  // one plain store per run: the row is this lane's alone
    if (GRAD) {
      const uint32_t grow = gC_a + cur_row * ROWB;
      if (RP % 4 == 0) {
#pragma unroll
        for (int r = 0; r < RP; r += 4) sts128(grow + r * 4, make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]));
      } else {
#pragma unroll
        for (int r = 0; r < RP; ++r) sts32(grow + r * 4, acc[r]);
      }
    }
  };
  auto next_run = [&]() {
    flush_run();
    const uint32_t e = lds32u(run_a);
    run_a += 128;
    cur_row = e & LW_RUN_ROW_MASK;
    run_left = (int)(e >> LW_RUN_LEN_SHIFT);
    const uint32_t crow = C_a + ((e >> LW_RUN_BAND_SHIFT) & 0x1FFu) * ROWB;
    if (RP % 4 == 0) {
#pragma unroll
      for (int r = 0; r < RP; r += 4) {
        const float4 c4 = lds128_ro(crow + r * 4);  // read-only tile: free to move above the pending gS updates
        c[r] = c4.x; c[r + 1] = c4.y; c[r + 2] = c4.z; c[r + 3] = c4.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < RP; ++r) c[r] = lds32_ro(crow + r * 4);
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) acc[r] = 0.0f;
  };

  // constants of the packed one-bit epilogue: everything in units scaled by s = sqrt(log2 e), so that
  // exp(-u^2) = 2^(-(s u)^2) and log P comes out in log2 units
  constexpr float kS = 1.2011224087864498f;  // sqrt(log2(e))
  constexpr float ecf[QMC_ERFCX_DEG + 1] = QMC_ERFCX_COEFFS;
  const f2 nthr2 = bc2(-prm.thr), inva2 = bc2(prm.inv_a * kS);
  const f2 kg2 = bc2(kInvSqrtPi * prm.inv_a);

  // word decoding.  cw: the word with its level field at the top (bit 31 = the one-bit level)
  constexpr int SH = RP == 1 ? 2 : RP == 2 ? 3 : RP == 4 ? 4 : RP == 8 ? 5 : RP == 16 ? 6 : 7;  // log2(ROWB)
  const int lvl_bits = (W16 && EPI != EPI_ONEBIT) ? prm.lvl_bits : 1;
  const uint32_t pixm16 = (1u << (15 - lvl_bits)) - 1u;
  auto w_ok = [&](uint32_t cw) -> bool { return W16 ? ((cw >> (31 - lvl_bits)) & 1u) == 0u : cw < 0xFF000000u; };
  auto w_off = [&](uint32_t cw) -> uint32_t {  // byte offset of the pixel's row in the S / gS tiles
    return W16 ? ((cw >> (16 - SH)) & (pixm16 << SH)) : ((cw & LW_PIX_MASK) << SH);
  };
  auto w_level = [&](uint32_t cw) -> int { return W16 ? (int)(cw >> (32 - lvl_bits)) : (int)(((cw >> 23) & 0xFEu) | (cw >> 31)); };

  // The gS updates of a group are applied one group late, in the same straight-line block as the next
  // group's likelihood: the read-modify-write chain (LDS -> FFMA2 -> STS, four in a row, ordered) then
  // overlaps the arithmetic instead of stalling the warp on the short scoreboard.
  uint32_t pw[4];   // pending group: row offsets
  float pg[4];      // their g = dNLL/dx; 0 for padding, and a zero g has nothing to store
#pragma unroll
  for (int j = 0; j < 4; ++j) { pw[j] = 0u; pg[j] = 0.0f; }
  auto apply_pending = [&](const float (&cp)[RP]) {  // cp: the C row of the pending group
    if (!do_gs) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t rs = S_a + gS_delta + pw[j];
      const bool pok = pg[j] != 0.0f;  // padding lanes (g = 0) must not store: their row is a real lane's
      if (RP % 4 == 0) {
        const f2 g2 = bc2(pg[j]);
#pragma unroll
        for (int r = 0; r < RP; r += 4) {
          float4 v = lds128(rs + r * 4);  // padding lanes read a real lane's row; only the store is predicated
          const f2 a = fma2(g2, mk2(cp[r], cp[r + 1]), mk2(v.x, v.y));
          const f2 bq = fma2(g2, mk2(cp[r + 2], cp[r + 3]), mk2(v.z, v.w));
          un2(a, v.x, v.y);
          un2(bq, v.z, v.w);
          sts128_if(rs + r * 4, v, pok);
        }
      } else {
#pragma unroll
        for (int r = 0; r < RP; ++r) sts32_if(rs + r * 4, fmaf(pg[j], cp[r], lds32(rs + r * 4)), pok);
      }
    }
  };

  auto group = [&](const uint32_t (&w)[4]) {
    float g[4], sv[4][RP];
    uint32_t soff[4];
    bool ok[4];
    float t[4];
    // program order matters to ptxas (shared-memory loads are not moved above stores that may alias): this
    // group's S rows and, on a run change, its C row are loaded first, then the pending gS updates are
    // issued, and the arithmetic below fills their latency
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ok[j] = w_ok(w[j]);
      soff[j] = w_off(w[j]);
      const uint32_t srow = S_a + soff[j];
      if (RP % 4 == 0) {
#pragma unroll
        for (int r = 0; r < RP; r += 4) {
          const float4 s4 = lds128_ro(srow + r * 4);  // padding words point at a real lane's row (broadcast)
          sv[j][r] = s4.x; sv[j][r + 1] = s4.y; sv[j][r + 2] = s4.z; sv[j][r + 3] = s4.w;
        }
      } else {
#pragma unroll
        for (int r = 0; r < RP; ++r) sv[j][r] = lds32_ro(srow + r * 4);
      }
    }
    float cp[RP];
#pragma unroll
    for (int r = 0; r < RP; ++r) cp[r] = c[r];
    if (run_left == 0) next_run();  // lanes change band only at group boundaries
    __syncwarp();  // reconverged: the gS updates below are applied in step order by the whole warp
    --run_left;
    apply_pending(cp);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (RP % 2 == 0) {
        f2 d = mul2(mk2(sv[j][0], sv[j][1]), mk2(c[0], c[1]));
#pragma unroll
        for (int r = 2; r < RP; r += 2) d = fma2(mk2(sv[j][r], sv[j][r + 1]), mk2(c[r], c[r + 1]), d);
        float lo, hi;
        un2(d, lo, hi);
        t[j] = lo + hi;
      } else {
        t[j] = 0.0f;
#pragma unroll
        for (int r = 0; r < RP; ++r) t[j] = fmaf(sv[j][r], c[r], t[j]);
      }
    }
    if (EPI == EPI_ONEBIT && !LOGD) {
      // P = 0.5 erfc(u), u = (x - thr) * (level ? -1 : +1) / a.  Padding: x := huge, level bit set (padding words
      // have all level bits set) -> u = -inf -> E = 0, P = 1, log P = 0, g = 0 with no further selects.
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const f2 x2 = mk2(ok[j] ? t[j] : 3.0e38f, ok[j + 1] ? t[j + 1] : 3.0e38f);
        const f2 v2 = mul2(add2(x2, nthr2), inva2);  // s * |u| up to sign
        float va, vb;
        un2(v2, va, vb);
        const f2 t2 = mk2(rcp_approx(fabsf(va) + QMC_ERFCX_C * kS), rcp_approx(fabsf(vb) + QMC_ERFCX_C * kS));
        const f2 q2 = fma2(t2, bc2(-2.0f * QMC_ERFCX_C * kS), bc2(1.0f));
        f2 p2 = bc2(-0.5f * kS * ecf[QMC_ERFCX_DEG]);
#pragma unroll
        for (int i = QMC_ERFCX_DEG - 1; i >= 0; --i) p2 = fma2(p2, q2, bc2(-0.5f * kS * ecf[i]));
        const f2 hn2 = mul2(p2, t2);  // -0.5 erfcx(|u|)
        const f2 w22 = mul2(v2, v2);  // (s u)^2
        float wa2, wb2, hna, hnb;
        un2(w22, wa2, wb2);
        un2(hn2, hna, hnb);
        const float Ea = ex2_approx(-wa2), Eb = ex2_approx(-wb2);
        const f2 om2 = fma2(mk2(Ea, Eb), hn2, bc2(1.0f));  // 1 - E * 0.5 erfcx
        float oma, omb;
        un2(om2, oma, omb);
        // sign(u) = sign(v) ^ level bit (bit 31 of the word)
        const bool posa = (int)(__float_as_uint(va) ^ w[j]) >= 0, posb = (int)(__float_as_uint(vb) ^ w[j + 1]) >= 0;
        const float arga = posa ? -hna : oma, argb = posb ? -hnb : omb;
        const f2 lg = mk2(lg2_approx(arga), lg2_approx(argb));
        const f2 sel = mk2(posa ? wa2 : 0.0f, posb ? wb2 : 0.0f);
        nl2 = add2(nl2, lg);
        nl2 = fma2(sel, bc2(-1.0f), nl2);
        const f2 fr = mul2(mk2(posa ? 1.0f : Ea, posb ? 1.0f : Eb), mk2(rcp_approx(arga), rcp_approx(argb)));
        const f2 ga = mul2(fr, kg2);
        float gaa, gab;
        un2(ga, gaa, gab);
        g[j] = __uint_as_float(__float_as_uint(gaa) ^ (w[j] & 0x80000000u));
        g[j + 1] = __uint_as_float(__float_as_uint(gab) ^ (w[j + 1] & 0x80000000u));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int lv = ok[j] ? w_level(w[j]) : 0;
        float dxdt;
        const BinEval ev = eval_entry<EPI, LOGD, true>(prm, NEED_BND ? bnd_sm : prm.bounds, t[j], (EPI == EPI_ONEBIT) ? (lv & 1) : lv, dxdt);
        nll_part -= ok[j] ? ev.logp : 0.0f;
        g[j] = ok[j] ? ev.gx * dxdt : 0.0f;
      }
    }
    auto gc_step = [&](int j) {  // gC: registers only
      if (RP % 2 == 0) {
        const f2 g2 = bc2(g[j]);
#pragma unroll
        for (int r = 0; r < RP; r += 2) {
          const f2 a = fma2(g2, mk2(sv[j][r], sv[j][r + 1]), mk2(acc[r], acc[r + 1]));
          un2(a, acc[r], acc[r + 1]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < RP; ++r) acc[r] = fmaf(g[j], sv[j][r], acc[r]);
      }
    };
    if (do_gc) {
#pragma unroll
      for (int j = 0; j < 4; ++j) gc_step(j);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { pw[j] = soff[j]; pg[j] = g[j]; }
  };

  {
    const int nslots = W16 ? (ngroups + 1) >> 1 : ngroups;
    const int last = max(nslots - 1, 0);
    int slot = 0, gleft = ngroups;
#pragma unroll 1
    for (int s = 0; s < nslots; ++s) {
      cp_async_wait<SLOTS - 1>();  // slot s has landed (this lane reads only what it copied itself)
      const uint32_t sa = ring_a + slot * 512;
      const uint4 wv = lds128_u4(sa);
      cp_async16(sa, gp + (size_t)min(s + SLOTS, last) * 32);  // refill the slot just read
      cp_async_commit();
      slot = (slot + 1) & (SLOTS - 1);
      if (W16) {
        // two groups per slot; 16-bit words moved to the top half of a 32-bit one
        const int ge = min(2, gleft);
        gleft -= 2;
#pragma unroll 1
        for (int q = 0; q < ge; ++q) {
          const uint32_t a = q ? wv.z : wv.x, bq = q ? wv.w : wv.y;
          const uint32_t w4[4] = {a << 16, a, bq << 16, bq};
          group(w4);
        }
      } else {
        const uint32_t w4[4] = {wv.x, wv.y, wv.z, wv.w};
        group(w4);
      }
    }
  }
  cp_async_wait<0>();
  __syncwarp();
  apply_pending(c);
  __syncwarp();
  flush_run();  // the last run
  if (do_gc && prm.has_cont) {
    // merge this warp's continuation rows into their bands' rows (a band may be continued by several lanes:
    // lanes whose bands coincide take turns)
    __syncwarp();
    const uint32_t e0 = lds32u(hdr_a);
    const uint32_t row0 = e0 & LW_RUN_ROW_MASK, band0 = (e0 >> LW_RUN_BAND_SHIFT) & 0x1FFu;
    bool todo = row0 > (uint32_t)K;
    while (__any_sync(0xffffffffu, todo)) {
      const unsigned m = __match_any_sync(0xffffffffu, todo ? band0 : (0x10000u + lane));
      const bool turn = todo && (__ffs(m) - 1) == lane;
      if (turn) {
        const uint32_t dst = gC_a + band0 * ROWB, src = gC_a + row0 * ROWB;
        if (RP % 4 == 0) {
#pragma unroll
          for (int r = 0; r < RP; r += 4) {
            float4 d = lds128(dst + r * 4);
            const float4 e = lds128(src + r * 4);
            d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
            sts128(dst + r * 4, d);
          }
        } else {
#pragma unroll
          for (int r = 0; r < RP; ++r) sts32(dst + r * 4, lds32(dst + r * 4) + lds32(src + r * 4));
        }
      }
      todo = todo && !turn;
      __syncwarp();
    }
  }

  // ---- epilogue: own gS slice out, then the last warp folds gC and the NLL -----------------------
  if (EPI == EPI_ONEBIT && !LOGD) {
    float lo, hi;
    un2(nl2, lo, hi);
    nll_part = -(lo + hi) * kLn2;
  }
  const double wsumv = warp_sum((double)nll_part);
  if (lane == 0) wsum[warp] = wsumv;
  if (FUSE && sln > 0) {
    // Adam + regulariser + projection on this warp's slice, straight from the S and gS tiles in shared
    // memory: p, m, v are contiguous in the pixel-major storage (row stride RP == R, checked by the host)
    __syncwarp();
    const int tstep = prm.step + (prm.step_dev ? *prm.step_dev : 0);
    const float bc1 = 1.0f - powf(prm.beta1, (float)tstep), bc2 = 1.0f - powf(prm.beta2, (float)tstep);
    const float step_size = prm.lr / bc1, bc2_sqrt = 1.0f / sqrtf(bc2);  // reciprocal: see adam_one
    const float nrm = (float)sqrt(prm.ss_in ? prm.ss_in[b] : 0.0);
    const float coef = (prm.lam != 0.0f && nrm > 0.0f) ? prm.lam / nrm : 0.0f;
    const float omb1 = 1.0f - prm.beta1, omb2 = 1.0f - prm.beta2;
    const bool proj = prm.project != 0;
    const int64_t goff = b * prm.sB + (int64_t)(p0 + sl0) * RP;
    float4* __restrict__ pg4 = reinterpret_cast<float4*>(prm.S_rw + goff);
    float4* __restrict__ mg4 = reinterpret_cast<float4*>(prm.adam_m + goff);
    float4* __restrict__ vg4 = reinterpret_cast<float4*>(prm.adam_v + goff);
    const float4* ps4 = reinterpret_cast<const float4*>(Sw);
    const float4* gs4 = reinterpret_cast<const float4*>(gSw);
    const int nvec = sln * (RP / 4);
    double accsq = 0.0;
    constexpr int FU = 4;  // vectors per lane in flight
    for (int i0 = lane; i0 < nvec; i0 += 32 * FU) {
      float4 mm[FU], vv[FU];
#pragma unroll
      for (int u = 0; u < FU; ++u) {
        const int i = i0 + 32 * u;
        if (i < nvec) { mm[u] = mg4[i]; vv[u] = vg4[i]; }
      }
#pragma unroll
      for (int u = 0; u < FU; ++u) {
        const int i = i0 + 32 * u;
        if (i < nvec) {
          float4 pp = ps4[i];
          const float4 gg = gs4[i];
          pp.x = adam_one(pp.x, gg.x, mm[u].x, vv[u].x, coef, omb1, prm.beta2, omb2, step_size, bc2_sqrt, prm.eps, proj);
          pp.y = adam_one(pp.y, gg.y, mm[u].y, vv[u].y, coef, omb1, prm.beta2, omb2, step_size, bc2_sqrt, prm.eps, proj);
          pp.z = adam_one(pp.z, gg.z, mm[u].z, vv[u].z, coef, omb1, prm.beta2, omb2, step_size, bc2_sqrt, prm.eps, proj);
          pp.w = adam_one(pp.w, gg.w, mm[u].w, vv[u].w, coef, omb1, prm.beta2, omb2, step_size, bc2_sqrt, prm.eps, proj);
          pg4[i] = pp; mg4[i] = mm[u]; vg4[i] = vv[u];
          accsq += (double)pp.x * pp.x + (double)pp.y * pp.y + (double)pp.z * pp.z + (double)pp.w * pp.w;
        }
      }
    }
    accsq = warp_sum(accsq);
    if (lane == 0 && prm.ss_out) atomicAdd(prm.ss_out + b, accsq);
  }
  if (do_gs && !FUSE && sln > 0) {
    if (bulk) {
      fence_async_smem();  // generic-proxy writes to the slice -> visible to the bulk-copy engine
      __syncwarp();
      if (lane == 0) bulk_s2g(gSb + (int64_t)(p0 + sl0) * RP, gSw, (uint32_t)sln * RP * sizeof(float));
    } else {
      __syncwarp();
      if (prm.sP == 1 && (RP % 4) == 0) {
        // emitter-major rows: lane = pixel, one 16-byte read of the tile row per 4 emitters, R coalesced 4-byte stores;
        // one base pointer per emitter row, immediate offsets from there
        float* drow[RP];
#pragma unroll
        for (int r = 0; r < RP; ++r) drow[r] = gSb + (int64_t)(r < prm.R ? r : 0) * prm.sR + (p0 + sl0) + lane;
        const float* grow = gSw + lane * RP;
        const int rem = sln - lane;
        for (int base = 0; base < sln; base += 32 * NU) {
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            if (base + u * 32 < rem) {
#pragma unroll
              for (int r = 0; r < RP; r += 4) {
                const float4 v = *reinterpret_cast<const float4*>(grow + (base + u * 32) * RP + r);
                if (prm.R == RP) {
                  __stcg(drow[r] + base + u * 32, v.x);
                  __stcg(drow[r + 1] + base + u * 32, v.y);
                  __stcg(drow[r + 2] + base + u * 32, v.z);
                  __stcg(drow[r + 3] + base + u * 32, v.w);
                } else {
                  if (r < prm.R) __stcg(drow[r] + base + u * 32, v.x);
                  if (r + 1 < prm.R) __stcg(drow[r + 1] + base + u * 32, v.y);
                  if (r + 2 < prm.R) __stcg(drow[r + 2] + base + u * 32, v.z);
                  if (r + 3 < prm.R) __stcg(drow[r + 3] + base + u * 32, v.w);
                }
              }
            }
          }
        }
      } else {
        // any other strides: RP consecutive lanes take the RP emitters of one pixel -- conflict-free 4-byte reads
        // of the tile
        constexpr int PPI = RP <= 32 ? 32 / RP : 1;
        const int r = lane % RP, pq = lane / RP;
        if (r < prm.R) {
          float* dst = gSb + (int64_t)r * prm.sR + (int64_t)(p0 + sl0 + pq) * prm.sP;
          const int64_t dstep = (int64_t)PPI * prm.sP;
          const float* srcp = gSw + pq * RP + r;
          for (int pl = pq; pl < sln; pl += PPI, dst += dstep, srcp += PPI * RP) *dst = *srcp;
        }
      }
    }
  }
  __threadfence_block();
  __syncwarp();
  int prev = 0;
  if (lane == 0) prev = atomicAdd(&done, 1);
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev == W - 1) {
    __threadfence_block();
    if (lane == 0) {
      double tot = 0.0;
      for (int i = 0; i < W; ++i) tot += wsum[i];
      if (prm.tiles_per_map == 1) prm.nll[b] = tot;
      else atomicAdd(prm.nll + b, tot);
    }
    if (do_gc) {
      float* gCb = prm.gC + (int64_t)b * prm.R * K;
#pragma unroll
      for (int r = 0; r < RP; ++r) {
        if (r >= prm.R) break;
        for (int k = lane; k < K; k += 32) {
          float v = 0.0f;
          for (int w2 = 0; w2 < W; ++w2) v += gCw[((size_t)w2 * (K + XROWS) + k) * RP + r];
          if (prm.tiles_per_map == 1) gCb[r * K + k] = v;
          else atomicAdd(gCb + r * K + k, v);
        }
      }
    }
  }
  if (do_gs && !FUSE && bulk && sln > 0 && lane == 0) bulk_wait_all();  // the slice must stay alive until the engine has read it
}

template <int RP, int EPI, bool LOGD, bool GRAD, bool W16>
static int launch_lanes_one(const GatherParams& prm, cudaStream_t st) {
  const size_t smem = lanes_smem_bytes(prm.K, RP, prm.sub_pixels, prm.tile_warps, GRAD, prm.n_runs, W16);
  auto kern = gather_lanes_kernel<RP, EPI, LOGD, 0, W16>;
  if (GRAD) {
    if (prm.fuse_update && RP % 4 == 0) kern = gather_lanes_kernel<RP, EPI, LOGD, (GRAD && RP % 4 == 0) ? 4 : 0, W16>;
    else if (prm.want_gs && prm.want_gc) kern = gather_lanes_kernel<RP, EPI, LOGD, GRAD ? 1 : 0, W16>;
    else if (prm.want_gc) kern = gather_lanes_kernel<RP, EPI, LOGD, GRAD ? 2 : 0, W16>;
    else kern = gather_lanes_kernel<RP, EPI, LOGD, GRAD ? 3 : 0, W16>;
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
int launch_lanes_rp(const GatherParams& prm, int epi, bool logd, bool grad, cudaStream_t st) {
#define QMC_GO(E, L, G)                                                   \
  do {                                                                    \
    if (prm.word16) return launch_lanes_one<RP, E, L, G, true>(prm, st);  \
    return launch_lanes_one<RP, E, L, G, false>(prm, st);                 \
  } while (0)
  QMC_GATHER_SWITCH(QMC_GO);
#undef QMC_GO
  return QMC_OK;
}

}  // namespace qmc
