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
//   lanes : as tiled, on the lane-stream layout: every lane walks one band at a time with the band's C
//           row and gC accumulator in registers; gS rows are conflict-free shared-memory updates.
// The kernels live in qmc_gather_{flat,tiled,lanes}.cuh and are instantiated per padded rank in
// qmc_gather_inst.cu; this file is the C-ABI entry and the dispatch.
#include "qmc_gather_common.cuh"

using namespace qmc;

extern "C" int64_t qmc_lanes_smem_bytes(int K, int R, int sub_pixels, int tile_warps, int n_runs, int word_bits) {
  if (K <= 0 || R <= 0 || R > QMC_MAX_RANK || sub_pixels <= 0 || tile_warps <= 0 || tile_warps > 8 || n_runs <= 0) return 0;
  int RP = 1;
  while (RP < R) RP <<= 1;
  const size_t b = lanes_smem_bytes(K, RP, sub_pixels, tile_warps, true, n_runs, word_bits == 16);
  return b <= 227 * 1024 ? (int64_t)b : 0;
}

extern "C" int64_t qmc_tiled_smem_bytes(int K, int R, int sub_pixels, int tile_warps) {
  if (K <= 0 || R <= 0 || R > QMC_MAX_RANK || sub_pixels <= 0 || tile_warps <= 0 || tile_warps > 8) return 0;
  int RP = 1;
  while (RP < R) RP <<= 1;
  const size_t b = tiled_smem_bytes(K, RP, sub_pixels, tile_warps, true);
  return b <= 227 * 1024 ? (int64_t)b : 0;
}

namespace {
struct FusedUpdate {  // qmc_solver_s_step_fused
  float* S_rw;
  float* m;
  float* v;
  const double* ss_in;
  double* ss_out;
  float lr, beta1, beta2, eps, lam;
  int project, step;
  const int32_t* step_dev;
};
}  // namespace

static int gather_impl(const float* S_dev, int64_t s_stride_b, int64_t s_stride_r, int64_t s_stride_p,
                       const float* C_dev, const qmc_obs_view_t* obs, const qmc_likelihood_t* lik, int B, int IJ,
                       int K, int R, int algo, int tile_warps, double* nll_out_dev, float* gS_out_dev,
                       float* gC_out_dev, void* stream, const FusedUpdate* fu) {
  QMC_REQUIRE(S_dev && C_dev && obs && lik && nll_out_dev, "null argument");
  const bool lanes = obs->words_dev != nullptr;
  QMC_REQUIRE(lanes ? ((obs->stream_off_dev || obs->stream_stride > 0) && obs->nrows_dev)
                    : (obs->idx_dev && obs->lvl_dev && obs->row_off_dev),
              "null observation arrays");
  QMC_REQUIRE(B > 0 && IJ > 0 && K > 0 && R > 0, "bad sizes B=%d IJ=%d K=%d R=%d", B, IJ, K, R);
  QMC_REQUIRE(R <= QMC_MAX_RANK, "rank %d > %d", R, QMC_MAX_RANK);
  QMC_REQUIRE((int64_t)K * IJ < (1LL << 31), "K*IJ = %lld does not fit the int32 linear index", (long long)K * IJ);
  QMC_REQUIRE(lik->n_bounds >= 2 && lik->n_bounds <= QMC_MAX_BOUNDS, "n_bounds %d out of range", lik->n_bounds);
  QMC_REQUIRE(lik->noise_std > 0.0f || (lik->flags & QMC_EPI_LSQ), "noise_std must be positive");
  QMC_REQUIRE(obs->n_sub > 0 && obs->sub_pixels > 0 && (int64_t)obs->n_sub * obs->sub_pixels >= IJ,
              "sub-tiles (%d x %d) do not cover IJ=%d", obs->n_sub, obs->sub_pixels, IJ);
  const bool grad = !(lik->flags & QMC_FORWARD_ONLY);
  const bool lanes_set = obs->words_dev != nullptr;
  const bool want_gs = grad && !(lanes_set && (lik->flags & QMC_SKIP_GS));
  const bool want_gc = grad && !(lanes_set && (lik->flags & QMC_SKIP_GC));
  QMC_REQUIRE(fu || ((!want_gs || gS_out_dev) && (!want_gc || gC_out_dev)), "a requested gradient output is NULL");
  cudaStream_t st = (cudaStream_t)stream;

  GatherParams prm;
  prm.S = S_dev; prm.sB = s_stride_b; prm.sR = s_stride_r; prm.sP = s_stride_p;
  prm.C = C_dev; prm.idx = obs->idx_dev; prm.lvl = obs->lvl_dev; prm.row_off = obs->row_off_dev;
  prm.words = obs->words_dev; prm.stream_off = obs->stream_off_dev; prm.nrows = obs->nrows_dev;
  prm.stream_stride = lanes ? obs->stream_stride : 0;
  prm.n_runs = lanes ? obs->n_runs : 0;
  prm.word16 = lanes && obs->word_bits == 16;
  prm.lvl_bits = lanes ? obs->lvl_bits : 0;
  prm.has_cont = lanes ? obs->has_cont : 0;
  prm.map_mod = 0;
  if (obs->map_modulo > 0) {
    QMC_REQUIRE(lanes && !grad, "map_modulo (maps sharing one observation set) needs a lane-stream observation set and QMC_FORWARD_ONLY");
    QMC_REQUIRE(B % obs->map_modulo == 0, "B = %d is not a multiple of map_modulo = %d", B, obs->map_modulo);
    prm.map_mod = obs->map_modulo;
  }
  prm.lookahead = 0;
  prm.want_gs = want_gs; prm.want_gc = want_gc;
  prm.fuse_update = fu != nullptr;
  prm.S_rw = nullptr; prm.adam_m = nullptr; prm.adam_v = nullptr; prm.ss_in = nullptr; prm.ss_out = nullptr;
  prm.lr = prm.beta1 = prm.beta2 = prm.eps = prm.lam = 0.0f;
  prm.project = prm.step = 0; prm.step_dev = nullptr;
  if (fu) {
    prm.S_rw = fu->S_rw; prm.adam_m = fu->m; prm.adam_v = fu->v; prm.ss_in = fu->ss_in; prm.ss_out = fu->ss_out;
    prm.lr = fu->lr; prm.beta1 = fu->beta1; prm.beta2 = fu->beta2; prm.eps = fu->eps; prm.lam = fu->lam;
    prm.project = fu->project; prm.step = fu->step; prm.step_dev = fu->step_dev;
    prm.want_gs = 1; prm.want_gc = 0;
  }
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
  const float a = (lik->flags & QMC_EPI_LSQ) ? 1.0f                   // least squares has no noise model
                  : (lik->flags & QMC_EPI_LOGISTIC) ? lik->noise_std  // logistic scale, no sqrt(2)
                                                    : probit_scale(lik->noise_std);
  prm.inv_a = 1.0f / a;
  prm.offset = lik->offset;
  prm.n_bounds = lik->n_bounds;
  for (int i = 0; i < lik->n_bounds; ++i) prm.bounds[i] = lik->bounds[i];
  prm.thr = lik->n_bounds >= 3 ? lik->bounds[1] : 0.0f;

  // epilogue: reference-literal, one-bit fast path (both outer bounds numerically infinite), or general
  int epi = EPI_STABLE;
  if (lik->flags & QMC_EPI_LSQ) epi = EPI_LSQ;
  else if (lik->flags & QMC_EPI_LOGISTIC) epi = EPI_LOGISTIC;
  else if (lik->flags & QMC_EPI_REFERENCE) epi = EPI_REFERENCE;
  else if (lik->n_bounds == 3) {
    // erfc(z) == 0 exactly (even in double) for z > 27; require the sentinel to sit that far out for
    // any |x| < half its magnitude
    const float lo = lik->bounds[0], hi = lik->bounds[2];
    const bool inf_lo = lo < 0 && (-lo * 0.5f) * prm.inv_a > 30.0f;
    const bool inf_hi = hi > 0 && (hi * 0.5f) * prm.inv_a > 30.0f;
    if (inf_lo && inf_hi && fabsf(lo) >= 1e4f && fabsf(hi) >= 1e4f) epi = EPI_ONEBIT;
  }
  prm.one_sided = 0;
  if (epi == EPI_LOGISTIC && lik->n_bounds == 3) {
    // e^{-|z|} == 0 in fp32 (ex2 flushes) for |z| > 104: the sentinel must sit that far out for any |x| < half its size
    const float lo = lik->bounds[0], hi = lik->bounds[2];
    prm.one_sided = lo <= -1e4f && hi >= 1e4f && (-lo * 0.5f) * prm.inv_a > 110.0f && (hi * 0.5f) * prm.inv_a > 110.0f;
  }
  const bool logd = (lik->flags & QMC_LOG_DOMAIN) != 0;

  int RP = 1;
  while (RP < R) RP <<= 1;

  if (lanes) {
    QMC_REQUIRE(algo == QMC_ALGO_AUTO || algo == QMC_ALGO_LANES, "a lane-stream observation set needs QMC_ALGO_LANES");
    QMC_REQUIRE(K <= 256 && lik->n_bounds <= 256, "lane-stream layout: K <= 256 and at most 255 levels");
    QMC_REQUIRE(obs->n_runs >= 1 && obs->n_runs <= 64 && (obs->word_bits == 16 || obs->word_bits == 32),
                "lane-stream layout: n_runs %d / word_bits %d out of range", obs->n_runs, obs->word_bits);
    QMC_REQUIRE(obs->word_bits == 32 || (obs->lvl_bits >= 1 && obs->lvl_bits <= 8 &&
                                         (int64_t)obs->sub_pixels * tile_warps <= (1LL << (15 - obs->lvl_bits))),
                "lane-stream layout: 16-bit words cannot hold %d level bits and a tile of %lld pixels", obs->lvl_bits,
                (long long)obs->sub_pixels * tile_warps);
    QMC_REQUIRE(obs->word_bits == 32 || epi != EPI_ONEBIT || obs->lvl_bits == 1,
                "lane-stream layout: a one-bit model needs observation words with a 1-bit level field");
    algo = QMC_ALGO_LANES;
  } else {
    QMC_REQUIRE(algo != QMC_ALGO_LANES, "QMC_ALGO_LANES needs a lane-stream observation set (qmc_obs_build_lanes)");
  }
  if (algo == QMC_ALGO_AUTO) {
    const int64_t smem = qmc_tiled_smem_bytes(K, R, obs->sub_pixels, tile_warps > 0 ? tile_warps : 1);
    algo = (tile_warps > 0 && smem > 0 && obs->n_sub % tile_warps == 0 && (int64_t)B * (obs->n_sub / tile_warps) >= 64)
               ? QMC_ALGO_TILED : QMC_ALGO_FLAT;
  }
  if (algo == QMC_ALGO_LANES) {
    QMC_REQUIRE(tile_warps > 0 && tile_warps <= 8, "tile_warps %d out of range [1, 8]", tile_warps);
    QMC_REQUIRE(obs->n_sub % tile_warps == 0, "n_sub %d is not a multiple of tile_warps %d", obs->n_sub, tile_warps);
    QMC_REQUIRE(lanes_smem_bytes(K, RP, obs->sub_pixels, tile_warps, grad, obs->n_runs, obs->word_bits == 16) <= 227 * 1024,
                "tile of %d pixels x rank %d (+ %d private gC copies) does not fit shared memory",
                obs->sub_pixels * tile_warps, RP, tile_warps);
    prm.tile_warps = tile_warps;
    prm.tiles_per_map = obs->n_sub / tile_warps;
    {
      int dev = 0, sms = 0;
      QMC_CUDA_CHECK(cudaGetDevice(&dev));
      QMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      const size_t per_cta = lanes_smem_bytes(K, RP, obs->sub_pixels, tile_warps, grad, obs->n_runs, obs->word_bits == 16) + 1024 + 256;
      int per_sm = (int)((228 * 1024) / per_cta);
      per_sm = per_sm < 1 ? 1 : per_sm;
      const int by_threads = 2048 / (tile_warps * 32), by_regs = 65536 / (128 * tile_warps * 32);
      per_sm = per_sm < by_threads ? per_sm : by_threads;
      per_sm = per_sm < (by_regs < 1 ? 1 : by_regs) ? per_sm : (by_regs < 1 ? 1 : by_regs);
      prm.lookahead = sms * per_sm;
    }
  } else if (algo == QMC_ALGO_TILED) {
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
    if (want_gc) QMC_CUDA_CHECK(cudaMemsetAsync(gC_out_dev, 0, sizeof(float) * (size_t)B * R * K, st));
  }
  if (grad && algo == QMC_ALGO_FLAT) {
    // gS may be strided: zero the dense extent it spans only when it is compact
    const bool dense = (s_stride_p == 1 && s_stride_r == IJ) || (s_stride_r == 1 && s_stride_p == R);
    QMC_REQUIRE(dense && s_stride_b == (int64_t)R * IJ, "flat kernel needs a compact gS layout");
    QMC_CUDA_CHECK(cudaMemsetAsync(gS_out_dev, 0, sizeof(float) * (size_t)B * R * IJ, st));
  }

#define QMC_RP_CASE(N)                                                                          \
  case N:                                                                                       \
    return algo == QMC_ALGO_FLAT    ? launch_flat_rp<N>(prm, epi, logd, grad, st)               \
           : algo == QMC_ALGO_TILED ? launch_tiled_rp<N>(prm, epi, logd, grad, st)              \
                                    : launch_lanes_rp<N>(prm, epi, logd, grad, st)
  switch (RP) {
    QMC_RP_CASE(1);
    QMC_RP_CASE(2);
    QMC_RP_CASE(4);
    QMC_RP_CASE(8);
    QMC_RP_CASE(16);
    QMC_RP_CASE(32);
  }
#undef QMC_RP_CASE
  return set_error(QMC_ERR_UNSUPPORTED, "rank %d", R);
}

extern "C" int qmc_nll_fwd_bwd_gather(const float* S_dev, int64_t s_stride_b, int64_t s_stride_r,
                                      int64_t s_stride_p, const float* C_dev, const qmc_obs_view_t* obs,
                                      const qmc_likelihood_t* lik, int B, int IJ, int K, int R, int algo,
                                      int tile_warps, double* nll_out_dev, float* gS_out_dev,
                                      float* gC_out_dev, void* stream) {
  return gather_impl(S_dev, s_stride_b, s_stride_r, s_stride_p, C_dev, obs, lik, B, IJ, K, R, algo, tile_warps,
                     nll_out_dev, gS_out_dev, gC_out_dev, stream, nullptr);
}

extern "C" int qmc_solver_s_step_fused(float* S_dev, int64_t s_stride_b, int64_t s_stride_r, int64_t s_stride_p,
                                       const float* C_dev, const qmc_obs_view_t* obs, const qmc_likelihood_t* lik,
                                       int B, int IJ, int K, int R, int tile_warps, double* nll_out_dev,
                                       float* m_dev, float* v_dev, const double* sumsq_in_dev,
                                       double* sumsq_out_dev, float lr, float beta1, float beta2, float eps,
                                       float lam, int project, int step, const int32_t* step_dev, void* stream) {
  QMC_REQUIRE(S_dev && C_dev && obs && lik && nll_out_dev && m_dev && v_dev, "null argument");
  QMC_REQUIRE(lam == 0.0f || sumsq_in_dev, "the Frobenius regulariser needs the squared norms of S");
  QMC_REQUIRE(step >= 0 && (step > 0 || step_dev), "Adam steps count from 1");
  QMC_REQUIRE(!sumsq_out_dev || sumsq_out_dev != sumsq_in_dev, "sumsq_out must not alias sumsq_in");
  // the fused tail addresses S, m and v with the padded rank as the row stride: R must be its own padded
  // rank (a power of two)
  const bool layout_ok = obs->words_dev && R % 4 == 0 && (R & (R - 1)) == 0 && R <= QMC_MAX_RANK && s_stride_r == 1 && s_stride_p == R &&
                         s_stride_b == (int64_t)R * IJ && tile_warps > 0 && obs->n_sub == tile_warps &&
                         !(lik->flags & QMC_FORWARD_ONLY) &&
                         ((reinterpret_cast<uintptr_t>(S_dev) | reinterpret_cast<uintptr_t>(m_dev) |
                           reinterpret_cast<uintptr_t>(v_dev)) & 15) == 0;
  if (!layout_ok)
    return set_error(QMC_ERR_UNSUPPORTED, "fused S-step needs a lane-stream observation set with one tile per map, "
                                          "pixel-major S/m/v (16-byte aligned) and a rank of 4, 8, 16 or 32");
  if (sumsq_out_dev) QMC_CUDA_CHECK(cudaMemsetAsync(sumsq_out_dev, 0, sizeof(double) * B, (cudaStream_t)stream));
  FusedUpdate fu{S_dev, m_dev, v_dev, sumsq_in_dev, sumsq_out_dev, lr, beta1, beta2, eps, lam, project, step, step_dev};
  return gather_impl(S_dev, s_stride_b, s_stride_r, s_stride_p, C_dev, obs, lik, B, IJ, K, R, QMC_ALGO_LANES, tile_warps,
                     nll_out_dev, nullptr, nullptr, stream, &fu);
}
