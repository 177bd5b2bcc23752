/*
 * qmc_b200.h -- C ABI of libqmc_b200.so: the quantized-matrix-completion (QMC) maximum-likelihood
 * hot path of shresthasagar/quantized_spectrum_cartography on NVIDIA B200 (sm_100a).
 *
 * Plain pointers and sizes only: no torch types, no C++ types, no exceptions across the boundary.
 * Every function returns 0 on success and a QMC_ERR_* code otherwise; qmc_last_error() gives the
 * message for the calling thread.  All `*_dev` pointers are device pointers on the current CUDA
 * device, `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  The library
 * never allocates device memory (except qmc_peer_alloc, whose job it is): the caller owns every buffer.  Its process-wide state is the thread-local error
 * string, the launch counter, and -- for qmc_nll_fwd_bwd_gather_host only -- one lazily created set of internal
 * copy/compute streams per device, guarded by a per-device mutex (calls on one device are serialised, different
 * devices proceed independently).  There is no CPU implementation behind these entry points.
 *
 * The reference is a Python module namespace, not an FFI; each entry point
 * names the reference function(s) (file:line under /root/reference) whose arithmetic it replaces.
 * The reference-side binding a maintainer would add is in INTEGRATION.md.
 *
 * Tensor layouts (the reference's, after the notebook's permutes, qmc/qmc.ipynb c1:75-81):
 *   S  [B][R][IJ] fp32  spatial loss fields, one row per emitter   (reference: S[R,1,I,J])
 *   C  [B][R][K]  fp32  power spectra                              (reference: C[R,K])
 *   X  [B][K][IJ] fp32  band-major unfolding  X[k][p] = sum_r S[r][p] * C[r][k]
 *   Y  [B][K][IJ] int64 or uint8 level indices;  Wx [B][K][IJ] fp32 0/1 mask
 * B is the number of independent maps (1 for the reference's single-instance calls).  S may be
 * given with arbitrary element strides (s_stride_b/r/p) so that a pixel-major [IJ][R] storage, which
 * the batched kernel can stage with one bulk copy, is accepted as well.
 */
#ifndef QMC_B200_H
#define QMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define QMC_API __attribute__((visibility("default")))
#else
#define QMC_API
#endif

#define QMC_ABI_VERSION 3
#define QMC_MAX_BOUNDS 257 /* uint8 levels: at most 256 levels = 257 boundaries (qmc/utils.py:24) */
#define QMC_MAX_RANK 32

enum {
  QMC_OK = 0,
  QMC_ERR_INVALID = 1,     /* bad argument (shape, null pointer, unsupported size) */
  QMC_ERR_CUDA = 2,        /* a CUDA runtime call failed; message carries cudaGetErrorString */
  QMC_ERR_UNSUPPORTED = 3  /* valid request this build has no kernel for */
};

/* flags of qmc_likelihood_t */
enum {
  QMC_LOG_DOMAIN = 1u << 0,    /* x = log(t + offset) before the likelihood (qmc.ipynb c1:149) */
  QMC_EPI_REFERENCE = 1u << 1, /* evaluate P literally as the reference does, 0.5*(1+erf(zu)) -
                                  0.5*(1+erf(zl)) in fp32 (quantization_model.py:38,61): loses the
                                  tails exactly where the reference does.  Default is the numerically
                                  stable log-difference of scaled complementary error functions. */
  QMC_FORWARD_ONLY = 1u << 2,  /* NLL only, no gradients (random-restart search, qmc.ipynb c1:168-197) */
  QMC_SKIP_GS = 1u << 3,       /* the caller does not need gS (the C-step of the alternating solver, qmc.ipynb
                                  c1:140-154: S is detached): the lane-stream kernel skips its shared-memory
                                  updates and the write; other kernels ignore the flag.  gS_out may then be
                                  NULL for the lane-stream kernel only. */
  QMC_SKIP_GC = 1u << 4,       /* likewise for gC (the S-step, c1:199-212) */
  QMC_EPI_LSQ = 1u << 5,       /* masked least-squares baseline instead of the likelihood: every observed entry
                                  contributes (x - (bounds[l] + bounds[l+1])/2)^2, the de-quantised mid-point of
                                  get_quantized_obs_from_ordinal (quantization_model_log.py:43-51) under the cost
                                  norm(Wx*(T_hat - Obs))**2 of qmc_dowjons.ipynb c1:112,130.  noise_std is not
                                  used; nll_out receives the cost. */
  QMC_EPI_LOGISTIC = 1u << 6   /* logistic instead of Gaussian noise: P = F((hi - x)/s) - F((lo - x)/s) with
                                  F = F_sigmoid (quantization_model.py:43-47) and s = noise_std (s = 1 is the
                                  reference's F_sigmoid as it stands), evaluated as a stable log-difference. */
};

/*
 * Likelihood model.  `bounds` are the boundaries as prob_probit sees them, i.e. AFTER the
 * file-specific treatment of the two outer ones: quantization_model.py:31-33 overwrites them with
 * -1e5/+1e5, quantization_model_log.py:32-34 leaves them alone.  The host wrapper applies that.
 * noise_std is sigma of F_probit (quantization_model.py:57-61); the kernels use
 * a = (float)(sigma * 1.414213) exactly like the reference (truncated sqrt 2).
 */
typedef struct qmc_likelihood {
  int32_t n_bounds;  /* n boundaries => n-1 levels 0..n-2 */
  uint32_t flags;
  float noise_std;
  float offset;      /* log-link offset, used when QMC_LOG_DOMAIN */
  float bounds[QMC_MAX_BOUNDS];
} qmc_likelihood_t;

/*
 * Compact observation set: the information content of (Y, Wx) the likelihood actually uses
 * (only Wx != 0 entries contribute, qmc.ipynb c1:150).  Per observed entry a 4-byte linear index
 * idx = k*IJ + p into the reference's [K][IJ] layout and a 1-byte level.  Entries are grouped into
 * rows (map b, pixel sub-tile s, band k) -- sub-tile s covers pixels [s*sub_pixels, (s+1)*sub_pixels)
 * -- in row-major order of (b, s, k); the order inside a row is free (see qmc_obs_fill).  row_off has
 * B*n_sub*K + 1 entries.  With n_sub == 1 this is plain band-major order.
 */
typedef struct qmc_obs_view {
  const int32_t* idx_dev;
  const uint8_t* lvl_dev;
  const int64_t* row_off_dev;
  int32_t n_sub;       /* pixel sub-tiles per map */
  int32_t sub_pixels;  /* pixels per sub-tile */
  /* Lane-stream layout (qmc_obs_build_lanes), NULL otherwise: the entries of every (map, sub-tile)
   * stream re-cut so that each lane of the owning warp walks a list of runs -- a band and a number of
   * 4-step groups -- and the 32 entries of a step hit 32 different pixels; a lane changes band only at a
   * multiple of four steps.  A stream starts with its run table, n_runs entries per lane stored
   * [entry][lane]: bits 0..8 = row of the warp's gC copy the run's partial goes to (the band, or K+1+lane for
   * a piece that continues a band begun by another lane), bits 9..17 = band (K = none), bits 18..31 = groups
   * (the last entry of a lane never ends).  The words follow in 512-byte slots of 32 lanes x 16 bytes:
   *   word_bits == 32: one group per slot; bit 31 = level & 1, bits 24..30 = level >> 1, bits 0..14 = tile-local
   *                    pixel; level 0xFF (bits 24..31 all set) marks padding;
   *   word_bits == 16: two groups per slot (halfwords 0..3 and 4..7 of a lane's 16 bytes); the top lvl_bits bits
   *                    = level, the next bit = padding flag (padding also has all level bits set), the rest =
   *                    tile-local pixel.
   * idx/lvl/row_off are not used by the kernel. */
  const uint32_t* words_dev;
  const int64_t* stream_off_dev; /* B*n_sub + 1 offsets in 32-bit words, multiples of 32 */
  const int32_t* nrows_dev;      /* B*n_sub steps per stream, multiples of 4 */
  int64_t stream_stride;         /* > 0: stream s starts at word s * stream_stride (a multiple of 32) and
                                    stream_off_dev is not read -- lets a CTA prefetch a later CTA's data */
  int32_t n_runs;                /* run-table entries per lane */
  int32_t word_bits;             /* 16 or 32 */
  int32_t lvl_bits;              /* word_bits == 16: width of the level field */
  int32_t has_cont;              /* != 0: continuation rows are in use (bands split over lanes) */
  int32_t map_modulo;            /* > 0, lane-stream layout with QMC_FORWARD_ONLY: map b of the launch uses the
                                    observation streams and the C of map b % map_modulo -- D candidate factors S per map
                                    (the random-restart latent search, qmc.ipynb c1:168-197) are scored in ONE launch of
                                    B = D * map_modulo maps against one resident observation set */
} qmc_obs_view_t;

QMC_API int qmc_abi_version(void);
QMC_API const char* qmc_last_error(void);
/* number of kernels this library has launched on behalf of the calling process (all threads) */
QMC_API int64_t qmc_launch_count(void);

/* ---- a8: quantizer -------------------------------------------------------------------------- */

/* noisy = x + noise*noise_std, or log(x + offset) + noise*noise_std when log_domain != 0, with the
 * reference's rounding (separate multiply and add, no FMA): quantization_model.py:13,
 * quantization_model_log.py:14.  noise_dev may be NULL (noisy = x or log(x+offset)). */
QMC_API int qmc_noisy_signal(const float* x_dev, const float* noise_dev, float noise_std, float offset,
                     int log_domain, int64_t n, float* noisy_out_dev, void* stream);

/* Level assignment of an already-noisy signal, bit-exact with the loop of
 * quantization_model.py:14-20: level 0 unless bounds[i] < v <= bounds[i+1] for some i in [1, n-2]
 * (the last such i wins), the last boundary acting as +inf; NaN -> 0.  Either output may be NULL.
 * bounds_host: the caller's ORIGINAL table (this function applies the +inf itself). */
QMC_API int qmc_quantize_levels(const float* noisy_dev, int64_t n, const float* bounds_host, int n_bounds,
                        uint8_t* lvl_out_dev, int64_t* y_out_dev, void* stream);

/* ---- compact observation builder -------------------------------------------------------------- */

/* Pass 1: count observed entries per row (b, s, k) and exclusive-scan them into row_off
 * (B*n_sub*K + 1 int64).  wx_dev may be NULL (every entry observed).  scan_ws_dev: workspace of at
 * least qmc_obs_scan_ws_elems(n_rows) int64.  The total is row_off[n_rows]. */
QMC_API int64_t qmc_obs_scan_ws_elems(int64_t n_rows);
QMC_API int qmc_obs_count_scan(const float* wx_dev, int B, int K, int IJ, int n_sub, int sub_pixels,
                       int64_t* row_off_dev, int64_t* scan_ws_dev, void* stream);
/* Pass 2: write idx/lvl in row order.  y is int64 (y_is_int64 != 0, the reference's dtype) or uint8.
 * bank_mod: 0/1 = pixels increasing inside a row; M (power of two <= 32) = entries of a row dealt
 * round-robin over the residue classes p mod M, which makes the tiled kernel's shared-memory
 * gathers conflict-free (M = 128 / (4*R_padded)); the kernels accept any order inside a row. */
QMC_API int qmc_obs_fill(const void* y_dev, int y_is_int64, const float* wx_dev, int B, int K, int IJ,
                 int n_sub, int sub_pixels, int bank_mod, const int64_t* row_off_dev,
                 int32_t* idx_out_dev, uint8_t* lvl_out_dev, void* stream);

/* Re-cut a row-ordered observation set (n_sub, sub_pixels as built by qmc_obs_count_scan/qmc_obs_fill
 * with bank_mod = 0; read only) into the lane-stream layout for tiles of tile_warps sub-tiles: per stream the
 * bands are laid end to end in 4-step groups and cut into 32 equal quotas, one per lane (see qmc_obs_view_t and
 * csrc/qmc_lanes_build.cu).  Every stream occupies stream_stride 32-bit words of words_out_dev (a multiple of 32,
 * at least qmc_lanes_stream_words(max_entries_per_stream, K, n_runs, word_bits, extra_groups));
 * nrows_out_dev receives the steps actually used.  *overflow_dev comes back non-zero if some stream could not be
 * laid out: bits 0, 2, 3 = not every entry found a slot (capacity / too many short pieces / repair failed: retry
 * with more extra_groups), bit 1 = a lane needs more than n_runs - 1 runs (retry with a larger n_runs).  K <= 256, levels <= 254, tile_warps * sub_pixels + 32 <= 32768;
 * word_bits == 16 needs lvl_bits + 1 + ceil(log2(tile pixels)) <= 16. */
QMC_API int64_t qmc_lanes_stream_words(int64_t max_entries_per_stream, int K, int n_runs, int word_bits, int extra_groups);
QMC_API int qmc_obs_build_lanes(const int32_t* idx_rows_dev, const uint8_t* lvl_rows_dev, const int64_t* row_off_dev,
                        int B, int K, int IJ, int n_sub, int sub_pixels, int tile_warps,
                        int64_t max_entries_per_stream, int extra_groups, int split_bands, int64_t stream_stride,
                        uint32_t* words_out_dev, int32_t* nrows_out_dev, int32_t* overflow_dev, int n_runs,
                        int word_bits, int lvl_bits, void* stream);
/* Shared-memory bytes of the lanes kernel for a geometry (0 if it cannot run it). */
QMC_API int64_t qmc_lanes_smem_bytes(int K, int R, int sub_pixels, int tile_warps, int n_runs, int word_bits);

/* ---- a2..a7: fused masked low-rank reconstruction + quantized NLL + factor gradients ---------- */

/* algorithms of qmc_nll_fwd_bwd_gather */
enum {
  QMC_ALGO_AUTO = 0,
  QMC_ALGO_FLAT = 1,  /* one thread per observed entry, factors gathered through L2, gradients by
                         warp-aggregated global atomics: single small/medium instance */
  QMC_ALGO_TILED = 2, /* one CTA per (map, pixel tile): factor rows staged in shared memory, band
                         segments reduced in registers, no global atomics on gS: batched maps */
  QMC_ALGO_LANES = 3  /* as TILED, on a lane-stream observation set: each lane keeps its band's C row and gC
                         accumulator in registers; gS rows are plain shared-memory updates */
};

/*
 * Replaces, for every map b, the reference idiom (qmc.ipynb c1:145-153)
 *     T_hat = get_tensor(S, C).unsqueeze(1)            quantization_model.py:70-86
 *     T_hat = torch.log(T_hat + offset)                (QMC_LOG_DOMAIN)
 *     nll   = -torch.sum(Wx * torch.log(prob_probit(Y, T_hat, bb, std)))   :22-39, :57-61
 *     nll.backward()  ->  S.grad, C.grad
 * visiting only the observed entries.  nll_out_dev: B doubles; gS_out_dev: same strides as S;
 * gC_out_dev: [B][R][K].  Gradient outputs may be NULL with QMC_FORWARD_ONLY.  Outputs are
 * overwritten (the function zero-fills what its kernels accumulate into).
 * tile_warps: warps per CTA of the tiled kernel = sub-tiles per pixel tile (obs.n_sub must be a
 * multiple of it); ignored by the flat kernel.
 */
QMC_API int qmc_nll_fwd_bwd_gather(const float* S_dev, int64_t s_stride_b, int64_t s_stride_r,
                           int64_t s_stride_p, const float* C_dev, const qmc_obs_view_t* obs,
                           const qmc_likelihood_t* lik, int B, int IJ, int K, int R, int algo,
                           int tile_warps, double* nll_out_dev, float* gS_out_dev,
                           float* gC_out_dev, void* stream);

/* Shared-memory bytes the tiled kernel needs for a geometry (0 if it cannot run it). */
QMC_API int64_t qmc_tiled_smem_bytes(int K, int R, int sub_pixels, int tile_warps);

/*
 * End-to-end form for callers that hold the factors in HOST memory (pinned for full speed): copies
 * S and C to the device scratch buffers, runs qmc_nll_fwd_bwd_gather, copies nll/gS/gC back and
 * synchronises the stream.  S_host/gS_host are dense [B][R][IJ]; scratch buffers are caller-owned
 * device memory of the same sizes as the host arrays.
 */
QMC_API int qmc_nll_fwd_bwd_gather_host(const float* S_host, const float* C_host, float* S_scratch_dev,
                                float* C_scratch_dev, const qmc_obs_view_t* obs,
                                const qmc_likelihood_t* lik, int B, int IJ, int K, int R, int algo,
                                int tile_warps, double* nll_scratch_dev, float* gS_scratch_dev,
                                float* gC_scratch_dev, double* nll_host, float* gS_host,
                                float* gC_host, void* stream);

/* ---- dense-sampling path: tcgen05 tensor cores, likelihood as the GEMM epilogue ------------------ */

/* code8[b][p][k] = (Wx[b][k][p] != 0) ? Y[b][k][p] : 255 -- one byte per dense entry, pixel-major, the
 * observation format of the dense kernel (levels 0..254).  wx_dev may be NULL (everything observed). */
QMC_API int qmc_dense_pack(const void* y_dev, int y_is_int64, const float* wx_dev, int B, int K, int IJ,
                   uint8_t* code_out_dev, void* stream);

/* Same contract as qmc_nll_fwd_bwd_gather for ONE instance (B = 1) with S [R][IJ], C [R][K] contiguous:
 * X = S*C^T, gS = G*C and gC = G^T*S are formed by tcgen05.mma (kind::tf32, 3xTF32 operand split, fp32
 * accumulators in tensor memory); X and G = dNLL/dX never touch HBM.  gs_row_stride: elements between the rows of
 * gS_out_dev (0 = IJ; larger when the rows are written straight into a wider buffer, e.g. the pixel block of one
 * rank inside the all-reduce buffer of a sharded instance).  Supported: K a multiple of 32 and
 * <= 256, R <= 16, levels <= 255; otherwise QMC_ERR_UNSUPPORTED (use the gather entry point). */
QMC_API int qmc_nll_fwd_bwd_dense(const float* S_dev, const float* C_dev, const uint8_t* code_dev,
                          const qmc_likelihood_t* lik, int IJ, int K, int R, double* nll_out_dev,
                          float* gS_out_dev, int64_t gs_row_stride, float* gC_out_dev, void* stream);
/* Shared-memory bytes of the dense kernel for a geometry (0 = unsupported). */
QMC_API int64_t qmc_dense_smem_bytes(int K, int R);

/* ---- sharded instance: factor-gradient exchange fused into the dense kernel (peer memory over NVLink) --------
 * Replaces, for the pixel-block sharding of one oversized instance (BASELINE configs[3]), the separate collective
 * that follows the reference's `backward()` when the sum of qmc/qmc.ipynb c1:150 is split over devices: every rank
 * evaluates its pixel block, and the kernel's last CTA writes the rank's partial [gC | nll] into a slot of every
 * peer's exchange region (plain stores through NVLink peer mappings), publishes a flag, waits for the peers' flags
 * and adds the slots up in rank order -- every rank ends with the same bits, no NCCL call, no second launch.
 *
 * The partial sums of a launch's CTAs go through a scratch slot of the rank's own region that the kernel leaves
 * zeroed, so the call is one kernel launch and nothing else (no memset of the outputs).
 *
 * An exchange region is device memory of qmc_peer_region_bytes() bytes allocated by qmc_peer_alloc (which also
 * returns the 64-byte CUDA IPC handle the other ranks open with qmc_peer_open; exchange the handles with any host
 * mechanism, e.g. torch.distributed.all_gather_object).  All ranks must issue the same sequence of exchange calls
 * on a region set, with the same world, slot_floats, K and R; a peer that never arrives makes the kernel give up after
 * ~2 s, hand back this rank's own partial sums and set the region's status word (qmc_peer_status != 0).  Regions must
 * outlive every launch that names them: synchronise all ranks before qmc_peer_close / qmc_peer_free. */
#define QMC_PEER_MAX_WORLD 8
typedef struct qmc_peer_exchange {
  int32_t rank, world;                 /* 1 <= world <= QMC_PEER_MAX_WORLD */
  int32_t slot_floats;                 /* capacity of one slot in floats (>= R*K + 2) */
  int32_t reserved;
  void* region[QMC_PEER_MAX_WORLD];    /* region[q]: rank q's exchange region as mapped into THIS process */
} qmc_peer_exchange_t;
QMC_API int64_t qmc_peer_region_bytes(int world, int slot_floats);
QMC_API int qmc_peer_alloc(int64_t bytes, void** region_out_dev, void* ipc_handle_out64);
QMC_API int qmc_peer_open(const void* ipc_handle64, void** region_out_dev);
QMC_API int qmc_peer_close(void* region_dev);   /* a region opened with qmc_peer_open */
QMC_API int qmc_peer_free(void* region_dev);    /* a region allocated with qmc_peer_alloc */
QMC_API int qmc_peer_status(const void* own_region_dev, int* status_out, void* stream);   /* synchronises the stream */
/* qmc_nll_fwd_bwd_dense on this rank's pixel block + the fused exchange: on return of the kernel nll_out_dev and
 * gC_out_dev hold the sums over all ranks (gS_out_dev: this rank's block, complete by construction). */
QMC_API int qmc_nll_fwd_bwd_dense_exchange(const float* S_dev, const float* C_dev, const uint8_t* code_dev,
                                   const qmc_likelihood_t* lik, int IJ, int K, int R, double* nll_out_dev,
                                   float* gS_out_dev, int64_t gs_row_stride, float* gC_out_dev,
                                   const qmc_peer_exchange_t* px, void* stream);

/* ---- a1/a2/a10 helpers kept importable by the reference's call surface ------------------------ */

/* X[b][k][p] = sum_r S[b][r][p]*C[b][r][k] (+ optional log link): get_tensor, quantization_model.py:79-86 */
QMC_API int qmc_get_tensor(const float* S_dev, const float* C_dev, int B, int IJ, int K, int R,
                   float* X_out_dev, void* stream);

/* sum over entries of (X_hat - X_ref)^2 and X_ref^2 for X_hat = S*C^T formed on the fly (optionally
 * after log(. + offset)): the two Frobenius norms of NMSE / NMSE_LOG (quantization_model.py:88-92,
 * quantization_model_log.py:104-111) without materialising X_hat.  out_dev: 2*B doubles. */
QMC_API int qmc_nmse_terms(const float* S_dev, const float* C_dev, const float* X_ref_dev, int B, int IJ,
                   int K, int R, int log_domain, float offset, double* out_dev, void* stream);

/* a9, the one-bit BCE form of the likelihood: NegLikelihood.forward, quantization_model.py:97-113 --
 * nn.BCELoss()(F_probit(T - mean, std) or F_sigmoid(T - mean), target): mean over the n elements, logarithms
 * clamped at -100, in the reference's own fp32 arithmetic.  loss_out_dev: one double; gx_out_dev (may be NULL):
 * d loss / d x as torch's BCELoss backward forms it. */
QMC_API int qmc_bce_one_bit(const float* x_dev, const float* target_dev, int64_t n, float mean, float noise_std,
                    int probit, double* loss_out_dev, float* gx_out_dev, void* stream);

/* ---- solver step: what the notebook does with torch.optim.Adam around the path ---- */

/* out[b] = sum of squares of map b of x (n_per_map contiguous floats per map): ||.||_F^2 of
 * torch.norm(., 'fro'), qmc/qmc.ipynb c1:151,209. */
QMC_API int qmc_sumsq_per_map(const float* x_dev, int B, int64_t n_per_map, double* out_dev, void* stream);

/* One fused factor update per map b, elementwise over n_per_map contiguous floats:
 *   g' = g + lam * p / sqrt(sumsq_in[b])        gradient of lam*||p||_F (0 where the norm is 0)
 *   Adam(lr, beta1, beta2, eps) moments m, v and step number t, torch.optim.Adam arithmetic
 *   p  = max(p - step, 0) if project             the notebook's `X[X < 0] = 0`
 *   sumsq_out[b] = sum p^2 after the update      (NULL to skip; sumsq_in may be NULL when lam == 0)
 * t = step + (*step_dev if step_dev else 0), t >= 1.  Replaces opt.zero_grad / cost.backward's norm
 * term / opt.step / projection of qmc/qmc.ipynb c1:126-128,151-157,209-212. */
QMC_API int qmc_adam_frob_project(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, int B,
                          int64_t n_per_map, const double* sumsq_in_dev, double* sumsq_out_dev, float lr,
                          float beta1, float beta2, float eps, float lam, int project, int step,
                          const int32_t* step_dev, void* stream);

/* The S-step of the alternating solver in ONE launch (lane-stream observation set, S stored pixel-major,
 * R a multiple of 4, one tile per map): evaluates the likelihood at (S, C), keeps the gS tile in shared
 * memory and applies qmc_adam_frob_project's update to S, m, v in place from there -- gS never goes to
 * HBM.  nll_out_dev: NLL at the S the call was entered with.  sumsq_in/out as in qmc_adam_frob_project
 * (sumsq_out must not alias sumsq_in).  Returns QMC_ERR_UNSUPPORTED for any other layout: run
 * qmc_nll_fwd_bwd_gather + qmc_adam_frob_project instead.  qmc/qmc.ipynb c1:199-212 with S in place of
 * generator(Z). */
QMC_API int qmc_solver_s_step_fused(float* S_dev, int64_t s_stride_b, int64_t s_stride_r, int64_t s_stride_p,
                            const float* C_dev, const qmc_obs_view_t* obs, const qmc_likelihood_t* lik,
                            int B, int IJ, int K, int R, int tile_warps, double* nll_out_dev, float* m_dev,
                            float* v_dev, const double* sumsq_in_dev, double* sumsq_out_dev, float lr,
                            float beta1, float beta2, float eps, float lam, int project, int step,
                            const int32_t* step_dev, void* stream);

/* *counter_dev += add on the stream (the step counter of a CUDA-graph-captured solver iteration). */
QMC_API int qmc_counter_add(int32_t* counter_dev, int add, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QMC_B200_H */
