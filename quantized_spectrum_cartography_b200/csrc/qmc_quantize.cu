// Quantizer (a8) and the compact-observation builder.
//   qmc_noisy_signal    : quantization_model.py:13 / quantization_model_log.py:14
//   qmc_quantize_levels : quantization_model.py:14-20 (bit-exact level assignment)
//   qmc_obs_count_scan / qmc_obs_fill : (Y, Wx) dense [B][K][IJ] -> (idx, lvl, row_off)
// (the lane-stream builder lives in qmc_lanes_build.cu)
#include "qmc_common.cuh"

namespace qmc {

struct BoundsTable {
  int n;
  float b[QMC_MAX_BOUNDS];
};

__global__ void noisy_kernel(const float* __restrict__ x, const float* __restrict__ noise, float std,
                             float offset, int log_domain, int64_t n, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float base = x[i];
    if (log_domain) base = logf(__fadd_rn(base, offset));
    // multiply, round, add, round: what the CPU reference does (no contraction into an FMA)
    out[i] = noise ? __fadd_rn(base, __fmul_rn(noise[i], std)) : base;
  }
}

// Replays the reference's overwrite loop: for i = 1..n-2, "if b[i] < v <= b[i+1] then level = i",
// with b[n-1] treated as +inf.  For a sorted table at most one i matches; for an unsorted one the
// last match wins, exactly as in the reference.  NaN compares false everywhere -> level 0.
__global__ void quantize_kernel(const float* __restrict__ v, int64_t n, const BoundsTable tab,
                                uint8_t* __restrict__ lvl, int64_t* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float x = v[i];
    int level = 0;
    for (int j = 1; j < tab.n - 1; ++j) {
      const float up = (j + 1 == tab.n - 1) ? __int_as_float(0x7f800000) : tab.b[j + 1];
      if (tab.b[j] < x && x <= up) level = j;
    }
    if (lvl) lvl[i] = (uint8_t)level;
    if (y) y[i] = level;
  }
}

// ---- observation builder ---------------------------------------------------------------------
// One warp per row (b, s, k): the row's pixels are [s*SP, min((s+1)*SP, IJ)).
__global__ void obs_count_kernel(const float* __restrict__ wx, int64_t n_rows, int K, int IJ, int n_sub,
                                 int SP, int64_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int k = (int)(row % K);
  const int64_t bs = row / K;
  const int s = (int)(bs % n_sub);
  const int64_t b = bs / n_sub;
  const int pb = s * SP, pe = min(pb + SP, IJ);
  int cnt = 0;
  if (!wx) {
    cnt = max(pe - pb, 0);
  } else {
    const float* w = wx + ((int64_t)b * K + k) * IJ;
    for (int p = pb + lane; p < pe; p += 32) cnt += (w[p] != 0.0f);
    cnt = (int)warp_sum((float)cnt);  // exact: counts << 2^24
  }
  if (lane == 0) counts[row] = cnt;
}

// Order of the entries inside a row.  bank_mod <= 1: increasing pixel.  bank_mod = M > 1: the
// shared-memory-friendly order used by the tiled kernel -- entries are dealt round-robin over the
// residue classes c = p mod M (level t holds the t-th entry of every class that has one, classes in
// increasing order), so that any M consecutive entries of a row touch M different bank groups when
// the kernel gathers the [pixel][R] rows of S / gS.  Within a row the kernel is order-agnostic.
template <typename YT>
__global__ void obs_fill_kernel(const YT* __restrict__ y, const float* __restrict__ wx, int64_t n_rows,
                                int K, int IJ, int n_sub, int SP, int bank_mod,
                                const int64_t* __restrict__ row_off, int32_t* __restrict__ idx,
                                uint8_t* __restrict__ lvl) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int k = (int)(row % K);
  const int64_t bs = row / K;
  const int s = (int)(bs % n_sub);
  const int64_t b = bs / n_sub;
  const int pb = s * SP, pe = min(pb + SP, IJ);
  const int64_t plane = ((int64_t)b * K + k) * IJ;
  const int64_t out0 = row_off[row];
  if (bank_mod <= 1) {
    int64_t out = out0;
    for (int p0 = pb; p0 < pe; p0 += 32) {
      const int p = p0 + lane;
      const bool on = p < pe && (!wx || wx[plane + p] != 0.0f);
      const unsigned m = __ballot_sync(0xffffffffu, on);
      if (on) {
        const int64_t o = out + __popc(m & ((1u << lane) - 1u));
        idx[o] = k * IJ + p;
        lvl[o] = (uint8_t)y[plane + p];
      }
      out += __popc(m);
    }
    return;
  }
  // pass 1: class sizes.  Lane c (c < M) keeps the running count of class c.
  const int M = bank_mod;  // power of two <= 32
  int cnt = 0;
  for (int p0 = pb; p0 < pe; p0 += 32) {
    const int p = p0 + lane;
    const bool on = p < pe && (!wx || wx[plane + p] != 0.0f);
    const unsigned m = __ballot_sync(0xffffffffu, on);
    // lanes whose pixel has residue `lane` (for lane < M): l with (p0 + l) % M == lane
    unsigned cls = 0;
    if (lane < M) {
      const int first = ((lane - p0) % M + M) % M;
      for (int l = first; l < 32; l += M) cls |= 1u << l;
    }
    cnt += __popc(m & cls);
  }
  // pass 2: position = #entries in lower levels + #lower classes present in my level
  int seen = 0;  // lane c: entries of class c placed so far
  for (int p0 = pb; p0 < pe; p0 += 32) {
    const int p = p0 + lane;
    const bool on = p < pe && (!wx || wx[plane + p] != 0.0f);
    const unsigned m = __ballot_sync(0xffffffffu, on);
    const int c = p & (M - 1);
    unsigned same = 0;  // lanes of this chunk with my residue
    {
      const int first = lane & (M - 1);
      for (int l = first; l < 32; l += M) same |= 1u << l;
    }
    const int rank = __shfl_sync(0xffffffffu, seen, c) + __popc(m & same & ((1u << lane) - 1u));
    int pos = 0;
    for (int c2 = 0; c2 < M; ++c2) {
      const int n2 = __shfl_sync(0xffffffffu, cnt, c2);
      pos += min(n2, rank) + ((c2 < c && n2 > rank) ? 1 : 0);
    }
    if (on) {
      const int64_t o = out0 + pos;
      idx[o] = k * IJ + p;
      lvl[o] = (uint8_t)y[plane + p];
    }
    unsigned cls = 0;
    if (lane < M) {
      const int first = ((lane - p0) % M + M) % M;
      for (int l = first; l < 32; l += M) cls |= 1u << l;
    }
    seen += __popc(m & cls);
  }
}

// ---- exclusive scan of int64 counts (in place: counts[i] -> offset, plus total at [n]) ----------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t& total) {
  __shared__ int64_t wtot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  int64_t wbase = 0, tot = 0;
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    if (w < warp) wbase += wtot[w];
    tot += wtot[w];
  }
  __syncthreads();
  total = tot;
  return wbase + inc - v;
}

// pass A: per-chunk sums
__global__ void scan_chunk_sums(const int64_t* __restrict__ data, int64_t n, int64_t* __restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK;
  int64_t v = 0;
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const int64_t i = base + (int64_t)threadIdx.x * SCAN_ITEMS + j;
    if (i < n) v += data[i];
  }
  int64_t tot;
  block_exclusive_scan(v, tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
// pass B: scan the chunk sums with one block (in place, exclusive; grand total at sums[n_chunks])
__global__ void scan_sums(int64_t* __restrict__ sums, int64_t n_chunks) {
  int64_t carry = 0;
  for (int64_t base = 0; base < n_chunks; base += SCAN_THREADS) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < n_chunks ? sums[i] : 0;
    int64_t tot;
    const int64_t ex = block_exclusive_scan(v, tot);
    if (i < n_chunks) sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) sums[n_chunks] = carry;
}
// pass C: final offsets.  `data` holds counts in [0, n) and receives offsets in [0, n]; processed
// chunk by chunk, each thread reads its items before anything of that chunk is overwritten.
__global__ void scan_apply(int64_t* __restrict__ data, int64_t n, const int64_t* __restrict__ sums, int64_t n_chunks) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK;
  int64_t item[SCAN_ITEMS];
  int64_t v = 0;
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const int64_t i = base + (int64_t)threadIdx.x * SCAN_ITEMS + j;
    item[j] = i < n ? data[i] : 0;
    v += item[j];
  }
  int64_t tot;
  int64_t ex = block_exclusive_scan(v, tot) + sums[blockIdx.x];
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const int64_t i = base + (int64_t)threadIdx.x * SCAN_ITEMS + j;
    if (i < n) data[i] = ex;
    ex += item[j];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) data[n] = sums[n_chunks];
}

}  // namespace qmc

using namespace qmc;

static int grid_for(int64_t n, int threads, int cap = 148 * 16) {
  int64_t g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  return (int)(g < cap ? g : cap);
}

extern "C" int qmc_noisy_signal(const float* x_dev, const float* noise_dev, float noise_std, float offset,
                                int log_domain, int64_t n, float* noisy_out_dev, void* stream) {
  QMC_REQUIRE(x_dev && noisy_out_dev && n >= 0, "bad arguments");
  if (n == 0) return QMC_OK;
  noisy_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x_dev, noise_dev, noise_std, offset,
                                                                   log_domain, n, noisy_out_dev);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

extern "C" int qmc_quantize_levels(const float* noisy_dev, int64_t n, const float* bounds_host, int n_bounds,
                                   uint8_t* lvl_out_dev, int64_t* y_out_dev, void* stream) {
  QMC_REQUIRE(noisy_dev && bounds_host && n >= 0, "bad arguments");
  QMC_REQUIRE(n_bounds >= 2 && n_bounds <= QMC_MAX_BOUNDS, "n_bounds %d out of range [2, %d]", n_bounds, QMC_MAX_BOUNDS);
  QMC_REQUIRE(lvl_out_dev || y_out_dev, "no output requested");
  if (n == 0) return QMC_OK;
  BoundsTable tab;
  tab.n = n_bounds;
  for (int i = 0; i < n_bounds; ++i) tab.b[i] = bounds_host[i];
  quantize_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(noisy_dev, n, tab, lvl_out_dev, y_out_dev);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

extern "C" int64_t qmc_obs_scan_ws_elems(int64_t n_rows) {
  return (n_rows + SCAN_CHUNK - 1) / SCAN_CHUNK + 2;
}

extern "C" int qmc_obs_count_scan(const float* wx_dev, int B, int K, int IJ, int n_sub, int sub_pixels,
                                  int64_t* row_off_dev, int64_t* scan_ws_dev, void* stream) {
  QMC_REQUIRE(row_off_dev && scan_ws_dev, "null argument");
  QMC_REQUIRE(B > 0 && K > 0 && IJ > 0 && n_sub > 0 && sub_pixels > 0, "bad sizes");
  QMC_REQUIRE((int64_t)n_sub * sub_pixels >= IJ, "sub-tiles (%d x %d) do not cover IJ=%d", n_sub, sub_pixels, IJ);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_rows = (int64_t)B * n_sub * K;
  const int64_t cblocks = (n_rows * 32 + 255) / 256;
  QMC_REQUIRE(cblocks <= 0x7fffffff, "too many rows");
  obs_count_kernel<<<(unsigned)cblocks, 256, 0, st>>>(wx_dev, n_rows, K, IJ, n_sub, sub_pixels, row_off_dev);
  const int64_t n_chunks = (n_rows + SCAN_CHUNK - 1) / SCAN_CHUNK;
  scan_chunk_sums<<<(unsigned)n_chunks, SCAN_THREADS, 0, st>>>(row_off_dev, n_rows, scan_ws_dev);
  scan_sums<<<1, SCAN_THREADS, 0, st>>>(scan_ws_dev, n_chunks);
  scan_apply<<<(unsigned)n_chunks, SCAN_THREADS, 0, st>>>(row_off_dev, n_rows, scan_ws_dev, n_chunks);
  count_launch(4);
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

extern "C" int qmc_obs_fill(const void* y_dev, int y_is_int64, const float* wx_dev, int B, int K, int IJ,
                            int n_sub, int sub_pixels, int bank_mod, const int64_t* row_off_dev,
                            int32_t* idx_out_dev, uint8_t* lvl_out_dev, void* stream) {
  QMC_REQUIRE(bank_mod >= 0 && bank_mod <= 32 && (bank_mod & (bank_mod - 1)) == 0, "bank_mod %d must be 0 or a power of two <= 32", bank_mod);
  QMC_REQUIRE(y_dev && row_off_dev && idx_out_dev && lvl_out_dev, "null argument");
  QMC_REQUIRE(B > 0 && K > 0 && IJ > 0 && n_sub > 0 && sub_pixels > 0, "bad sizes");
  QMC_REQUIRE((int64_t)K * IJ < (1LL << 31), "K*IJ does not fit the int32 linear index");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_rows = (int64_t)B * n_sub * K;
  const int64_t blocks = (n_rows * 32 + 255) / 256;
  QMC_REQUIRE(blocks <= 0x7fffffff, "too many rows");
  if (y_is_int64)
    obs_fill_kernel<int64_t><<<(unsigned)blocks, 256, 0, st>>>((const int64_t*)y_dev, wx_dev, n_rows, K, IJ, n_sub,
                                                               sub_pixels, bank_mod, row_off_dev, idx_out_dev, lvl_out_dev);
  else
    obs_fill_kernel<uint8_t><<<(unsigned)blocks, 256, 0, st>>>((const uint8_t*)y_dev, wx_dev, n_rows, K, IJ, n_sub,
                                                               sub_pixels, bank_mod, row_off_dev, idx_out_dev, lvl_out_dev);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}
