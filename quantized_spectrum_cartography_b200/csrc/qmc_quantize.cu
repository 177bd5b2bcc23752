// Quantizer (a8) and the compact-observation builder.
//   qmc_noisy_signal    : quantization_model.py:13 / quantization_model_log.py:14
//   qmc_quantize_levels : quantization_model.py:14-20 (bit-exact level assignment)
//   qmc_obs_count_scan / qmc_obs_fill : (Y, Wx) dense [B][K][IJ] -> (idx, lvl, row_off)
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

// ---- lane streams -------------------------------------------------------------------------------
// One warp per (map, sub-tile) stream.  The stream's rows (bands) are dealt to the 32 lanes in snake
// order of their sizes (largest, ..., 32nd | 64th, ..., 33rd | ...), so every lane owns about the same
// number of entries, and a lane walks its bands one after the other: in the gather kernel the lane
// keeps C[band] and the band's gC accumulator in registers.  Step t of the stream takes one entry
// from every lane; the entries of a step have pairwise distinct pixels (hard: the kernel updates gS
// rows without atomics) and, where the lane still has a choice, distinct shared-memory bank groups
// inside each quarter-warp (soft).  A lane that cannot comply, or has nothing left, emits a padding
// word; a lane changes band only at a multiple of four steps.  Word: bit 31 = level & 1, bits 24..30 =
// level >> 1, bits 15..23 = band, bits 0..14 = tile-local pixel; padding has bits 24..31 all set (level
// 0xFF) and carries the lane's current band.  Steps are stored in groups of four, lane-interleaved:
// word(t, lane) at ((t / 4) * 32 + lane) * 4 + t % 4.
constexpr int LANE_GROUP = 4;
constexpr uint32_t LANE_PAD_LEVEL = 0xFFu;

// Bank-group assignment inside one quarter-warp: lane i may take bank group r (pixel row mod 8) if it
// still has a candidate pixel in that group (byte r of cnt[i] > 0).  Maximum bipartite matching by
// augmenting paths (8 x 8), each lane trying its best-stocked groups first so that the groups it keeps
// for later steps stay diverse.  owner[r] = lane of the quarter that takes group r, or -1.
__device__ bool lanes_augment(int i, const unsigned long long* cnt, int* owner, unsigned& visited) {
  unsigned cand = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if ((cnt[i] >> (8 * r)) & 0xFFull) cand |= 1u << r;
  cand &= ~visited;
  while (cand) {
    int best = 0, bc = -1;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int cr = (int)((cnt[i] >> (8 * r)) & 0xFFull);
      if (((cand >> r) & 1u) && cr > bc) { bc = cr; best = r; }
    }
    cand &= ~(1u << best);
    visited |= 1u << best;
    if (owner[best] < 0 || lanes_augment(owner[best], cnt, owner, visited)) {
      owner[best] = i;
      return true;
    }
  }
  return false;
}

template <int G>
__global__ void obs_lanes_kernel(int32_t* __restrict__ idx, uint8_t* __restrict__ lvl,
                                 const int64_t* __restrict__ row_off, int64_t n_streams, int K, int IJ, int n_sub,
                                 int sub_pixels, int tile_warps, const int64_t* __restrict__ stream_off,
                                 uint32_t* __restrict__ words, int32_t* __restrict__ nrows,
                                 int32_t* __restrict__ overflow, int n_runs, int word16, int lvl_bits) {
  extern __shared__ __align__(16) unsigned char lanes_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= n_streams) return;
  // per warp: hist[32][G] u64 | band_of_rank[K16] i16 | lane_bands[32][G] i16 | assign[32] i16
  const int K16 = (K + 15) & ~15;
  const size_t per_warp = (size_t)32 * G * 8 + (size_t)K16 * 2 + (size_t)32 * G * 2 + 64;
  unsigned char* wbase = lanes_smem + wib * per_warp;
  unsigned long long* sh_hist = reinterpret_cast<unsigned long long*>(wbase);
  int16_t* band_of_rank = reinterpret_cast<int16_t*>(wbase + (size_t)32 * G * 8);
  int16_t* sh_bands = band_of_rank + K16;
  int16_t* sh_assign = sh_bands + 32 * G;
  const int64_t row0 = s * K;
  const int64_t beg = row_off[row0];
  const int st = (int)(s % n_sub);
  const int TP = tile_warps * sub_pixels;
  const int p0 = (st / tile_warps) * TP;  // first pixel of the tile this sub-tile belongs to
  const uint32_t own0 = (uint32_t)((st % tile_warps) * sub_pixels);  // tile-local: first pixel of this sub-tile (idle padding points here)
  // rank the bands by size (descending, ties by band index)
  for (int k = lane; k < K; k += 32) {
    const int64_t ck = row_off[row0 + k + 1] - row_off[row0 + k];
    int rank = 0;
    for (int j = 0; j < K; ++j) {
      const int64_t cj = row_off[row0 + j + 1] - row_off[row0 + j];
      rank += (cj > ck) || (cj == ck && j < k);
    }
    band_of_rank[rank] = (int16_t)k;
  }
  __syncwarp();
  int bands[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int r = (g & 1) ? 32 * g + 31 - lane : 32 * g + lane;
    bands[g] = r < K ? band_of_rank[r] : -1;
  }
  // Regroup the 32 band sets into the four quarter-warps so that, phase by phase (g-th band of every
  // lane), each quarter's supply of pixels per shared-memory bank group is as even as possible: a
  // quarter-warp can avoid bank conflicts only while all eight groups are still in stock.
#pragma unroll
  for (int g = 0; g < G; ++g) {
    unsigned long long h = 0;
    if (bands[g] >= 0) {
      const int c0 = (int)(row_off[row0 + bands[g]] - beg), e0 = (int)(row_off[row0 + bands[g] + 1] - beg);
      for (int qq = c0; qq < e0; ++qq) {
        const int r = (idx[beg + qq] - bands[g] * IJ - p0) & 7;
        if (((h >> (8 * r)) & 0xFFull) < 255) h += 1ull << (8 * r);
      }
    }
    sh_hist[lane * G + g] = h;
  }
  __syncwarp();
  {
    const int myq = lane >> 3, myr = lane & 7;  // this lane keeps the running supply of (quarter, bank group)
    int supply[G];
#pragma unroll
    for (int g = 0; g < G; ++g) supply[g] = 0;
    int fill = 0;  // lanes already placed in this lane's quarter
    for (int i = 0; i < 32; ++i) {
      int add[G];
      int cost = 0;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        add[g] = (int)((sh_hist[i * G + g] >> (8 * myr)) & 0xFFull);
        int mx = supply[g] + add[g], mn = mx;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        cost += mx - mn;
      }
      if (fill >= 8) cost = 0x3fffffff;
      int best = 0, bc = 0x7fffffff;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ck = __shfl_sync(0xffffffffu, cost, 8 * k);
        if (ck < bc) { bc = ck; best = k; }
      }
      const int slot = __shfl_sync(0xffffffffu, fill, 8 * best);
      if (myq == best) {
#pragma unroll
        for (int g = 0; g < G; ++g) supply[g] += add[g];
        ++fill;
      }
      if (lane == 0) sh_assign[i] = (int16_t)(8 * best + slot);
    }
  }
  __syncwarp();
  {
    const int dst = sh_assign[lane];
#pragma unroll
    for (int g = 0; g < G; ++g) sh_bands[dst * G + g] = (int16_t)bands[g];
  }
  __syncwarp();
  int left = 0;  // entries this lane still has to emit
#pragma unroll
  for (int g = 0; g < G; ++g) {
    bands[g] = sh_bands[lane * G + g];
    if (bands[g] >= 0) left += (int)(row_off[row0 + bands[g] + 1] - row_off[row0 + bands[g]]);
  }
  int g = -1, band = K, cur = 0, end = 0;  // band K: the dummy band of a lane that owns nothing
  unsigned long long cnt = 0;              // candidates left per bank group, for band cnt_band
  int cnt_band = -1;
  // stream: run table (n_runs entries per lane, [entry][lane]) followed by the words in 512-byte slots
  // (32 lanes x 16 bytes: four 32-bit words = one group, or eight 16-bit words = two groups)
  uint32_t* const table = words + stream_off[s];
  const int64_t out0 = stream_off[s] + (int64_t)n_runs * 32;
  const int cap = (int)((stream_off[s + 1] - out0) >> 7) * (word16 ? 8 : 4);  // steps
  uint16_t* const words16 = reinterpret_cast<uint16_t*>(words + out0);
  auto put_word = [&](int t, uint32_t lv, bool pad, uint32_t pix) {
    const int gi = t >> 2;
    if (word16) {
      const uint32_t lvf = pad ? ((1u << lvl_bits) - 1u) : lv;  // padding: all level bits set (the one-bit epilogue relies on it)
      words16[((int64_t)(gi >> 1) * 32 + lane) * 8 + (gi & 1) * 4 + (t & 3)] =
          (uint16_t)((lvf << (16 - lvl_bits)) | ((pad ? 1u : 0u) << (15 - lvl_bits)) | pix);
    } else {
      const uint32_t lvf = pad ? LANE_PAD_LEVEL : lv;
      words[out0 + ((int64_t)gi * 32 + lane) * LANE_GROUP + (t & 3)] = ((lvf & 1u) << 31) | ((lvf >> 1) << 24) | pix;
    }
  };
  int nrun = 0, run_band = -1, run_beg = 0;  // run table of this lane
  for (int i = 0; i < n_runs; ++i) table[i * 32 + lane] = 0u;
  int step = 0;
  while (__any_sync(0xffffffffu, left > 0)) {
    // move on to the lane's next non-empty band -- only at a group boundary, so that the gather
    // kernel sees one band per lane and group
    if ((step & (LANE_GROUP - 1)) == 0) {
      while (cur >= end && g < G) {
        ++g;
        int nb = -1;
#pragma unroll
        for (int gg = 0; gg < G; ++gg)
          if (gg == g) nb = bands[gg];
        if (g < G && nb >= 0) {
          const int c0 = (int)(row_off[row0 + nb] - beg), e0 = (int)(row_off[row0 + nb + 1] - beg);
          if (e0 > c0) { band = nb; cur = c0; end = e0; }
        }
      }
      if (band != run_band) {  // a new run: close the previous entry, open the next
        if (nrun > 0 && nrun <= n_runs)
          table[(nrun - 1) * 32 + lane] = (uint32_t)run_band | ((uint32_t)run_band << 9) | ((uint32_t)((step >> 2) - run_beg) << 18);
        run_band = band;
        run_beg = step >> 2;
        ++nrun;
      }
    }
    const bool has = cur < end;
    // per-bank-group counts of the lane's remaining candidates (8 x 8 bits, saturating at 255; a
    // saturated field is recounted after the next take)
    if (has && cnt_band != band) {
      cnt = 0;
      for (int qq = cur; qq < end; ++qq) {
        const int r = (idx[beg + qq] - band * IJ - p0) & 7;
        if (((cnt >> (8 * r)) & 0xFFull) < 255) cnt += 1ull << (8 * r);
      }
      cnt_band = band;
    }
    // matching of the quarter-warp's lanes to bank groups (every lane of the quarter computes the same)
    unsigned long long qc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qc[i] = __shfl_sync(0xffffffffu, has ? cnt : 0ull, (lane & 24) + i);
    int owner[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) owner[r] = -1;
    for (int i = 0; i < 8; ++i) {
      unsigned visited = 0;
      if (qc[i]) lanes_augment(i, qc, owner, visited);
    }
    int want = -1;  // the bank group this lane should use in this step
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (owner[r] == (lane & 7)) want = r;
    // pick a pixel: pass 0 walks the candidates of group `want`, pass 1 any candidate; a candidate is
    // taken if no other lane of the warp holds the same pixel (hard)
    bool active = has, won = false;
    int pass = want >= 0 ? 0 : 1, q = cur - 1, id = 0, pix = -(lane + 2);
    auto advance = [&]() {  // next candidate of the current pass, or give up
      while (active) {
        ++q;
        if (q >= end) {
          if (pass == 1) { active = false; pix = -(lane + 2); break; }
          pass = 1;
          q = cur - 1;
          continue;
        }
        id = idx[beg + q];
        pix = id - band * IJ - p0;
        if (pass == 1 || (pix & 7) == want) break;
      }
    };
    advance();
    for (int it = 0; it < 4 * 96; ++it) {
      const unsigned wonmask = __ballot_sync(0xffffffffu, won);
      const unsigned und = __ballot_sync(0xffffffffu, active && !won);
      if (!und) break;
      const unsigned m = __match_any_sync(0xffffffffu, pix);
      if (active && !won) {
        if ((m & wonmask) == 0 && lane == __ffs(m & und) - 1) won = true;
        else advance();
      }
    }
    if (!won) pix = -(lane + 2);
    // padding re-reads the S row of a real lane of the same quarter-warp (same address: no extra
    // shared-memory wavefront); its updates are predicated off in the kernel
    const unsigned wonmask = __ballot_sync(0xffffffffu, won);
    const unsigned wonq = wonmask & (0xffu << (lane & 24));
    const int srcl = wonq ? __ffs(wonq) - 1 : (wonmask ? __ffs(wonmask) - 1 : lane);
    const int padpix = __shfl_sync(0xffffffffu, pix, srcl);
    int lvw = 0;
    if (won) {
      const int lv = lvl[beg + q];
      if (q != cur) {  // move the chosen entry to the front of what is left of the row
        idx[beg + q] = idx[beg + cur];
        lvl[beg + q] = lvl[beg + cur];
        idx[beg + cur] = id;
        lvl[beg + cur] = (uint8_t)lv;
      }
      ++cur;
      --left;
      if (((cnt >> (8 * (pix & 7))) & 0xFFull) == 255) cnt_band = -1;  // saturated: recount
      else cnt -= 1ull << (8 * (pix & 7));
      lvw = lv;
    }
    if (step < cap) put_word(step, (uint32_t)lvw, !won, won ? (uint32_t)pix : (wonmask ? (uint32_t)padpix : own0));
    ++step;
    if (step > cap + 4096) break;  // hopeless: report and stop
  }
  // pad the last group; idle padding points at the first pixel of the stream's own sub-tile (rows the warp itself staged)
  while (step & (LANE_GROUP - 1)) {
    if (step < cap) put_word(step, 0u, true, own0);
    ++step;
  }
  // the last run never ends (a lane that has run out keeps walking padding words of its last band); a lane
  // that never opened a run walks the dummy band K
  if (nrun == 0) { run_band = K; nrun = 1; }
  if (nrun <= n_runs) table[(nrun - 1) * 32 + lane] = (uint32_t)run_band | ((uint32_t)run_band << 9) | (0x3FFFu << 18);
  const bool bad_runs = __any_sync(0xffffffffu, nrun > n_runs);
  if (lane == 0) {
    nrows[s] = step <= cap ? step : cap;
    if (step > cap || bad_runs) atomicExch(overflow, 1);
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

extern "C" int qmc_obs_build_lanes(int32_t* idx_rows_dev, uint8_t* lvl_rows_dev, const int64_t* row_off_dev,
                                   int B, int K, int IJ, int n_sub, int sub_pixels, int tile_warps,
                                   const int64_t* stream_off_dev, uint32_t* words_out_dev,
                                   int32_t* nrows_out_dev, int32_t* overflow_dev, int n_runs, int word_bits,
                                   int lvl_bits, void* stream) {
  QMC_REQUIRE(idx_rows_dev && lvl_rows_dev && row_off_dev && stream_off_dev && words_out_dev && nrows_out_dev &&
              overflow_dev, "null argument");
  QMC_REQUIRE(B > 0 && K > 0 && K <= 256 && IJ > 0 && n_sub > 0, "bad sizes (K must be <= 256)");
  QMC_REQUIRE(tile_warps > 0 && n_sub % tile_warps == 0 && sub_pixels > 0, "bad tiling");
  QMC_REQUIRE((int64_t)tile_warps * sub_pixels + 32 <= 32768, "tile of %lld pixels does not fit the 15-bit pixel field",
              (long long)tile_warps * sub_pixels);
  QMC_REQUIRE(word_bits == 16 || word_bits == 32, "word_bits must be 16 or 32");
  QMC_REQUIRE(word_bits == 32 || (lvl_bits >= 1 && lvl_bits <= 8 && (int64_t)tile_warps * sub_pixels <= (1LL << (15 - lvl_bits))),
              "16-bit words cannot hold %d level bits and a tile of %lld pixels", lvl_bits, (long long)tile_warps * sub_pixels);
  QMC_REQUIRE(n_runs >= (K + 31) / 32 && n_runs <= 64, "n_runs %d: need at least ceil(K/32) run-table entries per lane", n_runs);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_streams = (int64_t)B * n_sub;
  const int warps = 4;
  const int64_t blocks = (n_streams + warps - 1) / warps;
  QMC_REQUIRE(blocks <= 0x7fffffff, "too many streams");
  QMC_CUDA_CHECK(cudaMemsetAsync(overflow_dev, 0, sizeof(int32_t), st));
  const int G = (K + 31) / 32;
  const int Gt = G > 8 ? 8 : G;
  const size_t smem = (size_t)warps * ((size_t)32 * Gt * 8 + (size_t)((K + 15) & ~15) * 2 + (size_t)32 * Gt * 2 + 64);
#define QMC_LANES_GO(GG)                                                                                       \
  obs_lanes_kernel<GG><<<(unsigned)blocks, warps * 32, smem, st>>>(idx_rows_dev, lvl_rows_dev, row_off_dev,     \
                                                                   n_streams, K, IJ, n_sub, sub_pixels, tile_warps, \
                                                                   stream_off_dev, words_out_dev, nrows_out_dev,   \
                                                                   overflow_dev, n_runs, word_bits == 16, lvl_bits)
  switch (G) {
    case 1: QMC_LANES_GO(1); break;
    case 2: QMC_LANES_GO(2); break;
    case 3: QMC_LANES_GO(3); break;
    case 4: QMC_LANES_GO(4); break;
    case 5: QMC_LANES_GO(5); break;
    case 6: QMC_LANES_GO(6); break;
    case 7: QMC_LANES_GO(7); break;
    default: QMC_LANES_GO(8); break;
  }
#undef QMC_LANES_GO
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
