// Lane-stream builder: re-cuts the row-ordered compact observations of every (map, pixel sub-tile) stream
// into per-lane band walks for gather_lanes_kernel (layout: include/qmc_b200.h, qmc_obs_view_t).
//
// One warp per stream, the whole stream staged in shared memory.
//   plan   : the bands are laid end to end in units of 4-step groups (band k takes ceil(n_k / 4) groups), in an
//            order that makes most quota boundaries fall between two bands (see below), and the
//            sequence is cut into 32 quotas of equal length (+-1 group); a band that straddles a cut is split --
//            at a multiple of four entries, so the split costs no padding -- and the piece that continues it is
//            the first run of the next lane (it gets a gC row of its own, K+1+lane).  Entries are handed to the
//            pieces of a band from the back, so the rounding slack of a band sits in the piece at the END of a
//            quota: lanes have their slack where the candidate pools are smallest.
//   steps  : step by step, every lane takes one entry of its current piece.  Hard rule: the 32 entries of a step
//            have pairwise distinct pixels (the kernel updates gS rows without atomics).  Soft rule: inside a
//            quarter-warp the pixels fall into different shared-memory bank groups (pixel mod 8): a greedy
//            matching lanes x bank groups per step (most constrained lane first: lanes without slack, then the
//            fewest free groups; each takes its best-stocked free group); a lane the matching leaves out skips the
//            step if it has slack and takes any pixel otherwise.  Pixel collisions are won by the lane with the
//            least slack.
//   repair : a piece at the end of its planned groups with an entry left over first tries to put it into one of its
//            padding slots, directly or by moving one or two of its other entries along.
//            Otherwise a piece is walked until its entries are placed (run lengths are written when a run ends): a lane that
//            falls behind its plan -- dense sampling of a small sub-tile, where all lanes want the same few pixels
//            -- just takes longer, and the stream is as long as its slowest lane.  Slack is measured against the
//            planned end of the stream.
// A stream that outgrows its capacity is reported through *overflow and the caller retries with more room.
#include "qmc_gather_common.cuh"

namespace qmc {

constexpr uint32_t E_PLACED = 0x80000000u;   // entry word: bits 0..14 pixel, 15..22 level
constexpr uint32_t O_REAL = 0x80000000u;     // placed word: bits 0..14 pixel, 15..22 level; 0 = padding
constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total) {
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  total = __shfl_sync(FULL, inc, 31);
  return inc - v;
}

struct LaneBuildParams {
  const int32_t* idx;
  const uint8_t* lvl;
  const int64_t* row_off;
  int64_t n_streams;
  int K, IJ, n_sub, sub_pixels, tile_warps;
  int64_t stream_stride;  // 32-bit words
  uint32_t* words;
  int32_t* nrows;
  int32_t* overflow;
  int n_runs, word16, lvl_bits;
  int e_cap, t_cap, extra_groups, split_bands;
};

__global__ void __launch_bounds__(32) obs_lanes_quota_kernel(const LaneBuildParams prm) {
  extern __shared__ __align__(16) uint32_t bsm[];
  const int lane = threadIdx.x;
  const int64_t s = blockIdx.x;
  const int K = prm.K, IJ = prm.IJ;
  uint32_t* ent = bsm;                               // [e_cap]
  uint32_t* out = ent + prm.e_cap;                   // [t_cap][32]
  const int G = (K + 31) >> 5, NS = 32 * G;          // bands per lane set, sequence positions
  const int UW = (prm.sub_pixels + 31) >> 5;
  uint32_t* used = out + (size_t)prm.t_cap * 32;     // [t_cap][UW] pixels (relative to the sub-tile) taken per step
  uint32_t* boff = used + (size_t)prm.t_cap * UW;    // [K+1] first entry of every band
  uint32_t* gsz = boff + (K + 1);                    // [K+1] groups of every band (entry K: the empty band)
  uint32_t* seq = gsz + (K + 1);                     // [NS] band at every position of the end-to-end sequence (K = none)
  uint32_t* gpre = seq + NS;                         // [NS+1] first group of every sequence position

  const int64_t row0 = s * K;
  const int64_t beg = prm.row_off[row0];
  const int st = (int)(s % prm.n_sub);
  const int TP = prm.tile_warps * prm.sub_pixels;
  const int p0 = (st / prm.tile_warps) * TP;                              // first pixel of the tile
  const uint32_t own0 = (uint32_t)((st % prm.tile_warps) * prm.sub_pixels);  // tile-local first pixel of this sub-tile
  uint32_t* const table = prm.words + s * prm.stream_stride;
  uint4* const slots = reinterpret_cast<uint4*>(table + (size_t)prm.n_runs * 32);
  const int n_slots_cap = (int)((prm.stream_stride - (int64_t)prm.n_runs * 32) >> 7);
  const int steps_cap = min(n_slots_cap * (prm.word16 ? 8 : 4), prm.t_cap);

  // ---- band sizes -> entry and group prefixes ------------------------------------------------------
  uint32_t ebase = 0, gbase = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    const uint32_t nk = k < K ? (uint32_t)(prm.row_off[row0 + k + 1] - prm.row_off[row0 + k]) : 0u;
    const uint32_t gk = (nk + 3u) >> 2;
    uint32_t te, tg;
    const uint32_t xe = warp_excl_scan(nk, lane, te), xg = warp_excl_scan(gk, lane, tg);
    if (k < K) { boff[k] = ebase + xe; gsz[k] = gk; }
    ebase += te;
    gbase += tg;
  }
  if (lane == 0) { boff[K] = ebase; gsz[K] = 0u; }
  __syncwarp();
  // ---- order of the bands in the end-to-end sequence -------------------------------------------------------
  // The bands are dealt to 32 sets in snake order of their sizes (largest ... 32nd | 64th ... 33rd | ...), so the
  // sets are about equally long, and the sets are lined up so that the running total stays as close as possible
  // to a multiple of the quota: most lanes then walk exactly one set -- their run changes fall into the same few
  // groups of the stream (start, middle, end), which is what the gather kernel's run switch, executed by the
  // whole warp whenever any lane switches, wants -- and only the few groups by which a set is longer or shorter
  // than the quota move to a neighbour.
  {
    uint32_t* bor = gpre;  // band of rank, temporary
    for (int k = lane; k < K; k += 32) {
      const uint32_t gk = gsz[k];
      int rank = 0;
      for (int j = 0; j < K; ++j) {
        const uint32_t gj = gsz[j];
        rank += (gj > gk) || (gj == gk && j < k);
      }
      bor[rank] = (uint32_t)k;
    }
    __syncwarp();
    uint32_t load = 0;
    for (int g = 0; g < G; ++g) {
      const int r = (g & 1) ? 32 * g + 31 - lane : 32 * g + lane;
      if (r < K) load += gsz[bor[r]];
    }
    bool taken = false;
    int my_slot = 0;
    uint32_t posg = 0;
    for (int i = 0; i < 32; ++i) {
      const long long d = 32ll * (long long)(posg + load) - (long long)(i + 1) * (long long)gbase;
      const unsigned long long ad = (unsigned long long)(d < 0 ? -d : d);
      const unsigned key = taken ? 0xffffffffu : (((unsigned)(ad > 0x3ffffffull ? 0x3ffffffull : ad) << 5) | (unsigned)lane);
      const unsigned best = __reduce_min_sync(FULL, key);
      const int bl = (int)(best & 31u);
      if (lane == bl) { taken = true; my_slot = i; }
      posg += __shfl_sync(FULL, load, bl);
    }
    uint32_t mine[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int r = (g & 1) ? 32 * g + 31 - lane : 32 * g + lane;
      mine[g] = (g < G && r < K) ? bor[r] : (uint32_t)K;
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 8; ++g)
      if (g < G) seq[my_slot * G + g] = mine[g];
    __syncwarp();
    uint32_t gb = 0;
    for (int i0 = 0; i0 < NS; i0 += 32) {
      const uint32_t gk = gsz[seq[i0 + lane]];
      uint32_t tg;
      const uint32_t xg = warp_excl_scan(gk, lane, tg);
      gpre[i0 + lane] = gb + xg;
      gb += tg;
    }
    if (lane == 0) gpre[NS] = gb;
    __syncwarp();
  }
  const uint32_t n_total = ebase, M = gbase;
  // quotas: equal shares of the sequence (bands split where a quota ends), or -- split_bands == 0 -- every lane
  // walks exactly its own set of whole bands and the stream is as long as the longest set
  uint32_t myq0, qa0;
  if (prm.split_bands) {
    const uint32_t qbase = M >> 5, qrem = M & 31u;
    myq0 = qbase + ((((lane + 1) * qrem) >> 5) != ((lane * qrem) >> 5) ? 1u : 0u);
    uint32_t tq;
    qa0 = warp_excl_scan(myq0, lane, tq);
  } else {
    qa0 = gpre[lane * G];
    myq0 = gpre[(lane + 1) * G] - qa0;
  }
  uint32_t qmax = myq0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qmax = max(qmax, __shfl_xor_sync(FULL, qmax, o));
  const int len0 = M ? (int)qmax : 0;   // planned groups of the stream (extra_groups is room, not plan)
  if (n_total > (uint32_t)prm.e_cap || 4 * len0 > steps_cap) {          // does not fit: report, leave an empty stream
    if (lane == 0) {
      atomicOr(prm.overflow, 1);
      prm.nrows[s] = 0;
    }
    for (int i = 1; i < prm.n_runs; ++i) table[i * 32 + lane] = 0u;
    table[lane] = (uint32_t)K | ((uint32_t)K << LW_RUN_BAND_SHIFT) | (0x3FFFu << LW_RUN_LEN_SHIFT);
    return;
  }
  for (uint32_t e = lane; e < n_total; e += 32) {
    const int id = prm.idx[beg + e];
    ent[e] = (uint32_t)(id % IJ - p0) | ((uint32_t)prm.lvl[beg + e] << 15);
  }
  const int len = len0;                       // planned groups; the stream runs longer if lanes fall behind
  const int T_plan = 4 * len;
  for (int i = 0; i < prm.n_runs; ++i) table[i * 32 + lane] = 0u;
  for (int i = lane; i < prm.t_cap * 32; i += 32) out[i] = 0u;
  for (int i = lane; i < prm.t_cap * UW; i += 32) used[i] = 0u;
  __syncwarp();

  // ---- this lane's pieces: groups [qa, qb) of the sequence -----------------------------------------------
  const uint32_t myq = myq0, qa = qa0, qb = qa + myq;
  int kb = 0;                                   // sequence position the next piece is looked for at
  {
    int lo = 0, hi = NS;  // first sequence position with gpre[k + 1] > qa
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (gpre[mid + 1] > qa) hi = mid; else lo = mid + 1;
    }
    kb = lo;
  }
  uint32_t pos = qa;     // next group of the sequence this lane has to cover
  int spare = T_plan - 4 * (int)myq;  // slots between the planned end of this lane's pieces and the planned end of the stream
  int nslot = 0;                      // planned slots left in the current piece
  int nrun = 0;
  bool bad_runs = false;

  // current piece: walked until its entries are placed (a piece that falls behind its plan just takes longer:
  // run lengths are written when the run ends)
  int band = -1, ne = 0, run_beg = 0;
  uint32_t e0 = 0, e1 = 0, run_word = 0;
  unsigned long long cnt = 0;
  bool recount = false, finished = myq == 0;

  int t = 0;
  for (;; ++t) {
    if ((t & 3) == 0) {
      // A piece at the end of its planned groups with entries left over (a step in which all its candidates'
      // pixels were taken) would cost its lane, and with it the stream, a whole group: first try to put the entry
      // into a padding slot of the piece -- directly (a), by moving the entry of another step there (b), or
      // through a chain of two moves (c).
      {
        unsigned fm = __ballot_sync(FULL, band >= 0 && ne > 0 && nslot <= 0 && spare < 4);
        while (fm) {
          const int fl = __ffs(fm) - 1;
          fm &= fm - 1;
          const int fb = __shfl_sync(FULL, band, fl), a0 = 4 * __shfl_sync(FULL, run_beg, fl), a1 = t;
          const int fc = __shfl_sync(FULL, ne, fl);
          const uint32_t be0 = boff[fb], be1 = boff[fb + 1];
          int fixed = 0;
          for (int c = 0; c < fc; ++c) {
            int found = -1;  // an entry of the band nobody has placed
            for (uint32_t base = be0; base < be1 && found < 0; base += 32) {
              const uint32_t e = base + lane;
              const unsigned bm = __ballot_sync(FULL, e < be1 && !(ent[e] & E_PLACED));
              if (bm) found = (int)base + __ffs(bm) - 1;
            }
            if (found < 0) break;
            const uint32_t w = ent[found];
            const uint32_t rel = (w & LW_PIX_MASK) - own0;
            bool placed = false;
            for (int base = a0; base < a1 && !placed; base += 32) {   // (a)
              const int u = base + lane;
              const bool ok = u < a1 && !(out[u * 32 + fl] & O_REAL) && !((used[u * UW + (rel >> 5)] >> (rel & 31u)) & 1u);
              const unsigned bm = __ballot_sync(FULL, ok);
              if (bm) {
                const int uu = base + __ffs(bm) - 1;
                if (lane == 0) {
                  out[uu * 32 + fl] = O_REAL | (w & 0x7FFFFFu);
                  used[uu * UW + (rel >> 5)] |= 1u << (rel & 31u);
                  ent[found] = w | E_PLACED;
                }
                placed = true;
              }
            }
            for (int base = a0; base < a1 && !placed; base += 32) {   // (b), (c)
              const int v = base + lane;
              int v2 = -1, u2 = -1;  // v2 < 0: nothing; v2 == v: direct move v -> u2; else chain v -> v2 -> u2
              if (v < a1 && (out[v * 32 + fl] & O_REAL) && !((used[v * UW + (rel >> 5)] >> (rel & 31u)) & 1u)) {
                const uint32_t relb = (out[v * 32 + fl] & LW_PIX_MASK) - own0;
                for (int u = a0; u < a1; ++u)
                  if (!(out[u * 32 + fl] & O_REAL) && !((used[u * UW + (relb >> 5)] >> (relb & 31u)) & 1u)) { v2 = v; u2 = u; break; }
                for (int x = a0; x < a1 && v2 < 0; ++x) {
                  if (x == v || !(out[x * 32 + fl] & O_REAL) || ((used[x * UW + (relb >> 5)] >> (relb & 31u)) & 1u)) continue;
                  const uint32_t relc = (out[x * 32 + fl] & LW_PIX_MASK) - own0;
                  for (int u = a0; u < a1; ++u)
                    if (!(out[u * 32 + fl] & O_REAL) && !((used[u * UW + (relc >> 5)] >> (relc & 31u)) & 1u)) { v2 = x; u2 = u; break; }
                }
              }
              const unsigned bm = __ballot_sync(FULL, v2 >= 0);
              if (bm) {
                if (lane == __ffs(bm) - 1) {
                  const uint32_t wb = out[v * 32 + fl];
                  const uint32_t relb = (wb & LW_PIX_MASK) - own0;
                  if (v2 != v) {  // the entry of step v2 moves to the padding slot first
                    const uint32_t wc = out[v2 * 32 + fl];
                    const uint32_t relc = (wc & LW_PIX_MASK) - own0;
                    out[u2 * 32 + fl] = wc;
                    used[u2 * UW + (relc >> 5)] |= 1u << (relc & 31u);
                    used[v2 * UW + (relc >> 5)] &= ~(1u << (relc & 31u));
                    u2 = v2;
                  }
                  out[u2 * 32 + fl] = wb;
                  used[u2 * UW + (relb >> 5)] |= 1u << (relb & 31u);
                  used[v * UW + (relb >> 5)] &= ~(1u << (relb & 31u));
                  out[v * 32 + fl] = O_REAL | (w & 0x7FFFFFu);
                  used[v * UW + (rel >> 5)] |= 1u << (rel & 31u);
                  ent[found] = w | E_PLACED;
                }
                placed = true;
              }
            }
            __syncwarp();
            if (!placed) break;
            ++fixed;
          }
          if (lane == fl) { ne -= fixed; recount = true; }
        }
      }
      if (band >= 0 && ne == 0) {   // piece done: close its run; groups it did not need go to the spare
        if (nrun <= prm.n_runs) table[(nrun - 1) * 32 + lane] = run_word | ((uint32_t)((t >> 2) - run_beg) << LW_RUN_LEN_SHIFT);
        spare += max(nslot, 0);
        band = -1;
      } else if (band >= 0 && nslot <= 0) {   // behind plan: borrow a group
        nslot += 4;
        spare -= 4;
      }
      if (band < 0 && !finished) {
        while (kb < NS && gpre[kb + 1] <= pos) ++kb;
        if (pos >= qb || kb >= NS) {
          finished = true;
        } else {
          const uint32_t g0 = gpre[kb], g1 = gpre[kb + 1];
          const uint32_t pe = min(qb, g1);
          const uint32_t ra = pos - g0, rb = pe - g0;
          band = (int)seq[kb];
          e0 = boff[band];
          e1 = boff[band + 1];
          const int rem = (int)(4u * (g1 - g0) - (e1 - e0));   // rounding slack of the band: in its first piece
          const int hi = 4 * (int)rb - rem, lo = ra > 0 ? 4 * (int)ra - rem : 0;
          ne = max(hi - max(lo, 0), 0);
          nslot = 4 * (int)(pe - pos);
          run_beg = t >> 2;
          run_word = (uint32_t)(ra > 0 ? K + 1 + lane : band) | ((uint32_t)band << LW_RUN_BAND_SHIFT);
          recount = true;
          pos = pe;
          if (nrun >= prm.n_runs) bad_runs = true;
          ++nrun;
          if (ne == 0) {   // cannot happen for a well-formed plan; be safe
            if (nrun <= prm.n_runs) table[(nrun - 1) * 32 + lane] = run_word;
            band = -1;
          }
        }
      }
      // done when every lane has walked all its pieces (checked at group boundaries only: streams are whole groups)
      if (!__any_sync(FULL, band >= 0 || !finished)) break;
      if (t + 4 > steps_cap) break;   // out of room: reported below
    }
    const bool has = band >= 0 && ne > 0;
    // pieces of one band active in several lanes share the pool: their counts go stale with every take
    {
      const unsigned mb = __match_any_sync(FULL, has ? band : (K + 1 + lane));
      if (__popc(mb) > 1) recount = true;
    }
    if (has && recount) {
      cnt = 0;
      for (uint32_t q = e0; q < e1; ++q) {
        const uint32_t w = ent[q];
        if (w & E_PLACED) continue;
        const int r = (int)(w & 7u);
        if (((cnt >> (8 * r)) & 0xFFull) < 255) cnt += 1ull << (8 * r);
      }
      recount = false;
    }
    // slack: steps this lane can idle without running past the planned end of the stream -- what its piece has
    // beyond its entries (a skip that pushes a piece over a group boundary costs a whole group) plus whole spare groups
    const int slack = has ? max(nslot - ne, 0) + max(spare, 0) : 0x7fff;
    // Bank groups for the quarter-warp's lanes: a greedy matching run by the eight lanes together, all in
    // registers.  Eight rounds; in each the most constrained unassigned lane -- lanes that must place an entry
    // in every remaining step first, then the fewest free bank groups among its candidates -- takes its
    // best-stocked free group (so that the groups it keeps for later steps stay diverse).
    int want = -1;  // the bank group this lane should use in this step
    {
      unsigned avail = 0;
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (has && ((cnt >> (8 * r)) & 0xFFull)) avail |= 1u << r;
      unsigned freec = 0xFFu;
      bool assigned = avail == 0;
#pragma unroll 1
      for (int round = 0; round < 8; ++round) {
        const unsigned a = avail & freec;
        const int key = (assigned || a == 0) ? 0x7fffffff : (((slack == 0 ? 0 : 1) << 8) | (__popc(a) << 4) | (lane & 7));
        int mn = key;
        mn = min(mn, __shfl_xor_sync(FULL, mn, 4));
        mn = min(mn, __shfl_xor_sync(FULL, mn, 2));
        mn = min(mn, __shfl_xor_sync(FULL, mn, 1));
        int cls = 0;
        if (mn == key && key != 0x7fffffff) {
          int bc = -1;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int cr = (int)((cnt >> (8 * r)) & 0xFFull);
            if (((a >> r) & 1u) && cr > bc) { bc = cr; cls = r; }
          }
          want = cls;
          assigned = true;
        }
        cls = __shfl_sync(FULL, cls, (lane & 24) + (mn & 7));
        if (mn != 0x7fffffff) freec &= ~(1u << cls);
      }
    }
    // pick a pixel: pass 0 walks the candidates of group `want`; pass 1 -- lanes without slack only -- any candidate;
    // a candidate is taken if no other lane of the warp holds the same pixel in this step (hard)
    bool active = has && (want >= 0 || slack == 0), won = false;
    int pass = want >= 0 ? 0 : 1;
    uint32_t q = e0 - 1u, w = 0;
    int pix = -(lane + 2);
    auto advance = [&]() {  // next candidate of the current pass, or give up
      while (active) {
        ++q;
        if (q >= e1) {
          if (pass == 1 || slack > 0) { active = false; pix = -(lane + 2); break; }
          pass = 1;
          q = e0 - 1u;
          continue;
        }
        w = ent[q];
        if (w & E_PLACED) continue;
        pix = (int)(w & LW_PIX_MASK);
        if (pass == 1 || (pix & 7) == want) break;
      }
    };
    advance();
    const int prio = (min(slack, 1023) << 5) | lane;
    for (int it = 0; it < 8 * 1024; ++it) {
      const unsigned wonmask = __ballot_sync(FULL, won);
      const unsigned und = __ballot_sync(FULL, active && !won);
      if (!und) break;
      const unsigned m = __match_any_sync(FULL, pix);
      if (active && !won) {
        const unsigned cm = m & und;  // undecided lanes proposing this pixel
        const int best = __reduce_min_sync(cm, prio);
        if ((m & wonmask) == 0 && best == prio) won = true;
        else advance();
      }
    }
    if (won) {
      ent[q] = w | E_PLACED;
      out[t * 32 + lane] = O_REAL | (w & 0x7FFFFFu);
      const uint32_t rel = (uint32_t)pix - own0;
      atomicOr(&used[t * UW + (rel >> 5)], 1u << (rel & 31u));
      const int r = pix & 7;
      if (((cnt >> (8 * r)) & 0xFFull) == 255) recount = true;  // saturated: recount
      else cnt -= 1ull << (8 * r);
      --ne;
    }
    if (band >= 0) --nslot;
    __syncwarp();
  }
  const int T = t;   // a multiple of four
  const bool incomplete = __any_sync(FULL, band >= 0 || !finished);
  // a lane that has run out keeps walking padding words of its last run; a lane that never had one walks the
  // dummy band K
  if (nrun == 0) table[lane] = (uint32_t)K | ((uint32_t)K << LW_RUN_BAND_SHIFT) | (0x3FFFu << LW_RUN_LEN_SHIFT);
  else if (nrun <= prm.n_runs) table[(nrun - 1) * 32 + lane] = run_word | (0x3FFFu << LW_RUN_LEN_SHIFT);
  const bool bad_any = __any_sync(FULL, bad_runs);
  if (lane == 0) {
    if (incomplete) atomicOr(prm.overflow, 1);
    if (bad_any) atomicOr(prm.overflow, 2);
    prm.nrows[s] = incomplete ? 0 : T;
  }
  __syncwarp();

  // ---- write the words: one 16-byte store per lane and slot ------------------------------------------------
  // padding re-reads the S row of a real lane of the same quarter-warp (same address: no extra shared-memory
  // wavefront); its updates are predicated off in the kernel
  const int spp = prm.word16 ? 8 : 4;  // steps per slot
  const uint32_t lvmask = (1u << prm.lvl_bits) - 1u;
  for (int sl = 0; sl * spp < T; ++sl) {
    uint32_t hw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hw[j] = 0;
      const int t = sl * spp + j;
      if (j < spp && t < T) {   // warp-uniform
        const uint32_t w = out[t * 32 + lane];
        const bool real = (w & O_REAL) != 0;
        const unsigned rm = __ballot_sync(FULL, real);
        const unsigned rq = rm & (0xffu << (lane & 24));
        const int srcl = rq ? __ffs(rq) - 1 : (rm ? __ffs(rm) - 1 : lane);
        const uint32_t ppix = __shfl_sync(FULL, w & LW_PIX_MASK, srcl);
        const uint32_t pix = real ? (w & LW_PIX_MASK) : (rm ? ppix : own0);
        const uint32_t lv = (w >> 15) & 0xFFu;
        if (prm.word16)
          hw[j] = ((real ? lv : lvmask) << (16 - prm.lvl_bits)) | ((real ? 0u : 1u) << (15 - prm.lvl_bits)) | pix;
        else
          hw[j] = real ? (((lv & 1u) << 31) | ((lv >> 1) << 24) | pix) : (0xFF000000u | pix);
      } else if (j < spp) {
        hw[j] = prm.word16 ? ((lvmask << (16 - prm.lvl_bits)) | (1u << (15 - prm.lvl_bits)) | own0) : (0xFF000000u | own0);
      }
    }
    uint4 v;
    if (prm.word16) v = make_uint4(hw[0] | (hw[1] << 16), hw[2] | (hw[3] << 16), hw[4] | (hw[5] << 16), hw[6] | (hw[7] << 16));
    else v = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    slots[(size_t)sl * 32 + lane] = v;
  }
}

}  // namespace qmc

using namespace qmc;

extern "C" int64_t qmc_lanes_stream_words(int64_t max_entries_per_stream, int K, int n_runs, int word_bits, int extra_groups) {
  if (max_entries_per_stream < 0 || K <= 0 || n_runs <= 0 || (word_bits != 16 && word_bits != 32) || extra_groups < 0) return 0;
  // groups of a stream <= ceil((entries + 3 per non-empty band) / 4 / 32) + extra; at least four groups of room
  const int64_t nb = max_entries_per_stream < K ? max_entries_per_stream : K;
  int64_t groups = ((max_entries_per_stream + 3 * nb + 3) / 4 + 31) / 32 + extra_groups + 2;  // + room for lanes that fall behind
  if (groups < 4) groups = 4;
  const int64_t slots = word_bits == 16 ? (groups + 1) / 2 : groups;
  return (int64_t)n_runs * 32 + slots * 128;
}

extern "C" int qmc_obs_build_lanes(const int32_t* idx_rows_dev, const uint8_t* lvl_rows_dev, const int64_t* row_off_dev,
                                   int B, int K, int IJ, int n_sub, int sub_pixels, int tile_warps,
                                   int64_t max_entries_per_stream, int extra_groups, int split_bands, int64_t stream_stride,
                                   uint32_t* words_out_dev, int32_t* nrows_out_dev, int32_t* overflow_dev, int n_runs,
                                   int word_bits, int lvl_bits, void* stream) {
  QMC_REQUIRE(idx_rows_dev && lvl_rows_dev && row_off_dev && words_out_dev && nrows_out_dev && overflow_dev, "null argument");
  QMC_REQUIRE(B > 0 && K > 0 && K <= 256 && IJ > 0 && n_sub > 0, "bad sizes (K must be <= 256)");
  QMC_REQUIRE(tile_warps > 0 && n_sub % tile_warps == 0 && sub_pixels > 0, "bad tiling");
  QMC_REQUIRE((int64_t)tile_warps * sub_pixels + 32 <= 32768, "tile of %lld pixels does not fit the 15-bit pixel field",
              (long long)tile_warps * sub_pixels);
  QMC_REQUIRE(word_bits == 16 || word_bits == 32, "word_bits must be 16 or 32");
  QMC_REQUIRE(word_bits == 32 || (lvl_bits >= 1 && lvl_bits <= 8 && (int64_t)tile_warps * sub_pixels <= (1LL << (15 - lvl_bits))),
              "16-bit words cannot hold %d level bits and a tile of %lld pixels", lvl_bits, (long long)tile_warps * sub_pixels);
  QMC_REQUIRE(n_runs >= 2 && n_runs <= 64, "n_runs %d out of range [2, 64]", n_runs);
  QMC_REQUIRE(max_entries_per_stream >= 0 && extra_groups >= 0 && extra_groups < 4096, "bad capacity arguments");
  QMC_REQUIRE(stream_stride >= qmc_lanes_stream_words(max_entries_per_stream, K, n_runs, word_bits, extra_groups) &&
              stream_stride % 32 == 0, "stream_stride %lld is too small (see qmc_lanes_stream_words)", (long long)stream_stride);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_streams = (int64_t)B * n_sub;
  QMC_REQUIRE(n_streams <= 0x7fffffff, "too many streams");
  QMC_CUDA_CHECK(cudaMemsetAsync(overflow_dev, 0, sizeof(int32_t), st));
  LaneBuildParams p;
  p.idx = idx_rows_dev; p.lvl = lvl_rows_dev; p.row_off = row_off_dev; p.n_streams = n_streams;
  p.K = K; p.IJ = IJ; p.n_sub = n_sub; p.sub_pixels = sub_pixels; p.tile_warps = tile_warps;
  p.stream_stride = stream_stride; p.words = words_out_dev; p.nrows = nrows_out_dev; p.overflow = overflow_dev;
  p.n_runs = n_runs; p.word16 = word_bits == 16; p.lvl_bits = word_bits == 16 ? lvl_bits : 8;
  p.e_cap = (int)(max_entries_per_stream > 0 ? max_entries_per_stream : 1);
  const int64_t slots = (stream_stride - (int64_t)n_runs * 32) >> 7;
  p.t_cap = (int)(slots * (word_bits == 16 ? 8 : 4));
  p.extra_groups = extra_groups;
  p.split_bands = split_bands != 0;
  const size_t smem = ((size_t)p.e_cap + (size_t)p.t_cap * 32 + (size_t)p.t_cap * ((sub_pixels + 31) / 32) + 2 * (size_t)(K + 1) +
                       2 * (size_t)(32 * ((K + 31) / 32)) + 1) * 4;
  QMC_REQUIRE(smem <= 227 * 1024, "a stream of %lld entries does not fit the builder's shared memory (%zu bytes)",
              (long long)max_entries_per_stream, smem);
  QMC_CUDA_CHECK(cudaFuncSetAttribute(obs_lanes_quota_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  obs_lanes_quota_kernel<<<(unsigned)n_streams, 32, smem, st>>>(p);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}
