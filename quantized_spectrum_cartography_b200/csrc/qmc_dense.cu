// Dense-sampling path (cfg4: one large instance, ~50 % of the entries observed): the reconstruction
// X = S*C^T and both gradient contractions run on the 5th-generation tensor cores (tcgen05, fp32
// accumulators in tensor memory), with the quantized likelihood as the epilogue between them, so
// neither X nor g = dNLL/dX ever exists in HBM.  Replaces the same reference idiom as the gather
// kernels (qmc/quantization_model.py:22-39,57-61,70-86; qmc/qmc.ipynb c1:145-153).
//
// Per CTA (16 epilogue warps + 2 warps that only issue tensor-core instructions, persistent over tiles of up to 128 pixels):
//   MMA1  D1[128 px x K bands]  = S_tile * C^T          kind::tf32, 3xTF32 split (hi*hi + hi*lo + lo*hi: three
//         descriptor walks over the same [Sh|Sl] and [Ch|Cl] operand tiles)
//   epilogue (16 warps): tcgen05.ld an 8-column slab of D1 per thread, read the 1-byte codes of the
//         same entries, evaluate log P and g, store g (hi/lo TF32 parts) into shared memory twice:
//         G[128 px x 32 bands] with bands contiguous and G^T[32 bands x 128 px] with pixels contiguous,
//         both as K-major UMMA operands (MN-major TF32 operands in the no-swizzle layout read as zeros
//         on this part, measured; hence the explicit transposed copy)
//   MMA2  D2[128 px x 16]      += G * C                 (gS tile)
//   MMA3  D3[32(64) bands x 16] += G^T * S_tile          (gC block; M = 64 instruction, rows 32..63 unused)
// D2 is written to gS after the last band block of a tile, D3 accumulates over all tiles of the CTA
// and is added to gC once at the end.
// A tcgen05.mma of these shapes costs ~65 cycles whatever its size (measured: tools/dense_probe.py), so the 48
// instructions of MMA3 are a 1.6 us chain per band block: the G^T operand is double-buffered and MMA3 of block b runs
// under the likelihood arithmetic of block b+1 (MMA2, 12 instructions, keeps a single G buffer).
//
// Observation format: one byte per dense entry, pixel-major code8[IJ][K], 255 = not observed
// (qmc_dense_pack builds it from the reference's Y / Wx).
#include "qmc_common.cuh"

namespace qmc {

constexpr int DT_PIX = 128;    // pixels per tile = TMEM lanes
constexpr int DT_BLK = 32;     // bands per G block
constexpr int DT_RP = 16;      // padded rank of the gradient MMAs (N)
constexpr int DT_THREADS = 512;   // 16 epilogue warps: 4 per TMEM lane quadrant, each takes an 8-band slab of a block
constexpr int DT_LAUNCH = DT_THREADS + 64;   // + two warps that only issue tensor-core instructions (MMA1 + MMA2, MMA3)
constexpr int DT_SLAB = DT_BLK / (DT_THREADS / 128);
constexpr uint32_t TMEM_COLS = 512;
#ifndef QMC_DENSE_NCH
#define QMC_DENSE_NCH 4
#endif
#ifndef QMC_DENSE_NCH_ALT
#define QMC_DENSE_NCH_ALT 2
#endif
constexpr uint32_t COL_D2 = 256, COL_D3 = 272;

struct DenseParams {
  const float* S;        // [R][IJ]
  const float* C;        // [R][K]
  const uint8_t* code;   // [IJ][K]
  double* nll;
  float* gS;             // [R][IJ], row stride gs_stride
  int64_t gs_stride;
  float* gC;             // [R][K]
  float* gC_acc;         // where the CTAs add their partial gC and NLL: the outputs themselves (zeroed by the host), or,
  double* nll_acc;       // with the fused exchange, the self-cleaning scratch slot of this rank's exchange region
  int IJ, K, R, Rp8, n_tiles;
  int tile_pix;          // pixels a tile really holds (<= DT_PIX; the rest of the 128 TMEM lanes idles), chosen by the host so
                         // that the tiles divide evenly over the persistent CTAs
  int n_bounds, one_sided;
  int debug;             // measurement hook (QMC_DENSE_DEBUG): 1 skip MMA3, 2 skip MMA2, 4 skip the likelihood, 8 skip the G stores
  float inv_a, offset, thr;
  // fused factor-gradient exchange of a sharded instance (px_world == 0: none); see peer_exchange()
  int px_rank, px_world, px_slot_floats;
  uint8_t* px_region[QMC_PEER_MAX_WORLD];
  float bounds[QMC_MAX_BOUNDS];
};

// ---- exchange region (one per rank, device memory mapped into every peer through CUDA IPC) -------------------------
//   [0, 256)  header: epoch (exchanges completed), status (0 = fine, 1 = a peer never arrived), done (CTA counter)
//             and, from byte 128, flag[q] = the last epoch rank q has delivered into this region
//   then a scratch slot the CTAs of a launch add their partials into (the last CTA reads it and leaves it zeroed for
//   the next launch: no memset nodes around the kernel), then
//   two sets (epoch parity) of `world` slots of slot_floats floats: slot q = rank q's partial [gC | nll as a double]
// A rank can run at most one exchange ahead of a peer (it needs the peer's flag of epoch e to finish e, and the peer
// raises it only after it has read everything of epoch e-1), hence two slot sets and a monotonic flag are enough.
struct PeerHeader {
  uint32_t epoch, status, done, pad[29];
  uint32_t flag[32];
};
static_assert(sizeof(PeerHeader) == 256, "PeerHeader is the first 256 bytes of a region");
__host__ __device__ inline size_t peer_slot_offset(int world, int slot_floats, int parity, int q) {
  return sizeof(PeerHeader) + (1 + (size_t)parity * world + q) * (size_t)slot_floats * 4;   // slot 0 is the scratch
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Run by the last CTA of the grid to finish (every CTA's gC / NLL atomics are complete and visible): push this rank's
// partial [gC | nll] into slot `rank` of every rank's region with plain stores over NVLink, raise the flags, wait for
// the peers' flags in the own region and add the slots up in rank order (every rank gets the same bits).
__device__ void peer_exchange(const DenseParams& prm, int nthreads) {
  const int tid = threadIdx.x, world = prm.px_world, rank = prm.px_rank;
  const int n = prm.R * prm.K, n4 = n >> 2;            // K is a multiple of 32: n % 4 == 0
  PeerHeader* own = reinterpret_cast<PeerHeader*>(prm.px_region[rank]);
  const uint32_t e = own->epoch + 1;                   // written only by this code path, one kernel at a time
  const int par = (int)(e & 1u);
  // this rank's partial sums: out of the scratch slot into registers, and the scratch slot back to zero
  constexpr int PV = 4;                                // float4 per thread: R*K <= 16*256 floats = 1024 float4 <= 4 * 512
  float4* part = reinterpret_cast<float4*>(prm.gC_acc);
  float4 pv[PV];
#pragma unroll
  for (int u = 0; u < PV; ++u) {
    const int i = tid + u * nthreads;
    if (i < n4) {
      pv[u] = __ldcg(part + i);
      __stcg(part + i, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  double nll_part = 0.0;
  if (tid == 0) {
    nll_part = __ldcg(prm.nll_acc);
    __stcg(prm.nll_acc, 0.0);
  }
  for (int q0 = 0; q0 < world; ++q0) {
    const int q = (rank + 1 + q0) % world;             // start with the neighbour: the ranks do not all hit rank 0 first
    float4* dst = reinterpret_cast<float4*>(prm.px_region[q] + peer_slot_offset(world, prm.px_slot_floats, par, rank));
#pragma unroll
    for (int u = 0; u < PV; ++u) {
      const int i = tid + u * nthreads;
      if (i < n4) dst[i] = pv[u];
    }
    if (tid == 0) *reinterpret_cast<double*>(dst + n4) = nll_part;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int px_failed;
  if (tid == 0) px_failed = 0;
  __syncthreads();
  if (tid < world) {
    st_release_sys(&reinterpret_cast<PeerHeader*>(prm.px_region[tid])->flag[rank], e);
    const uint64_t t0 = globaltimer_ns();
    // flags are monotonic; the comparison is wrap-safe
    while ((int32_t)(ld_acquire_sys(&own->flag[tid]) - e) < 0) {
      if (globaltimer_ns() - t0 > 2000000000ull) { px_failed = 1; break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (px_failed) {   // a peer never arrived: flag it and hand back this rank's own partial sums
#pragma unroll
    for (int u = 0; u < PV; ++u) {
      const int i = tid + u * nthreads;
      if (i < n4) reinterpret_cast<float4*>(prm.gC)[i] = pv[u];
    }
    if (tid == 0) { *prm.nll = nll_part; own->status = 1; own->done = 0; own->epoch = e; }
    return;
  }
  const uint8_t* slots = prm.px_region[rank] + peer_slot_offset(world, prm.px_slot_floats, par, 0);
  const size_t pitch = (size_t)prm.px_slot_floats * 4;
  for (int i = tid; i < n4; i += nthreads) {
    // all the slots' loads in flight at once (a load from a region the peers write is an L2 round trip), then the sum
    // in rank order
    float4 v[QMC_PEER_MAX_WORLD];
#pragma unroll
    for (int q = 0; q < QMC_PEER_MAX_WORLD; ++q)
      if (q < world) v[q] = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(slots + q * pitch) + i);
    float4 a = v[0];
#pragma unroll
    for (int q = 1; q < QMC_PEER_MAX_WORLD; ++q)
      if (q < world) { a.x += v[q].x; a.y += v[q].y; a.z += v[q].z; a.w += v[q].w; }
    reinterpret_cast<float4*>(prm.gC)[i] = a;
  }
  if (tid == 0) {
    double tot = 0.0;
    for (int q = 0; q < world; ++q) {
      const unsigned long long* p = reinterpret_cast<const unsigned long long*>(slots + q * pitch + (size_t)n * 4);
      unsigned long long bits;
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(bits) : "l"(p) : "memory");
      tot += __longlong_as_double((long long)bits);
    }
    *prm.nll = tot;
    own->done = 0;
    own->epoch = e;
  }
}

enum : int { DEPI_STABLE = 0, DEPI_REFERENCE = 1, DEPI_ONEBIT = 2, DEPI_LSQ = 3, DEPI_LOGISTIC = 4 };

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // shared-memory matrix descriptor, no swizzle (interleaved 8x16-byte core matrices), sm_100 version 1
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N, bool a_mn_major) {
  return (1u << 4) /*D = f32*/ | (2u << 7) /*A = tf32*/ | (2u << 10) /*B = tf32*/ |
         ((a_mn_major ? 1u : 0u) << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Both are executed by the WHOLE issuing warp, converged; one elected lane issues the instruction.  The operands are
// then warp-uniform values the compiler keeps in uniform registers: an MMA costs two or three uniform integer
// instructions plus UTCHMMA.  (Issued from inside an `if (lane == 0)` branch instead, every operand went through an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop: ~12 dependent instructions, ~115 cycles per MMA -- the issuing thread,
// not the tensor pipe, bounded the kernel.)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void dmbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void dmbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "DW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DD_%=;\n\t"
      "bra DW_%=;\n\t"
      "DD_%=:\n\t}" ::"r"(s_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// The same load in two halves: issue (the registers are written asynchronously) and wait (the registers pass through
// the wait as read-write operands, so nothing that uses them can be scheduled above it)
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_wait(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :: "memory");
}

// TF32 split: hi keeps the 10 explicit mantissa bits the tensor core reads, lo is the exact rest
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// bnd: the boundary table in shared memory (per-lane levels: a shared-memory gather instead of a divergent
// constant-bank index)
template <int EPI, bool LOGD>
__device__ __forceinline__ BinEval dense_eval(const DenseParams& prm, const float* bnd, float t, int lvl, float& dxdt) {
  float x = t;
  dxdt = 1.0f;
  if (LOGD) {
    // SFU grade like the rest of the epilogue: lg2.approx is good to 2^-22 absolute in log2 units, far below what the
    // bin widths (>= 0.1 in log units) can see; a non-positive argument gives NaN / -inf exactly as logf does
    const float u = t + prm.offset;
    x = (EPI == DEPI_REFERENCE) ? logf(u) : kLn2 * lg2_approx(u);
    dxdt = (EPI == DEPI_REFERENCE) ? 1.0f / u : rcp_approx(u);
  }
  if (EPI == DEPI_LSQ) {
    // masked least squares on the bin mid-point (quantization_model_log.py:43-51, qmc_dowjons.ipynb c1:112):
    // "logp" = -(x - mid)^2, so that nll -= logp accumulates the squared residual
    const float d = x - 0.5f * (bnd[lvl] + bnd[lvl + 1]);
    BinEval o;
    o.logp = -d * d;
    o.gx = 2.0f * d;
    return o;
  }
  if (EPI == DEPI_LOGISTIC) {  // F_sigmoid in place of F_probit (quantization_model.py:43-47)
    if (prm.one_sided) return logistic_one_sided_fast(prm.thr, lvl ? prm.inv_a : -prm.inv_a, x);
    return logistic_bin(bnd[lvl], bnd[lvl + 1], x, prm.inv_a);
  }
  if (EPI == DEPI_ONEBIT) return probit_one_sided_fast(prm.thr, lvl ? -prm.inv_a : prm.inv_a, x);
  if (EPI == DEPI_REFERENCE) return probit_bin_reference(bnd[lvl], bnd[lvl + 1], x, prm.inv_a);
  return probit_bin_stable<true>(bnd[lvl], bnd[lvl + 1], x, prm.inv_a);
}

// Shared memory map (bytes).  All operand regions are 128-byte aligned.
//   A1  : S tile split [Sh|Sl], K-major [chunk j][pixel][16 B], chunks = 2*Rp8/4   128*16*chunks
//   B1  : C split [Ch|Cl],      K-major [chunk j][band ][16 B]                     K*16*chunks
//   B2h/B2l : C as [chunk of 4 bands][r-group][r%8][16 B]  (N = 16, K-major)       K/4*256 each
//   B3h/B3l : S tile as [chunk of 4 pixels][r-group][r%8][16 B]                    32*256 each
//   Gh/Gl   : g block  [chunk of 4 bands][pixel][16 B]                             8*2048 each
//   GTh/GTl : g block transposed [chunk of 4 pixels][band][16 B], chunk pitch 528 B (32 rows + 16 B of
//             padding so that the scalar stores of 32 consecutive pixels hit 32 different banks); two buffers
//             (even / odd band blocks), and 512 B of slack behind the last one because the M = 64 instruction
//             reads 64 rows per chunk
//   queue   : per warp 32 x DT_SLAB values (fp32) and as many codes (bytes): the observed entries of a warp's
//             slab, compacted
constexpr uint32_t GT_PITCH = 528;
constexpr uint32_t GT_BYTES = 32 * GT_PITCH;   // one G^T operand (hi or lo) of one band block
constexpr uint32_t QUEUE_WARP_BYTES = 32 * DT_SLAB * 5;
struct DenseSmem {
  uint32_t a1, b1, b2h, b2l, b3h, b3l, gh, gl, gt, queue, total;   // gt: [buffer 0: h, l][buffer 1: h, l][slack]
};
__host__ __device__ inline DenseSmem dense_smem_map(int K, int Rp8) {
  const uint32_t chunks = 2 * Rp8 / 4;
  DenseSmem m;
  uint32_t o = 0;
  m.a1 = o; o += DT_PIX * 16 * chunks;
  m.b1 = o; o += (uint32_t)K * 16 * chunks;
  m.b2h = o; o += (uint32_t)(K / 4) * 256;
  m.b2l = o; o += (uint32_t)(K / 4) * 256;
  m.b3h = o; o += 32 * 256;
  m.b3l = o; o += 32 * 256;
  m.gh = o; o += 8 * 2048;
  m.gl = o; o += 8 * 2048;
  m.gt = o; o += 4 * GT_BYTES + 512;
  m.queue = o; o += (DT_THREADS / 32) * QUEUE_WARP_BYTES;
  m.total = o;
  return m;
}

// NCH: independent likelihood evaluations per lane and pass of the compacted queue (instruction-level parallelism for
// the 18 warps of the CTA: 4 passes a slab's ~128 observed entries in one go)
template <int EPI, bool LOGD, bool GRAD, int NCH>
__global__ void __launch_bounds__(DT_LAUNCH, 1) dense_kernel(const DenseParams prm) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ uint64_t bar1, bar2, bar3[2], gfull;
  __shared__ uint32_t tmem_base_sh;
  __shared__ double wsum[DT_LAUNCH / 32];
  __shared__ float bnd[QMC_MAX_BOUNDS + 1];

  const int K = prm.K, R = prm.R, Rp8 = prm.Rp8;
  const int chunks1 = 2 * Rp8 / 4;
  const DenseSmem map = dense_smem_map(K, Rp8);
  // (the warp index through a shuffle: the compiler then knows it is the same in all lanes, and the issuing warp's
  //  branch is warp-uniform code that may use the uniform datapath)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;   // TMEM lane quadrant, which 8-band slab of a 32-band block
  // Warp specialisation: warps 0..15 evaluate the likelihood and write the G operands; warps 16 and 17 issue the
  // tcgen05.mma instructions (16: MMA1 and MMA2, 17: MMA3).  The epilogue warps hand a block's G over through the `gfull`
  // mbarrier and go straight on to the next block's arithmetic -- they never wait for each other (no CTA barrier
  // inside a tile) nor for the issue of ~60 MMA instructions; they only wait for the previous block's MMA2 (bar2)
  // before overwriting the G buffer and for MMA3 of the block before that (bar3) before overwriting its G^T buffer.
  // Two issuing warps: the 48 instructions of a block's MMA3 (~15 issue-side instructions each) do not hold up the 12
  // of the next MMA2, and a tcgen05.commit tracks the committing thread's own MMAs, so bar2 and bar3 separate cleanly.
  const bool issuer = warp >= DT_THREADS / 32, issuer2 = warp == DT_THREADS / 32;
  const int row = quad * 32 + lane;              // pixel row of this thread inside the tile
  const uint32_t sbase = s_u32(dsm);

  // ---- one-time setup ---------------------------------------------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base_sh)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    dmbar_init(&bar1, 1);
    dmbar_init(&bar2, 1);
    dmbar_init(&bar3[0], 1);
    dmbar_init(&bar3[1], 1);
    dmbar_init(&gfull, DT_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < prm.n_bounds; i += DT_THREADS) bnd[i] = prm.bounds[i];
  // B1 (C split for MMA1) and B2h/B2l (C for MMA2): built once, C is the same for every tile
  for (int k = tid; k < K; k += DT_THREADS) {
    float ch[16], cl[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float c = (r < R) ? __ldg(prm.C + (size_t)r * K + k) : 0.0f;
      ch[r] = tf32_hi(c);
      cl[r] = c - ch[r];
    }
    // B1 row = band k, elements [Ch(0..Rp8) | Cl]
    for (int j = 0; j < chunks1; ++j) {
      const int e = 4 * j, seg = e / Rp8, r0 = e % Rp8;
      const float* src = (seg == 1) ? cl : ch;
      *reinterpret_cast<float4*>(dsm + map.b1 + (size_t)j * K * 16 + k * 16) = make_float4(src[r0], src[r0 + 1], src[r0 + 2], src[r0 + 3]);
    }
    // B2: element (n = r, kk = band k) at chunk k/4, group r/8, row r%8, word k%4
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const uint32_t o = (k >> 2) * 256 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4;
      *reinterpret_cast<float*>(dsm + map.b2h + o) = ch[r];
      *reinterpret_cast<float*>(dsm + map.b2l + o) = cl[r];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_sh;
  const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);

  const uint32_t idesc1 = umma_idesc_tf32(128, K, false);
  const uint32_t idesc2 = umma_idesc_tf32(128, DT_RP, false);
  const uint32_t idesc3 = umma_idesc_tf32(64, DT_RP, false);

  float nll_part = 0.0f;   // per tile (<= 256 terms) in fp32, folded into nll_acc in fp64 after every tile
  double nll_acc = 0.0;
  uint32_t ph1 = 0, ph2 = 0, ph3[2] = {0, 0}, phg = 0;
  bool d3_started = false;   // D3 blocks accumulate over all tiles of this CTA
  const int nblk = K / DT_BLK;
  const bool short_tile = prm.tile_pix < DT_PIX;

  // ---- staging of a tile's S rows: A1 (MMA1) and B3h/B3l (MMA3); threads 0..255: operand row p (= TMEM lane), 8 ranks
  // each.  A tile's rows are loaded and its A1 written while the previous tile's last MMAs drain (A1 is free as soon
  // as MMA1 has run); B3 follows once those MMAs are done.
  float st_h[8], st_l[8];
  auto stage_load_a1 = [&](int t) {
    if (tid < 256) {
      const int p = tid & 127, rh = tid >> 7;
      const int ppix = short_tile ? ((p & 31) << 2 | (p >> 5)) : p;
      const int pg = t * prm.tile_pix + ppix;
      const bool in = ppix < prm.tile_pix && pg < prm.IJ;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rh * 8 + i;
        const float sv = (in && r < R) ? __ldg(prm.S + (size_t)r * prm.IJ + pg) : 0.0f;
        st_h[i] = tf32_hi(sv);
        st_l[i] = sv - st_h[i];
      }
      if (rh * 8 < Rp8) {   // A1 row = pixel p, elements [Sh | Sl]
        const int per = Rp8 / 4;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int jr = rh * 2 + q;   // chunk index inside one Rp8-wide segment
          *reinterpret_cast<float4*>(dsm + map.a1 + (size_t)(jr) * 2048 + p * 16) =
              make_float4(st_h[4 * q], st_h[4 * q + 1], st_h[4 * q + 2], st_h[4 * q + 3]);
          *reinterpret_cast<float4*>(dsm + map.a1 + (size_t)(per + jr) * 2048 + p * 16) =
              make_float4(st_l[4 * q], st_l[4 * q + 1], st_l[4 * q + 2], st_l[4 * q + 3]);
        }
      }
    }
  };
  auto stage_b3 = [&]() {
    if (GRAD && tid < 256) {
      const int p = tid & 127, rh = tid >> 7;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rh * 8 + i;
        const uint32_t o = (p >> 2) * 256 + (r >> 3) * 128 + (r & 7) * 16 + (p & 3) * 4;
        *reinterpret_cast<float*>(dsm + map.b3h + o) = st_h[i];
        *reinterpret_cast<float*>(dsm + map.b3l + o) = st_l[i];
      }
    }
  };
  // code bytes of a tile's first block for this thread's row (fetched with the staging, a tile ahead)
  auto first_codes = [&](int t) -> uint2 {
    const int pixl = short_tile ? ((row & 31) << 2 | (row >> 5)) : row;
    const int pg = t * prm.tile_pix + pixl;
    uint2 c = make_uint2(0xffffffffu, 0xffffffffu);
    if (!issuer && t < prm.n_tiles && pixl < prm.tile_pix && pg < prm.IJ)
      c = __ldg(reinterpret_cast<const uint2*>(prm.code + (size_t)pg * K + half * DT_SLAB));
    return c;
  };
  uint2 cfirst = first_codes(blockIdx.x);
  if (blockIdx.x < prm.n_tiles) {
    stage_load_a1(blockIdx.x);
    stage_b3();
  }
  fence_async();
  tc_fence_before();
  __syncthreads();

  for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
    const int p0 = tile * prm.tile_pix;
    // the code bytes of a block are fetched one block ahead (the first block's here, before the tile is staged):
    // their DRAM latency hides behind the staging, MMA1 and, later, a block's arithmetic
    // A short tile spreads its pixels evenly over the four TMEM lane quadrants (pixel = 4 * lane + quadrant): every
    // warp then compacts the same number of observed entries
    const int pix = short_tile ? ((row & 31) << 2 | (row >> 5)) : row;
    const bool inside = pix < prm.tile_pix && p0 + pix < prm.IJ;
    const uint8_t* crow = prm.code + (size_t)(p0 + pix) * K;
    uint2 cnext = cfirst;
    // ---- MMA1: D1 = Sh*Ch^T + Sh*Cl^T + Sl*Ch^T ---------------------------------------------------
    if (issuer2) {
      tc_fence_after();
      const int per = Rp8 / 4;   // 16-byte chunks per segment
      for (int term = 0; term < 3; ++term) {
        const int aseg = term == 2 ? per : 0, bseg = term == 1 ? per : 0;
        for (int kk = 0; kk < Rp8 / 8; ++kk) {
          const uint64_t ad = umma_desc(sbase + map.a1 + (aseg + kk * 2) * 2048, 2048, 128);
          const uint64_t bd = umma_desc(sbase + map.b1 + (bseg + kk * 2) * K * 16, K * 16, 128);
          umma_tf32(tmem, ad, bd, idesc1, (term | kk) != 0);
        }
      }
      umma_commit(&bar1);
    }
    if (!issuer) {
      dmbar_wait(&bar1, ph1);
      tc_fence_after();
    }
    ph1 ^= 1;

    // ---- epilogue, one 32-band block at a time ------------------------------------------------------
    // a warp's slab of D1 is loaded one block ahead of its use: the TMEM round trip hides behind a block's arithmetic
    uint32_t xnext[DT_SLAB];
    if (!issuer) tmem_ld8_issue(tlane + (uint32_t)(half * DT_SLAB), xnext);
    for (int blk = 0; blk < nblk; ++blk) {
      if (issuer) {
        if (GRAD) {
          dmbar_wait(&gfull, phg);   // all sixteen epilogue warps have written this block's G operands
          tc_fence_after();
          // MMA2: D2[128 x 16] += G_blk * C_blk   (K = 32 bands of this block, 4 steps of 8).  A descriptor's start
          // address field counts 16-byte units: stepping through an operand is an add on the descriptor.
          if (issuer2) {
          if (!(prm.debug & 2)) {
#pragma unroll
            for (int term = 0; term < 3; ++term) {
              const uint64_t ad0 = umma_desc(sbase + ((term == 2) ? map.gl : map.gh), 2048, 128);
              const uint64_t bd0 = umma_desc(sbase + ((term == 1) ? map.b2l : map.b2h) + blk * (DT_BLK / 4) * 256, 256, 128);
#pragma unroll
              for (int ks = 0; ks < DT_BLK / 8; ++ks)
                umma_tf32(tmem + COL_D2, ad0 + (uint64_t)(ks * 2 * 2048 / 16), bd0 + (uint64_t)(ks * 2 * 256 / 16), idesc2,
                          (blk | term | ks) != 0);
            }
          }
          umma_commit(&bar2);   // the G buffer is free again
          } else {
          // MMA3: D3_blk[bands x 16] += G_blk^T * S_tile   (K = 128 pixels, 16 steps of 8; M = 64
          // instruction whose rows 32..63 read the following chunk and are never looked at)
          const uint32_t gtb = map.gt + (uint32_t)(blk & 1) * 2 * GT_BYTES;
          if (!(prm.debug & 1)) {
#pragma unroll
            for (int term = 0; term < 3; ++term) {
              const uint64_t ad0 = umma_desc(sbase + ((term == 2) ? gtb + GT_BYTES : gtb), GT_PITCH, 128);
              const uint64_t bd0 = umma_desc(sbase + ((term == 1) ? map.b3l : map.b3h), 256, 128);
#pragma unroll
              for (int ks = 0; ks < DT_PIX / 8; ++ks)
                umma_tf32(tmem + COL_D3 + blk * DT_RP, ad0 + (uint64_t)(ks * 2 * GT_PITCH / 16), bd0 + (uint64_t)(ks * 2 * 256 / 16),
                          idesc3, d3_started || (term | ks) != 0);
            }
          }
          umma_commit(&bar3[blk & 1]);   // this block's G^T buffer is free again
          }
        }
        phg ^= 1;
        continue;
      }
      const int k0 = blk * DT_BLK + half * DT_SLAB;   // first band of this thread's 8-column slab
      float x[DT_SLAB];
      tmem_ld8_wait(xnext);
#pragma unroll
      for (int i = 0; i < DT_SLAB; ++i) x[i] = __uint_as_float(xnext[i]);
      if (blk + 1 < nblk) tmem_ld8_issue(tlane + (uint32_t)(k0 + DT_BLK), xnext);
      uint32_t cw[DT_SLAB / 4];
      cw[0] = cnext.x; cw[1] = cnext.y;
      if (inside && blk + 1 < nblk) cnext = __ldg(reinterpret_cast<const uint2*>(crow + k0 + DT_BLK));
      // Likelihood of the slab: x[] is overwritten by g = dNLL/dt (0 where nothing was observed).  Only part of
      // the entries is observed (half at cfg4), so the warp first compacts its observed (value, code) pairs into a
      // queue in shared memory and then evaluates 32 of them per round with all lanes busy, instead of walking its
      // DT_SLAB columns with the unobserved lanes idle; the results go back through the queue.
      {
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < DT_SLAB; ++i) cnt += ((cw[i >> 2] >> (8 * (i & 3))) & 0xffu) != 255u;
        // exclusive prefix sum of the counts (0..DT_SLAB = 8: four bits) out of four ballots -- a shuffle scan compiles
        // into five calls of an out-of-line collective routine here (~140 instructions per block)
        const uint32_t lt = (1u << lane) - 1u;
        const uint32_t b0 = __ballot_sync(0xffffffffu, cnt & 1), b1 = __ballot_sync(0xffffffffu, cnt & 2);
        const uint32_t b2 = __ballot_sync(0xffffffffu, cnt & 4), b3 = __ballot_sync(0xffffffffu, cnt & 8);
        const int base = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt) + 8 * __popc(b3 & lt);
        const int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2) + 8 * __popc(b3);
        float* qv = reinterpret_cast<float*>(dsm + map.queue + warp * QUEUE_WARP_BYTES);
        uint8_t* qc = reinterpret_cast<uint8_t*>(qv + 32 * DT_SLAB);
        {
          int w = base;
#pragma unroll
          for (int i = 0; i < DT_SLAB; ++i) {
            const uint32_t code = (cw[i >> 2] >> (8 * (i & 3))) & 0xffu;
            if (code != 255u) {
              qv[w] = x[i];
              qc[w++] = (uint8_t)code;
            }
          }
        }
        __syncwarp();
        for (int id0 = 0; id0 < ((prm.debug & 4) ? 0 : total); id0 += 32 * NCH) {   // warp-uniform; NCH entries per lane and pass
          const int rem = total - id0;
          if (rem <= 32) {                                  // at most 32 entries left: a single evaluation each
            if (lane < rem) {
              float d0;
              const BinEval v0 = dense_eval<EPI, LOGD>(prm, bnd, qv[id0 + lane], (int)qc[id0 + lane], d0);
              nll_part -= v0.logp;
              qv[id0 + lane] = v0.gx * d0;
            }
            break;
          }
          float ev[NCH];
          int ec[NCH];
          BinEval v[NCH];
          float d[NCH];
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const int idx = id0 + lane + ((c * 32 + lane < rem) ? c * 32 : 0);   // lane < 32 < rem
            ev[c] = qv[idx];
            ec[c] = (int)qc[idx];
          }
#pragma unroll
          for (int c = 0; c < NCH; ++c) v[c] = dense_eval<EPI, LOGD>(prm, bnd, ev[c], ec[c], d[c]);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (c * 32 + lane < rem) {
              nll_part -= v[c].logp;
              qv[id0 + lane + c * 32] = v[c].gx * d[c];
            }
          }
        }
        __syncwarp();
        int w = base;
#pragma unroll
        for (int i = 0; i < DT_SLAB; ++i) {
          const uint32_t code = (cw[i >> 2] >> (8 * (i & 3))) & 0xffu;
          float g = 0.0f;
          if (code != 255u) g = qv[w++];
          x[i] = g;
        }
      }
      if (GRAD) {
        if (blk > 0) {
          // the previous block's MMA2 must have consumed the G buffer before it is overwritten,
          dmbar_wait(&bar2, ph2);
          ph2 ^= 1;
        }
        if (blk > 1) {
          // and MMA3 of the block before that this block's G^T buffer
          dmbar_wait(&bar3[blk & 1], ph3[blk & 1]);
          ph3[blk & 1] ^= 1;
        }
        const uint32_t gth = map.gt + (uint32_t)(blk & 1) * 2 * GT_BYTES, gtl = gth + GT_BYTES;
#pragma unroll
        for (int c4 = 0; c4 < ((prm.debug & 8) ? 0 : DT_SLAB / 4); ++c4) {
          const float4 h4 = make_float4(tf32_hi(x[4 * c4]), tf32_hi(x[4 * c4 + 1]), tf32_hi(x[4 * c4 + 2]), tf32_hi(x[4 * c4 + 3]));
          const float4 l4 = make_float4(x[4 * c4] - h4.x, x[4 * c4 + 1] - h4.y, x[4 * c4 + 2] - h4.z, x[4 * c4 + 3] - h4.w);
          const uint32_t o = (uint32_t)(half * (DT_SLAB / 4) + c4) * 2048 + row * 16;
          *reinterpret_cast<float4*>(dsm + map.gh + o) = h4;
          *reinterpret_cast<float4*>(dsm + map.gl + o) = l4;
          // transposed copy: element (band, pixel) at chunk pixel/4, row band, word pixel%4
          const uint32_t ot = (uint32_t)(row >> 2) * GT_PITCH + (uint32_t)(half * DT_SLAB + 4 * c4) * 16 + (row & 3) * 4;
          *reinterpret_cast<float*>(dsm + gth + ot) = h4.x;
          *reinterpret_cast<float*>(dsm + gth + ot + 16) = h4.y;
          *reinterpret_cast<float*>(dsm + gth + ot + 32) = h4.z;
          *reinterpret_cast<float*>(dsm + gth + ot + 48) = h4.w;
          *reinterpret_cast<float*>(dsm + gtl + ot) = l4.x;
          *reinterpret_cast<float*>(dsm + gtl + ot + 16) = l4.y;
          *reinterpret_cast<float*>(dsm + gtl + ot + 32) = l4.z;
          *reinterpret_cast<float*>(dsm + gtl + ot + 48) = l4.w;
        }
        fence_async();       // generic-proxy writes of G -> visible to the tensor core's async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&gfull)) : "memory");
      }
    }
    d3_started = true;
    const int ntile = tile + gridDim.x;
    if (ntile < prm.n_tiles) {
      stage_load_a1(ntile);
      cfirst = first_codes(ntile);
    }
    if (GRAD && !issuer) {
      // ---- gS tile out of D2 ---------------------------------------------------------------------
      dmbar_wait(&bar2, ph2);   // the last block's MMA2: D2 is complete
      ph2 ^= 1;
      // every MMA3 still in flight (the last one or two blocks'): the next tile re-stages their S operand, and D3 is
      // read after the last tile
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i < nblk) {
          dmbar_wait(&bar3[i], ph3[i]);
          ph3[i] ^= 1;
        }
      }
      tc_fence_after();
      {   // the four warps of a lane quadrant take four ranks each
        float v[4];
        tmem_ld4(tlane + COL_D2 + (uint32_t)(half * 4), v);
        if (inside) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = half * 4 + i;
            if (r < R) prm.gS[(size_t)r * prm.gs_stride + p0 + pix] = v[i];
          }
        }
      }
    }
    nll_acc += (double)nll_part;
    nll_part = 0.0f;
    if (ntile < prm.n_tiles) stage_b3();   // (the epilogue warps have just waited for the last MMA3)
    // the next tile's operands are staged and all TMEM reads of this tile are done before its MMAs overwrite D1 / D2
    fence_async();
    tc_fence_before();
    __syncthreads();
  }

  // ---- gC out of D3, NLL -----------------------------------------------------------------------------
  if (GRAD && blockIdx.x < prm.n_tiles && !issuer) {
    tc_fence_after();
    if (half == 0 && quad < 2) {
      for (int blk = 0; blk < nblk; ++blk) {
        float v[16];
        // warp-wide load; M = 64 accumulator: rows 0..15 in lanes 0..15 of quadrant 0, rows 16..31 in
        // lanes 0..15 of quadrant 1 (rows 32..63, quadrants 2 and 3, are the unused half)
        tmem_ld16(tlane + COL_D3 + blk * DT_RP, v);
        const int band = blk * DT_BLK + quad * 16 + lane;
        if (lane < 16) {
#pragma unroll
          for (int r = 0; r < 16; ++r)
            if (r < R) atomicAdd(prm.gC_acc + (size_t)r * K + band, v[r]);
        }
      }
    }
  }
  double w = warp_sum(nll_acc);
  if (lane == 0) wsum[warp] = w;
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int i = 0; i < DT_LAUNCH / 32; ++i) tot += wsum[i];
    if (tot != 0.0) atomicAdd(prm.nll_acc, tot);
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
  }
  if (GRAD && prm.px_world > 0) {
    // the last CTA to get here owns the exchange (its peers' atomics are ordered before their counter increment)
    __shared__ int px_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      PeerHeader* own = reinterpret_cast<PeerHeader*>(prm.px_region[prm.px_rank]);
      px_last = atomicAdd(&own->done, 1u) == gridDim.x - 1;
      __threadfence();
    }
    __syncthreads();
    if (px_last) peer_exchange(prm, DT_LAUNCH);
  }
}

// code8[b][p][k] = Wx != 0 ? Y : 255, from the reference's band-major [K][IJ] arrays
template <typename YT>
__global__ void dense_pack_kernel(const YT* __restrict__ y, const float* __restrict__ wx, int K, int IJ,
                                  uint8_t* __restrict__ code) {
  __shared__ uint8_t t[32][33];
  const int p0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const size_t plane = (size_t)blockIdx.z * K * IJ;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, p = p0 + threadIdx.x;
    uint8_t c = 255;
    if (k < K && p < IJ) {
      const size_t o = plane + (size_t)k * IJ + p;
      if (!wx || wx[o] != 0.0f) c = (uint8_t)y[o];
    }
    t[i][threadIdx.x] = c;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, k = k0 + threadIdx.x;
    if (p < IJ && k < K) code[plane + (size_t)p * K + k] = t[threadIdx.x][i];
  }
}

template <int EPI>
static int dense_launch(const DenseParams& prm, bool logd, bool grad, int grid, size_t smem, cudaStream_t st) {
#define QMC_DENSE_GO(L, G)                                                                                   \
  do {                                                                                                       \
    auto kern = dense_kernel<EPI, L, G, QMC_DENSE_NCH>;                                                      \
    if (nch_alt) kern = dense_kernel<EPI, L, G, QMC_DENSE_NCH_ALT>;                                          \
    QMC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    kern<<<grid, DT_LAUNCH, smem, st>>>(prm);                                                               \
  } while (0)
  static const bool nch_alt = getenv("QMC_DENSE_NCH_ALT") != nullptr;   // measurement hook
  if (logd) { if (grad) QMC_DENSE_GO(true, true); else QMC_DENSE_GO(true, false); }
  else { if (grad) QMC_DENSE_GO(false, true); else QMC_DENSE_GO(false, false); }
#undef QMC_DENSE_GO
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

}  // namespace qmc

using namespace qmc;

extern "C" int qmc_dense_pack(const void* y_dev, int y_is_int64, const float* wx_dev, int B, int K, int IJ,
                              uint8_t* code_out_dev, void* stream) {
  QMC_REQUIRE(y_dev && code_out_dev, "null argument");
  QMC_REQUIRE(B > 0 && K > 0 && IJ > 0 && B <= 65535 && (K + 31) / 32 <= 65535, "bad sizes");
  dim3 grid((IJ + 31) / 32, (K + 31) / 32, B), block(32, 8);
  if (y_is_int64)
    dense_pack_kernel<int64_t><<<grid, block, 0, (cudaStream_t)stream>>>((const int64_t*)y_dev, wx_dev, K, IJ, code_out_dev);
  else
    dense_pack_kernel<uint8_t><<<grid, block, 0, (cudaStream_t)stream>>>((const uint8_t*)y_dev, wx_dev, K, IJ, code_out_dev);
  count_launch();
  QMC_CUDA_CHECK(cudaGetLastError());
  return QMC_OK;
}

extern "C" int64_t qmc_dense_smem_bytes(int K, int R) {
  if (K <= 0 || K > 256 || (K % DT_BLK) != 0 || R <= 0 || R > DT_RP) return 0;
  return (int64_t)dense_smem_map(K, R <= 8 ? 8 : 16).total;
}

static int dense_entry(const float* S_dev, const float* C_dev, const uint8_t* code_dev, const qmc_likelihood_t* lik,
                       int IJ, int K, int R, double* nll_out_dev, float* gS_out_dev, int64_t gs_row_stride,
                       float* gC_out_dev, const qmc_peer_exchange_t* px, void* stream) {
  QMC_REQUIRE(S_dev && C_dev && code_dev && lik && nll_out_dev, "null argument");
  QMC_REQUIRE(IJ > 0 && R > 0, "bad sizes");
  QMC_REQUIRE(gs_row_stride == 0 || gs_row_stride >= IJ, "gs_row_stride %lld is smaller than IJ", (long long)gs_row_stride);
  if (K <= 0 || K > 256 || (K % DT_BLK) != 0)
    return set_error(QMC_ERR_UNSUPPORTED, "dense path needs K a multiple of %d and <= 256 (got %d)", DT_BLK, K);
  if (R > DT_RP) return set_error(QMC_ERR_UNSUPPORTED, "dense path needs rank <= %d (got %d)", DT_RP, R);
  QMC_REQUIRE(lik->n_bounds >= 2 && lik->n_bounds <= QMC_MAX_BOUNDS - 1, "n_bounds %d out of range (255 is the 'unobserved' code)", lik->n_bounds);
  QMC_REQUIRE(lik->noise_std > 0.0f || (lik->flags & QMC_EPI_LSQ), "noise_std must be positive");
  const bool grad = !(lik->flags & QMC_FORWARD_ONLY);
  QMC_REQUIRE(!grad || (gS_out_dev && gC_out_dev), "gradient outputs are NULL without QMC_FORWARD_ONLY");
  cudaStream_t st = (cudaStream_t)stream;

  DenseParams prm;
  prm.S = S_dev; prm.C = C_dev; prm.code = code_dev; prm.nll = nll_out_dev; prm.gS = gS_out_dev; prm.gC = gC_out_dev;
  prm.gs_stride = gs_row_stride > 0 ? gs_row_stride : IJ;
  prm.IJ = IJ; prm.K = K; prm.R = R; prm.Rp8 = R <= 8 ? 8 : 16;
  prm.px_rank = 0; prm.px_world = 0; prm.px_slot_floats = 0;
  for (int q = 0; q < QMC_PEER_MAX_WORLD; ++q) prm.px_region[q] = nullptr;
  if (px) {
    QMC_REQUIRE(grad, "the fused exchange combines gradients: QMC_FORWARD_ONLY is not supported with it");
    QMC_REQUIRE(px->world >= 1 && px->world <= QMC_PEER_MAX_WORLD && px->rank >= 0 && px->rank < px->world,
                "bad rank %d / world %d", px->rank, px->world);
    QMC_REQUIRE(px->slot_floats >= R * K + 2 && (px->slot_floats % 4) == 0,
                "slot_floats %d: need a multiple of 4 that is >= R*K + 2 = %d", px->slot_floats, R * K + 2);
    for (int q = 0; q < px->world; ++q) {
      QMC_REQUIRE(px->region[q] != nullptr, "region[%d] is NULL", q);
      prm.px_region[q] = (uint8_t*)px->region[q];
    }
    prm.px_rank = px->rank; prm.px_world = px->world; prm.px_slot_floats = px->slot_floats;
  }
  prm.gC_acc = gC_out_dev;
  prm.nll_acc = nll_out_dev;
  if (px) {
    float* scratch = reinterpret_cast<float*>((uint8_t*)px->region[px->rank] + sizeof(PeerHeader));
    prm.gC_acc = scratch;
    prm.nll_acc = reinterpret_cast<double*>(scratch + R * K);   // R*K is a multiple of 32: 8-byte aligned
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  {
    // The work of a tile is its observed entries (the epilogue compacts them), not its 128 TMEM lanes: size the tiles
    // so that every persistent CTA gets the same number of them.  256 full tiles on 148 CTAs (cfg4 on one of eight
    // GPUs) take two rounds of 128 pixels; 296 tiles of 111 pixels take two rounds of 111.
    const int64_t full = ((int64_t)IJ + DT_PIX - 1) / DT_PIX;
    const int64_t rounds = (full + sms - 1) / sms;
    int64_t tp = ((int64_t)IJ + rounds * sms - 1) / (rounds * sms);
    tp = (tp + 7) & ~(int64_t)7;
    if (const char* e = getenv("QMC_DENSE_TILE_PIX")) tp = atoi(e);   // measurement hook
    prm.tile_pix = (int)(tp < 8 ? 8 : tp > DT_PIX ? DT_PIX : tp);
  }
  prm.n_tiles = (IJ + prm.tile_pix - 1) / prm.tile_pix;
  prm.n_bounds = lik->n_bounds;
  prm.inv_a = (lik->flags & QMC_EPI_LSQ) ? 1.0f                              // least squares has no noise model
              : (lik->flags & QMC_EPI_LOGISTIC) ? 1.0f / lik->noise_std      // logistic scale, no sqrt(2)
                                                : 1.0f / probit_scale(lik->noise_std);
  prm.one_sided = 0;
  {
    const char* e = getenv("QMC_DENSE_DEBUG");
    prm.debug = e ? atoi(e) : 0;
  }
  prm.offset = lik->offset;
  for (int i = 0; i < lik->n_bounds; ++i) prm.bounds[i] = lik->bounds[i];
  prm.thr = lik->n_bounds >= 3 ? lik->bounds[1] : 0.0f;
  int epi = DEPI_STABLE;
  if (lik->flags & QMC_EPI_LSQ) epi = DEPI_LSQ;
  else if (lik->flags & QMC_EPI_LOGISTIC) {
    epi = DEPI_LOGISTIC;
    if (lik->n_bounds == 3) {
      const float lo = lik->bounds[0], hi = lik->bounds[2];
      prm.one_sided = lo <= -1e4f && hi >= 1e4f && (-lo * 0.5f) * prm.inv_a > 110.0f && (hi * 0.5f) * prm.inv_a > 110.0f;
    }
  } else if (lik->flags & QMC_EPI_REFERENCE) epi = DEPI_REFERENCE;
  else if (lik->n_bounds == 3) {
    const float lo = lik->bounds[0], hi = lik->bounds[2];
    if (lo <= -1e4f && hi >= 1e4f && (-lo * 0.5f) * prm.inv_a > 30.0f && (hi * 0.5f) * prm.inv_a > 30.0f) epi = DEPI_ONEBIT;
  }
  const bool logd = (lik->flags & QMC_LOG_DOMAIN) != 0;

  if (!px) {   // (with the exchange the partials go through the region's scratch slot, which the kernel leaves zeroed)
    QMC_CUDA_CHECK(cudaMemsetAsync(nll_out_dev, 0, sizeof(double), st));
    if (grad) QMC_CUDA_CHECK(cudaMemsetAsync(gC_out_dev, 0, sizeof(float) * (size_t)R * K, st));
  }
  const size_t smem = dense_smem_map(K, prm.Rp8).total;
  const int grid = prm.n_tiles < sms ? prm.n_tiles : sms;
  switch (epi) {
    case DEPI_ONEBIT: return dense_launch<DEPI_ONEBIT>(prm, logd, grad, grid, smem, st);
    case DEPI_REFERENCE: return dense_launch<DEPI_REFERENCE>(prm, logd, grad, grid, smem, st);
    case DEPI_LSQ: return dense_launch<DEPI_LSQ>(prm, logd, grad, grid, smem, st);
    case DEPI_LOGISTIC: return dense_launch<DEPI_LOGISTIC>(prm, logd, grad, grid, smem, st);
    default: return dense_launch<DEPI_STABLE>(prm, logd, grad, grid, smem, st);
  }
}

extern "C" int qmc_nll_fwd_bwd_dense(const float* S_dev, const float* C_dev, const uint8_t* code_dev,
                                     const qmc_likelihood_t* lik, int IJ, int K, int R, double* nll_out_dev,
                                     float* gS_out_dev, int64_t gs_row_stride, float* gC_out_dev, void* stream) {
  return dense_entry(S_dev, C_dev, code_dev, lik, IJ, K, R, nll_out_dev, gS_out_dev, gs_row_stride, gC_out_dev, nullptr, stream);
}

// ---- fused exchange: regions and the entry point -----------------------------------------------------------------
extern "C" int qmc_nll_fwd_bwd_dense_exchange(const float* S_dev, const float* C_dev, const uint8_t* code_dev,
                                              const qmc_likelihood_t* lik, int IJ, int K, int R, double* nll_out_dev,
                                              float* gS_out_dev, int64_t gs_row_stride, float* gC_out_dev,
                                              const qmc_peer_exchange_t* px, void* stream) {
  QMC_REQUIRE(px, "null exchange descriptor");
  return dense_entry(S_dev, C_dev, code_dev, lik, IJ, K, R, nll_out_dev, gS_out_dev, gs_row_stride, gC_out_dev, px, stream);
}

extern "C" int64_t qmc_peer_region_bytes(int world, int slot_floats) {
  if (world < 1 || world > QMC_PEER_MAX_WORLD || slot_floats < 4 || (slot_floats % 4) != 0) return 0;
  return (int64_t)peer_slot_offset(world, slot_floats, 2, 0);
}

extern "C" int qmc_peer_alloc(int64_t bytes, void** region_out_dev, void* ipc_handle_out64) {
  QMC_REQUIRE(region_out_dev && bytes >= (int64_t)sizeof(PeerHeader), "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI promises a 64-byte handle");
  void* p = nullptr;
  QMC_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
  QMC_CUDA_CHECK(cudaMemset(p, 0, (size_t)bytes));
  QMC_CUDA_CHECK(cudaDeviceSynchronize());
  if (ipc_handle_out64) {
    cudaIpcMemHandle_t h;
    const cudaError_t err = cudaIpcGetMemHandle(&h, p);
    if (err != cudaSuccess) {
      cudaFree(p);
      return set_error(QMC_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(err));
    }
    memcpy(ipc_handle_out64, &h, 64);
  }
  *region_out_dev = p;
  return QMC_OK;
}
extern "C" int qmc_peer_open(const void* ipc_handle64, void** region_out_dev) {
  QMC_REQUIRE(ipc_handle64 && region_out_dev, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64, 64);
  void* p = nullptr;
  QMC_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *region_out_dev = p;
  return QMC_OK;
}
extern "C" int qmc_peer_close(void* region_dev) {
  if (region_dev) QMC_CUDA_CHECK(cudaIpcCloseMemHandle(region_dev));
  return QMC_OK;
}
extern "C" int qmc_peer_free(void* region_dev) {
  if (region_dev) QMC_CUDA_CHECK(cudaFree(region_dev));
  return QMC_OK;
}
extern "C" int qmc_peer_status(const void* own_region_dev, int* status_out, void* stream) {
  QMC_REQUIRE(own_region_dev && status_out, "null argument");
  uint32_t st = 0;
  QMC_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  QMC_CUDA_CHECK(cudaMemcpy(&st, (const uint8_t*)own_region_dev + offsetof(PeerHeader, status), 4, cudaMemcpyDeviceToHost));
  *status_out = (int)st;
  return QMC_OK;
}
