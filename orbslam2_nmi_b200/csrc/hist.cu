// hist.cu -- fused joint histogram + entropy + NMI score; one CTA holds one evaluation at a time
// (default build: one persistent CTA per SM walking the pair schedule, see
// joint_hist_score_persistent_kernel; the other builds launch one CTA per evaluation).
//
// Replaces, per (render, warped frame) pair, the reference's six launches
//   histogram256Kernel + mergeHistogram256Kernel + mergeJointHistogram256Kernel
//   (Thirdparty/CUDA_Functions/NMI.cu:52-161, driver :188-226),
//   ComputeEntropyKernel (:230-267), AddvectorParwiseMidKernel (:270-287),
//   AddVectorPairwiseKernel (:290-363)
// and the 4 MiB cudaMemset + 10 cudaMalloc/cudaFree + blocking D2H around them
// (kernel.cu:49-114).  Semantics follow SURVEY.md Appendix A.6/A.7:
//   J[a*B+b]++ for a = render[p], b = warped[p]  (BG rule NMI.cu:85),
//   e(c) = (c/L) log2f(c/L), L = W*H always (kernel.cu:85),
//   fp32 pairwise trees in the reference's order (strides B/2..1 per row, then
//   over the row sums), SUC / ENMI with the all-zero guard (NMI.cu:342-362).
//
// B200 design: the whole 256x256 joint histogram of one evaluation lives in the
// CTA's shared memory and never touches HBM; the marginals are its row / column
// sums (integers, exact); only one float leaves the SM.  Both images are streamed
// through a 4-stage shared-memory ring filled by the TMA engine
// (cp.async.bulk + mbarrier complete_tx), issued by a dedicated producer warp,
// so the 16 or 32 consumer warps spend their issue slots on shared-memory atomics.
// The epilogue of the default build walks the reference's summation trees depth-first in
// registers, two threads per histogram row (rows_epilogue_fast).
// The hot loop is branch-free: each thread fires all its atomics of a chunk back to
// back (independent ATOMS in flight) and only afterwards inspects the returned values.
//
// Histogram storage policies (template POLICY):
//   P_U16G  256 bins, single pass. 65 536 counters, two 16-bit fields per 32-bit
//           word (128 KiB).  The increment that takes a field across a multiple of
//           4096 ("crossing") is unique; its thread subtracts 4096 again and logs a
//           (bin) event, and the epilogue adds 4096 per event back.  The hot loop
//           does not test every pixel: it ORs the post-increment words of a chunk
//           (one LOP3 per two pixels) and only when a bit >= 12 of either field shows
//           up does it look for its own crossings.  All updates are commutative adds,
//           so the counts are exact provided a field never reaches 2^16, i.e. at most
//           15 crossings are pending at once.  A crossing is repaid before its thread
//           leaves the chunk, and while a thread sits in one chunk the TMA ring lets
//           the other warps process at most 2*kStages-1 = 7 chunks = 57 344 pixels
//           < 15*4096.  (The LDG variant re-synchronises every two chunks instead.)
//           Hot-bin skipping (template SKIPCAP, BG mode): shared-memory atomics on ONE word
//           retire at roughly one per ~100 cycles however many warps issue them, so a flat
//           region (render background 255, saturated sky, warp border) that piles pixels
//           onto a few words makes a pair 2-4x slower.  A sampled histogram per image
//           (image_mode_kernel) names the most frequent level a* of the render and b* of
//           the warp; when together they cover >= 1/6 of the pixels, a thread whose 16
//           pixels all have a == a* keeps them out of the joint histogram and counts
//           their b values into a small per-warp-pair table instead (no return value, so
//           equal lanes merge); likewise all b == b* -> a table over a; both -> one
//           register counter.  The epilogue adds the tables to row a* / column b*.  Every
//           count stays an exact integer; only where it is accumulated changes.
//           Batched searches go one step further (marginal_side_counts): with the exact
//           256-bin histogram of every render and warp (image_hist_kernel, once per image and
//           search, not per pair) those pixels are not counted at all -- row a* and column b*
//           of the joint histogram follow from the marginals.  What remains of a flat region
//           is the streaming of its pixels (C2 on a constant frame: 4.58 -> 3.29 ms per 4 096
//           evaluations, the L2 -> SM bandwidth of 17 GB; sky frame 4.90 -> 4.49 ms).
//   P_U32X2 256 bins, two passes over the pixels, 128 render rows x 256 u32 per
//           pass (128 KiB); no overflow logic, twice the L2->SM traffic.
//   P_B64   64 bins (value >> 2), 8 replicated 64x64 u32 sub-histograms.
#include "nmi_internal.h"

namespace nmi {
namespace {

#ifndef NMI_HIST_CHUNK
#define NMI_HIST_CHUNK 8192
#endif
#ifndef NMI_HIST_STAGES
#define NMI_HIST_STAGES 4
#endif
constexpr int kChunk = NMI_HIST_CHUNK;  // pixels per stage and image
constexpr int kStages = NMI_HIST_STAGES;
static_assert((2 * kStages - 1) * kChunk < 15 * 4096, "in-flight pixels must stay below 15 crossings");
constexpr int kEvCap = 2048;       // >= npix / 4096 drain events (npix <= 8.3M)
constexpr uint32_t kFlagMask = 0xF000F000u;  // a field >= 4096 has one of these bits set
constexpr uint32_t kCross = 0x0FFFu;         // low 12 bits of a field: 0 right after a crossing
constexpr int kHistWords = 32768;  // 128 KiB
#ifndef NMI_B64_COPIES
#define NMI_B64_COPIES 8
#endif
constexpr int kB64Copies = NMI_B64_COPIES;  // replicated 64 x 64 sub-histograms (one per pair of consumer warps)
constexpr int kTermTab = 1024;     // counts below this take their entropy term from a table

enum Policy { P_U16G = 0, P_U32X2 = 2, P_B64 = 3 };

struct __align__(16) Smem {
  uint32_t hist[kHistWords];
  uint8_t rbuf[kStages][kChunk];
  uint8_t wbuf[kStages][kChunk];
  unsigned long long full[kStages];
  unsigned long long empty[kStages];
  float rowE[256];
  uint32_t HA[256];
  uint32_t HB[256];
  float sums[4];
  float term_tab[kTermTab];  // e(c) for small counts, filled per CTA with term() itself
  uint32_t ev_count;
  uint32_t rowmask[8];       // fast epilogue: rows that hold a repaid crossing (bit = render level)
  uint16_t ev_list[kEvCap];  // bin of every repaid crossing (+4096 each)
};

constexpr int kSkipCopies = 8;  // side tables: one per pair of consumer warps
struct __align__(16) SmemSkip {   // follows Smem in the SKIPCAP builds
  uint32_t n1[kSkipCopies][256];  // b-histogram of the uncounted pixels of all-(a == a*) threads
  uint32_t n2[kSkipCopies][256];  // a-histogram of the uncounted pixels of all-(b == b*) threads
  uint32_t nboth;                 // pixels of threads with all a == a* and all b == b*
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) __trap();  // never hang the GPU on a logic error
  } while (!done);
}
// TMA engine 1-D bulk copy global -> shared, completion counted on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- entropy pieces (NMI.cu:240-266) ---------------------------------------
// log2f is the one libm call on the path.  SUC of two nearly independent images is 2(1 - x) with
// x ~ 0.99: one fp32 step of x is 1.8e-5 of a 0.007 score, so "within 1e-5 of the reference" means
// the same bits, and that needs the reference's own logarithm: CUDA's libdevice log2f, the function
// its ComputeEntropyKernel calls (NMI.cu:248).  Measured against the reference's kernels compiled
// for this GPU (oracle/_ref, tests/test_gpu_reference_kernels.py): every term and every score
// bit-identical.  The CPU oracle carries a transcription of the same routine
// (oracle/nmi_oracle.c: log2f_cuda).  term_table_kernel tabulates e(c) for every possible count
// 0..L once per image size, so the histogram kernel itself never evaluates a logarithm.
__device__ __forceinline__ float term(uint32_t c, float L) {
  if (c == 0) return 0.0f;
  const float p = __fdiv_rn((float)c, L);
  return __fmul_rn(p, log2f(p));
}
// Same value as term(c, L): small counts (the vast majority of non-empty bins) are looked up in
// the per-CTA table that was filled with term() itself, so the result is bit-identical.
__device__ __forceinline__ float term_t(const float* tab, const float* __restrict__ gtab, uint32_t c, float L) {
  if (c < (uint32_t)kTermTab) return tab[c];   // shared-memory copy of the first entries
  return gtab != nullptr ? __ldg(gtab + c) : term(c, L);  // full table (L2), else compute
}
// Pairwise tree of NMI.cu:270-287 / :295-338 for n = 32*K values, lane l holding
// x[l + 32k]: strides 16K..32 fold k, strides 16..1 are shuffles. Result in lane 0.
template <int K>
__device__ __forceinline__ float tree_lanes(float (&v)[K]) {
#pragma unroll
  for (int h = K / 2; h >= 1; h /= 2)
#pragma unroll
    for (int k = 0; k < h; k++) v[k] = __fadd_rn(v[k], v[k + h]);
  float x = v[0];
#pragma unroll
  for (int d = 16; d >= 1; d /= 2) x = __fadd_rn(x, __shfl_down_sync(0xffffffffu, x, d));
  return x;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t x) {
#pragma unroll
  for (int d = 16; d >= 1; d /= 2) x += __shfl_xor_sync(0xffffffffu, x, d);
  return x;
}
__device__ __forceinline__ float finish_score(float sa, float sb, float sab, int mode) {
  if (sa == 0.0f && sb == 0.0f && sab == 0.0f) return 0.0f;  // NMI.cu:344,353
  if (mode == NMI_SCORE_ENMI) return __fdiv_rn(__fadd_rn(-sa, -sb), -sab);  // NMI.cu:348
  // NMI.cu:357  2*(1-((-Hab)/((-Ha)+(-Hb))))
  return __fmul_rn(2.0f, __fsub_rn(1.0f, __fdiv_rn(-sab, __fadd_rn(-sa, -sb))));
}

// ---- bin addressing ----------------------------------------------------------
// t = (a << 8) | b (upper bits of t may hold garbage).  U16G: word w = t >> 1 (15 bits: a in the upper
// 8, m = b >> 1 in the lower 7), field t & 1.  With SWZ m is XORed with the low 7 bits of a: the bank
// then depends on both images, so a flat region in one of them no longer piles a warp onto one bank.
// The map stays a bijection (a itself sits untouched in w's upper bits).  Written on the pixel BYTES
// the swizzle is b' = b ^ ((a & 0x7F) << 1) -- one LOP3 on a word of four pixels (accum_fast).
template <bool SWZ>
__device__ __forceinline__ uint32_t u16g_word(uint32_t t) {
  const uint32_t w = (t >> 1) & 0x7FFFu;
  return SWZ ? (w ^ ((t >> 8) & 0x7Fu)) : w;
}
// The thread whose increment made field (t & 1) of word w cross a multiple of 4096 repays
// it.  `nw` is the word right after that increment.
template <bool SWZ>
__device__ __forceinline__ void u16g_repay_if_crossed(Smem& sm, uint32_t t, uint32_t nw) {
  const uint32_t sh = (t & 1u) << 4;
  if (((nw >> sh) & kCross) == 0) {
    atomicSub(sm.hist + u16g_word<SWZ>(t), 0x1000u << sh);
    const uint32_t e = atomicAdd(&sm.ev_count, 1u);
    if (e < kEvCap) sm.ev_list[e] = (uint16_t)t;
  }
}

// 64-bin policy: word of bin (a >> 2, b >> 2) inside a 64 x 64 sub-histogram, t = a << 8 | b.  Row-major with the
// low five bits of the column XORed with the row: the bank of a bin then mixes both images (as the packed-u16
// swizzle does), instead of being the warp image's value alone -- a smooth camera frame put the 32 lanes of an
// atomic on a handful of banks.
__device__ __forceinline__ uint32_t b64_word(uint32_t t) {
  return (((t >> 4) & 0xFC0u) | ((t >> 2) & 0x3Fu)) ^ ((t >> 10) & 0x1Fu);
}

// generic (branchy) per-pixel path: partial chunks and BG == false
template <int POLICY, bool SWZ>
__device__ __forceinline__ void accum_one(Smem& sm, uint32_t t, int pass, int warp) {
  if (POLICY == P_U16G) {
    const uint32_t inc = (t & 1u) ? 0x10000u : 1u;
    const uint32_t old = atomicAdd(sm.hist + u16g_word<SWZ>(t), inc);
    u16g_repay_if_crossed<SWZ>(sm, t, old + inc);
  } else if (POLICY == P_U32X2) {
    if ((int)(t >> 15) == pass) atomicAdd(sm.hist + (t & 0x7FFFu), 1u);
  } else {
    atomicAdd(sm.hist + (warp & (kB64Copies - 1)) * 4096 + b64_word(t), 1u);
  }
}

template <int POLICY, bool SWZ, int NW>
__device__ __forceinline__ void accum_slow(Smem& sm, const uint32_t (&rw)[NW], const uint32_t (&ww)[NW],
                                           int nvalid, int bg, int pass, int warp) {
#pragma unroll
  for (int j = 0; j < NW; j++)
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (4 * j + k < nvalid) {
        const uint32_t t = __byte_perm(ww[j], rw[j], 0x4440 + k * 0x11) & 0xFFFFu;
        if (bg || ((t & 0xFF00u) != 0 && (t & 0xFFu) != 0)) accum_one<POLICY, SWZ>(sm, t, pass, warp);
      }
}

// fast path: full chunk, every pixel counted (nmi_prop_BG == true, the reference default)
// Per pixel the hot loop costs 8 issue slots (was 10.5): the swizzle is applied to four pixels at once on
// the warp-image word (b' = b ^ ((a & 0x7F) << 1): one shift + one LOP3 per word), PRMT then delivers
// t' = a << 8 | b' directly, and the byte address is ONE IMAD, (t' & 0xFFFE) * 2 + base, instead of
// shift / extract / shift / xor / add.  The kernel keeps the issue slots as busy as the ATOMS pipe
// (14.25 instructions per 32-lane atomic = 3.6 clk of issue against 3.5 clk of ATOMS, ncu round 1), so
// instructions are time here.  SCHED: round-2 experiment (source-level grouping of the ATOMS, variants
// 11-13) measured no difference (4.543 vs 4.547 ms) and is gone; the parameter is kept so that the
// kernel templates' signatures do not change.
template <int POLICY, bool SWZ, int NW, int SCHED = 0>
__device__ __forceinline__ void accum_fast(Smem& sm, const uint32_t (&rw)[NW], const uint32_t (&ww)[NW],
                                           int pass, int warp) {
  constexpr int N = NW * 4;
  if (POLICY == P_U16G) {
    // 16 (or 8) independent ATOMS in flight; the returned words are only ORed together.
    // Integer work is split between the ALU pipe (PRMT, LOP3) and the FMA pipe (IMAD).
    uint32_t tt[N], nw[N], ws[NW];
    const uint32_t base = smem_u32(sm.hist);
#pragma unroll
    for (int j = 0; j < NW; j++) ws[j] = SWZ ? (ww[j] ^ ((rw[j] << 1) & 0xFEFEFEFEu)) : ww[j];
#pragma unroll
    for (int i = 0; i < N; i++) {
      const uint32_t t = __byte_perm(ws[i >> 2], rw[i >> 2], 0x4440 + (i & 3) * 0x11);  // a << 8 | b' (+ garbage above)
      uint32_t addr, inc;
      asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr) : "r"(t & 0xFFFEu), "r"(base));  // byte address of word t' >> 1
      asm("mad.lo.u32 %0, %1, 0xFFFF, 1;" : "=r"(inc) : "r"(t & 1u));  // 1 or 0x10000; a real IMAD (FMA pipe)
      uint32_t old;
      asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(inc) : "memory");
      // the bin's un-swizzled name, for the rare repay path: b = b' ^ ((a & 0x7F) << 1)
      tt[i] = t;
      nw[i] = old + inc;  // the word right after this thread's increment
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) acc |= nw[i];
    if (acc & kFlagMask) {  // rare: some field of a touched word is >= 4096
#pragma unroll
      for (int i = 0; i < N; i++) {
        const uint32_t ts = tt[i] & 0xFFFFu;
        u16g_repay_if_crossed<SWZ>(sm, SWZ ? (ts ^ ((ts >> 7) & 0xFEu)) : ts, nw[i]);
      }
    }
  } else if (POLICY == P_U32X2) {
    const uint32_t base = smem_u32(sm.hist);
#pragma unroll
    for (int i = 0; i < N; i++) {
      const uint32_t t = __byte_perm(ww[i >> 2], rw[i >> 2], 0x4440 + (i & 3) * 0x11);
      // predicated reduction, no branch: only rows of this pass's half are counted
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.eq.u32 p, %1, %2;\n\t"
          "@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(base + ((t & 0x7FFFu) << 2)),
          "r"((t >> 15) & 1u), "r"((uint32_t)pass)
          : "memory");
    }
  } else {
    uint32_t* h = sm.hist + (warp & (kB64Copies - 1)) * 4096;
#pragma unroll
    for (int i = 0; i < N; i++) {
      const uint32_t t = __byte_perm(ww[i >> 2], rw[i >> 2], 0x4440 + (i & 3) * 0x11);
      atomicAdd(h + b64_word(t), 1u);
    }
  }
}

// ---- epilogue over the rows of one pass --------------------------------------
// U16G: one pass, 256 rows of 128 packed words.  U32X2: 128 rows of 256 words per
// pass.  B64: 64 rows of 64 words after folding the copies.
// ZERO (packed-u16 only): every word is cleared right after it has been read, so a persistent
// CTA finds an empty histogram for its next evaluation without a separate clearing pass.
// COLS = false / rowmask != nullptr (packed-u16 only): the fast epilogue below has taken the column
// sums and the rows without repaid crossings; only the rows named in rowmask are left for this one.
template <int POLICY, bool SWZ, int NWARPS, bool ZERO = false, bool COLS = true>
__device__ __forceinline__ void rows_epilogue(Smem& sm, int pass, float L, const HistArgs& a,
                                              bool dump, int warp, int lane,
                                              const SmemSkip* sk = nullptr, uint32_t skipT = 0xFFFFFFFFu,
                                              const uint32_t* rowmask = nullptr) {
  constexpr int kConsumers = NWARPS * 32;
  if (POLICY == P_U16G) {
    // skip mode (skipT = a* << 8 | b*): sk->n1[0] / n2[0] / nboth hold the folded side counts
    const bool skip = sk != nullptr && skipT != 0xFFFFFFFFu;
    const uint32_t astar = skip ? (skipT >> 8) & 0xFFu : 0x100u, bstar = skipT & 0xFFu;
    const bool own_bstar = ((bstar >> 1) & 31u) == (uint32_t)lane;  // this lane holds column b* ...
    const uint32_t i_bstar = ((bstar >> 6) << 1) | (bstar & 1u);     // ... in c[i_bstar]
    const uint32_t nev = min(sm.ev_count, (uint32_t)kEvCap);
    uint32_t col[8];
#pragma unroll
    for (int i = 0; i < 8; i++) col[i] = 0;
    for (int row = warp; row < 256; row += NWARPS) {
      if (rowmask != nullptr && ((rowmask[row >> 5] >> (row & 31)) & 1u) == 0u) continue;
      uint32_t c[8];  // c[2k+h] = J[row][2(lane+32k)+h]
#pragma unroll
      for (int k = 0; k < 4; k++) {
        // word of bins (row, 2(lane+32k)) / (row, 2(lane+32k)+1), bank swizzle undone
        const uint32_t wi = u16g_word<SWZ>(((uint32_t)row << 8) | (2u * (lane + 32 * k)));
        const uint32_t wv = sm.hist[wi];
        if (ZERO) sm.hist[wi] = 0u;
        c[2 * k] = wv & 0xFFFFu;
        c[2 * k + 1] = wv >> 16;
      }
      for (uint32_t e = 0; e < nev; e++) {  // repaid crossings of this row: +4096 each
        const uint32_t t = sm.ev_list[e];
        if ((int)(t >> 8) == row) {
          const uint32_t b = t & 0xFFu;
          if (((b >> 1) & 31u) == (uint32_t)lane) {
#pragma unroll
            for (int i = 0; i < 8; i++)
              if ((uint32_t)i == (((b >> 6) << 1) | (b & 1u))) c[i] += 4096u;
          }
        }
      }
      if (skip) {  // the pixels that bypassed the joint histogram
        if ((uint32_t)row == astar) {
#pragma unroll
          for (int i = 0; i < 8; i++) c[i] += sk->n1[0][2 * (lane + 32 * (i >> 1)) + (i & 1)];
        }
        if (own_bstar) {
          const uint32_t add = sk->n2[0][row] + ((uint32_t)row == astar ? sk->nboth : 0u);
#pragma unroll
          for (int i = 0; i < 8; i++)
            if ((uint32_t)i == i_bstar) c[i] += add;
        }
      }
      uint32_t rs = 0;
      float vl[4], vh[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        rs += c[2 * k] + c[2 * k + 1];
        if (COLS) {
          col[2 * k] += c[2 * k];
          col[2 * k + 1] += c[2 * k + 1];
        }
        vl[k] = term_t(sm.term_tab, a.term_tab, c[2 * k], L);
        vh[k] = term_t(sm.term_tab, a.term_tab, c[2 * k + 1], L);
        if (dump) {
          a.dumpJ[row * 256 + 2 * (lane + 32 * k)] = c[2 * k];
          a.dumpJ[row * 256 + 2 * (lane + 32 * k) + 1] = c[2 * k + 1];
        }
      }
      rs = warp_sum(rs);
      // tree over b: strides 128,64 fold k; 32..2 are lane shuffles (b = 2(lane+32k)+h);
      // stride 1 folds h.
      const float tl = tree_lanes<4>(vl);
      const float th = tree_lanes<4>(vh);
      if (lane == 0) {
        sm.rowE[row] = __fadd_rn(tl, th);
        sm.HA[row] = rs;
      }
    }
    if (COLS) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        atomicAdd(&sm.HB[2 * (lane + 32 * k)], col[2 * k]);
        atomicAdd(&sm.HB[2 * (lane + 32 * k) + 1], col[2 * k + 1]);
      }
    }
  } else if (POLICY == P_U32X2) {
    uint32_t col[8];
#pragma unroll
    for (int i = 0; i < 8; i++) col[i] = 0;
    for (int lr = warp; lr < 128; lr += NWARPS) {
      const int row = pass * 128 + lr;
      float v[8];
      uint32_t rs = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const uint32_t c = sm.hist[lr * 256 + lane + 32 * k];
        rs += c;
        col[k] += c;
        v[k] = term_t(sm.term_tab, a.term_tab, c, L);
        if (dump) a.dumpJ[row * 256 + lane + 32 * k] = c;
      }
      rs = warp_sum(rs);
      const float tr = tree_lanes<8>(v);
      if (lane == 0) {
        sm.rowE[row] = tr;
        sm.HA[row] = rs;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) atomicAdd(&sm.HB[lane + 32 * k], col[k]);
  } else {
    // fold the replicated sub-histograms into copy 0 (all consumers)
    const int tid = warp * 32 + lane;
    for (int i = tid; i < 4096; i += kConsumers) {
      uint32_t s = 0;
#pragma unroll
      for (int r = 0; r < kB64Copies; r++) s += sm.hist[r * 4096 + i];
      sm.hist[i] = s;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
    uint32_t col[2] = {0, 0};
    for (int row = warp; row < 64; row += NWARPS) {
      float v[2];
      uint32_t rs = 0;
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const uint32_t c = sm.hist[row * 64 + ((lane + 32 * k) ^ (row & 31))];  // b64_word's swizzle undone
        rs += c;
        col[k] += c;
        v[k] = term_t(sm.term_tab, a.term_tab, c, L);
        if (dump) a.dumpJ[row * 64 + lane + 32 * k] = c;
      }
      rs = warp_sum(rs);
      const float tr = tree_lanes<2>(v);
      if (lane == 0) {
        sm.rowE[row] = tr;
        sm.HA[row] = rs;
      }
    }
#pragma unroll
    for (int k = 0; k < 2; k++) atomicAdd(&sm.HB[lane + 32 * k], col[k]);
  }
}

// ---- fast epilogue: packed-u16, swizzled layout, 16 consumer warps --------------------------
// The warp-per-row epilogue above costs ~12 us per evaluation (ncu on 4096 evaluations of two-chunk
// images: ~175 instructions per row, 80 % of all instructions executed) -- 8 % of the kernel at
// C2 -- because every row pays 15 dependent shuffles for the trees and the row sum, and all 32
// lanes repeat the address and bookkeeping work.  Here a row belongs to TWO THREADS and the reference's
// tree (NMI.cu:270-287: strides 128, 64, ..., 1 over b) is walked depth-first in registers:
// with b = 2m + h, the last stride (1) joins the h = 0 and h = 1 halves of the packed words, the
// one before (2) joins even and odd word index m -- thread A of the row takes the even words,
// thread B the odd ones -- and strides 4..128 are bits 1..6 of m, the recursion below, deepest
// first.  fp32 addition is commutative, the pairing is the reference's: same bits.
// Bank conflicts: lanes 0-15 are the A threads of 16 consecutive rows, lanes 16-31 the B threads of
// the same rows.  The swizzle puts word m of row a in bank (m ^ a) & 31, so the A lanes read an
// aligned block of 16 distinct banks; B walks its subtree with bit 4 of m flipped (children swapped
// at one level -- same sums) and lands in the other block.  One wavefront per load.
// Rows that hold a repaid crossing (+4096 per event, rare) are left to the warp-per-row code, which
// knows how to add them; the marginal over the camera image comes from a column pass in which 32
// lanes read 32 consecutive words of one row (again one wavefront), plus 4096 per event.
// Entropy term from the tables only (the fast epilogue runs only when the per-image-size table
// exists): the first kTermTab entries sit in shared memory, the rest in global memory.  Keeps a
// leaf at a dozen instructions.  (Measured and dropped: a 4096-entry shared-memory table -- all
// a field can hold once every crossing is repaid -- and immediate-offset addressing of bits 5-6;
// both shorten the epilogue on small images and make the C2 kernel 1-4 % slower.)
__device__ __forceinline__ float term_lut(const float* tab, const float* __restrict__ gtab, uint32_t c) {
  return c < (uint32_t)kTermTab ? tab[c] : __ldg(gtab + c);
}
// subtree over the words wx ^ {bits BIT..6}: splits on bit BIT first, bit 6 deepest
template <int BIT, uint32_t M>
__device__ __forceinline__ void row_subtree(const uint32_t* hist, uint32_t wx, const float* tab,
                                            const float* __restrict__ gtab, uint32_t& rs, float& lo, float& hi) {
  if constexpr (BIT == 7) {
    const uint32_t wv = hist[wx ^ M];
    const uint32_t cl = wv & 0xFFFFu, ch = wv >> 16;
    rs += cl + ch;
    lo = term_lut(tab, gtab, cl);
    hi = term_lut(tab, gtab, ch);
  } else {
    float l0, h0, l1, h1;
    row_subtree<BIT + 1, M>(hist, wx, tab, gtab, rs, l0, h0);
    row_subtree<BIT + 1, (M | (1u << BIT))>(hist, wx, tab, gtab, rs, l1, h1);
    lo = __fadd_rn(l0, l1);
    hi = __fadd_rn(h0, h1);
  }
}

// Call with all 512 consumer threads after the barrier that ends the pixel loop; sm.rowmask must be
// zero on entry and is left set (the caller clears it if the CTA goes on to another pair).
// Leaves sm.rowE, sm.HA, sm.HB complete once the callers' next barrier has passed.
template <bool ZERO>
__device__ __forceinline__ void rows_epilogue_fast(Smem& sm, float L, const HistArgs& a, int warp, int lane) {
  constexpr int NWARPS = 16, kConsumers = NWARPS * 32;
  const int tid = warp * 32 + lane;
  const uint32_t nev = min(sm.ev_count, (uint32_t)kEvCap);
  if (nev != 0u) {  // CTA-uniform
    for (uint32_t e = tid; e < nev; e += kConsumers) {
      const uint32_t t = sm.ev_list[e];
      atomicOr(&sm.rowmask[t >> 13], 1u << ((t >> 8) & 31u));
      atomicAdd(&sm.HB[t & 0xFFu], 4096u);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
  }
  // rows: thread pair (lane, lane + 16) owns row 16 * warp + (lane & 15)
  {
    const uint32_t row = (uint32_t)warp * 16u + ((uint32_t)lane & 15u);
    const uint32_t pb = (uint32_t)lane >> 4;
    const uint32_t rowx = ((row << 7) ^ (row & 0x7Fu)) ^ (pb ? 17u : 0u);  // u16g_word<true>(row << 8); B: odd words, bit 4 swapped
    uint32_t rs = 0;
    // bits 1..3 of m: eight subtrees visited depth-first (bit 1 is the split nearest the root, so
    // the visiting order counts q = bits 3..1 bit-reversed) and folded like a binary counter --
    // a rolled loop, so the code stays a few hundred instructions; bits 4..6: unrolled recursion
    float s0l = 0.f, s0h = 0.f, s1l = 0.f, s1h = 0.f, s2l = 0.f, s2h = 0.f, lo = 0.f, hi = 0.f;
#pragma unroll 1
    for (uint32_t j = 0; j < 8u; j++) {
      const uint32_t q = ((j & 1u) << 2) | (j & 2u) | (j >> 2);
      row_subtree<4, 0u>(sm.hist, rowx ^ (q << 1), sm.term_tab, a.term_tab, rs, lo, hi);
      if (j & 1u) {
        lo = __fadd_rn(s0l, lo); hi = __fadd_rn(s0h, hi);
        if (j & 2u) {
          lo = __fadd_rn(s1l, lo); hi = __fadd_rn(s1h, hi);
          if (j & 4u) { lo = __fadd_rn(s2l, lo); hi = __fadd_rn(s2h, hi); }
          else { s2l = lo; s2h = hi; }
        } else { s1l = lo; s1h = hi; }
      } else { s0l = lo; s0h = hi; }
    }
    const float lo2 = __shfl_down_sync(0xffffffffu, lo, 16);
    const float hi2 = __shfl_down_sync(0xffffffffu, hi, 16);
    const uint32_t rs2 = __shfl_down_sync(0xffffffffu, rs, 16);
    if (pb == 0u && ((sm.rowmask[row >> 5] >> (row & 31u)) & 1u) == 0u) {
      sm.rowE[row] = __fadd_rn(__fadd_rn(lo, lo2), __fadd_rn(hi, hi2));  // strides 2, 2, then 1
      sm.HA[row] = rs + rs2;
    }
  }
  if (nev != 0u)  // the few rows with repaid crossings, warp per row
    rows_epilogue<P_U16G, true, NWARPS, false, false>(sm, 0, L, a, false, warp, lane, nullptr, 0xFFFFFFFFu, sm.rowmask);
  if (ZERO) asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // every row has been read
  // columns: thread = (word index, quarter of the rows)
  {
    const uint32_t cw = (uint32_t)tid & 127u, r0 = ((uint32_t)tid >> 7) * 64u;
    uint32_t slo = 0, shi = 0;
#pragma unroll 16
    for (uint32_t r = r0; r < r0 + 64u; r++) {
      const uint32_t idx = ((r << 7) ^ (r & 0x7Fu)) ^ cw;
      const uint32_t wv = sm.hist[idx];
      if (ZERO) sm.hist[idx] = 0u;
      slo += wv & 0xFFFFu;
      shi += wv >> 16;
    }
    atomicAdd(&sm.HB[2u * cw], slo);
    atomicAdd(&sm.HB[2u * cw + 1u], shi);
  }
}

// Parity dump for the builds that run the fast epilogue: the 65 536 counters are copied out of the
// swizzled packed-u16 words BEFORE the epilogue reads (and, in the persistent kernel, clears) them,
// repaid crossings added (+4096 each).  The epilogue then runs exactly as in a timed launch, so the
// marginals and the score that are read back are the fast path's own.  All 512 consumers call it.
__device__ __forceinline__ void dump_joint_swizzled(const Smem& sm, uint32_t* __restrict__ J, int tid) {
  constexpr int kConsumers = 16 * 32;
  for (uint32_t i = tid; i < (uint32_t)kHistWords; i += kConsumers) {
    const uint32_t r = i >> 7, m = i & 127u;
    const uint32_t wv = sm.hist[((r << 7) ^ (r & 0x7Fu)) ^ m];  // u16g_word<true>((r << 8) | 2m)
    J[r * 256u + 2u * m] = wv & 0xFFFFu;
    J[r * 256u + 2u * m + 1u] = wv >> 16;
  }
  __threadfence_block();
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
  const uint32_t nev = min(sm.ev_count, (uint32_t)kEvCap);
  for (uint32_t e = tid; e < nev; e += kConsumers) atomicAdd(J + sm.ev_list[e], 4096u);
}

template <int NW>
__device__ __forceinline__ void load_words(uint32_t (&dst)[NW], const uint8_t* p) {
  if (NW % 4 == 0) {
#pragma unroll
    for (int j = 0; j < NW / 4; j++) {
      const uint4 v = *reinterpret_cast<const uint4*>(p + 16 * j);
      dst[4 * j] = v.x; dst[4 * j + 1] = v.y; dst[4 * j + 2] = v.z; dst[4 * j + 3] = v.w;
    }
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    dst[0] = v.x; dst[NW - 1] = v.y;
  }
}
template <int NW>
__device__ __forceinline__ void ldg_words(uint32_t (&dst)[NW], const uint8_t* p) {
  if (NW % 4 == 0) {
#pragma unroll
    for (int j = 0; j < NW / 4; j++) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + 16 * j));
      dst[4 * j] = v.x; dst[4 * j + 1] = v.y; dst[4 * j + 2] = v.z; dst[4 * j + 3] = v.w;
    }
  } else {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    dst[0] = v.x; dst[NW - 1] = v.y;
  }
}

// ---- hot-bin skipping through the image marginals ---------------------------------------------
// With a.img_hist (the full 256-bin histogram of every render and every warp, one small launch per
// search) the pixels of a thread whose 16 render values all equal a*, or whose 16 warp values all equal
// b*, are not counted AT ALL: every such pixel lies in row a* or in column b* of the joint histogram,
// and those follow from the marginals.  With J' the counted part and S the skipped pixels,
//   column b != b*:  J_S[a*][b] = HB[b] - sum_a J'[a][b]      (only row a* of S reaches such a column)
//   row    a != a*:  J_S[a][b*] = HA[a] - sum_b J'[a][b]
//   J_S[a*][b*]    = HA[a*] - sum_b J'[a*][b] - sum_{b != b*} J_S[a*][b]
// -- exact integers, no sampling.  This pass forms the row and column sums of J' (repaid crossings
// included) and leaves J_S in the places the epilogue already reads the side counts from
// (sk.n1[0] = row a*, sk.n2[0] = column b*, sk.nboth).  Called by all consumers after the barrier that
// ends the pixel loop; the caller's next barrier orders it before the epilogue.
template <bool SWZ, int NWARPS>
__device__ __forceinline__ void marginal_side_counts(Smem& sm, SmemSkip& sk, const HistArgs& a, int2 pr,
                                                     uint32_t skipT, int warp, int lane) {
  constexpr int kConsumers = NWARPS * 32;
  const int tid = warp * 32 + lane;
  const uint32_t astar = (skipT >> 8) & 0xFFu, bstar = skipT & 0xFFu;
  const uint32_t nev = min(sm.ev_count, (uint32_t)kEvCap);
  uint32_t col[8];
#pragma unroll
  for (int i = 0; i < 8; i++) col[i] = 0;
  for (int row = warp; row < 256; row += NWARPS) {
    uint32_t c[8];  // c[2k+h] = J'[row][2(lane+32k)+h]
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t wv = sm.hist[u16g_word<SWZ>(((uint32_t)row << 8) | (2u * (lane + 32 * k)))];
      c[2 * k] = wv & 0xFFFFu;
      c[2 * k + 1] = wv >> 16;
    }
    for (uint32_t e = 0; e < nev; e++) {  // repaid crossings of this row: +4096 each
      const uint32_t t = sm.ev_list[e];
      if ((int)(t >> 8) == row) {
        const uint32_t b = t & 0xFFu;
        if (((b >> 1) & 31u) == (uint32_t)lane) {
#pragma unroll
          for (int i = 0; i < 8; i++)
            if ((uint32_t)i == (((b >> 6) << 1) | (b & 1u))) c[i] += 4096u;
        }
      }
    }
    uint32_t rs = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      rs += c[i];
      col[i] += c[i];
    }
    rs = warp_sum(rs);
    if (lane == 0) sm.HA[row] = rs;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    atomicAdd(&sm.HB[2 * (lane + 32 * k)], col[2 * k]);
    atomicAdd(&sm.HB[2 * (lane + 32 * k) + 1], col[2 * k + 1]);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
  const uint32_t* __restrict__ HAt = a.img_hist + 256 * (size_t)pr.x;
  const uint32_t* __restrict__ HBt = a.img_hist + 256 * (size_t)(a.nrenders + pr.y);
  if (tid < 256) {
    sk.n1[0][tid] = (uint32_t)tid != bstar ? __ldg(HBt + tid) - sm.HB[tid] : 0u;
    sk.n2[0][tid] = (uint32_t)tid != astar ? __ldg(HAt + tid) - sm.HA[tid] : 0u;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
  if (warp == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) t += sk.n1[0][lane + 32 * k];
    t = warp_sum(t);
    if (lane == 0) sk.nboth = __ldg(HAt + astar) - sm.HA[astar] - t;
  }
  if (tid < 256) sm.HB[tid] = 0;  // the epilogue's column sums start from zero
}

// NWARPS consumer warps.  With 16 of them a 17th warp is the dedicated TMA producer; with 32
// (the 1024-thread CTA limit) thread 0 doubles as producer: after releasing chunk k it waits
// until every warp has released it and refills that stage with chunk k + kStages.
// LDGD (USE_TMA == false only): chunks a thread keeps in flight in registers (ld.global.nc straight from
// L2, no shared-memory staging at all: the pixels then cross the shared-memory array zero times instead
// of twice -- round-2 experiment, variants 11-13).
template <int POLICY, bool USE_TMA, int NWARPS, bool SWZ, bool SKIPCAP, bool FASTEP = false, int LDGD = 1>
__global__ void __launch_bounds__(NWARPS == 32 ? 1024 : NWARPS * 32 + 32, 1)
joint_hist_score_kernel(const HistArgs a) {
  static_assert(!SKIPCAP || POLICY == P_U16G, "hot-bin skipping is built for the packed-u16 policy");
  static_assert(!FASTEP || (POLICY == P_U16G && SWZ && NWARPS == 16 && !SKIPCAP), "fast epilogue: packed-u16, swizzled, 16 warps");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  SmemSkip& sk = *reinterpret_cast<SmemSkip*>(smem_raw + sizeof(Smem));  // SKIPCAP builds only
  constexpr int BINS = POLICY == P_B64 ? 64 : 256;
  constexpr int NPASS = POLICY == P_U32X2 ? 2 : 1;
  constexpr int kConsumers = NWARPS * 32;
  constexpr bool INLINE_PRODUCER = NWARPS == 32;
  constexpr int kThreads = INLINE_PRODUCER ? kConsumers : kConsumers + 32;
  constexpr int PIX = kChunk / kConsumers;  // pixels per thread and chunk: 16 or 8
  constexpr int NW = PIX / 4;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int2 pr = a.pairs[blockIdx.x];
  const uint8_t* __restrict__ rimg = a.renders + (size_t)pr.x * a.render_pitch;
  const uint8_t* __restrict__ wimg = a.warps + (size_t)pr.y * a.warp_pitch;
  const uint32_t npix = a.npix;
  const int nchunks = (int)((npix + kChunk - 1) / kChunk);
  const int total = nchunks * NPASS;
  const float L = (float)a.length;
  const bool dump = a.dumpJ != nullptr && (int)blockIdx.x == a.dump_pair;

  // ---- init ----
  {
    uint4* h4 = reinterpret_cast<uint4*>(sm.hist);
    for (int i = tid; i < kHistWords / 4; i += kThreads) h4[i] = make_uint4(0, 0, 0, 0);
    if (tid < 256) sm.HB[tid] = 0;
    if (tid < 8) sm.rowmask[tid] = 0;
    if (SKIPCAP) {
      for (int i = tid; i < 2 * kSkipCopies * 256; i += kThreads) (&sk.n1[0][0])[i] = 0;
      if (tid == 0) sk.nboth = 0;
    }
    // e(c) for c < kTermTab: from the per-image-size table of term_table_kernel, or computed here
    if (a.term_tab != nullptr)
      for (int i = tid; i < kTermTab && (uint32_t)i <= a.length; i += kThreads) sm.term_tab[i] = __ldg(a.term_tab + i);
    else
      for (int i = tid; i < kTermTab; i += kThreads) sm.term_tab[i] = term((uint32_t)i, L);
    if (tid == 0) {
      sm.ev_count = 0;
      for (int s = 0; s < kStages; s++) {
        mbar_init(&sm.full[s], 1);
        mbar_init(&sm.empty[s], NWARPS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  // hot-bin skipping: the sampled modes of this pair's two images decide (CTA-uniform)
  uint32_t skipT = 0xFFFFFFFFu;
  if (SKIPCAP && a.img_mode != nullptr && a.bg && a.skip_mode != 0) {
    const uint32_t ka = __ldg(a.img_mode + pr.x), kb = __ldg(a.img_mode + a.nrenders + pr.y);
    if (a.skip_mode == 2 || ((unsigned long long)(ka >> 8) + (kb >> 8)) * 6ull >= a.sample_total)
      skipT = ((ka & 0xFFu) << 8) | (kb & 0xFFu);
  }
  const bool skip = SKIPCAP && skipT != 0xFFFFFFFFu;
  const bool marg = skip && a.img_hist != nullptr;  // skipped pixels come back through the image marginals
  const uint32_t a4 = ((skipT >> 8) & 0xFFu) * 0x01010101u, b4 = (skipT & 0xFFu) * 0x01010101u;
  uint32_t nboth = 0;
  __syncthreads();

  auto issue_chunk = [&](int k) {  // one elected thread: both images of chunk k -> stage k % kStages
    const int st = k % kStages;
    const int ck = k % nchunks;
    const uint32_t off = (uint32_t)ck * kChunk;
    uint32_t bytes = npix - off;
    bytes = bytes > (uint32_t)kChunk ? (uint32_t)kChunk : ((bytes + 15u) & ~15u);
    mbar_expect_tx(&sm.full[st], 2 * bytes);
    tma_load_1d(sm.rbuf[st], rimg + off, bytes, &sm.full[st]);
    tma_load_1d(sm.wbuf[st], wimg + off, bytes, &sm.full[st]);
  };

  if (!INLINE_PRODUCER && warp == NWARPS) {
    // ===== producer warp: TMA ring =====
    if (USE_TMA && lane == 0) {
      for (int k = 0; k < total; k++) {
        if (k >= kStages) mbar_wait(&sm.empty[k % kStages], ((k / kStages) - 1) & 1);
        issue_chunk(k);
      }
    }
  } else {
    if (INLINE_PRODUCER && USE_TMA && tid == 0)
      for (int k = 0; k < kStages && k < total; k++) issue_chunk(k);
    // ===== consumers =====
    uint32_t qr[LDGD][NW], qw[LDGD][NW];  // LDG build: the next LDGD chunks of this thread, in registers
#pragma unroll
    for (int d = 0; d < LDGD; d++)
#pragma unroll
      for (int j = 0; j < NW; j++) qr[d][j] = qw[d][j] = 0;
    if (!USE_TMA) {
#pragma unroll
      for (int d = 0; d < LDGD; d++) {
        const uint32_t off = (uint32_t)(d % nchunks) * kChunk + (uint32_t)tid * PIX;
        if (d < total && off < npix) {
          ldg_words<NW>(qr[d], rimg + off);
          ldg_words<NW>(qw[d], wimg + off);
        }
      }
    }
    for (int k0 = 0; k0 < total; k0 += (USE_TMA ? 1 : LDGD))
#pragma unroll
    for (int dd = 0; dd < (USE_TMA ? 1 : LDGD); dd++) {
      const int k = k0 + dd;
      if (k >= total) break;
      const int pass = k / nchunks;
      const int ck = k - pass * nchunks;
      const uint32_t off = (uint32_t)ck * kChunk + (uint32_t)tid * PIX;
      uint32_t r[NW], w[NW];
      if (USE_TMA) {
        const int st = k % kStages;
        mbar_wait(&sm.full[st], (k / kStages) & 1);
        load_words<NW>(r, sm.rbuf[st] + tid * PIX);
        load_words<NW>(w, sm.wbuf[st] + tid * PIX);
      } else {
#pragma unroll
        for (int j = 0; j < NW; j++) { r[j] = qr[dd][j]; w[j] = qw[dd][j]; }
        if (k + LDGD < total) {  // refill this register slot with the chunk LDGD ahead
          const int ck2 = (k + LDGD) % nchunks;
          const uint32_t off2 = (uint32_t)ck2 * kChunk + (uint32_t)tid * PIX;
          if (off2 < npix) {
            ldg_words<NW>(qr[dd], rimg + off2);
            ldg_words<NW>(qw[dd], wimg + off2);
          }
        }
      }
      const int nvalid = off >= npix ? 0 : (int)min((uint32_t)PIX, npix - off);
      if (nvalid == PIX && a.bg) {
        if (!skip) {
          accum_fast<POLICY, SWZ, NW>(sm, r, w, pass, warp);
        } else {
          // all of this thread's pixels at a*, or all at b*?  Then they stay out of the joint
          // histogram (see the header): side table over the other image's values instead.
          uint32_t dr = 0, dw = 0;
#pragma unroll
          for (int j = 0; j < NW; j++) {
            dr |= r[j] ^ a4;
            dw |= w[j] ^ b4;
          }
          if (dr != 0u && dw != 0u) {
            accum_fast<POLICY, SWZ, NW>(sm, r, w, pass, warp);
          } else if (marg) {
            // not counted at all: marginal_side_counts() reconstructs row a* and column b*
          } else if (dr == 0u && dw == 0u) {
            nboth += PIX;
          } else {
            const uint32_t (&x)[NW] = dr == 0u ? w : r;
            const uint32_t base = smem_u32(dr == 0u ? sk.n1[warp & (kSkipCopies - 1)] : sk.n2[warp & (kSkipCopies - 1)]);
#pragma unroll
            for (int i = 0; i < PIX; i++) {
              const uint32_t v = __byte_perm(x[i >> 2], 0u, 0x4440 + (i & 3));
              asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + v * 4u) : "memory");
            }
          }
        }
      } else if (nvalid > 0)
        accum_slow<POLICY, SWZ, NW>(sm, r, w, nvalid, a.bg, pass, warp);
      if (USE_TMA) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[k % kStages]);
        if (INLINE_PRODUCER && tid == 0 && k + kStages < total) {
          mbar_wait(&sm.empty[k % kStages], (k / kStages) & 1);
          issue_chunk(k + kStages);
        }
      } else if (POLICY == P_U16G && (k & 3) == 3) {
        // bound the in-flight pixels: between two barriers the fastest warp gets at most 4 chunks
        // = 32 768 pixels ahead of a thread that still has a crossing to repay (< 15 * 4096)
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
      }
      if (NPASS == 2 && ck == nchunks - 1) {
        // end of a pass: rows of this half are final
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
        rows_epilogue<POLICY, SWZ, NWARPS>(sm, pass, L, a, dump, warp, lane);
        if (pass == 0) {
          asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
          uint4* h4 = reinterpret_cast<uint4*>(sm.hist);
          for (int i = tid; i < kHistWords / 4; i += kConsumers) h4[i] = make_uint4(0, 0, 0, 0);
          asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
        }
      }
    }
    if (NPASS == 1) {
      if (SKIPCAP && marg) {
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // every increment of this pair has landed
        if constexpr (SKIPCAP) marginal_side_counts<SWZ, NWARPS>(sm, sk, a, pr, skipT, warp, lane);
      } else if (skip) {  // fold the side tables into copy 0
        if (nboth) atomicAdd(&sk.nboth, nboth);
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
        for (int i = tid; i < 512; i += kConsumers) {
          uint32_t* t0 = i < 256 ? &sk.n1[0][i] : &sk.n2[0][i - 256];
          uint32_t sum = 0;
#pragma unroll
          for (int c = 0; c < kSkipCopies; c++) sum += t0[c * 256];
          *t0 = sum;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
      if (FASTEP && a.term_tab != nullptr) {
        if (dump) dump_joint_swizzled(sm, a.dumpJ, tid);  // CTA-uniform
        rows_epilogue_fast<false>(sm, L, a, warp, lane);
      } else
        rows_epilogue<POLICY, SWZ, NWARPS>(sm, 0, L, a, dump, warp, lane, SKIPCAP ? &sk : nullptr, skipT);
    }
  }
  __syncthreads();

  // ---- final trees: SAB over row sums, SA over HA terms, SB over HB terms ----
  if (warp < 3) {
    constexpr int K = BINS / 32;
    float v[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int i = lane + 32 * k;
      v[k] = warp == 0 ? sm.rowE[i] : term_t(sm.term_tab, a.term_tab, warp == 1 ? sm.HA[i] : sm.HB[i], L);
    }
    const float s = tree_lanes<K>(v);
    if (lane == 0) sm.sums[warp] = s;
  }
  if (dump && tid < BINS) {
    a.dumpHA[tid] = sm.HA[tid];
    a.dumpHB[tid] = sm.HB[tid];
  }
  __syncthreads();
  if (tid == 0)
    a.scores[a.out_index ? a.out_index[blockIdx.x] : blockIdx.x] =
        finish_score(sm.sums[1], sm.sums[2], sm.sums[0], a.mode);
}

// ---- persistent build of the default configuration -------------------------------------
// Packed-u16 histogram, TMA ring, 16 consumer warps + producer warp, one CTA per SM that walks
// the pair schedule with stride gridDim.x.  The fixed cost per evaluation, measured as 4096
// evaluations of two-chunk images (tools/exp_hist_overhead.py): one CTA per pair with the
// warp-per-row epilogue 0.40 ms (14 us per evaluation -- 80 % of the instructions of such a launch
// are the epilogue's), with the fast epilogue 0.26 ms, persistent with the fast epilogue 0.22 ms.
// At C2 (254 chunks per image) that is 4.73 -> 4.55 ms per 4096 evaluations.  Here the CTA is set
// up once; the producer warp runs ahead into the NEXT pair's first kStages chunks while the
// consumers are still in the epilogue of the current one (the ring is idle then), the epilogue
// clears each histogram word as it reads it, and the only per-pair state left to reset is 1 KiB
// of marginals, the row mask and the event counter.
// The ring's stage / phase arithmetic runs on a chunk counter that never restarts (kk), so the
// mbarriers need no re-initialisation between pairs.  The schedule is static (pair i on CTA
// i mod gridDim.x): the builds with the side tables, whose pairs differ a lot in cost, stay on
// one CTA per pair and the hardware's dynamic block scheduler.
// Launched with 544 threads.  -DNMI_HIST_REGCAP_THREADS=1024 caps the kernel at 64 registers (ptxas:
// no spills) so that other kernels' CTAs fit next to this one on an SM.  Measured (round 2, two
// contexts alternating searches on one GPU, with and without a common L1/shared carve-out on all
// kernels): the kernel itself gets 2 % slower (4.55 -> 4.64 ms) and the overlap buys nothing,
// because this kernel already keeps 74 % of the issue slots and 91 % of the shared-memory pipe busy
// (profiles/r01_hist_ncu_summary.txt) -- the render-stage kernels are issue-bound too.  Default: 95
// registers.
#ifndef NMI_HIST_REGCAP_THREADS
#define NMI_HIST_REGCAP_THREADS (16 * 32 + 32)
#endif
// DUMP: the parity build (nmi_get_hist_path): the same source, plus the copy-out of the joint
// histogram in front of the epilogue.  A separate instantiation, because this kernel's speed moves
// by several per cent with its register allocation (the mere presence of the never-taken dump branch
// cost 4.55 -> 4.72 ms at C2): the timed build contains no parity code at all.
template <bool SWZ, bool SKIPCAP, bool DUMP = false, int SCHED = 0>
__global__ void __launch_bounds__(NMI_HIST_REGCAP_THREADS, 1)
joint_hist_score_persistent_kernel(const HistArgs a) {
  constexpr int NWARPS = 16;
  constexpr int kConsumers = NWARPS * 32;
  constexpr int kThreads = kConsumers + 32;
  constexpr int PIX = kChunk / kConsumers;  // 16 pixels per thread and chunk
  constexpr int NW = PIX / 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  SmemSkip& sk = *reinterpret_cast<SmemSkip*>(smem_raw + sizeof(Smem));  // SKIPCAP builds only

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t npix = a.npix;
  const int nchunks = (int)((npix + kChunk - 1) / kChunk);
  const float L = (float)a.length;

  // ---- once per CTA ----
  {
    uint4* h4 = reinterpret_cast<uint4*>(sm.hist);
    for (int i = tid; i < kHistWords / 4; i += kThreads) h4[i] = make_uint4(0, 0, 0, 0);
    if (tid < 256) sm.HB[tid] = 0;
    if (tid < 8) sm.rowmask[tid] = 0;
    if (SKIPCAP) {
      for (int i = tid; i < 2 * kSkipCopies * 256; i += kThreads) (&sk.n1[0][0])[i] = 0;
      if (tid == 0) sk.nboth = 0;
    }
    if (a.term_tab != nullptr)
      for (int i = tid; i < kTermTab && (uint32_t)i <= a.length; i += kThreads) sm.term_tab[i] = __ldg(a.term_tab + i);
    else
      for (int i = tid; i < kTermTab; i += kThreads) sm.term_tab[i] = term((uint32_t)i, L);
    if (tid == 0) {
      sm.ev_count = 0;
      for (int s = 0; s < kStages; s++) {
        mbar_init(&sm.full[s], 1);
        mbar_init(&sm.empty[s], NWARPS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  __syncthreads();

  if (warp == NWARPS) {
    // ===== producer warp: one TMA ring across all of this CTA's pairs =====
    if (lane == 0) {
      int kk = 0;
      for (int pi = blockIdx.x; pi < a.npairs; pi += gridDim.x) {
        const int2 pr = a.pairs[pi];
        const uint8_t* rimg = a.renders + (size_t)pr.x * a.render_pitch;
        const uint8_t* wimg = a.warps + (size_t)pr.y * a.warp_pitch;
        for (int k = 0; k < nchunks; k++, kk++) {
          const int st = kk % kStages;
          if (kk >= kStages) mbar_wait(&sm.empty[st], ((kk / kStages) - 1) & 1);
          const uint32_t off = (uint32_t)k * kChunk;
          uint32_t bytes = npix - off;
          bytes = bytes > (uint32_t)kChunk ? (uint32_t)kChunk : ((bytes + 15u) & ~15u);
          mbar_expect_tx(&sm.full[st], 2 * bytes);
          tma_load_1d(sm.rbuf[st], rimg + off, bytes, &sm.full[st]);
          tma_load_1d(sm.wbuf[st], wimg + off, bytes, &sm.full[st]);
        }
      }
    }
    return;  // the consumers synchronise among themselves with named barrier 1 from here on
  }

  // ===== consumers =====
  int kk = 0;
  for (int pi = blockIdx.x; pi < a.npairs; pi += gridDim.x) {
    const int2 pr = a.pairs[pi];
    const bool dump = a.dumpJ != nullptr && pi == (DUMP ? a.dump_pair : 0);
    // hot-bin skipping: the sampled modes of this pair's two images decide (CTA-uniform)
    uint32_t skipT = 0xFFFFFFFFu;
    if (SKIPCAP && a.img_mode != nullptr && a.bg && a.skip_mode != 0) {
      const uint32_t ka = __ldg(a.img_mode + pr.x), kb = __ldg(a.img_mode + a.nrenders + pr.y);
      if (a.skip_mode == 2 || ((unsigned long long)(ka >> 8) + (kb >> 8)) * 6ull >= a.sample_total)
        skipT = ((ka & 0xFFu) << 8) | (kb & 0xFFu);
    }
    const bool skip = SKIPCAP && skipT != 0xFFFFFFFFu;
    const bool marg = skip && a.img_hist != nullptr;  // skipped pixels come back through the image marginals
    const uint32_t a4 = ((skipT >> 8) & 0xFFu) * 0x01010101u, b4 = (skipT & 0xFFu) * 0x01010101u;
    uint32_t nboth = 0;

    for (int k = 0; k < nchunks; k++, kk++) {
      const int st = kk % kStages;
      const uint32_t off = (uint32_t)k * kChunk + (uint32_t)tid * PIX;
      uint32_t r[NW], w[NW];
      mbar_wait(&sm.full[st], (kk / kStages) & 1);
      load_words<NW>(r, sm.rbuf[st] + tid * PIX);
      load_words<NW>(w, sm.wbuf[st] + tid * PIX);
      const int nvalid = off >= npix ? 0 : (int)min((uint32_t)PIX, npix - off);
      if (nvalid == PIX && a.bg) {
        if (!skip) {
          accum_fast<P_U16G, SWZ, NW, SCHED>(sm, r, w, 0, warp);
        } else {
          uint32_t dr = 0, dw = 0;
#pragma unroll
          for (int j = 0; j < NW; j++) {
            dr |= r[j] ^ a4;
            dw |= w[j] ^ b4;
          }
          if (dr != 0u && dw != 0u) {
            accum_fast<P_U16G, SWZ, NW>(sm, r, w, 0, warp);
          } else if (marg) {
            // not counted at all: marginal_side_counts() reconstructs row a* and column b*
          } else if (dr == 0u && dw == 0u) {
            nboth += PIX;
          } else {
            const uint32_t (&x)[NW] = dr == 0u ? w : r;
            const uint32_t base = smem_u32(dr == 0u ? sk.n1[warp & (kSkipCopies - 1)] : sk.n2[warp & (kSkipCopies - 1)]);
#pragma unroll
            for (int i = 0; i < PIX; i++) {
              const uint32_t v = __byte_perm(x[i >> 2], 0u, 0x4440 + (i & 3));
              asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + v * 4u) : "memory");
            }
          }
        }
      } else if (nvalid > 0) {
        accum_slow<P_U16G, SWZ, NW>(sm, r, w, nvalid, a.bg, 0, warp);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[st]);
    }

    if (SKIPCAP && marg) {
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // every increment of this pair has landed
      if constexpr (SKIPCAP) marginal_side_counts<SWZ, NWARPS>(sm, sk, a, pr, skipT, warp, lane);
    } else if (skip) {  // fold the side tables into copy 0
      if (nboth) atomicAdd(&sk.nboth, nboth);
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
      for (int i = tid; i < 512; i += kConsumers) {
        uint32_t* t0 = i < 256 ? &sk.n1[0][i] : &sk.n2[0][i - 256];
        uint32_t sum = 0;
#pragma unroll
        for (int c = 0; c < kSkipCopies; c++) sum += t0[c * 256];
        *t0 = sum;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // B1: every increment of this pair has landed
    if constexpr (DUMP) {
      if (SWZ && !SKIPCAP && a.term_tab != nullptr) {
        if (dump) dump_joint_swizzled(sm, a.dumpJ, tid);  // CTA-uniform; the epilogue below is the timed one
        rows_epilogue_fast<true>(sm, L, a, warp, lane);
      } else {
        rows_epilogue<P_U16G, SWZ, NWARPS, true>(sm, 0, L, a, dump, warp, lane, SKIPCAP ? &sk : nullptr, skipT);
      }
    } else {  // the timed build: kept statement for statement as measured in round 1 (see DUMP above)
      if (SWZ && !SKIPCAP && !dump && a.term_tab != nullptr)
        rows_epilogue_fast<true>(sm, L, a, warp, lane);
      else
        rows_epilogue<P_U16G, SWZ, NWARPS, true>(sm, 0, L, a, dump, warp, lane, SKIPCAP ? &sk : nullptr, skipT);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // B2: rows read and cleared, marginals complete
    if (warp < 3) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int i = lane + 32 * k;
        v[k] = warp == 0 ? sm.rowE[i] : term_t(sm.term_tab, a.term_tab, warp == 1 ? sm.HA[i] : sm.HB[i], L);
      }
      const float s = tree_lanes<8>(v);
      if (lane == 0) sm.sums[warp] = s;
    } else {
      // what the next pair's pixel loop must find empty (the marginals are still being read)
      if (tid == 3 * 32) sm.ev_count = 0;
      if (tid >= 4 * 32 && tid < 4 * 32 + 8) sm.rowmask[tid - 4 * 32] = 0;
      if (SKIPCAP) {
        for (int i = tid - 3 * 32; i < 2 * kSkipCopies * 256; i += kConsumers - 3 * 32) (&sk.n1[0][0])[i] = 0;
        if (tid == 3 * 32) sk.nboth = 0;
      }
    }
    if (dump && tid < 256) {
      a.dumpHA[tid] = sm.HA[tid];
      a.dumpHB[tid] = sm.HB[tid];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // B3: totals done, marginals free again
    if (tid == 0)
      a.scores[a.out_index ? a.out_index[pi] : pi] = finish_score(sm.sums[1], sm.sums[2], sm.sums[0], a.mode);
    if (tid < 256) sm.HB[tid] = 0;  // the next epilogue's column sums start from zero (ordered by its B1)
  }
}

// ---- persistent build with the pixel ring staged through TENSOR MEMORY (variant 10) ---------
// 13 % of the LSU's shared-memory wavefronts in the kernel above only read the staged pixels
// (LDS.128).  Tensor memory has its own datapath: tcgen05.cp copies shared memory -> TMEM without
// the LSU and tcgen05.ld brings TMEM -> registers.  Here the TMA ring is unchanged, a second
// single-thread warp (the "issuer") forwards every landed stage with four tcgen05.cp.128x256b
// (4 KiB each: two per image) into a ring of 4 x 32 TMEM columns and commits to an mbarrier; the
// consumers read their 16 + 16 pixels with two tcgen05.ld.32x32b.x4 and never touch the staged
// bytes with the LSU.  Shared-memory descriptor: no swizzle, LBO 2048 B, SBO 128 B, for which lane t
// receives bytes [16 t, 16 t + 16) of the 4 KiB block in columns 0-3 and [2048 + 16 t, ...) in columns
// 4-7 (measured with tools/ubench_tmem.cu); a warp reads the lane quarter (warp % 4) and the column
// group (warp / 4), so every thread again owns 16 consecutive pixels of the chunk, the same ones
// in both images.  A stage of shared memory is free again as soon as its copies have committed.
// RESULT: bit-identical scores at every size, and 4.66 ms instead of 4.53 ms at C2.  The copy
// engine reads the same bytes out of the same shared-memory array the LDS would have: what bounds
// the kernel is that array's port, not the LSU in front of it -- LSU wavefronts (1172 M per launch)
// plus the TMA ring's writes (17 GB / 128 B = 133 M) already equal the SM's active cycles (1283 M).
// Kept as a tested variant, not the default.
constexpr int kTmemColsPerStage = 32;  // 16 render + 16 warp columns x 128 lanes x 4 B = 2 x 8192 B
static_assert(kChunk == 8192, "the TMEM staging is laid out for 8192-pixel chunks");
struct __align__(16) SmemTm {  // follows Smem
  unsigned long long tfull[kStages];   // copies of the stage into TMEM have completed (tcgen05.commit)
  unsigned long long tempty[kStages];  // all 16 consumer warps have read the TMEM stage
  uint32_t tbase;                      // TMEM base address from tcgen05.alloc
};
__device__ __forceinline__ uint64_t tm_desc(uint32_t saddr) {  // no swizzle, LBO 2048 B, SBO 128 B, version 1
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
         ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(16 * 32 + 64, 1)
joint_hist_score_tmem_kernel(const HistArgs a) {
  constexpr int NWARPS = 16;
  constexpr int kConsumers = NWARPS * 32;
  constexpr int kThreads = kConsumers + 64;
  constexpr int PIX = 16, NW = 4;
  constexpr bool SWZ = true;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  SmemTm& tm = *reinterpret_cast<SmemTm*>(smem_raw + sizeof(Smem));

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t npix = a.npix;
  const int nchunks = (int)((npix + kChunk - 1) / kChunk);
  const float L = (float)a.length;

  // ---- once per CTA ----
  {
    uint4* h4 = reinterpret_cast<uint4*>(sm.hist);
    for (int i = tid; i < kHistWords / 4; i += kThreads) h4[i] = make_uint4(0, 0, 0, 0);
    if (tid < 256) sm.HB[tid] = 0;
    if (tid < 8) sm.rowmask[tid] = 0;
    if (a.term_tab != nullptr)
      for (int i = tid; i < kTermTab && (uint32_t)i <= a.length; i += kThreads) sm.term_tab[i] = __ldg(a.term_tab + i);
    else
      for (int i = tid; i < kTermTab; i += kThreads) sm.term_tab[i] = term((uint32_t)i, L);
    if (tid == 0) {
      sm.ev_count = 0;
      for (int s = 0; s < kStages; s++) {
        mbar_init(&sm.full[s], 1);
        mbar_init(&tm.tfull[s], 1);
        mbar_init(&tm.tempty[s], NWARPS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == NWARPS + 1) {  // the issuer warp owns the TMEM allocation
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tm.tbase)),
                   "n"(kTmemColsPerStage * kStages)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tm.tbase;

  if (warp == NWARPS) {
    // ===== TMA producer: a shared-memory stage is free once its copies into TMEM have committed =====
    if (lane == 0) {
      int kk = 0;
      for (int pi = blockIdx.x; pi < a.npairs; pi += gridDim.x) {
        const int2 pr = a.pairs[pi];
        const uint8_t* rimg = a.renders + (size_t)pr.x * a.render_pitch;
        const uint8_t* wimg = a.warps + (size_t)pr.y * a.warp_pitch;
        for (int k = 0; k < nchunks; k++, kk++) {
          const int st = kk % kStages;
          if (kk >= kStages) mbar_wait(&tm.tfull[st], ((kk / kStages) - 1) & 1);
          const uint32_t off = (uint32_t)k * kChunk;
          uint32_t bytes = npix - off;
          bytes = bytes > (uint32_t)kChunk ? (uint32_t)kChunk : ((bytes + 15u) & ~15u);
          mbar_expect_tx(&sm.full[st], 2 * bytes);
          tma_load_1d(sm.rbuf[st], rimg + off, bytes, &sm.full[st]);
          tma_load_1d(sm.wbuf[st], wimg + off, bytes, &sm.full[st]);
        }
      }
    }
    return;
  }
  if (warp == NWARPS + 1) {
    // ===== issuer: shared memory -> TMEM, four 4 KiB copies per stage =====
    int total = 0;
    for (int pi = blockIdx.x; pi < a.npairs; pi += gridDim.x) total += nchunks;
    if (lane == 0) {
      for (int kk = 0; kk < total; kk++) {
        const int st = kk % kStages;
        mbar_wait(&sm.full[st], (kk / kStages) & 1);                                   // the TMA bytes have landed
        if (kk >= kStages) mbar_wait(&tm.tempty[st], ((kk / kStages) - 1) & 1);        // the TMEM stage has been read
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t col = tbase + (uint32_t)(st * kTmemColsPerStage);
        const uint32_t rs = smem_u32(sm.rbuf[st]), ws = smem_u32(sm.wbuf[st]);
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(col + 0u), "l"(tm_desc(rs)) : "memory");
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(col + 8u), "l"(tm_desc(rs + 4096u)) : "memory");
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(col + 16u), "l"(tm_desc(ws)) : "memory");
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(col + 24u), "l"(tm_desc(ws + 4096u)) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&tm.tfull[st]))
                     : "memory");
      }
    }
    __syncwarp();
    asm volatile("bar.sync 2, %0;" ::"n"(kConsumers + 32));  // every consumer has finished its last read
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(kTmemColsPerStage * kStages) : "memory");
    return;
  }

  // ===== consumers =====
  // this thread's 16 pixels of a chunk: TMEM lane 32 (warp % 4) + lane, column group warp / 4
  const uint32_t tlane = 32u * ((uint32_t)warp & 3u) + (uint32_t)lane, cgrp = (uint32_t)warp >> 2;
  const uint32_t pix0 = 4096u * (cgrp >> 1) + 2048u * (cgrp & 1u) + 16u * tlane;  // its offset inside the chunk
  const uint32_t taddr0 = tbase + ((32u * ((uint32_t)warp & 3u)) << 16) + 8u * (cgrp >> 1) + 4u * (cgrp & 1u);
  int kk = 0;
  for (int pi = blockIdx.x; pi < a.npairs; pi += gridDim.x) {
    const bool dump = a.dumpJ != nullptr && pi == a.dump_pair;
    for (int k = 0; k < nchunks; k++, kk++) {
      const int st = kk % kStages;
      const uint32_t off = (uint32_t)k * kChunk + pix0;
      uint32_t r[NW], w[NW];
      mbar_wait(&tm.tfull[st], (kk / kStages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = taddr0 + (uint32_t)(st * kTmemColsPerStage);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ta));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(ta + 16u));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (lane == 0) mbar_arrive(&tm.tempty[st]);  // the registers hold the pixels: the TMEM stage may be refilled
      const int nvalid = off >= npix ? 0 : (int)min((uint32_t)PIX, npix - off);
      if (nvalid == PIX && a.bg)
        accum_fast<P_U16G, SWZ, NW>(sm, r, w, 0, warp);
      else if (nvalid > 0)
        accum_slow<P_U16G, SWZ, NW>(sm, r, w, nvalid, a.bg, 0, warp);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // B1: every increment of this pair has landed
    if (a.term_tab != nullptr) {
      if (dump) dump_joint_swizzled(sm, a.dumpJ, tid);
      rows_epilogue_fast<true>(sm, L, a, warp, lane);
    } else {
      rows_epilogue<P_U16G, SWZ, NWARPS, true>(sm, 0, L, a, dump, warp, lane);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // B2
    if (warp < 3) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int i = lane + 32 * q;
        v[q] = warp == 0 ? sm.rowE[i] : term_t(sm.term_tab, a.term_tab, warp == 1 ? sm.HA[i] : sm.HB[i], L);
      }
      const float s = tree_lanes<8>(v);
      if (lane == 0) sm.sums[warp] = s;
    } else {
      if (tid == 3 * 32) sm.ev_count = 0;
      if (tid >= 4 * 32 && tid < 4 * 32 + 8) sm.rowmask[tid - 4 * 32] = 0;
    }
    if (dump && tid < 256) {
      a.dumpHA[tid] = sm.HA[tid];
      a.dumpHB[tid] = sm.HB[tid];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // B3
    if (tid == 0)
      a.scores[a.out_index ? a.out_index[pi] : pi] = finish_score(sm.sums[1], sm.sums[2], sm.sums[0], a.mode);
    if (tid < 256) sm.HB[tid] = 0;
  }
  asm volatile("bar.arrive 2, %0;" ::"n"(kConsumers + 32));  // lets the issuer warp free the TMEM columns
}


// ---- one evaluation on a CLUSTER of eight CTAs (the single-evaluation entry points) ------------------
// CUDAF::NMIWithCuda_noMask scores ONE (render, warp) pair per call (kernel.cu:49-114); with one CTA per
// evaluation that call used 1 of 148 SMs: 36 us for a 752x480 pair.  Here the pair is split over a
// thread-block cluster: CTA r takes the chunks r, r + 8, ... into its own packed-u16 partial histogram
// (the same TMA ring, hot loop, crossing / repay scheme and side tables as above), then the partials are
// reduced through DISTRIBUTED SHARED MEMORY: CTA r owns render levels 32 r .. 32 r + 31, sums the eight
// partial rows with ld.shared::cluster (uint4, sixteen threads per row), adds the repaid crossings and
// the side tables, turns the 256 counts of a row into entropy terms and walks the reference's pairwise
// tree (strides 128..16 between the sixteen threads of a row with shfl.xor, 8..1 inside a thread: pairing
// b with b ^ stride at every level, in the reference's level order -- the swizzle only relabels the leaves
// by an XOR, and fp32 addition is commutative, so the sums keep their bits).  CTA 0 finally collects the
// 256 row sums and the marginals over DSMEM and forms the score.  Integers, terms, trees and score are
// the ones every other build produces (same parity tests).
constexpr int kClusterCtas = 8;
#ifdef NMI_CLUSTER_TIMING  // experiment build: phase timestamps of cluster 0 / CTA 0 / thread 0 (globaltimer, ns)
#define NMI_CT(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); ct__[i] = t__; } } while (0)
#else
#define NMI_CT(i) do { } while (0)
#endif
__device__ __forceinline__ void im_count16(uint32_t* h, const uint4 v, bool valid);  // defined with image_mode_kernel below

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t ldc_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ldc_f32(uint32_t addr) { return __uint_as_float(ldc_u32(addr)); }

__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(16 * 32 + 32, 1)
joint_hist_score_cluster_kernel(const HistArgs a) {
  constexpr int NWARPS = 16;
  constexpr int kConsumers = NWARPS * 32;
  constexpr int kThreads = kConsumers + 32;
  constexpr int PIX = kChunk / kConsumers;  // 16 pixels per thread and chunk
  constexpr int NW = PIX / 4;
  constexpr bool SWZ = true;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  SmemSkip& sk = *reinterpret_cast<SmemSkip*>(smem_raw + sizeof(Smem));
  uint32_t* extra = reinterpret_cast<uint32_t*>(sm.rbuf);  // [32][256] u32 after the pixel loop: +4096 per repaid crossing
  unsigned long long* rx = reinterpret_cast<unsigned long long*>(smem_raw + sizeof(Smem) + sizeof(SmemSkip));  // reduce-add arrivals
  constexpr uint32_t kSliceBytes = 32u * 128u * 4u;  // the 32 render levels one CTA owns: 16 KiB of packed words
  static_assert(sizeof(sm.rbuf) >= 32 * 256 * sizeof(uint32_t), "the ring's render half holds the crossing table");

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
#ifdef NMI_CLUSTER_TIMING
  unsigned long long ct__[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  NMI_CT(0);
  const uint32_t rank = cluster_rank();
  const int pi = blockIdx.x / kClusterCtas;  // evaluation of this cluster
  const int2 pr = a.pairs[pi];
  const uint8_t* __restrict__ rimg = a.renders + (size_t)pr.x * a.render_pitch;
  const uint8_t* __restrict__ wimg = a.warps + (size_t)pr.y * a.warp_pitch;
  const uint32_t npix = a.npix;
  const int nchunks = (int)((npix + kChunk - 1) / kChunk);
  const int mine = nchunks > (int)rank ? (nchunks - (int)rank + kClusterCtas - 1) / kClusterCtas : 0;  // chunks of this CTA
  const bool dump = a.dumpJ != nullptr && pi == a.dump_pair;

  {
    uint4* h4 = reinterpret_cast<uint4*>(sm.hist);
    for (int i = tid; i < kHistWords / 4; i += kThreads) h4[i] = make_uint4(0, 0, 0, 0);
    if (tid < 256) sm.HB[tid] = 0;
    for (int i = tid; i < 2 * kSkipCopies * 256; i += kThreads) (&sk.n1[0][0])[i] = 0;
    if (tid == 0) sk.nboth = 0;
    for (int i = tid; i < kTermTab && (uint32_t)i <= a.length; i += kThreads) sm.term_tab[i] = __ldg(a.term_tab + i);
    if (tid == 0) {
      sm.ev_count = 0;
      for (int s = 0; s < kStages; s++) {
        mbar_init(&sm.full[s], 1);
        mbar_init(&sm.empty[s], NWARPS);
      }
      mbar_init(rx, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(rx, (kClusterCtas - 1) * kSliceBytes);  // seven partial slices will be added into this CTA's own
    }
  }
  // hot-bin skipping: the sampled most frequent grey level of either image decides (CTA-uniform).  Either from
  // image_mode_kernel (a.img_mode), or -- a single call must not pay two more launches for it -- sampled right
  // here: every CTA takes the same 1/64 of the pixels image_mode_kernel takes (a few KB), into two 256-bin tables
  // that live in sm.HA / sm.rowE until the reduce phase needs those arrays.
  uint32_t skipT = 0xFFFFFFFFu;
  if (a.img_mode != nullptr && a.bg && a.skip_mode != 0) {
    const uint32_t ka = __ldg(a.img_mode + pr.x), kb = __ldg(a.img_mode + a.nrenders + pr.y);
    if (a.skip_mode == 2 || ((unsigned long long)(ka >> 8) + (kb >> 8)) * 6ull >= a.sample_total)
      skipT = ((ka & 0xFFu) << 8) | (kb & 0xFFu);
  } else if (a.sample_in_kernel && a.bg && a.skip_mode != 0) {
    uint32_t* hr = sm.HA;
    uint32_t* hw = reinterpret_cast<uint32_t*>(sm.rowE);
    uint32_t* best = reinterpret_cast<uint32_t*>(sm.sums);  // [0] render, [1] warp: count << 8 | 255 - level
    if (tid < 256) hr[tid] = hw[tid] = 0;
    if (tid < 2) best[tid] = 0;
    __syncthreads();
    const uint32_t ngroups = npix / 16, nsamp = (ngroups + 63) / 64;
    const uint4* r4 = reinterpret_cast<const uint4*>(rimg);
    const uint4* w4 = reinterpret_cast<const uint4*>(wimg);
    for (uint32_t i0 = 0; i0 < nsamp; i0 += kThreads - 32) {  // warp-uniform trip count (im_count16 votes)
      const uint32_t i = i0 + (uint32_t)tid;
      const bool valid = tid < kConsumers && i < nsamp && (size_t)i * 64 < ngroups;
      const uint4 vr = valid ? __ldg(r4 + (size_t)i * 64) : make_uint4(0, 0, 0, 0);
      const uint4 vw = valid ? __ldg(w4 + (size_t)i * 64) : make_uint4(0, 0, 0, 0);
      im_count16(hr, vr, valid);
      im_count16(hw, vw, valid);
    }
    __syncthreads();
    if (tid < 256) {
      atomicMax(&best[0], (hr[tid] << 8) | (255u - (uint32_t)tid));  // largest count, lowest level among equals
      atomicMax(&best[1], (hw[tid] << 8) | (255u - (uint32_t)tid));
    }
    __syncthreads();
    const uint32_t ka = best[0], kb = best[1];
    const uint32_t total = nsamp * 16u;  // == image_mode_sample_total(npix)
    if (a.skip_mode == 2 || ((unsigned long long)(ka >> 8) + (kb >> 8)) * 6ull >= total)
      skipT = ((255u - (ka & 0xFFu)) << 8) | (255u - (kb & 0xFFu));
  }
  const bool skip = skipT != 0xFFFFFFFFu;
  const uint32_t a4 = ((skipT >> 8) & 0xFFu) * 0x01010101u, b4 = (skipT & 0xFFu) * 0x01010101u;
  __syncthreads();
  NMI_CT(1);  // init + mode sampling done

  if (warp == NWARPS) {
    // ===== producer warp: this CTA's chunks through the TMA ring =====
    if (lane == 0) {
      for (int kk = 0; kk < mine; kk++) {
        const int st = kk % kStages;
        if (kk >= kStages) mbar_wait(&sm.empty[st], ((kk / kStages) - 1) & 1);
        const uint32_t off = (uint32_t)(rank + (uint32_t)kk * kClusterCtas) * kChunk;
        uint32_t bytes = npix - off;
        bytes = bytes > (uint32_t)kChunk ? (uint32_t)kChunk : ((bytes + 15u) & ~15u);
        mbar_expect_tx(&sm.full[st], 2 * bytes);
        tma_load_1d(sm.rbuf[st], rimg + off, bytes, &sm.full[st]);
        tma_load_1d(sm.wbuf[st], wimg + off, bytes, &sm.full[st]);
      }
    }
    __syncwarp();  // the cluster barriers below are warp-aligned
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (this warp zeroed part of the histogram)
  } else {
    // ===== consumers: the pixel loop of the persistent build over this CTA's chunks =====
    uint32_t nboth = 0;
    for (int kk = 0; kk < mine; kk++) {
      const int st = kk % kStages;
      const uint32_t off = (uint32_t)(rank + (uint32_t)kk * kClusterCtas) * kChunk + (uint32_t)tid * PIX;
      uint32_t r[NW], w[NW];
      mbar_wait(&sm.full[st], (kk / kStages) & 1);
      load_words<NW>(r, sm.rbuf[st] + tid * PIX);
      load_words<NW>(w, sm.wbuf[st] + tid * PIX);
      const int nvalid = off >= npix ? 0 : (int)min((uint32_t)PIX, npix - off);
      if (nvalid == PIX && a.bg) {
        if (!skip) {
          accum_fast<P_U16G, SWZ, NW>(sm, r, w, 0, warp);
        } else {
          uint32_t dr = 0, dw = 0;
#pragma unroll
          for (int j = 0; j < NW; j++) {
            dr |= r[j] ^ a4;
            dw |= w[j] ^ b4;
          }
          if (dr != 0u && dw != 0u) {
            accum_fast<P_U16G, SWZ, NW>(sm, r, w, 0, warp);
          } else if (dr == 0u && dw == 0u) {
            nboth += PIX;
          } else {
            const uint32_t (&x)[NW] = dr == 0u ? w : r;
            const uint32_t base = smem_u32(dr == 0u ? sk.n1[warp & (kSkipCopies - 1)] : sk.n2[warp & (kSkipCopies - 1)]);
#pragma unroll
            for (int i = 0; i < PIX; i++) {
              const uint32_t v = __byte_perm(x[i >> 2], 0u, 0x4440 + (i & 3));
              asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + v * 4u) : "memory");
            }
          }
        }
      } else if (nvalid > 0) {
        accum_slow<P_U16G, SWZ, NW>(sm, r, w, nvalid, a.bg, 0, warp);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[st]);
    }
    if (skip) {  // fold the side tables into copy 0
      if (nboth) atomicAdd(&sk.nboth, nboth);
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
      for (int i = tid; i < 512; i += kConsumers) {
        uint32_t* t0 = i < 256 ? &sk.n1[0][i] : &sk.n2[0][i - 256];
        uint32_t sum = 0;
#pragma unroll
        for (int c = 0; c < kSkipCopies; c++) sum += t0[c * 256];
        *t0 = sum;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this thread's increments, before the bulk engine reads or adds to them
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));  // every increment of this CTA has landed, the ring is idle
    for (int i = tid; i < 32 * 256; i += kConsumers) extra[i] = 0;
    NMI_CT(2);  // pixel loop done
  }
  cluster_sync_all();  // C1: all eight partial histograms, event lists and side tables are complete

  NMI_CT(3);  // cluster barrier 1 passed
  if (warp < NWARPS) {
    // ---- the partial histograms: CTA r ADDS its slice of rows 32 c .. 32 c + 31 into CTA c's own copy of that slice
    //      with one bulk reduce per destination (cp.reduce.async.bulk .add.u32 through the cluster's shared-memory
    //      network, completion counted in bytes on the destination's mbarrier).  Packed fields are < 4096 once every
    //      crossing is repaid, so eight of them add without a carry into the neighbouring field, and the swizzle
    //      depends on the row alone: word for word the same layout in all eight CTAs.
    if (warp < kClusterCtas && (uint32_t)warp != rank && lane == 0) {
      const uint32_t dst_cta = (uint32_t)warp;
      const uint32_t src = smem_u32(sm.hist) + dst_cta * kSliceBytes;
      asm volatile(
          "cp.reduce.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes.add.u32 [%0], [%1], %2, [%3];" ::
              "r"(map_to_cta(sm.hist, dst_cta) + dst_cta * kSliceBytes),
          "r"(src), "r"(kSliceBytes), "r"(map_to_cta(rx, dst_cta))
          : "memory");
    }
    // ---- what is not in the packed words, into extra[local row][b] (u32): +4096 per repaid crossing of a row this
    //      CTA owns (all eight event lists), and the pixels the hot-bin side tables kept out of the histogram:
    //      n1 -> row a*, n2 -> column b*, nboth -> (a*, b*), summed over the eight CTAs.  Rolled loops, few threads.
    uint32_t nevs[kClusterCtas], nev_all = 0;
#pragma unroll
    for (uint32_t c = 0; c < (uint32_t)kClusterCtas; c++) nevs[c] = ldc_u32(map_to_cta(&sm.ev_count, c));  // eight loads in flight
#pragma unroll
    for (uint32_t c = 0; c < (uint32_t)kClusterCtas; c++) {
      nevs[c] = min(nevs[c], (uint32_t)kEvCap);
      nev_all += nevs[c];
    }
    if (nev_all != 0u) {  // CTA-uniform; most pairs have no bin above 4095 counts per CTA
#pragma unroll 1
      for (uint32_t c = 0; c < (uint32_t)kClusterCtas; c++) {
        const uint32_t evl = map_to_cta(sm.ev_list, c);
        for (uint32_t e = tid; e < nevs[c]; e += kConsumers) {
          const uint32_t word = ldc_u32(evl + (e >> 1) * 4u);  // two 16-bit entries per word
          const uint32_t t = (e & 1u) ? (word >> 16) : (word & 0xFFFFu);
          if ((t >> 13) == rank) atomicAdd(&extra[((t >> 8) & 31u) * 256u + (t & 0xFFu)], 4096u);
        }
      }
    }
    if (skip) {
      const uint32_t astar = (skipT >> 8) & 0xFFu, bstar = skipT & 0xFFu;
      if ((astar >> 5) == rank && tid < 256) {  // row a* lives here: its b-histogram of skipped pixels
        uint32_t n1 = 0;
#pragma unroll
        for (uint32_t q = 0; q < (uint32_t)kClusterCtas; q++) n1 += ldc_u32(map_to_cta(&sk.n1[0][tid], q));
        if (n1) atomicAdd(&extra[(astar & 31u) * 256u + (uint32_t)tid], n1);
      }
      if (tid >= 256 && tid < 288) {  // column b* of this CTA's 32 rows
        const uint32_t row = rank * 32u + (uint32_t)(tid - 256);
        uint32_t n2 = 0;
#pragma unroll
        for (uint32_t q = 0; q < (uint32_t)kClusterCtas; q++) n2 += ldc_u32(map_to_cta(&sk.n2[0][row], q));
        if (row == astar) {
#pragma unroll
          for (uint32_t q = 0; q < (uint32_t)kClusterCtas; q++) n2 += ldc_u32(map_to_cta(&sk.nboth, q));
        }
        if (n2) atomicAdd(&extra[(row & 31u) * 256u + bstar], n2);
      }
    }
    const bool use_extra = nev_all != 0u || skip;  // CTA-uniform
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers));
    // ---- row = 32 rank + tid / 16; the thread sums physical words 8u .. 8u + 7 of that row over the CTAs ----
    const uint32_t row = rank * 32u + ((uint32_t)tid >> 4), u = (uint32_t)tid & 15u;
    const uint32_t cr = row & 0x7Fu;                      // swizzle: logical word m = physical ^ cr
    const uint32_t pbase = ((row << 7) | (8u * u)) * 4u;  // byte offset of the thread's first physical word
    uint32_t cnt[16];
    mbar_wait(rx, 0);  // the seven other partial slices have been added into this CTA's own
    {
      const uint4* hp = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(sm.hist) + pbase);
      const uint4 v0 = hp[0], v1 = hp[1];
      const uint32_t wv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int j = 0; j < 8; j++) {
        cnt[2 * j] = wv[j] & 0xFFFFu;
        cnt[2 * j + 1] = wv[j] >> 16;
      }
    }
    uint32_t rs = 0;
    float v[16];
    const uint32_t* ex_row = extra + ((uint32_t)tid >> 4) * 256u;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t b = 2u * ((8u * u + (uint32_t)(i >> 1)) ^ cr) + (uint32_t)(i & 1);  // the camera level this register counts
      uint32_t c = cnt[i];
      if (use_extra) c += ex_row[b];
      rs += c;
      if (dump) a.dumpJ[row * 256u + b] = c;
      v[i] = term_lut(sm.term_tab, a.term_tab, c);  // tables only (the dispatcher requires the per-size table)
    }
    // the reference's tree over the 256 terms of the row (NMI.cu:270-287): strides 128, 64, 32, 16 pair the
    // sixteen threads of the row (bits 3..0 of the word group), 8, 4, 2, 1 pair registers (bits 3..1 of the word
    // inside the group, then the two halves of a word)
#pragma unroll
    for (int d = 8; d >= 1; d /= 2)
#pragma unroll
      for (int i = 0; i < 16; i++) v[i] = __fadd_rn(v[i], __shfl_xor_sync(0xffffffffu, v[i], d));
#pragma unroll
    for (int h = 8; h >= 1; h /= 2)
#pragma unroll
      for (int i = 0; i < h; i++) v[i] = __fadd_rn(v[i], v[i + h]);
#pragma unroll
    for (int d = 8; d >= 1; d /= 2) rs += __shfl_xor_sync(0xffffffffu, rs, d);
    if (u == 0u) {
      sm.rowE[row] = v[0];
      sm.HA[row] = rs;
    }
    {
      // this CTA's share of the camera marginal: column sums over its 32 rows.  Thread -> logical word tid & 127
      // (camera levels 2 m, 2 m + 1) of eight rows; consecutive lanes read consecutive (XOR-permuted) words.
      const uint32_t m = (uint32_t)tid & 127u, g = (uint32_t)tid >> 7;
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (uint32_t k = 0; k < 8u; k++) {
        const uint32_t lr = 8u * g + k, r = rank * 32u + lr;
        const uint32_t wv = sm.hist[(r << 7) | (m ^ (r & 0x7Fu))];
        lo += wv & 0xFFFFu;
        hi += wv >> 16;
        if (use_extra) {
          const uint2 e = *reinterpret_cast<const uint2*>(extra + lr * 256u + 2u * m);
          lo += e.x;
          hi += e.y;
        }
      }
      atomicAdd(&sm.HB[2u * m], lo);
      atomicAdd(&sm.HB[2u * m + 1u], hi);
    }
    NMI_CT(4);  // reduce + row trees done
  }
  cluster_sync_all();  // C2: every CTA's 32 row sums / render marginals and its share of the camera marginal are final

  NMI_CT(5);  // cluster barrier 2 passed
  if (rank == 0 && warp < 3) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t i = (uint32_t)lane + 32u * (uint32_t)k, owner = i >> 5;
      if (warp == 0) {
        v[k] = ldc_f32(map_to_cta(&sm.rowE[i], owner));
      } else if (warp == 1) {
        const uint32_t ha = ldc_u32(map_to_cta(&sm.HA[i], owner));
        v[k] = term_lut(sm.term_tab, a.term_tab, ha);
        if (dump) a.dumpHA[i] = ha;
      } else {
        uint32_t hb = 0;
#pragma unroll
        for (uint32_t c = 0; c < (uint32_t)kClusterCtas; c++) hb += ldc_u32(map_to_cta(&sm.HB[i], c));
        v[k] = term_lut(sm.term_tab, a.term_tab, hb);
        if (dump) a.dumpHB[i] = hb;
      }
    }
    const float s = tree_lanes<8>(v);
    if (lane == 0) sm.sums[warp] = s;
  }
  __syncthreads();
  if (rank == 0 && tid == 0)
    a.scores[a.out_index ? a.out_index[pi] : pi] = finish_score(sm.sums[1], sm.sums[2], sm.sums[0], a.mode);
  NMI_CT(6);
  cluster_sync_all();  // C3: nobody leaves while CTA 0 may still be reading its shared memory
#ifdef NMI_CLUSTER_TIMING
  NMI_CT(7);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    printf("[cluster ns] init+sample %llu | pixels %llu | C1 %llu | reduce %llu | C2 %llu | final %llu | C3 %llu | total %llu\n",
           ct__[1] - ct__[0], ct__[2] - ct__[1], ct__[3] - ct__[2], ct__[4] - ct__[3], ct__[5] - ct__[4], ct__[6] - ct__[5],
           ct__[7] - ct__[6], ct__[7] - ct__[0]);
#endif
}

// e(c) for every count 0..length (counts cannot exceed the pixel count)
__global__ void term_table_kernel(float* __restrict__ tab, uint32_t length) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= length) tab[i] = term(i, (float)length);
}

// ---- sampled per-image mode (hot-bin skipping) ----------------------------------------
// One CTA per image: histogram of every 64th group of 16 pixels, then
// img_mode[img] = sampled count of the most frequent level << 8 | that level (lowest level
// among equals) and hot[0] / hot[1] = the largest such count over the renders / the warps
// (read back by the host one search later to pick the SKIPCAP build).
constexpr int kImThreads = 256;

// 16 pixels (one uint4) into the warp's private histogram.  Flat runs -- the very images
// this pass exists for -- would hammer one word, so 16 equal pixels are one add of 16, and
// a warp whose lanes all hold the same flat value adds 512 once.
__device__ __forceinline__ void im_count16(uint32_t* h, const uint4 v, bool valid) {
  const uint32_t b0 = v.x & 0xFFu;
  const bool flat = valid && v.x == b0 * 0x01010101u && v.y == v.x && v.z == v.x && v.w == v.x;
  const uint32_t first = __shfl_sync(0xffffffffu, b0, 0);
  if (__all_sync(0xffffffffu, flat && b0 == first)) {
    if ((threadIdx.x & 31) == 0) atomicAdd(h + b0, 512u);
    return;
  }
  if (!valid) return;
  if (flat) {
    atomicAdd(h + b0, 16u);
    return;
  }
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int k = 0; k < 4; k++) atomicAdd(h + ((w[j] >> (8 * k)) & 0xFFu), 1u);
}

__global__ void __launch_bounds__(kImThreads)
image_mode_kernel(const uint8_t* __restrict__ renders, size_t rpitch, int nr,
                  const uint8_t* __restrict__ warps, size_t wpitch, uint32_t npix,
                  uint32_t* __restrict__ img_mode, uint32_t* __restrict__ hot) {
  __shared__ uint32_t s_h[kImThreads / 32][256];
  __shared__ uint32_t s_best;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (kImThreads / 32) * 256; i += kImThreads) (&s_h[0][0])[i] = 0;
  if (tid == 0) s_best = 0;
  __syncthreads();
  const int img = blockIdx.x;
  const uint4* img4 = reinterpret_cast<const uint4*>(
      img < nr ? renders + (size_t)img * rpitch : warps + (size_t)(img - nr) * wpitch);
  const uint32_t nsamp = (npix / 16 + 63) / 64;
  for (uint32_t i0 = 0; i0 < nsamp; i0 += 4 * kImThreads) {  // warp-uniform trip count, 4 loads in flight
    uint4 v[4];
    bool valid[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t i = i0 + u * kImThreads + tid;
      valid[u] = i < nsamp && (size_t)i * 64 < npix / 16;
      v[u] = valid[u] ? __ldg(img4 + (size_t)i * 64) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) im_count16(s_h[warp], v[u], valid[u]);
  }
  __syncthreads();
  uint32_t c = 0;
  for (int w = 0; w < kImThreads / 32; w++) c += s_h[w][tid];
  atomicMax(&s_best, (c << 8) | (255u - (uint32_t)tid));  // largest count, lowest level among equals
  __syncthreads();
  if (tid == 0) {
    const uint32_t cnt = s_best >> 8, level = 255u - (s_best & 0xFFu);
    img_mode[img] = (cnt << 8) | level;
    atomicMax(hot + (img < nr ? 0 : 1), cnt);
  }
}

// ---- full per-image histograms (the marginals hot-bin skipping reconstructs from) ----------------
// grid (image, part): 256-bin histogram of ALL npix pixels of every render and every warp, per-warp
// private tables with the flat-run shortcuts of im_count16, merged with one global atomic per bin and
// CTA.  img_hist must be zero at launch.
constexpr int kIhParts = 8;
__global__ void __launch_bounds__(kImThreads)
image_hist_kernel(const uint8_t* __restrict__ renders, size_t rpitch, int nr,
                  const uint8_t* __restrict__ warps, size_t wpitch, uint32_t npix,
                  uint32_t* __restrict__ img_hist) {
  __shared__ uint32_t s_h[kImThreads / 32][256];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (kImThreads / 32) * 256; i += kImThreads) (&s_h[0][0])[i] = 0;
  __syncthreads();
  const int img = blockIdx.x;
  const uint8_t* base = img < nr ? renders + (size_t)img * rpitch : warps + (size_t)(img - nr) * wpitch;
  const uint4* img4 = reinterpret_cast<const uint4*>(base);
  const uint32_t ngroups = npix / 16;  // whole groups of 16 pixels; the tail is counted bytewise below
  const uint32_t per = (ngroups + kIhParts - 1) / kIhParts;
  const uint32_t g0 = blockIdx.y * per, g1 = min(ngroups, g0 + per);
  for (uint32_t i0 = g0; i0 < g1; i0 += 4 * kImThreads) {  // warp-uniform trip count, 4 loads in flight
    uint4 v[4];
    bool valid[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t i = i0 + u * kImThreads + tid;
      valid[u] = i < g1;
      v[u] = valid[u] ? __ldg(img4 + i) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) im_count16(s_h[warp], v[u], valid[u]);
  }
  if (blockIdx.y == 0 && tid < (int)(npix - ngroups * 16)) atomicAdd(&s_h[0][base[ngroups * 16 + tid]], 1u);
  __syncthreads();
  uint32_t c = 0;
  for (int w = 0; w < kImThreads / 32; w++) c += s_h[w][tid];
  if (c) atomicAdd(img_hist + 256 * (size_t)img + tid, c);
}

template <int POLICY, bool USE_TMA, int NWARPS, bool SWZ, bool SKIPCAP = false, bool FASTEP = false, int LDGD = 1>
int launch_t(const HistArgs& a, cudaStream_t st) {
  auto kern = joint_hist_score_kernel<POLICY, USE_TMA, NWARPS, SWZ, SKIPCAP, FASTEP, LDGD>;
  constexpr size_t smem = sizeof(Smem) + (SKIPCAP ? sizeof(SmemSkip) : 0);
  static bool configured[64] = {false};  // the attribute is per device (contexts of several GPUs in one process)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return -1;
    configured[dev] = true;
  }
  kern<<<a.npairs, NWARPS == 32 ? 1024 : NWARPS * 32 + 32, smem, st>>>(a);
  return 1;
}

template <bool SWZ, bool SKIPCAP, bool DUMP = false, int SCHED = 0>
int launch_persistent(const HistArgs& a, cudaStream_t st) {
  auto kern = joint_hist_score_persistent_kernel<SWZ, SKIPCAP, DUMP, SCHED>;
  constexpr size_t smem = sizeof(Smem) + (SKIPCAP ? sizeof(SmemSkip) : 0);
  static int sms[64] = {0};  // per device: one CTA per SM (the histogram takes most of an SM's shared memory)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return -1;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return -1;
    sms[dev] = n;
  }
  const int nsm = sms[dev];
  kern<<<a.npairs < nsm ? a.npairs : nsm, 16 * 32 + 32, smem, st>>>(a);
  return 1;
}

int launch_tmem(const HistArgs& a, cudaStream_t st) {
  auto kern = joint_hist_score_tmem_kernel;
  constexpr size_t smem = sizeof(Smem) + sizeof(SmemTm);
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return -1;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return -1;
    sms[dev] = n;
  }
  const int nsm = sms[dev];
  kern<<<a.npairs < nsm ? a.npairs : nsm, 16 * 32 + 64, smem, st>>>(a);
  return 1;
}

int launch_cluster(const HistArgs& a, cudaStream_t st) {
  auto kern = joint_hist_score_cluster_kernel;
  constexpr size_t smem = sizeof(Smem) + sizeof(SmemSkip) + 16;  // + the mbarrier the partial slices arrive on
  static bool configured[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured[dev] = true;
  }
  kern<<<a.npairs * kClusterCtas, 16 * 32 + 32, smem, st>>>(a);
  return 1;
}

}  // namespace

void launch_term_table(float* tab, uint32_t length, cudaStream_t st) {
  term_table_kernel<<<(length + 256) / 256, 256, 0, st>>>(tab, length);
}

uint32_t image_mode_sample_total(uint32_t npix) { return ((npix / 16 + 63) / 64) * 16; }

int launch_image_modes(const uint8_t* renders, size_t rpitch, int nr, const uint8_t* warps, size_t wpitch,
                       int nw, uint32_t npix, uint32_t* img_mode, uint32_t* hot, cudaStream_t st, bool clear_hot) {
  if (nr + nw == 0) return 0;
  if (clear_hot) cudaMemsetAsync(hot, 0, 2 * sizeof(uint32_t), st);
  prefer_max_shared((const void*)image_mode_kernel);
  image_mode_kernel<<<nr + nw, kImThreads, 0, st>>>(renders, rpitch, nr, warps, wpitch, npix, img_mode, hot);
  return 1;
}

int launch_image_hists(const uint8_t* renders, size_t rpitch, int nr, const uint8_t* warps, size_t wpitch,
                       int nw, uint32_t npix, uint32_t* img_hist, cudaStream_t st) {
  if (nr + nw == 0) return 0;
  cudaMemsetAsync(img_hist, 0, (size_t)(nr + nw) * 256 * sizeof(uint32_t), st);
  prefer_max_shared((const void*)image_hist_kernel);
  image_hist_kernel<<<dim3((unsigned)(nr + nw), kIhParts), kImThreads, 0, st>>>(renders, rpitch, nr, warps, wpitch, npix,
                                                                              img_hist);
  return 1;
}

// a handful of evaluations (the per-pair drop-in, tiny grids): one 8-CTA cluster per evaluation instead of
// one CTA -- $NMI_EVAL_CLUSTER=0 keeps the one-CTA builds (A/B switch)
bool hist_uses_cluster(const HistArgs& a) {
  static const bool use_cluster = [] {
    const char* e = getenv("NMI_EVAL_CLUSTER");
    return !(e && atoi(e) == 0);
  }();
  return use_cluster && !a.force_batched && a.bins == 256 && a.variant == 0 && a.npairs >= 1 &&
         a.npairs * kClusterCtas <= 64 && a.term_tab != nullptr && a.npix >= (uint32_t)(2 * kChunk);
}

int launch_joint_hist_score(const HistArgs& a, cudaStream_t st) {
  if (a.npairs <= 0) return 0;
  if (hist_uses_cluster(a)) return launch_cluster(a, st);
  if (a.bins == 256 && a.skipcap && a.bg && a.img_mode != nullptr) {
    // with the image marginals the skipped pixels cost nothing and the pairs of a search differ less in cost than
    // with the side tables: the persistent CTAs' static schedule does ($NMI_SKIP_PERSISTENT=0: one CTA per pair)
    static const bool skip_persistent = [] {
      const char* e = getenv("NMI_SKIP_PERSISTENT");
      return !(e && atoi(e) == 0);
    }();
    if (a.variant == 0 && a.img_hist != nullptr && a.dumpJ == nullptr && skip_persistent)
      return launch_persistent<true, true>(a, st);
    switch (a.variant) {  // the packed-u16 builds that carry the side tables
      case 0: case 8: case 9: return launch_t<P_U16G, true, 16, true, true>(a, st);
      case 4: return launch_t<P_U16G, true, 32, false, true>(a, st);
      case 5: return launch_t<P_U16G, true, 16, false, true>(a, st);
      default: break;
    }
  }
  if (a.bins == 64)
    return (a.variant & 1) ? launch_t<P_B64, false, 16, false>(a, st)
                           : launch_t<P_B64, true, 16, false>(a, st);
  switch (a.variant) {
    case 1: return launch_t<P_U16G, false, 16, false>(a, st);
    case 2: return launch_t<P_U32X2, true, 16, false>(a, st);
    case 3: return launch_t<P_U32X2, false, 16, false>(a, st);
    case 4: return launch_t<P_U16G, true, 32, false>(a, st);
    case 5: return launch_t<P_U16G, true, 16, false>(a, st);
    case 6: return launch_t<P_U16G, true, 32, true>(a, st);
    case 7: return launch_t<P_U32X2, true, 32, false>(a, st);
    case 8: return launch_t<P_U16G, true, 16, true>(a, st);  // variant 0's configuration: one CTA per pair, warp-per-row epilogue
    case 9:  // one CTA per pair, fast epilogue
      return a.term_tab != nullptr ? launch_t<P_U16G, true, 16, true, false, true>(a, st)
                                   : launch_t<P_U16G, true, 16, true>(a, st);
    case 10: return launch_tmem(a, st);  // variant 0 with the pixel ring staged through tensor memory
    // variant 9 without any shared-memory staging: ld.global.nc into registers, three chunks in flight per
    // thread.  Measured at C2 (round 2): 4.82 / 4.83 / 4.67 ms with 1 / 2 / 3 chunks in flight against 4.42
    // for variant 9 -- a global load costs the L1 pipe more than the TMA write + LDS it replaces.  Kept as a
    // tested variant (the deepest one), not a default.
    case 11: return a.term_tab != nullptr ? launch_t<P_U16G, false, 16, true, false, true, 3>(a, st) : launch_t<P_U16G, false, 16, true>(a, st);
    default:  // variant 0: TMA ring, bank swizzle, persistent CTAs with the fast epilogue
      return a.dumpJ != nullptr ? launch_persistent<true, false, true>(a, st) : launch_persistent<true, false>(a, st);
  }
}

}  // namespace nmi
