// tcgen05 / TMEM encoder self-attention for passages of up to 256 tokens (bf16, d_kv = 64).
// (NKB = 1: passages of <= 128 tokens, the shipped datasets; NKB = 2: up to 256 tokens = two 128-key blocks and up to two
//  128-row query blocks per passage, BASELINE.json configs[4])
//
// Reference semantics: per-passage bidirectional T5 self-attention, scores = Q K^T (no 1/sqrt(d) scale) + relative
// position bias (layer-0 table shared by all layers) + key padding mask, fp32 softmax, P V
// (reference src/model/gram_t5_modeling.py:572-621, bias :452-477 and :1249, mask :1130).
//
// Work item = one (passage, head); a persistent CTA (two per SM) walks its items, see the comment at the kernel:
//   TMA      three 128-row x 64-col boxes (Q, K, V of this head) straight out of the packed qkv activation, 128B swizzle,
//            into a 2-stage ring
//   MMA 1    S[128 q x 128 keys] = Q K^T : 4 x tcgen05.mma (M=128, N=128, K=16), both operands K-major, fp32 in one of
//            two TMEM accumulators
//   softmax  thread r owns query row r (TMEM lane r): tcgen05.ld of its 128 scores, bias LUT, mask, max, exp2, sum --
//            a row-wise softmax with no cross-thread traffic -- and writes P (bf16) into shared memory in the K-major
//            128B-swizzled operand layout, over the dead Q and K tiles of the stage
//   MMA 2    O[128 q x 64 d] = P V : 8 x tcgen05.mma (M=128, N=64, K=16), A = P (K-major), B = V read in place in its
//            MN-major (d-contiguous) layout; O re-uses the TMEM columns of S
//   epilogue tcgen05.ld of the 64 outputs, 1/sum, bf16, staged through the dead P rows to coalesced 128-byte stores
// Rows past the passage length in the boxes belong to the next passage (finite values): masked as keys, and not
// stored as queries.
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <map>
#include <tuple>

#include "common.cuh"
#include "kernels.h"

namespace gram {
namespace ta {

constexpr int DK = 64, LQ = 128;
constexpr uint32_t BOX = LQ * DK * 2;      // 16 KiB
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TA_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TA_WAIT_DONE;\n"
      "bra TA_WAIT_LOOP;\n"
      "TA_WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, 128B swizzle, 8-row (K-major) / 8-k (MN-major) groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                   // leading byte offset: unused (one 64-element block in the leading mode)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset = 1024 B
  d |= (uint64_t)1 << 46;                   // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}
// D = f32, A = B = bf16, M = 128
constexpr uint32_t IDESC_QK = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);              // N=128, B K-major
constexpr uint32_t IDESC_PV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // N=64, B MN-major

constexpr int CONSUMERS = 128;            // warps 0-3: softmax + epilogue, thread r = query row r = TMEM lane r
constexpr int THREADS = CONSUMERS + 32;   // warp 4: one elected thread issues every TMA load and every tcgen05.mma
constexpr int NSTAGE = 2;
// Stage of one work item with NKB key blocks: Q | K_0 .. K_{NKB-1} | (pad) | V_0 .. V_{NKB-1}.  P (128 rows x 128*NKB keys,
// bf16, K-major boxes of 64 keys) later overwrites everything in front of V: Q | K (NKB = 1), Q | K_0 | K_1 | pad (NKB = 2).
template <int NKB> struct Geo {
  static constexpr uint32_t P_BYTES = 2u * NKB * BOX;              // 32 / 64 KiB
  static constexpr uint32_t V_OFF = P_BYTES;
  static constexpr uint32_t STAGE = P_BYTES + NKB * BOX;           // 48 / 96 KiB
  static constexpr int TMEM_COLS = 2 * 128 * NKB;                  // two S/O accumulators of 128*NKB columns
  static constexpr int ACC = 128 * NKB;
};
enum { B_FULL = 0, B_EMPTY = 2, B_SFULL = 4, B_SFREE = 6, B_PREADY = 8, B_OFULL = 10, B_COUNT = 12 };

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Persistent, warp-specialised kernel: CTA b walks passages b, b + grid, ... and the H heads of each; two CTAs per SM.
// Work item k of a CTA (the k-th (passage, head) it processes) uses shared-memory stage k & 1 and TMEM accumulator
// k & 1; its barrier parity is (k >> 1) & 1.
//   producer thread     TMA(k) when stage free (EMPTY: epilogue of item k-2 has left the stage)  -> FULL
//                       MMA 1(k) when FULL and the accumulator is drained (SFREE, item k-2)      -> SFULL
//                       MMA 2(k) when the softmax warps have written P (PREADY)                  -> OFULL
//                       the three are polled (mbarrier.test_wait) so that whichever becomes possible first is issued:
//                       the loads and S = Q K^T of item k+1 run under the softmax of item k
//   softmax warps       pass 2 of item k (exp2, row sum, P -> shared), arrive PREADY; pass 1 (row maximum) of item
//                       k+1 while the tensor core computes O(k); wait OFULL, read O, arrive SFREE; O rows through the
//                       dead P rows of the stage to coalesced 128-byte stores; arrive EMPTY
// The non-persistent version (one CTA per item, four CTAs per SM) spent 5.9 us per item on a serial chain of CTA start,
// TMEM allocation, LUT load, TMA latency, MMA, softmax, MMA, store for 0.4 us of issue work (ncu: 28 % issue
// utilisation, long-scoreboard stalls).
// NKB = 2 (passages of 129-256 tokens): a work item is (passage, 128-row query block, head); it loads the query block's Q
// rows and the passage's key blocks that hold tokens (one or two), S is 128 x 128*nkb, P covers 128*nkb keys; the
// 96 KiB stages and 512 TMEM columns leave room for one CTA per SM.
template <int NKB>
__global__ void __launch_bounds__(THREADS, NKB == 1 ? 2 : 1)
enc_attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, bf16* __restrict__ out,
                        const int* __restrict__ plen, const int* __restrict__ poff,
                        const uint8_t* __restrict__ tok_valid, const float* __restrict__ bias_lut, int Lb, int H, int P) {
  using G = Geo<NKB>;
  constexpr uint32_t STAGE = G::STAGE;
  constexpr uint32_t ACC = (uint32_t)G::ACC;
  constexpr int NCH = 4 * NKB;                                                  // 32-key chunks of a score row
  const int HD = H * DK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int lut_n = 2 * Lb - 1;
  const uint32_t lut_bytes = (uint32_t)(((size_t)H * lut_n * 4 + 15) / 16 * 16);
  uint8_t* aux = smem + NSTAGE * STAGE;
  float* lut = reinterpret_cast<float*>(aux);                                  // [H][lut_n], log2 units
  float* bhi = reinterpret_cast<float*>(aux + lut_bytes);                      // [H] largest LUT entry of the head
  float* blo = bhi + 32;                                                       // [H] smallest
  uint32_t* masks = reinterpret_cast<uint32_t*>(aux + lut_bytes + 256);        // [8] key visibility of the current passage
  const uint32_t bars = base + NSTAGE * STAGE + lut_bytes + 256 + 32;          // B_COUNT mbarriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + lut_bytes + 256 + 32 + 8 * B_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
    for (int i = 0; i < B_COUNT; ++i)
      mbar_init(bar(i), ((i >= B_SFREE && i < B_OFULL) || (i >= B_EMPTY && i < B_SFULL)) ? CONSUMERS : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(G::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // bias LUTs of every head in log2 units (the softmax runs on exp2), and each head's range
  for (int i = tid; i < H * lut_n; i += THREADS) lut[i] = bias_lut[i] * LOG2E;
  __syncthreads();
  for (int hh = warp; hh < H; hh += THREADS / 32) {
    float bmax = -INFINITY, bmin = INFINITY;
    for (int i = lane; i < lut_n; i += 32) {
      const float b = lut[hh * lut_n + i];
      bmax = fmaxf(bmax, b);
      bmin = fminf(bmin, b);
    }
    bmax = warp_max(bmax);
    bmin = -warp_max(-bmin);
    if (lane == 0) { bhi[hh] = bmax; blo[hh] = bmin; }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ===================== producer: TMA + MMA issue =====================
    if (lane == 0) {
      // items of this CTA in the order (passage, query block, head); nkb = key blocks of the passage that hold tokens
      int p_it = (int)blockIdx.x - (int)gridDim.x, h_it = H - 1, qb_it = 0, nqb_it = 1, row0_it = 0, len_it = 0;
      bool tma_done = false;
      auto advance = [&]() {
        if (++h_it < H) return;
        h_it = 0;
        if (++qb_it < nqb_it) return;
        qb_it = 0;
        for (;;) {
          p_it += (int)gridDim.x;
          if (p_it >= P) { tma_done = true; return; }
          const int l = plen[p_it];
          if (l > 0) { row0_it = poff[p_it]; nqb_it = (l + LQ - 1) / LQ; len_it = l; return; }
        }
      };
      advance();
      uint32_t k_tma = 0, k_m1 = 0, k_m2 = 0;              // items loaded / S issued / O issued
      uint32_t nkb_of[NSTAGE] = {1u, 1u};                  // key blocks of the item in each stage
      uint32_t keys_of[NSTAGE] = {128u, 128u};             // keys of the item's passage, rounded up to the MMA's 16
      while (!(tma_done && k_m2 == k_tma)) {
        if (!tma_done && k_tma < k_m2 + NSTAGE) {
          const uint32_t st = k_tma & 1u, ph = (k_tma >> 1) & 1u;
          if (mbar_test(bar(B_EMPTY + st), ph ^ 1u)) {
            const uint32_t sb = base + st * STAGE;
            const uint32_t nkb = (uint32_t)nqb_it;         // = ceil(len / 128)
            nkb_of[st] = nkb;
            keys_of[st] = (uint32_t)((len_it + 15) & ~15);
            mbar_arrive_expect_tx(bar(B_FULL + st), BOX * (1u + 2u * nkb));
            tma_load_2d(sb, &map_qkv, bar(B_FULL + st), h_it * DK, row0_it + qb_it * LQ);
            for (uint32_t kb = 0; kb < nkb; ++kb) {
              tma_load_2d(sb + BOX * (1u + kb), &map_qkv, bar(B_FULL + st), HD + h_it * DK, row0_it + (int)kb * LQ);
              tma_load_2d(sb + G::V_OFF + BOX * kb, &map_qkv, bar(B_FULL + st), 2 * HD + h_it * DK, row0_it + (int)kb * LQ);
            }
            ++k_tma;
            advance();
          }
        }
        if (k_m1 < k_tma) {
          const uint32_t st = k_m1 & 1u, ph = (k_m1 >> 1) & 1u;
          if (mbar_test(bar(B_FULL + st), ph) && mbar_test(bar(B_SFREE + st), ph ^ 1u)) {
            tcgen05_fence_after();
            const uint32_t sb = base + st * STAGE;
            const uint64_t dq = make_desc(sb), dk = make_desc(sb + BOX);     // key blocks are contiguous 8-row groups
            // S only as wide as the passage (N = keys rounded up to 16): columns past it are never read unmasked, and
            // every score is the same dot product whatever N is
            const uint32_t idesc = (IDESC_QK & ~(0x3Fu << 17)) | ((keys_of[st] >> 3) << 17);
#pragma unroll
            for (int kk = 0; kk < DK / 16; ++kk)
              umma_bf16(tmem + st * ACC, dq + (uint64_t)(kk * 2), dk + (uint64_t)(kk * 2), idesc, kk ? 1u : 0u);
            umma_commit(bar(B_SFULL + st));
            ++k_m1;
          }
        }
        if (k_m2 < k_m1) {
          const uint32_t st = k_m2 & 1u, ph = (k_m2 >> 1) & 1u;
          if (mbar_test(bar(B_PREADY + st), ph)) {
            tcgen05_fence_after();
            const uint32_t sb = base + st * STAGE;
            // 16-key steps over the passage's keys only: the steps past them would add P = 0 times (finite) rows of the
            // next passage, i.e. exact zeros -- leaving them out changes no bit
            const int nks = (int)(keys_of[st] >> 4);
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t dp = make_desc(sb + (uint32_t)(ks >> 2) * BOX + (uint32_t)(ks & 3) * 32u);
              const uint64_t dv = make_desc(sb + G::V_OFF + (uint32_t)ks * 2048u);
              umma_bf16(tmem + st * ACC, dp, dv, IDESC_PV, ks ? 1u : 0u);      // O re-uses the (fully read) S columns
            }
            umma_commit(bar(B_OFULL + st));
            ++k_m2;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue: thread r owns query row r of the block =====================
    const int r = tid;
    const uint32_t tlane = (uint32_t)(warp * 32) << 16;
    // ---- iteration over this CTA's items; the key-visibility masks are rebuilt when the passage changes ----
    int p = (int)blockIdx.x - (int)gridDim.x, h = H - 1, qb = 0, nqb = 1, len = 0, row0 = 0;
    uint32_t mkc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) mkc[c] = 0u;
    auto next_item = [&]() -> bool {
      if (++h < H) return true;
      h = 0;
      if (++qb < nqb) return true;
      qb = 0;
      for (;;) {
        p += (int)gridDim.x;
        if (p >= P) return false;
        len = plen[p];
        if (len > 0) break;
      }
      row0 = poff[p];
      nqb = (len + LQ - 1) / LQ;
      unsigned m[NKB];
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        const int key = kb * LQ + r;
        const bool vis = key < len && tok_valid[row0 + key] != 0;
        m[kb] = __ballot_sync(0xffffffffu, vis);
      }
      consumer_bar();                                    // every warp is done with the previous passage's masks
      if (lane == 0) {
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) masks[kb * 4 + warp] = m[kb];
      }
      consumer_bar();
#pragma unroll
      for (int c = 0; c < NCH; ++c) mkc[c] = masks[c];
      return true;
    };
    // Pass 1 of item k (head hh): an upper bound of the row maximum in log2 units.  Softmax is invariant to the shift
    // as long as nothing overflows, so the exact maximum is not needed: for chunks whose 32 keys are all visible it
    // takes max(raw s) -- one FMNMX per score -- and bounds the bias by the head's largest LUT entry; the exact
    // bias + mask walk is kept for the passage's tail chunk, and for every chunk when the bias table spans more than
    // 2^64 (then the bound could push small terms into underflow).
    // A warp whose 32 query rows all lie past the passage (a 46-token passage leaves warps 2 and 3 without rows) keeps the
    // barrier protocol -- it waits and arrives exactly like the others, so it can never run a phase ahead -- but reads no
    // scores, writes no P rows (the stale bf16 there is finite; its O rows are never stored) and stores nothing.
    auto pass1 = [&](uint32_t k, int hh, int qrow) -> float {
      const uint32_t st = k & 1u, ph = (k >> 1) & 1u;
      const uint32_t trow = tmem + st * ACC + tlane;
      const float* lrow = lut + hh * lut_n + (Lb - 1 - min(qrow, Lb - 1));     // rows past the table are rows past the passage
      const float bias_hi = bhi[hh];
      const bool exact = !(bias_hi - blo[hh] <= 64.f);
      mbar_wait(bar(B_SFULL + st), ph);
      tcgen05_fence_after();
      if (qrow - r + warp * 32 >= len) return 0.f;       // no row of this warp belongs to the passage (warp-uniform)
      float mraw = -INFINITY, mex = -INFINITY;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const uint32_t mk = mkc[c];
        if (mk == 0u) continue;                          // chunk past the passage (uniform across the CTA)
        uint32_t v[32];
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
        if (mk == 0xffffffffu && !exact) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if ((mk >> j) & 1u) mex = fmaxf(mex, fmaf(__uint_as_float(v[j]), LOG2E, lrow[c * 32 + j]));
        }
      }
      return fmaxf(fmaf(mraw, LOG2E, bias_hi), mex);     // at least one key is visible (len > 0)
    };

    uint32_t k = 0;
    bool have = next_item();
    float mb = have ? pass1(0u, h, qb * LQ + r) : 0.f;
    while (have) {
      const uint32_t st = k & 1u, ph = (k >> 1) & 1u;
      const uint32_t trow = tmem + st * ACC + tlane;
      const uint32_t sQ = base + st * STAGE;
      const int nch = (len + 31) >> 5;                   // 32-key chunks that hold keys of the passage (P past them is not multiplied)
      // bias2 of (query q, key j) = lut[h][j - q + Lb - 1]; only read for j < len <= Lb
      const float* lrow = lut + h * lut_n + (Lb - 1 - min(qb * LQ + r, Lb - 1));
      // ---- pass 2: p = 2^(s*log2e + bias2 - mb), row sum, P (bf16) into the K-major 128B-swizzled operand layout:
      //      box b = keys [64b, 64b+64), row r, 16-byte chunk j ^ (r & 7) ----
      float sum = 0.f;
      const bool warp_live = qb * LQ + warp * 32 < len;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c >= nch || !warp_live) break;               // key block without tokens: not loaded, not multiplied
        const uint32_t mk = mkc[c];
        float pr[32];
        if (mk == 0u) {
#pragma unroll
          for (int j = 0; j < 32; ++j) pr[j] = 0.f;
        } else {
          uint32_t v[32];
          tmem_ld32(trow + c * 32, v);
          tmem_ld_wait();
          if (mk == 0xffffffffu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              pr[j] = ex2_ftz(fmaf(__uint_as_float(v[j]), LOG2E, lrow[c * 32 + j] - mb));
              sum += pr[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float pj = 0.f;
              if ((mk >> j) & 1u) pj = ex2_ftz(fmaf(__uint_as_float(v[j]), LOG2E, lrow[c * 32 + j] - mb));
              pr[j] = pj;
              sum += pj;
            }
          }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(pr[q4 * 8 + 2 * e], pr[q4 * 8 + 2 * e + 1]);
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
          const int ch = c * 4 + q4;                     // 16-byte chunk of the key row
          const uint32_t box = sQ + (uint32_t)(ch >> 3) * BOX;
          const uint32_t addr = box + (uint32_t)r * 128u + (uint32_t)((((ch & 7) ^ (r & 7)) & 7) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
      }
      tcgen05_fence_before();                            // S has been read out of TMEM by this thread
      fence_async_smem();                                // P is visible to the tensor core's (async) proxy
      mbar_arrive(bar(B_PREADY + st));
      // ---- while the tensor core computes O = P V: move on to the next item and run its pass 1 ----
      const int cur_h = h, cur_len = len, cur_row0 = row0, cur_q0 = qb * LQ;
      have = next_item();
      float mb_next = 0.f;
      if (have) mb_next = pass1(k + 1u, h, qb * LQ + r);
      // ---- epilogue of the current item ----
      mbar_wait(bar(B_OFULL + st), ph);
      tcgen05_fence_after();
      uint32_t ov[2][32];
      const bool store_live = cur_q0 + warp * 32 < cur_len;
      if (store_live) {
        tmem_ld32(trow, ov[0]);
        tmem_ld32(trow + 32, ov[1]);
        tmem_ld_wait();
      }
      tcgen05_fence_before();
      mbar_arrive(bar(B_SFREE + st));                    // the accumulator may be overwritten by item k + 2
      // O rows go through the (dead) P rows of this warp in the stage -- row r, 16-byte chunk g ^ (r & 7) -- and leave as
      // full 128-byte lines: a row-per-thread store writes 32 half sectors per instruction and its slow drain held the
      // registers of the next item (ncu: 12.7 % of the samples on that write-after-read)
      if (store_live) {
        const float inv = 1.0f / sum;
        const uint32_t srow = sQ + (uint32_t)r * 128u;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = g * 8 + 2 * e;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(ov[i >> 5][i & 31]) * inv,
                                                      __uint_as_float(ov[(i + 1) >> 5][(i + 1) & 31]) * inv);
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)(((g ^ (r & 7)) & 7) << 4)),
                       "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
        __syncwarp();
        const int chunk = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = warp * 32 + i * 4 + (lane >> 3);        // row of the tile
          uint4 val;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                       : "r"(sQ + (uint32_t)rr * 128u + (uint32_t)(((chunk ^ (rr & 7)) & 7) << 4)) : "memory");
          if (cur_q0 + rr < cur_len)
            *reinterpret_cast<uint4*>(out + (size_t)(cur_row0 + cur_q0 + rr) * HD + cur_h * DK + chunk * 8) = val;
        }
      }
      mbar_arrive(bar(B_EMPTY + st));                    // the stage (P/O staging rows and V) may be reloaded
      ++k;
      mb = mb_next;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(G::TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
std::mutex g_mu;
EncodeTiledFn g_encode = nullptr;
std::map<std::tuple<const void*, size_t, size_t>, CUtensorMap> g_maps;

bool get_map(const void* ptr, size_t rows, size_t cols, CUtensorMap* out) {
  auto key = std::make_tuple(ptr, rows, cols);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return true; }
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess || !fn) { cudaGetLastError(); return false; }
    g_encode = (EncodeTiledFn)fn;
  }
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)DK, (cuuint32_t)LQ};
  cuuint32_t estr[2] = {1, 1};
  if (g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
  if (g_maps.size() > 1024) g_maps.clear();
  g_maps[key] = m;
  *out = m;
  return true;
}

}  // namespace ta

static size_t tc_smem_bytes(int nkb, int H, int Lb) {
  const size_t stage = nkb == 1 ? ta::Geo<1>::STAGE : ta::Geo<2>::STAGE;
  return (size_t)ta::NSTAGE * stage + (((size_t)H * (2 * Lb - 1) * 4 + 15) / 16) * 16 + 256 + 32 + 8 * ta::B_COUNT + 16 + 1024;
}

// Lmax <= 128: two CTAs per SM with 48 KiB stages; Lmax <= 256: one CTA per SM, 96 KiB stages + the bias LUT must fit 227 KiB
bool enc_attention_tc_supported(int dk, int Lmax, int Lb, int H) {
  if (dk != ta::DK || Lb > 256 || H > 32 || Lmax > 2 * ta::LQ) return false;
  return tc_smem_bytes(Lmax <= ta::LQ ? 1 : 2, H, Lb) <= 227 * 1024;
}

cudaError_t enc_attention_tc(const void* qkv, size_t qkv_rows, void* out, const int* plen, const int* poff,
                             const uint8_t* tok_valid, const float* bias_lut, int Lb, int P, int H, int Lmax, cudaStream_t s) {
  if (P <= 0) return cudaSuccess;
  if (!enc_attention_tc_supported(ta::DK, Lmax, Lb, H)) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(ta::g_mu);
  CUtensorMap map;
  if (!ta::get_map(qkv, qkv_rows, (size_t)3 * H * ta::DK, &map)) return cudaErrorUnknown;
  const int nkb = Lmax <= ta::LQ ? 1 : 2;
  const size_t smem = tc_smem_bytes(nkb, H, Lb);
  static SmemAttr attr[2];
  {
    cudaError_t e = nkb == 1 ? attr[0].ensure(ta::enc_attention_tc_kernel<1>, smem) : attr[1].ensure(ta::enc_attention_tc_kernel<2>, smem);
    if (e != cudaSuccess) return e;
  }
  static int sms[64];
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (sms[dev] == 0 && cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return cudaGetLastError();
  if (nkb == 1) {
    const int grid = P < 2 * sms[dev] ? P : 2 * sms[dev];          // two persistent CTAs per SM
    ta::enc_attention_tc_kernel<1><<<grid, ta::THREADS, smem, s>>>(map, (bf16*)out, plen, poff, tok_valid, bias_lut, Lb, H, P);
  } else {
    const int grid = P < sms[dev] ? P : sms[dev];
    ta::enc_attention_tc_kernel<2><<<grid, ta::THREADS, smem, s>>>(map, (bf16*)out, plen, poff, tok_valid, bias_lut, Lb, H, P);
  }
  return cudaGetLastError();
}

}  // namespace gram
