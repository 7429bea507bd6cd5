// Tensor-core attention kernels for bf16 mode (mma.sync m16n8k16, fp32 accumulation, online softmax).
//
// cross_attention_mma  -- kernel (b) of the north star: decoder cross-attention over the fused FiD memory
//   (reference src/model/gram_t5_modeling.py:670-705 with cached K/V :549, attention :572-621).  The K beams of
//   a user are the M rows of ONE flash-decoding problem, so the user's K/V is streamed from HBM exactly once
//   per layer per step (the reference streams it K times after `_expand_inputs_for_generation`, and copies it
//   again in `_reorder_cache`, src/model/gram_t5.py:320-348).  K/V is read IN PLACE from the buffer the
//   projection GEMM wrote ([token][layer][K|V][head][dk], no concat copy): a TMA producer warp streams
//   64-row x 64-col boxes per head into a 3-stage 128B-swizzled shared-memory ring; four consumer warps (one
//   head each) run QK^T and PV on the tensor cores straight out of the swizzled tiles via ldmatrix.
//   This op is HBM-bound by design: per 2 KB row of K|V it does 2*2*32*512 padded flops, i.e. ~64 flop/B.
//
// enc_attention_mma    -- bidirectional self-attention of the encoder, one CTA per passage walking its heads
//   (reference src/model/gram_t5_modeling.py:572-621 with the layer-0 relative-position bias :452-477 shared by
//   all layers :1249 and the key padding mask :1130).  Same tile math; Q/K/V of head h+1 are prefetched with
//   cp.async while head h is computed.
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <map>
#include <tuple>
#include <string>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace gram {
namespace fa {

constexpr int TS = 64;                 // memory rows per tile
constexpr int DK = 64;                 // head dim (these kernels are specialised for d_kv = 64)
constexpr int BOX_BYTES = TS * DK * 2; // 8 KiB: one head's 64x64 bf16 tile, 128-byte rows
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FA_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra FA_WAIT_DONE;\n"
      "bra FA_WAIT_LOOP;\n"
      "FA_WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// address of 16-byte chunk `chunk` of row `row` inside a 128B-swizzled [rows][64 bf16] tile (base 1024-aligned)
__device__ __forceinline__ uint32_t swz(uint32_t base, int row, int chunk) {
  return base + (uint32_t)(row * 128) + (uint32_t)(((chunk ^ (row & 7)) & 7) << 4);
}

// One tile of flash attention for MT 16-row query tiles held by one warp: NKT*8 keys starting at key row
// `row_off` of a [64 keys][64 d] swizzled tile.
//   qf      Q fragments [MT][4 k-steps][4]
//   kbase   shared address of the K tile, vbase of the V tile
//   BiasFn  additive score term bias(row_in_warp_tile, key_in_tile) (0 for cross-attention)
//   kmask   bit j set = key (row_off + j) of the tile is visible
template <int MT, int NKT, bool SKIPS, typename BiasFn>
__device__ __forceinline__ void flash_tile(const uint32_t (&qf)[MT][4][4], uint32_t kbase, uint32_t vbase, int row_off,
                                           unsigned long long kmask, BiasFn bias, float (&o)[MT][8][4],
                                           float (&m_run)[MT][2], float (&l_run)[MT][2], int lane) {
  const int g = lane >> 2, q = lane & 3;
  float s[MT][NKT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
  // ---- S = Q K^T ----
#pragma unroll
  for (int p = 0; p < NKT / 2; ++p) {      // pairs of 8-key n-tiles
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {       // 16-wide steps over d
      uint32_t b0, b1, b2, b3;
      const int row = row_off + p * 16 + (lane & 7) + ((lane >> 4) << 3);
      const int chunk = ks * 2 + ((lane >> 3) & 1);
      ldsm_x4(swz(kbase, row, chunk), b0, b1, b2, b3);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        mma_bf16(s[mt][2 * p], qf[mt][ks], b0, b1);
        mma_bf16(s[mt][2 * p + 1], qf[mt][ks], b2, b3);
      }
    }
  }
  // ---- bias, mask, online softmax ----
  // SKIPS (encoder kernel, 16 warps per SM): warp-uniform branches drop work that is usually unnecessary -- masking
  // runs only for tiles that contain an invisible key (the tail of a passage) and the rescale of the output
  // accumulator only when some row of the warp raised its running maximum (corr is exactly 1 otherwise): -10 %.
  // The cross-attention kernel has ONE warp per scheduler and lives on instruction-level parallelism across the
  // unrolled (mt, nt) loops; the same branches fence the scheduler's reordering and cost it 13 %, so it keeps the
  // straight-line selects (measured both ways).
  if (SKIPS) {
    const bool all_visible = (NKT == 8) ? (kmask == ~0ull) : ((unsigned int)kmask == 0xffffffffu);
    if (!BiasFn::kZero) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            s[mt][nt][e] += bias(mt * 16 + g + ((e >> 1) << 3), row_off + nt * 8 + 2 * q + (e & 1));
    }
    if (!all_visible) {                            // after the bias: a masked column's bias index may be out of range
#pragma unroll
      for (int nt = 0; nt < NKT; ++nt) {
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          const bool vis = (kmask >> (nt * 8 + 2 * q + e2)) & 1ull;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            s[mt][nt][e2] = vis ? s[mt][nt][e2] : -INFINITY;
            s[mt][nt][2 + e2] = vis ? s[mt][nt][2 + e2] : -INFINITY;
          }
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NKT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[mt][nt][e]);
      }
      float corr[2];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float t = mx[hf];
        t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 1));
        t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 2));
        const float m_old = m_run[mt][hf];
        const float m_new = fmaxf(m_old, t);
        corr[hf] = (m_old == -INFINITY) ? 0.f : ex2_ftz((m_old - m_new) * LOG2E);
        m_run[mt][hf] = m_new;
        const float mb = (m_new == -INFINITY) ? 0.f : m_new * LOG2E;
        float psum = 0.f;
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) {
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int e = hf * 2 + e2;
            const float p = ex2_ftz(s[mt][nt][e] * LOG2E - mb);   // exp2(-inf) = 0 for masked keys
            s[mt][nt][e] = p;
            psum += p;
          }
        }
        l_run[mt][hf] = l_run[mt][hf] * corr[hf] + psum;
      }
      if (__any_sync(0xffffffffu, corr[0] != 1.f || corr[1] != 1.f)) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          o[mt][nt][0] *= corr[0];
          o[mt][nt][1] *= corr[0];
          o[mt][nt][2] *= corr[1];
          o[mt][nt][3] *= corr[1];
        }
      }
    }
  } else {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NKT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = nt * 8 + 2 * q + (e & 1);
          const int rr = mt * 16 + g + ((e >> 1) << 3);
          const bool vis = (kmask >> col) & 1ull;
          // (dropping the "+ 0" of NoBias changes ptxas's allocation of this 255-register kernel and spills 24 bytes)
          const float v = vis ? s[mt][nt][e] + bias(rr, row_off + col) : -INFINITY;
          s[mt][nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float t = mx[hf];
        t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 1));
        t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 2));
        const float m_old = m_run[mt][hf];
        const float m_new = fmaxf(m_old, t);
        const float corr = (m_old == -INFINITY) ? 0.f : ex2_ftz((m_old - m_new) * LOG2E);
        m_run[mt][hf] = m_new;
        l_run[mt][hf] *= corr;
        const float mb = (m_new == -INFINITY) ? 0.f : m_new * LOG2E;
        float psum = 0.f;
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) {
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int e = hf * 2 + e2;
            const float p = ex2_ftz(s[mt][nt][e] * LOG2E - mb);   // exp2(-inf) = 0 for masked keys
            s[mt][nt][e] = p;
            psum += p;
          }
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          o[mt][nt][hf * 2] *= corr;
          o[mt][nt][hf * 2 + 1] *= corr;
        }
        l_run[mt][hf] += psum;
      }
    }
  }
  // ---- O += P V ----
#pragma unroll
  for (int ks = 0; ks < NKT / 2; ++ks) {   // 16-key steps
    uint32_t pa[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      pa[mt][0] = pack_bf16(s[mt][2 * ks][0], s[mt][2 * ks][1]);
      pa[mt][1] = pack_bf16(s[mt][2 * ks][2], s[mt][2 * ks][3]);
      pa[mt][2] = pack_bf16(s[mt][2 * ks + 1][0], s[mt][2 * ks + 1][1]);
      pa[mt][3] = pack_bf16(s[mt][2 * ks + 1][2], s[mt][2 * ks + 1][3]);
    }
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {       // pairs of 8-wide d n-tiles
      uint32_t b0, b1, b2, b3;
      const int row = row_off + ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
      const int chunk = dp * 2 + (lane >> 4);
      ldsm_x4_t(swz(vbase, row, chunk), b0, b1, b2, b3);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        mma_bf16(o[mt][2 * dp], pa[mt], b0, b1);
        mma_bf16(o[mt][2 * dp + 1], pa[mt], b2, b3);
      }
    }
  }
}

// 8x8 transpose of a b16 matrix held in the ldmatrix / mma fragment distribution (lane = (row g, columns 2q, 2q+1))
__device__ __forceinline__ uint32_t movm_t(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

// Transposed flash tile for the cross-attention (no bias): the 64 keys of the tile are the M dimension and the beams the N
// dimension, S^T = K Q^T and O^T += V^T P^T, so that 20 beams cost NT = 3 eight-beam n-tiles instead of two 16-row m-tiles:
// 96 mma.sync and 48 exponentials per lane per tile instead of 128 and 64.  The kernel sits on the power cap even alone
// (scripts/exp_xattn_hot.py, exp_hot_memory.py): instructions it does not issue are bandwidth it gets back.
//   qb      B fragments of Q^T: [beam n-tile][16-wide step over d][2]
//   ot      O^T accumulators [d m-tile][beam n-tile][4]: rows d = 16 j + g (+ 8), columns beam = 8 nt + 2q (+ 1)
//   m_run / l_run  per beam column held by this lane; l_run is this lane's partial over its keys (summed over g at the end)
//   MASK    false: every key of the tile is visible (kmask == ~0, all tiles of a user but the last): no visibility selects
template <int NT, bool MASK>
__device__ __forceinline__ void flash_tile_t(const uint32_t (&qb)[NT][4][2], uint32_t kbase, uint32_t vbase,
                                             unsigned long long kmask, float (&ot)[4][NT][4], float (&m_run)[NT][2],
                                             float (&l_run)[NT][2], int lane) {
  const int g = lane >> 2;
  float st[4][NT][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) st[i][nt][e] = 0.f;
  // ---- S^T = K Q^T: A = 16 keys x 16 d straight out of the swizzled K tile ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      ldsm_x4(swz(kbase, i * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), ks * 2 + (lane >> 4)), a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_bf16(st[i][nt], a, qb[nt][ks][0], qb[nt][ks][1]);
    }
  }
  // ---- mask + online softmax over the keys (rows): this lane holds keys 16 i + g (+ 8) of columns 2q, 2q + 1 ----
  bool vis[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    vis[i][0] = !MASK || ((kmask >> (i * 16 + g)) & 1ull);
    vis[i][1] = !MASK || ((kmask >> (i * 16 + g + 8)) & 1ull);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e2 = 0; e2 < 2; ++e2) {
      float t = -INFINITY;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const float v = (!MASK || vis[i][hf]) ? st[i][nt][hf * 2 + e2] : -INFINITY;
          if (MASK) st[i][nt][hf * 2 + e2] = v;
          t = fmaxf(t, v);
        }
      }
      t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 4));
      t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 8));
      t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 16));
      const float m_old = m_run[nt][e2];
      const float m_new = fmaxf(m_old, t);
      const float corr = (m_old == -INFINITY) ? 0.f : ex2_ftz((m_old - m_new) * LOG2E);
      m_run[nt][e2] = m_new;
      const float mb = (m_new == -INFINITY) ? 0.f : m_new * LOG2E;
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const float pv = ex2_ftz(st[i][nt][hf * 2 + e2] * LOG2E - mb);     // exp2(-inf) = 0 for masked keys
          st[i][nt][hf * 2 + e2] = pv;
          psum += pv;
        }
      }
      l_run[nt][e2] = l_run[nt][e2] * corr + psum;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ot[j][nt][e2] *= corr;
        ot[j][nt][2 + e2] *= corr;
      }
    }
  }
  // ---- O^T += V^T P^T: B = P^T (the accumulator blocks transposed in registers), A = V^T via ldmatrix.trans ----
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {                 // 16-key steps
    uint32_t pb[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      pb[nt][0] = movm_t(pack_bf16(st[kk][nt][0], st[kk][nt][1]));     // keys 16 kk + 0..7
      pb[nt][1] = movm_t(pack_bf16(st[kk][nt][2], st[kk][nt][3]));     // keys 16 kk + 8..15
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {                  // 16-wide d m-tiles
      uint32_t a[4];
      ldsm_x4_t(swz(vbase, kk * 16 + (lane & 7) + (((lane >> 4) & 1) << 3), j * 2 + ((lane >> 3) & 1)), a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_bf16(ot[j][nt], a, pb[nt][0], pb[nt][1]);
    }
  }
}

struct NoBias {
  static constexpr bool kZero = true;
  __device__ __forceinline__ float operator()(int, int) const { return 0.f; }
};

// ================================================================================================
// cross attention
// ================================================================================================
constexpr int xa_threads(int warps) { return (warps + 1) * 32; }   // consumer warps + 1 TMA producer warp
constexpr int XA_STAGE_TARGET = 64 * 1024;

// HEADS heads per CTA, BH beam-halves (32 beams each) and KH key-halves (32 keys of every 64-key tile) per head:
// HEADS * BH * KH == 4 consumer warps.  Key-halves are merged once at the end through shared memory.
//   K <= 32: <4, 1, 1, 1>: one warp per head, 64 KiB stages x 3, one CTA per SM      (5.1-5.3 TB/s on B200)
//   K <= 64: <2, 2, 1, 1>: 2 heads x 2 beam-halves, 32 KiB stages x 6
//   Measured alternatives for K <= 32 (kept as template options): <2, 1, 2, 2> (key-halves, two CTAs per SM) 8 % slower;
//   <4, 1, 2, 1> (8 consumer warps per CTA) 40 % slower -- splitting a 64-key tile between two warps doubles the per-tile
//   softmax bookkeeping (running max / rescale of the 32 x 64 output tile) and spills; <4, 2, 1, 1, MT = 1> (8 warps,
//   16 beams each, no merge, 168 registers) within 0.6 % of <4, 1, 1, 1> -- the consumers are not what paces it.
//   Timed alone (scripts/exp_xattn_context.py) the kernel streams 6.3 TB/s back to back = 0.96 of the measured copy
//   bandwidth (6.0 with a sync between launches); inside a step, after the encoder phase has driven the chip into its
//   power cap, the same launches run at 4.9-5.3 TB/s.
template <int HEADS, int BH, int KH, int MINB, int MT = 2>
__global__ void __launch_bounds__(xa_threads(HEADS * BH * KH), MINB)
cross_attention_mma_kernel(const __grid_constant__ CUtensorMap map_kv, const bf16* __restrict__ qg,
                           bf16* __restrict__ out, const int* __restrict__ ustart, const int* __restrict__ uorder,
                           const uint8_t* __restrict__ tok_valid, int K_all, int H, int k_col0, int v_col0,
                           const int* __restrict__ live_start, const int* __restrict__ live_count) {
  constexpr int XA_HEADS = HEADS;
  constexpr uint32_t XA_STAGE_BYTES = 2 * HEADS * BOX_BYTES;   // K boxes then V boxes
  constexpr int XA_STAGES = MINB == 2 ? 3 : 3 * XA_STAGE_TARGET / (int)XA_STAGE_BYTES;
  constexpr int NKT = 8 / KH;
  constexpr int XA_WARPS = HEADS * BH * KH;                     // consumer warps (4 or 8)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + XA_STAGES * XA_STAGE_BYTES;     // full[ST], empty[ST]
  unsigned long long* masks = reinterpret_cast<unsigned long long*>(smem + XA_STAGES * XA_STAGE_BYTES + 128);

  // longest-first order: the block scheduler hands out blockIdx.x in order, so heavy users start first and the
  // short ones fill the tail of the last wave
  const int u = uorder ? uorder[blockIdx.x] : blockIdx.x, hg = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // query rows: the user's K beams, or (live-row compaction) its live beams only; a user whose beams are all dead or
  // whose hypotheses are final does not stream its K/V at all
  const int K = live_start ? live_count[u] : K_all;
  if (K == 0) return;
  const int qrow0 = live_start ? live_start[u] : u * K_all;
  // the kernel sits at the 255-register limit: the row range is parked in shared memory for the epilogue instead of
  // staying live across the key loop (two more live registers spill 24 bytes)
  __shared__ int s_rows[2];
  if (threadIdx.x == 0) { s_rows[0] = qrow0; s_rows[1] = K; }
  const int s_beg = ustart[u], s_end = ustart[u + 1];
  const int n_tiles = (s_end - s_beg + TS - 1) / TS;
  const int HD = H * DK;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    for (int s = 0; s < XA_STAGES; ++s) { mbar_init(bars + 8u * s, 1); mbar_init(bars + 8u * (XA_STAGES + s), XA_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == XA_WARPS) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(bars + 8u * (XA_STAGES + stage), phase ^ 1u);
      const int s0 = s_beg + t * TS;
      const int r0 = s0 + lane, r1 = s0 + 32 + lane;
      const bool v0 = (r0 < s_end) && (tok_valid == nullptr || tok_valid[r0] != 0);
      const bool v1 = (r1 < s_end) && (tok_valid == nullptr || tok_valid[r1] != 0);
      const unsigned lo = __ballot_sync(0xffffffffu, v0), hi = __ballot_sync(0xffffffffu, v1);
      if (lane == 0) {
        masks[stage] = ((unsigned long long)hi << 32) | lo;
        const uint32_t full = bars + 8u * stage;
        const uint32_t sb = base + stage * XA_STAGE_BYTES;
        mbar_arrive_expect_tx(full, XA_STAGE_BYTES);
#pragma unroll
        for (int hh = 0; hh < XA_HEADS; ++hh) {
          const int col = (hg * XA_HEADS + hh) * DK;
          tma_load_2d(sb + hh * BOX_BYTES, &map_kv, full, k_col0 + col, s0);
          tma_load_2d(sb + (XA_HEADS + hh) * BOX_BYTES, &map_kv, full, v_col0 + col, s0);
        }
      }
      __syncwarp();
      if (++stage == XA_STAGES) { stage = 0; phase ^= 1u; }
    }
    return;
  }

  // ===================== consumers: warp = (head, beam half, key half) =====================
  const int hl = warp / (BH * KH), b_off = ((warp / KH) % BH) * (MT * 16), kh = warp % KH;
  const int h = hg * XA_HEADS + hl;
  const int g = lane >> 2, q = lane & 3;
  uint32_t qf[MT][4][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int b = b_off + mt * 16 + g + ((e & 1) << 3);
        const int d = ks * 16 + 2 * q + ((e >> 1) << 3);
        uint32_t v = 0u;
        if (b < K) v = *reinterpret_cast<const uint32_t*>(qg + (size_t)(qrow0 + b) * HD + h * DK + d);
        qf[mt][ks][e] = v;
      }
    }
  }
  float o[MT][8][4];
  float m_run[MT][2], l_run[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    m_run[mt][0] = m_run[mt][1] = -INFINITY;
    l_run[mt][0] = l_run[mt][1] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.f;
  }
  int stage = 0; uint32_t phase = 0;
  for (int t = 0; t < n_tiles; ++t) {
    mbar_wait(bars + 8u * stage, phase);
    const unsigned long long kmask = masks[stage];
    const uint32_t sb = base + stage * XA_STAGE_BYTES;
    flash_tile<MT, NKT, false>(qf, sb + hl * BOX_BYTES, sb + (XA_HEADS + hl) * BOX_BYTES, kh * 32, kmask >> (kh * 32), NoBias(), o,
                        m_run, l_run, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8u * (XA_STAGES + stage));
    if (++stage == XA_STAGES) { stage = 0; phase ^= 1u; }
  }
  if (KH == 2) {
    // ---- merge the two key-halves of every (head, beam half): every tile has been consumed, so the ring is free ----
    asm volatile("bar.sync 1, %0;" ::"n"(XA_WARPS * 32) : "memory");
    float* xch = reinterpret_cast<float*>(smem) + (size_t)(warp >> 1) * (32 * 72);   // 72 floats per lane
    if (kh == 1) {
      float* dst = xch + lane;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          dst[(mt * 2 + hf) * 32] = m_run[mt][hf];
          dst[(4 + mt * 2 + hf) * 32] = l_run[mt][hf];
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[(8 + (mt * 8 + nt) * 4 + e) * 32] = o[mt][nt][e];
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(XA_WARPS * 32) : "memory");
    if (kh == 1) return;
    const float* src = xch + lane;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const float m0 = m_run[mt][hf], m1 = src[(mt * 2 + hf) * 32];
        const float mn = fmaxf(m0, m1);
        const float c0 = (m0 == -INFINITY) ? 0.f : ex2_ftz((m0 - mn) * LOG2E);
        const float c1 = (m1 == -INFINITY) ? 0.f : ex2_ftz((m1 - mn) * LOG2E);
        l_run[mt][hf] = l_run[mt][hf] * c0 + src[(4 + mt * 2 + hf) * 32] * c1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          o[mt][nt][hf * 2] = o[mt][nt][hf * 2] * c0 + src[(8 + (mt * 8 + nt) * 4 + hf * 2) * 32] * c1;
          o[mt][nt][hf * 2 + 1] = o[mt][nt][hf * 2 + 1] * c0 + src[(8 + (mt * 8 + nt) * 4 + hf * 2 + 1) * 32] * c1;
        }
      }
    }
  }
  // ---- normalise and store ----
  const int out_row0 = s_rows[0], out_rows = s_rows[1];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float l = l_run[mt][hf];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      const int b = b_off + mt * 16 + g + hf * 8;
      if (b < out_rows) {
        bf16* orow = out + (size_t)(out_row0 + b) * HD + h * DK;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          *reinterpret_cast<uint32_t*>(orow + nt * 8 + 2 * q) = pack_bf16(o[mt][nt][hf * 2] * inv, o[mt][nt][hf * 2 + 1] * inv);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant (default): one CTA per SM walks work items (user, head group) -- users longest first, items dealt
// round-robin so every CTA gets the same mix of long and short memories.  The TMA producer warp never stops: it streams the
// K/V tiles of item i+1 into the ring while the consumer warps finish item i (normalise, store) and fetch the next
// item's query fragments, so the per-CTA costs of the one-CTA-per-item kernel above -- CTA launch, barrier set-up, the
// first tiles' load latency with an empty ring, an idle memory system during the epilogue (~2-3 us of a ~25 us item) --
// are paid once per launch instead of once per item.
// ------------------------------------------------------------------------------------------------
template <int HEADS, int BH, int MT = 2, int NT = 0>      // NT > 0: transposed tiles (flash_tile_t), NT eight-beam n-tiles
__global__ void __launch_bounds__(xa_threads(HEADS * BH), 1)
cross_attention_persist_kernel(const __grid_constant__ CUtensorMap map_kv, const bf16* __restrict__ qg,
                               bf16* __restrict__ out, const int* __restrict__ ustart, const int* __restrict__ uorder,
                               const uint8_t* __restrict__ tok_valid, int K_all, int H, int k_col0, int v_col0,
                               const int* __restrict__ live_start, const int* __restrict__ live_count, int users, int fast_path) {
  constexpr int XA_HEADS = HEADS;
  constexpr uint32_t XA_STAGE_BYTES = 2 * HEADS * BOX_BYTES;   // K boxes then V boxes
  constexpr int XA_STAGES = 3 * XA_STAGE_TARGET / (int)XA_STAGE_BYTES;
  constexpr int XA_WARPS = HEADS * BH;                         // consumer warps
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bars = base + XA_STAGES * XA_STAGE_BYTES;     // full[ST], empty[ST]
  unsigned long long* masks = reinterpret_cast<unsigned long long*>(smem + XA_STAGES * XA_STAGE_BYTES + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_hg = H / XA_HEADS, n_items = users * n_hg;
  const int HD = H * DK;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    for (int s = 0; s < XA_STAGES; ++s) { mbar_init(bars + 8u * s, 1); mbar_init(bars + 8u * (XA_STAGES + s), XA_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- item schedule shared by the producer and the consumer warps -------------------------------------------------
  // The items of this CTA are sequence entries k = 0, 1, ...: item blockIdx.x + k * gridDim.x; users without live beams are
  // skipped.  Every warp keeps a WINDOW of 32 consecutive entries, one per lane (user's live rows, first query row, tile
  // count, K/V row range), loaded with one round of lane-parallel loads, and steps through it with a ballot + shuffles:
  // no dependent global load (uorder -> live_count / live_start / ustart) is left on any warp's per-item path: 751 -> 738 us
  // per launch alone, 36-37 -> 34.5-35.4 ms per step.  (What is left of the gap to a 1-row launch, 675 us, is not a code
  // path: with the query loads compiled out -- all-zero fragments -- the same 20-row launch takes 705 us, and staging the
  // query rows through shared memory a whole item ahead changes nothing (739 us): zero operands draw less power, and this
  // kernel sits on the power cap even when it runs alone, scripts/exp_xattn_hot.py.)
  int kwin = 0, w_item = 0, w_K = 0, w_qrow0 = 0, w_beg = 0, w_end = 0;
  auto load_window = [&]() {
    w_item = (int)blockIdx.x + (kwin + lane) * (int)gridDim.x;
    w_K = 0; w_qrow0 = 0; w_beg = 0; w_end = 0;
    if (w_item < n_items) {
      const int ui = w_item / n_hg;
      const int u = uorder ? uorder[ui] : ui;
      w_K = live_start ? live_count[u] : K_all;
      w_qrow0 = live_start ? live_start[u] : u * K_all;
      w_beg = ustart[u]; w_end = ustart[u + 1];
    }
  };
  struct Item { int K, qrow0, s_beg, s_end, n_tiles, hg, k; };
  // first entry with sequence index >= kmin whose user has live beams; the producer (need_tiles) also skips users without
  // memory tokens -- the consumers still visit those and store zeros, without touching the ring
  auto next_item = [&](int kmin, Item& it, bool need_tiles) -> bool {
    for (;;) {
      const bool ok = (kwin + lane >= kmin) && (w_item < n_items) && (w_K > 0) && (!need_tiles || w_end > w_beg);
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (m) {
        const int j = __ffs(m) - 1;
        it.K = __shfl_sync(0xffffffffu, w_K, j);
        it.qrow0 = __shfl_sync(0xffffffffu, w_qrow0, j);
        it.s_beg = __shfl_sync(0xffffffffu, w_beg, j);
        it.s_end = __shfl_sync(0xffffffffu, w_end, j);
        it.n_tiles = (it.s_end - it.s_beg + TS - 1) / TS;
        it.hg = __shfl_sync(0xffffffffu, w_item, j) % n_hg;
        it.k = kwin + j;
        return true;
      }
      if (__shfl_sync(0xffffffffu, w_item, 31) >= n_items) return false;   // the sequence ended inside this window
      kwin += 32;
      load_window();
    }
  };
  load_window();

  if (warp == XA_WARPS) {
    // ===================== TMA producer: one uninterrupted tile stream over all items of this CTA =====================
    // The key-validity bytes of tile t+1 are requested right after tile t's loads are issued, so their L2 round trip runs
    // under the wait for the next free ring slot instead of in front of every tile's TMA issue.
    int stage = 0; uint32_t phase = 0;
    Item it;
    int t = 0;
    bool more = next_item(0, it, true);
    bool v0 = false, v1 = false;
    auto fetch_valid = [&]() {
      const int r0 = it.s_beg + t * TS + lane, r1 = r0 + 32;
      v0 = (r0 < it.s_end) && (tok_valid == nullptr || tok_valid[r0] != 0);
      v1 = (r1 < it.s_end) && (tok_valid == nullptr || tok_valid[r1] != 0);
    };
    if (more) fetch_valid();
    while (more) {
      mbar_wait(bars + 8u * (XA_STAGES + stage), phase ^ 1u);
      const int s0 = it.s_beg + t * TS;
      const unsigned lo = __ballot_sync(0xffffffffu, v0), hi = __ballot_sync(0xffffffffu, v1);
      if (lane == 0) {
        masks[stage] = ((unsigned long long)hi << 32) | lo;
        const uint32_t full = bars + 8u * stage;
        const uint32_t sb = base + stage * XA_STAGE_BYTES;
        mbar_arrive_expect_tx(full, XA_STAGE_BYTES);
#pragma unroll
        for (int hh = 0; hh < XA_HEADS; ++hh) {
          const int col = (it.hg * XA_HEADS + hh) * DK;
          tma_load_2d(sb + hh * BOX_BYTES, &map_kv, full, k_col0 + col, s0);
          tma_load_2d(sb + (XA_HEADS + hh) * BOX_BYTES, &map_kv, full, v_col0 + col, s0);
        }
      }
      __syncwarp();
      if (++t == it.n_tiles) { more = next_item(it.k + 1, it, true); t = 0; }
      if (more) fetch_valid();
      if (++stage == XA_STAGES) { stage = 0; phase ^= 1u; }
    }
    return;
  }

  // ===================== consumers: warp = (head, beam half) =====================
  // The next item's query fragments are requested right after the tile loop -- qf is dead by then -- so that their round
  // trip runs under the normalise-and-store epilogue of the current item.
  const int hl = warp / BH, b_off = (warp % BH) * (MT * 16);
  const int g = lane >> 2, q = lane & 3;
  int stage = 0; uint32_t phase = 0;
  if constexpr (NT > 0) {
    // ---- transposed tiles (beams on the N dimension): one warp per head holds all of the user's beams ----
    static_assert(NT == 0 || BH == 1, "transposed tiles: one warp per head");
    uint32_t qb[NT][4][2];
    auto load_qb = [&](const Item& it) {
      const int h = it.hg * XA_HEADS + hl;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int b = nt * 8 + g;
        const bf16* qrow = qg + (size_t)(it.qrow0 + (b < it.K ? b : 0)) * HD + h * DK + 2 * q;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          qb[nt][ks][0] = b < it.K ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16) : 0u;
          qb[nt][ks][1] = b < it.K ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8) : 0u;
        }
      }
    };
    Item cur;
    bool have = next_item(0, cur, false);
    if (have) load_qb(cur);
    while (have) {
      float ot[4][NT][4];
      float m_run[NT][2], l_run[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        m_run[nt][0] = m_run[nt][1] = -INFINITY;
        l_run[nt][0] = l_run[nt][1] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) ot[j][nt][e] = 0.f;
      }
      for (int t = 0; t < cur.n_tiles; ++t) {
        mbar_wait(bars + 8u * stage, phase);
        const unsigned long long kmask = masks[stage];
        const uint32_t sb = base + stage * XA_STAGE_BYTES;
        if (kmask == ~0ull && fast_path)                 // warp-uniform: two straight-line copies of the tile
          flash_tile_t<NT, false>(qb, sb + hl * BOX_BYTES, sb + (XA_HEADS + hl) * BOX_BYTES, kmask, ot, m_run, l_run, lane);
        else
          flash_tile_t<NT, true>(qb, sb + hl * BOX_BYTES, sb + (XA_HEADS + hl) * BOX_BYTES, kmask, ot, m_run, l_run, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8u * (XA_STAGES + stage));
        if (++stage == XA_STAGES) { stage = 0; phase ^= 1u; }
      }
      const int out_row0 = cur.qrow0, out_rows = cur.K, h = cur.hg * XA_HEADS + hl;
      have = next_item(cur.k + 1, cur, false);
      if (have) load_qb(cur);                        // the next item's fragments: in flight under this epilogue
      // ---- row sums over the key lanes, normalise, transpose back to beam-major pairs of d, store ----
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float inv[2];
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          float l = l_run[nt][e2];
          l += __shfl_xor_sync(0xffffffffu, l, 4);
          l += __shfl_xor_sync(0xffffffffu, l, 8);
          l += __shfl_xor_sync(0xffffffffu, l, 16);
          inv[e2] = l > 0.f ? 1.0f / l : 0.f;
        }
        const int b = nt * 8 + g;                    // after the transpose this lane holds beam 8 nt + g, d pairs 2q, 2q + 1
        bf16* orow = out + (size_t)(out_row0 + b) * HD + h * DK + 2 * q;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t lo = movm_t(pack_bf16(ot[j][nt][0] * inv[0], ot[j][nt][1] * inv[1]));   // d = 16 j + 0..7
          const uint32_t hi = movm_t(pack_bf16(ot[j][nt][2] * inv[0], ot[j][nt][3] * inv[1]));   // d = 16 j + 8..15
          if (b < out_rows) {
            *reinterpret_cast<uint32_t*>(orow + j * 16) = lo;
            *reinterpret_cast<uint32_t*>(orow + j * 16 + 8) = hi;
          }
        }
      }
    }
    return;
  }
  uint32_t qf[MT][4][4];
  auto load_q = [&](const Item& it) {
    const int h = it.hg * XA_HEADS + hl;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int b = b_off + mt * 16 + g + ((e & 1) << 3);
          const int d = ks * 16 + 2 * q + ((e >> 1) << 3);
          uint32_t v = 0u;
          if (b < it.K) v = *reinterpret_cast<const uint32_t*>(qg + (size_t)(it.qrow0 + b) * HD + h * DK + d);
          qf[mt][ks][e] = v;
        }
      }
    }
  };
  Item cur, nxt;
  bool have = next_item(0, cur, false);
  if (have) load_q(cur);
  while (have) {
    float o[MT][8][4];
    float m_run[MT][2], l_run[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      m_run[mt][0] = m_run[mt][1] = -INFINITY;
      l_run[mt][0] = l_run[mt][1] = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.f;
    }
    for (int t = 0; t < cur.n_tiles; ++t) {
      mbar_wait(bars + 8u * stage, phase);
      const unsigned long long kmask = masks[stage];
      const uint32_t sb = base + stage * XA_STAGE_BYTES;
      flash_tile<MT, 8, false>(qf, sb + hl * BOX_BYTES, sb + (XA_HEADS + hl) * BOX_BYTES, 0, kmask, NoBias(), o, m_run, l_run, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8u * (XA_STAGES + stage));
      if (++stage == XA_STAGES) { stage = 0; phase ^= 1u; }
    }
    const int out_row0 = cur.qrow0, out_rows = cur.K, h = cur.hg * XA_HEADS + hl;
    const bool have_n = next_item(cur.k + 1, nxt, false);
    if (have_n) load_q(nxt);
    // ---- normalise and store ----
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float l = l_run[mt][hf];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        const int b = b_off + mt * 16 + g + hf * 8;
        if (b < out_rows) {
          bf16* orow = out + (size_t)(out_row0 + b) * HD + h * DK;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<uint32_t*>(orow + nt * 8 + 2 * q) = pack_bf16(o[mt][nt][hf * 2] * inv, o[mt][nt][hf * 2 + 1] * inv);
        }
      }
    }
    have = have_n;
    cur = nxt;
  }
}

// ================================================================================================
// encoder self-attention: shared definitions
// ================================================================================================
constexpr int EA_THREADS = 256;        // 8 warps x 16 query rows: low register count -> more resident CTAs per SM

struct RelBias {
  static constexpr bool kZero = false;
  const float* lut;   // shared memory, [2*Lb-1]
  int off;            // key_tile_start - query_warp_start + Lb - 1
  __device__ __forceinline__ float operator()(int rr, int col) const { return lut[off + col - rr]; }
};

// ================================================================================================
// encoder self-attention, pipelined version: one CTA per passage walks its H heads; Q, K and V of head h+1 are
// prefetched with cp.async (128B-swizzled, zero-filled past the passage length) while head h is computed, so the
// global-load latency that dominated the one-(passage, head)-per-CTA kernel is hidden.
// ================================================================================================
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;                                  // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int LT>
__global__ void __launch_bounds__(EA_THREADS, 2)
enc_attention_pipe_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, const int* __restrict__ plen,
                          const int* __restrict__ poff, const uint8_t* __restrict__ tok_valid,
                          const float* __restrict__ bias_lut, int Lb, int H) {
  const int p = blockIdx.x;
  const int len = plen[p];
  if (len == 0) return;
  const int row0 = poff[p];
  const int HD = H * DK;
  const size_t ld = (size_t)3 * HD;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  constexpr uint32_t BUF = 3 * LT * BOX_BYTES;                    // Q | K | V tiles of one head
  float* lut = reinterpret_cast<float*>(smem + 2 * BUF) + 32;     // [H][2*Lb-1] (+32 floats of slack below index 0)
  const int lut_n = 2 * Lb - 1;
  unsigned long long* masks = reinterpret_cast<unsigned long long*>(smem + 2 * BUF + 128 + (((size_t)H * lut_n * 4 + 15) / 16) * 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (len + TS - 1) / TS;

  auto prefetch = [&](int h, int buf) {
    const uint32_t qb = base + buf * BUF, kb = qb + LT * BOX_BYTES, vb = kb + LT * BOX_BYTES;
    for (int i = tid; i < n_tiles * TS * 8; i += EA_THREADS) {
      const int r = i >> 3, c = i & 7;
      const bool ok = r < len;
      const bf16* src = qkv + (size_t)(row0 + (ok ? r : 0)) * ld + h * DK + c * 8;
      const int tile = r >> 6, rr = r & 63;
      cp_async16(swz(qb + tile * BOX_BYTES, rr, c), src, ok);
      cp_async16(swz(kb + tile * BOX_BYTES, rr, c), src + HD, ok);
      cp_async16(swz(vb + tile * BOX_BYTES, rr, c), src + 2 * HD, ok);
    }
    cp_async_commit();
  };

  prefetch(0, 0);
  for (int i = tid; i < H * lut_n; i += EA_THREADS) lut[i] = bias_lut[i];
  if (warp < n_tiles) {
    const int r0 = warp * TS + lane, r1 = r0 + 32;
    const bool v0 = r0 < len && tok_valid[row0 + r0] != 0;
    const bool v1 = r1 < len && tok_valid[row0 + r1] != 0;
    const unsigned lo = __ballot_sync(0xffffffffu, v0), hi = __ballot_sync(0xffffffffu, v1);
    if (lane == 0) masks[warp] = ((unsigned long long)hi << 32) | lo;
  }
  const int g = lane >> 2, q = lane & 3;
  for (int h = 0; h < H; ++h) {
    const int buf = h & 1;
    if (h + 1 < H) { prefetch(h + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();                                              // head h's tiles (and lut/masks) are visible
    const uint32_t qb = base + buf * BUF, kb = qb + LT * BOX_BYTES, vb = kb + LT * BOX_BYTES;
    for (int q0 = warp * 16; q0 < len; q0 += (EA_THREADS / 32) * 16) {
      uint32_t qf[1][4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = q0 + (lane & 7) + (((lane >> 3) & 1) << 3);
        ldsm_x4(swz(qb + (r >> 6) * BOX_BYTES, r & 63, ks * 2 + (lane >> 4)), qf[0][ks][0], qf[0][ks][1], qf[0][ks][2],
                qf[0][ks][3]);
      }
      float o[1][8][4], m_run[1][2], l_run[1][2];
      m_run[0][0] = m_run[0][1] = -INFINITY;
      l_run[0][0] = l_run[0][1] = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[0][nt][e] = 0.f;
      for (int t = 0; t < n_tiles; ++t) {
        RelBias rb{lut + h * lut_n, t * TS - q0 + Lb - 1};
        flash_tile<1, 8, true>(qf, kb + t * BOX_BYTES, vb + t * BOX_BYTES, 0, masks[t], rb, o, m_run, l_run, lane);
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float l = l_run[0][hf];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        const int r = q0 + g + hf * 8;
        if (r < len) {
          bf16* orow = out + (size_t)(row0 + r) * HD + h * DK;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<uint32_t*>(orow + nt * 8 + 2 * q) = pack_bf16(o[0][nt][hf * 2] * inv, o[0][nt][hf * 2 + 1] * inv);
        }
      }
    }
    __syncthreads();                                              // buffer `buf` may be overwritten by prefetch(h + 2)
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
std::mutex g_mu;
EncodeTiledFn g_encode = nullptr;
std::map<std::tuple<const void*, size_t, size_t>, CUtensorMap> g_maps;

bool get_kv_map(const void* ptr, size_t rows, size_t cols, CUtensorMap* out) {
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
  cuuint32_t box[2] = {(cuuint32_t)DK, (cuuint32_t)TS};
  cuuint32_t estr[2] = {1, 1};
  if (g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
  if (g_maps.size() > 1024) g_maps.clear();
  g_maps[key] = m;
  *out = m;
  return true;
}

}  // namespace fa

bool cross_attention_mma_supported(int K, int H, int dk) { return dk == fa::DK && K <= 64 && (H % 2) == 0; }

cudaError_t cross_attention_mma(const void* q, const void* kv, size_t kv_rows, size_t kv_stride, int k_off, int v_off,
                                const int* ustart, const int* uorder, const uint8_t* tok_valid, void* out, int users,
                                int K, int H, const int* live_start, const int* live_count, int num_sms, cudaStream_t s) {
  if (users <= 0) return cudaSuccess;
  std::lock_guard<std::mutex> lk(fa::g_mu);
  CUtensorMap map;
  if (!fa::get_kv_map(kv, kv_rows, kv_stride, &map)) return cudaErrorUnknown;
  constexpr size_t smem = (size_t)3 * fa::XA_STAGE_TARGET + 1024 + 256;
  static SmemAttr attr[4];
  static const int xa_fast = [] { const char* e = getenv("GRAM_XATTN_NOMASK"); return (e && e[0] == '0') ? 0 : 1; }();   // A/B
  const bool persist = num_sms > 0;          // num_sms <= 0: one CTA per (user, head group), the round-1 kernel (A/B timing)
  if (K <= 32 && (H % 4) == 0) {
    if (persist) {
      const int items = users * (H / 4);
      // GRAM_XATTN_WARPS=8: two consumer warps per head, 16 beam rows each (A/B; measured equal in the step)
      static const bool eight = [] { const char* e = getenv("GRAM_XATTN_WARPS"); return e && e[0] == '8'; }();
      if (eight) {
        auto kern8 = fa::cross_attention_persist_kernel<4, 2, 1>;
        static SmemAttr attr8;
        cudaError_t e8 = attr8.ensure(kern8, smem);
        if (e8 != cudaSuccess) return e8;
        kern8<<<items < num_sms ? items : num_sms, fa::xa_threads(8), smem, s>>>(map, (const bf16*)q, (bf16*)out, ustart, uorder,
                                                                             tok_valid, K, H, k_off, v_off, live_start, live_count, users, xa_fast);
        return cudaGetLastError();
      }
      // up to 24 beams: transposed tiles, beams on the N dimension in units of 8 (GRAM_XATTN_T=0: the beams-on-M kernel, A/B)
      static const bool transposed = [] { const char* e = getenv("GRAM_XATTN_T"); return !(e && e[0] == '0'); }();
      if (transposed && K <= 24) {
        static SmemAttr attr_t[3];
        const int grid = items < num_sms ? items : num_sms;
#define GRAM_XA_T(NTV)                                                                                                  \
        {                                                                                                               \
          auto kt = fa::cross_attention_persist_kernel<4, 1, 2, NTV>;                                                   \
          cudaError_t et = attr_t[NTV - 1].ensure(kt, smem);                                                            \
          if (et != cudaSuccess) return et;                                                                             \
          kt<<<grid, fa::xa_threads(4), smem, s>>>(map, (const bf16*)q, (bf16*)out, ustart, uorder, tok_valid, K, H,    \
                                                   k_off, v_off, live_start, live_count, users, xa_fast);                        \
          return cudaGetLastError();                                                                                    \
        }
        if (K <= 8) GRAM_XA_T(1)
        if (K <= 16) GRAM_XA_T(2)
        GRAM_XA_T(3)
#undef GRAM_XA_T
      }
      auto kern = fa::cross_attention_persist_kernel<4, 1>;
      cudaError_t e = attr[2].ensure(kern, smem);
      if (e != cudaSuccess) return e;
      kern<<<items < num_sms ? items : num_sms, fa::xa_threads(4), smem, s>>>(map, (const bf16*)q, (bf16*)out, ustart, uorder, tok_valid,
                                                                          K, H, k_off, v_off, live_start, live_count, users, xa_fast);
      return cudaGetLastError();
    }
    auto kern = fa::cross_attention_mma_kernel<4, 1, 1, 1>;
    {
      cudaError_t e = attr[0].ensure(kern, smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<dim3(users, H / 4), fa::xa_threads(4), smem, s>>>(map, (const bf16*)q, (bf16*)out, ustart, uorder, tok_valid, K, H,
                                                          k_off, v_off, live_start, live_count);
  } else {
    if (persist) {
      auto kern = fa::cross_attention_persist_kernel<2, 2>;
      cudaError_t e = attr[3].ensure(kern, smem);
      if (e != cudaSuccess) return e;
      const int items = users * (H / 2);
      kern<<<items < num_sms ? items : num_sms, fa::xa_threads(4), smem, s>>>(map, (const bf16*)q, (bf16*)out, ustart, uorder, tok_valid,
                                                                          K, H, k_off, v_off, live_start, live_count, users, xa_fast);
      return cudaGetLastError();
    }
    auto kern = fa::cross_attention_mma_kernel<2, 2, 1, 1>;
    {
      cudaError_t e = attr[1].ensure(kern, smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<dim3(users, H / 2), fa::xa_threads(4), smem, s>>>(map, (const bf16*)q, (bf16*)out, ustart, uorder, tok_valid, K, H,
                                                          k_off, v_off, live_start, live_count);
  }
  return cudaGetLastError();
}

bool enc_attention_mma_supported(int dk, int Lmax) { return dk == fa::DK && Lmax <= 256; }

cudaError_t enc_attention_mma(const void* qkv, void* out, const int* plen, const int* poff, const uint8_t* tok_valid,
                              const float* bias_lut, int Lb, int P, int H, int Lmax, cudaStream_t s) {
  if (P <= 0) return cudaSuccess;
  const int LT = (Lmax + fa::TS - 1) / fa::TS;
  const size_t lut_bytes = ((size_t)H * (2 * Lb - 1) * 4 + 15) / 16 * 16;
  static SmemAttr attr[5];
#define GRAM_EA(LTV)                                                                                         \
  {                                                                                                          \
    const size_t smem = (size_t)2 * 3 * LTV * fa::BOX_BYTES + 128 + lut_bytes + LTV * 8 + 1024 + 64;         \
    auto kern = fa::enc_attention_pipe_kernel<LTV>;                                                          \
    {                                                                                                        \
      cudaError_t e = attr[LTV].ensure(kern, smem);                                                          \
      if (e != cudaSuccess) return e;                                                                        \
    }                                                                                                        \
    kern<<<P, fa::EA_THREADS, smem, s>>>((const bf16*)qkv, (bf16*)out, plen, poff, tok_valid, bias_lut, Lb, H); \
  }
  if (LT <= 1) GRAM_EA(1)
  else if (LT == 2) GRAM_EA(2)
  else if (LT == 3) GRAM_EA(3)
  else GRAM_EA(4)
#undef GRAM_EA
  return cudaGetLastError();
}

}  // namespace gram
