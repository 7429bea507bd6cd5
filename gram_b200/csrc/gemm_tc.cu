// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T, bf16 operands, fp32 accumulation.
//
// Every nn.Linear on the hot path runs through this kernel in bf16 mode (reference
// src/model/gram_t5_modeling.py:305-310,369-372,553-569,622; src/model/gram_t5.py:254): encoder q|k|v, o, wi,
// wo; the per-user cross-attention K/V projection of all decoder layers in one launch; decoder projections; and
// the 32128-wide vocabulary head.  Both operands are K-major (activations [M,K] row-major, nn.Linear weights
// [N,K] row-major), which is the native "TN" form of the 5th-generation tensor core.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor 128x64 bf16 boxes of A and W into a STAGES-deep shared-memory
//               ring with 128-byte swizzle, completion signalled on mbarriers (expect_tx)
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=128, K=16, kind::f16) reading the
//               swizzled tiles through shared-memory descriptors, accumulating in TMEM; tcgen05.commit releases
//               ring slots and publishes finished accumulators
//   warps 2-5   epilogue: tcgen05.ld of the fp32 accumulator (each warp owns its 32-lane TMEM quadrant), fused
//               ReLU / dtype conversion, staged through a 128B-swizzled shared-memory tile and written with TMA
//               bulk tensor stores; the residual add (x += A W^T) is a TMA reduce-add performed at the L2, so the
//               SM never reads the residual stream
// TMEM holds two accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// CTA pairs (CTAS = 2): two CTAs of a (2,1,1) cluster -- the two SMs of a TPC -- compute one 256 x 256 tile with
// tcgen05.mma.cta_group::2.  Each CTA loads its own 128 rows of A and HALF of the W tile (128 of the 256 output
// columns), so a CTA moves 32 KiB of operands per 128x256x64 MACs instead of 48 KiB: the single-CTA kernel is paced
// by the L2 -> SM feed (measured: tensor pipe 67 % busy = 64 of the 96 B/clk it asks for), the pair asks for 64.
// The leader CTA (cluster rank 0) issues every MMA; the TMA loads of both CTAs complete on the leader's `full`
// barrier; tcgen05.commit multicasts the `empty` / `accumulator full` arrivals to both CTAs; both epilogues hand
// the accumulator back on the leader's `accumulator empty` barrier.
// M may live in device memory (*m_ptr): the packed encoder token count is only known on the device, so the
// kernel derives its tile loop bound there and the host never synchronises.
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <map>
#include <tuple>
#include <string>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace gram {

namespace tc {

// Tile: 128 x BN x 64 with BN = 128 or 256.  BN = 256 moves (128 + 256) * 64 * 2 B of operands per 128*256*64 MACs
// (85 MAC/B) instead of 64 MAC/B: the K = 512 GEMMs of this model are L2-bandwidth-bound at 128 x 128, so the wide
// tile is used whenever the problem still yields at least two waves of tiles.
constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int ACC_STAGES = 2;
constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KiB
constexpr uint32_t CSTAGE_BYTES = 32 * 1024;               // epilogue staging: two 128-row x 128-byte boxes
// EPI_RESID_NORM stages the residual stream through the SM: two 32 KiB buffers (x of the next round is prefetched
// while the current round is updated in place) and one 16 KiB box for the bf16 copy; the operand ring keeps 4 stages.
// Measured alternative: the bf16 copy stored straight from registers (no box, 5 stages) is slower on both shapes
// (o projection 1.49 ms vs 1.39 ms, wo 2.35 ms vs 2.27 ms for 1.13 M rows): row-per-thread 16-byte stores are
// half-sector writes, and the K = 2048 GEMM is held by the power cap, not by the ring depth (77 % tensor-pipe
// activity at 1.16 GHz vs 65 % at 1.40 GHz)
constexpr uint32_t NORM_EPI_BYTES = 2 * CSTAGE_BYTES + 16 * 1024;
// EG = epilogue warp groups (4 warps each, one per TMEM lane quadrant).  With K = 512 a 256-column accumulator is
// produced in 8 k-blocks = 4096 tensor-pipe clocks, which one group barely drains (ncu source view of the q|k|v GEMM,
// round 2: the MMA warp waited for a free accumulator, not for operands; tensor pipe 60 % busy; 25 % of the samples on
// un-pipelined tcgen05.ld waits, 16 % on the MEMBAR.GPU of a release-arrive).  Fixed in the epilogue itself (pipelined
// TMEM reads, relaxed hand-back); a second group (EG = 2, alternate 32 KiB rounds through its own staging buffer) is
// what the bf16-stream epilogue uses and an opt-in for the plain ones (see pair_epilogue_groups()).
// MODE 0: plain epilogues, 1: EPI_RESID_NORM (fp32 stream through the SM), 2: EPI_RESID_BF16 (bf16 stream through the SM:
// three 16 KiB boxes per epilogue group -- the box of round r+1 is prefetched and the store of round r-1 may still be
// reading its box while round r is updated in place)
constexpr uint32_t XBOX_BYTES = 16 * 1024;
template <int BN, int CTAS, int MODE = 0, int EG = 1> struct Cfg {
  static constexpr int LOAD_N = BN / CTAS;                 // W rows each CTA loads per stage
  static constexpr uint32_t B_BYTES = LOAD_N * BLOCK_K * 2;   // 16 or 32 KiB
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // MODE 3: EPI_LSE with two epilogue groups -- nothing is staged (the statistics go straight to global memory)
  static constexpr uint32_t EPI_BYTES = MODE == 1 ? NORM_EPI_BYTES : MODE == 2 ? 3 * XBOX_BYTES * EG : MODE == 3 ? 0u : CSTAGE_BYTES * EG;
  static constexpr int STAGES = MODE == 1 ? (STAGE_BYTES == 32768 ? 4 : 2)
                                : MODE == 2 ? (STAGE_BYTES == 32768 ? (EG == 2 ? 4 : 5) : 3)
                                : MODE == 3 ? (STAGE_BYTES == 32768 ? 6 : 4)
                                            : (STAGE_BYTES == 32768 ? (EG == 2 ? 5 : 6) : (EG == 2 ? 3 : 4));
  static constexpr int THREADS = 64 + 128 * EG;
  static constexpr int TMEM_COLS = ACC_STAGES * BN;        // 256 or 512 (power of two)
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 232448, "227 KiB of shared memory per CTA");
  // instruction descriptor: D=f32, A=B=bf16, both K-major, M=128 per CTA (256 for a pair), N=BN
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                    ((uint32_t)((BLOCK_M * CTAS) >> 4) << 24);
};

// RMSNorm folded into the GEMMs around it (see kernels.h: GemmNormAux)
struct NormArgs {
  const float* row_ss;    // consumer: [M][ss_blocks] sums of squares of the A rows; nullptr = no row scale
  float* ss_out;          // producer (EPI_RESID_NORM): [M][N/128]
  const float* ln_w;      // producer: [N] weight of the next RMSNorm
  int ss_blocks;          // consumer: K / 128
  float inv_d, eps;       // consumer: 1 / K, epsilon
};

template <int EPI, int BLOCK_N, int CTAS, int EG = 1>
__global__ void __launch_bounds__(64 + 128 * EG, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_xb,
               float2* __restrict__ lse_partial, int M_imm, const int* __restrict__ m_ptr, int N, int K, NormArgs na) {
  using C_ = Cfg<BLOCK_N, CTAS, EPI == EPI_RESID_NORM ? 1 : EPI == EPI_RESID_BF16 ? 2 : (EPI == EPI_LSE && EG == 2) ? 3 : 0, EG>;
  static_assert(EG == 1 || EPI != EPI_RESID_NORM, "EPI_RESID_NORM runs one epilogue group");
  constexpr int STAGES = C_::STAGES;
  constexpr uint32_t STAGE_BYTES = C_::STAGE_BYTES;
  constexpr int TMEM_COLS = C_::TMEM_COLS;
  constexpr uint32_t kInstrDesc = C_::IDESC;
  constexpr int TILE_M = BLOCK_M * CTAS;                   // rows of one (pair) tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;           // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t cstage = base + STAGES * STAGE_BYTES;    // epilogue staging (1024-aligned)
  const uint32_t bars = cstage + C_::EPI_BYTES;           // full[STAGES], empty[STAGES], tfull[2], tempty[2], xfull[6], tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + ACC_STAGES + s); };
  auto xfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 * ACC_STAGES + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + C_::EPI_BYTES + 8 * (2 * STAGES + 2 * ACC_STAGES + 6));
  static_assert(8 * (2 * STAGES + 2 * ACC_STAGES + 6) + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = m_ptr ? *m_ptr : M_imm;
  const int num_m = (M + TILE_M - 1) / TILE_M, num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / BLOCK_K;
  const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;             // 0 = leader of the pair
  const int tile0 = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = (int)gridDim.x / CTAS;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * CTAS * EG); }
    for (int s = 0; s < 6; ++s) mbar_init(xfull_bar(s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all();      // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (lane == 0) {
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          if (CTAS == 2) {
            // both CTAs' bytes are counted on the leader's barrier
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
            const uint32_t fb = leader_addr(full_bar(stage));
            tma_load_2d_pair(sa, &map_a, fb, kb * BLOCK_K, m_blk * TILE_M + rank * BLOCK_M);
            tma_load_2d_pair(sb, &map_w, fb, kb * BLOCK_K, n_blk * BLOCK_N + rank * C_::LOAD_N);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_2d(sa, &map_a, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(sb, &map_w, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the leader CTA of a pair) =====================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = tile0; tile < num_tiles && rank == 0; tile += tile_step) {
      // the epilogue (of both CTAs) has drained this accumulator
      if (CTAS == 2) mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u);
      else mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);                 // TMA bytes have landed
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte address units
            if (CTAS == 2) umma_bf16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kInstrDesc, (kb | k) ? 1u : 0u);
            else umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kInstrDesc, (kb | k) ? 1u : 0u);
          }
          if (CTAS == 2) {
            umma_commit_pair(empty_bar(stage));            // frees the ring slot in both CTAs
            if (kb == num_kb - 1) umma_commit_pair(tfull_bar(acc));
          } else {
            umma_commit(empty_bar(stage));                 // frees the ring slot when these MMAs retire
            if (kb == num_kb - 1) umma_commit(tfull_bar(acc));   // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ===================== epilogue (warps 2..5; with EG = 2 also warps 6..9) =====================
    const int quad = warp & 3;                             // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;                        // row of the tile this thread owns
    const int grp = EG == 2 ? ((warp - 2) >> 2) : 0;       // epilogue group: takes rounds grp, grp + EG, ...
    const bool issuer = (warp == 2 + 4 * grp && lane == 0);
    const uint32_t gstage = cstage + (uint32_t)grp * CSTAGE_BYTES;   // this group's staging buffer
    constexpr bool kOutBf16 = (EPI == EPI_STORE || EPI == EPI_RELU);
    constexpr int CHUNKS_PER_ROUND = kOutBf16 ? 4 : 2;     // 32 KiB of staging = 128 bf16 or 64 fp32 columns
    constexpr int ROUNDS = (BLOCK_N / 32) / CHUNKS_PER_ROUND;
    static_assert(ROUNDS % EG == 0 || EPI == EPI_RESID_BF16, "every epilogue group takes part in every tile");
    int acc = 0; uint32_t acc_phase = 0;
    // Handing the accumulator back orders nothing but tcgen05 operations (the tcgen05.ld's have completed and
    // tcgen05.fence::before_thread_sync precedes the arrive), so the remote arrive is relaxed: the release form
    // costs MEMBAR.ALL.GPU + ERRBAR per warp per tile (11 % of the q|k|v GEMM's stall samples)
    auto release_acc = [&](int a) {
      if (CTAS == 2) mbar_arrive_cluster_relaxed(leader_addr(tempty_bar(a)));
      else mbar_arrive(tempty_bar(a));
    };
    auto group_bar = [&]() {
      if (EG == 2 && grp == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
      else epi_bar();
    };
    uint32_t xround = 0;                                   // EPI_RESID_NORM: rounds processed so far (buffer = parity)
    if constexpr (EPI == EPI_RESID_NORM) {
      if (issuer && tile0 < num_tiles) {                   // x of the very first round
        mbar_arrive_expect_tx(xfull_bar(0), CSTAGE_BYTES);
        tma_load_2d(cstage, &map_c, xfull_bar(0), (tile0 % num_n) * BLOCK_N, (tile0 / num_n) * TILE_M + rank * BLOCK_M);
        tma_load_2d(cstage + 16384u, &map_c, xfull_bar(0), (tile0 % num_n) * BLOCK_N + 32, (tile0 / num_n) * TILE_M + rank * BLOCK_M);
      }
    }
    if constexpr (EPI == EPI_RESID_BF16) {
      if (issuer && tile0 < num_tiles) {                   // this group's box of the very first round
        mbar_arrive_expect_tx(xfull_bar(grp * 3), XBOX_BYTES);
        tma_load_2d(cstage + (uint32_t)grp * 3u * XBOX_BYTES, &map_c, xfull_bar(grp * 3),
                    (tile0 % num_n) * BLOCK_N + grp * 128, (tile0 / num_n) * TILE_M + rank * BLOCK_M);
      }
    }
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int row0 = m_blk * TILE_M + rank * BLOCK_M;    // first row of this CTA's half of the tile
      // RMSNorm folded in from the producer side: this row of the output is scaled by rsqrt(mean(x^2) + eps) of its
      // input row (fetched while the accumulator is still being computed)
      // Rows past M inside the last tile are written as zeros: with a gain of rsqrt(eps) they would feed garbage back
      // through the residual stream of the next layers (and calls) until it overflows, and the attention kernel reads
      // -- masked, but 0 * inf = nan -- key/value rows past the last passage.
      float rs = 1.f;
      if (kOutBf16 && na.row_ss != nullptr) {
        const int row = row0 + r;
        rs = 0.f;
        if (row < M) {
          float t = 0.f;
          for (int b = 0; b < na.ss_blocks; ++b) t += na.row_ss[(size_t)row * na.ss_blocks + b];
          rs = 1.0f / sqrtf(t * na.inv_d + na.eps);
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      if (EPI == EPI_LSE) {
        // fused log-softmax statistics: per row, (max, sum exp(x - max)) over this tile's columns; the logits
        // themselves are never written (reference computes log_softmax over the materialised [rows, V] logits)
        float mx = -INFINITY, sum = 0.f;
        const int row = row0 + r;
        // (pipelining these TMEM reads as in the store epilogues was measured and is slower here: the vocabulary head went
        //  from 8.3-8.8 to 10.5-11.0 ms per step.  What paces this loop is its ALU issue slots: with one FFMA + MUFU + FADD
        //  per logit and the column select confined to the last tile the head went from 8.8 to 6.8 ms per step)
        // EG = 2: group g reduces the tile's column half g to its own (max, sum) partial
#pragma unroll 1
        for (int c = grp * (BLOCK_N / 32 / EG); c < (grp + 1) * (BLOCK_N / 32 / EG); ++c) {
          uint32_t v[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N + c * 32);
          tmem_ld32(taddr, v);
          tmem_ld_wait();
          const int col0 = n_blk * BLOCK_N + c * 32;
          float cm = -INFINITY;
          if (col0 + 32 > N) {                           // only the last column tile has columns past N (warp-uniform)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (col0 + e >= N) v[e] = __float_as_uint(-INFINITY);
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) cm = fmaxf(cm, __uint_as_float(v[e]));
          const float nm = fmaxf(mx, cm);
          if (nm > -INFINITY) {
            // one FFMA + one MUFU + one FADD per logit: 2^(x log2e - m log2e)
            const float nml = nm * 1.4426950408889634f;
            float cs = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) cs += ex2_ftz(fmaf(__uint_as_float(v[e]), 1.4426950408889634f, -nml));
            sum = sum * ex2_ftz((mx - nm) * 1.4426950408889634f) + cs;
            mx = nm;
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
        if (row < M) lse_partial[((size_t)row * num_n + n_blk) * EG + grp] = make_float2(mx, sum);
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
      if constexpr (EPI == EPI_RESID_NORM) {
        // x += A W^T with the next RMSNorm's input produced on the way: per round (64 columns) the old x tile arrives
        // by TMA (prefetched one round ahead into the other buffer), every thread updates its row in place, and
        // x (fp32), bf16(x * ln_w) and the row's sum of squares per 128-column block leave the SM.  The consumer GEMM
        // multiplies its output rows by rsqrt(mean x^2 + eps): y = ((x*w) W^T) * r = (w * x * r) W^T.
        float ssq = 0.f;
        const int row = row0 + r;
        const bool live_row = row < M;                     // rows past M in the last tile are kept at zero
#pragma unroll 1
        for (int rd = 0; rd < ROUNDS; ++rd, ++xround) {
          const uint32_t xb_cur = cstage + (xround & 1u) * CSTAGE_BYTES, xb_nxt = cstage + ((xround & 1u) ^ 1u) * CSTAGE_BYTES;
          if (issuer) {
            tma_store_wait_read();                         // every earlier store has read its staging buffer
            int nt = tile, nrd = rd + 1;                   // (tile, round) after this one
            if (nrd == ROUNDS) { nt = tile + tile_step; nrd = 0; }
            if (nt < num_tiles) {
              const int ncol = (nt % num_n) * BLOCK_N + nrd * 64, nrow = (nt / num_n) * TILE_M + rank * BLOCK_M;
              mbar_arrive_expect_tx(xfull_bar((xround & 1u) ^ 1u), CSTAGE_BYTES);
              tma_load_2d(xb_nxt, &map_c, xfull_bar((xround & 1u) ^ 1u), ncol, nrow);
              tma_load_2d(xb_nxt + 16384u, &map_c, xfull_bar((xround & 1u) ^ 1u), ncol + 32, nrow);
            }
          }
          uint32_t vv[2][32];                              // TMEM reads software-pipelined over the round's two chunks
          const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N + rd * 64);
          tmem_ld32(taddr0, vv[0]);
          epi_bar();                                       // the bf16 box is free again
          mbar_wait(xfull_bar(xround & 1u), (xround >> 1) & 1u);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = rd * 2 + cc;
            tmem_ld_wait();
            if (cc == 0) tmem_ld32(taddr0 + 32u, vv[1]);
            uint32_t (&v)[32] = vv[cc];
            const uint32_t xrow = xb_cur + (uint32_t)cc * 16384u + (uint32_t)r * 128u;
            const uint32_t brow = cstage + 2 * CSTAGE_BYTES + (uint32_t)r * 128u;
            const float* gw = na.ln_w + n_blk * BLOCK_N + c * 32;
#pragma unroll
            for (int g2 = 0; g2 < 4; ++g2) {
              uint32_t pk[4];
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int g = g2 * 2 + hh;
                const uint32_t addr = xrow + (uint32_t)((g ^ (r & 7)) << 4);
                float4 xo;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(xo.x), "=f"(xo.y), "=f"(xo.z), "=f"(xo.w) : "r"(addr) : "memory");
                xo.x = live_row ? xo.x + __uint_as_float(v[g * 4]) : 0.f;
                xo.y = live_row ? xo.y + __uint_as_float(v[g * 4 + 1]) : 0.f;
                xo.z = live_row ? xo.z + __uint_as_float(v[g * 4 + 2]) : 0.f;
                xo.w = live_row ? xo.w + __uint_as_float(v[g * 4 + 3]) : 0.f;
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(xo.x), "f"(xo.y), "f"(xo.z), "f"(xo.w) : "memory");
                ssq = fmaf(xo.x, xo.x, ssq); ssq = fmaf(xo.y, xo.y, ssq); ssq = fmaf(xo.z, xo.z, ssq); ssq = fmaf(xo.w, xo.w, ssq);
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(gw + g * 4));
                __nv_bfloat162 h0 = __floats2bfloat162_rn(xo.x * w4.x, xo.y * w4.y), h1 = __floats2bfloat162_rn(xo.z * w4.z, xo.w * w4.w);
                pk[hh * 2] = *reinterpret_cast<uint32_t*>(&h0);
                pk[hh * 2 + 1] = *reinterpret_cast<uint32_t*>(&h1);
              }
              const uint32_t piece = (uint32_t)((cc * 4 + g2) ^ (r & 7));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(brow + (piece << 4)), "r"(pk[0]), "r"(pk[1]),
                           "r"(pk[2]), "r"(pk[3]) : "memory");
            }
          }
          if ((rd & 1) == 1) {
            // rounds 2b and 2b+1 cover 128-column block b of the tile: one partial per (row, block), owned by this
            // thread alone -- no atomics, and the same summation order whatever the tile shape
            if (row < M && n_blk * (BLOCK_N >> 7) + (rd >> 1) < (N >> 7))
              na.ss_out[(size_t)row * (N >> 7) + n_blk * (BLOCK_N >> 7) + (rd >> 1)] = ssq;
            ssq = 0.f;
          }
          if (rd == ROUNDS - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
          fence_async_smem();
          epi_bar();
          if (issuer) {
            const int col0 = n_blk * BLOCK_N + rd * 64;
            if (col0 < N) {
              tma_store_2d(&map_c, xb_cur, col0, row0);
              tma_store_2d(&map_xb, cstage + 2 * CSTAGE_BYTES, col0, row0);
            }
            if (col0 + 32 < N) tma_store_2d(&map_c, xb_cur + 16384u, col0 + 32, row0);
            tma_store_commit();
          }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
      if constexpr (EPI == EPI_RESID_BF16) {
        // bf16 residual stream updated in place: per round one 128-row x 64-column box of the stream arrives by TMA
        // (prefetched a round ahead), thread = row adds its accumulator slice, rounds to bf16, sums the squares of the
        // ROUNDED values (what the consumer GEMM will read) and the box goes back by TMA.  Group g owns the tile's
        // 128-column block g, i.e. one sum-of-squares partial per (row, block) -- same owner rule as EPI_RESID_NORM.
        constexpr int RPG = (BLOCK_N / 64) / EG;           // rounds per group per tile (2)
        static_assert(RPG == 2, "a group owns one 128-column block of the tile");
        const uint32_t gx = cstage + (uint32_t)grp * 3u * XBOX_BYTES;
        const int row = row0 + r;
        const bool live_row = row < M;
        float ssq = 0.f;
#pragma unroll 1
        for (int rl = 0; rl < RPG; ++rl, ++xround) {
          const uint32_t bsel = xround % 3u, nsel = (xround + 1u) % 3u;
          const uint32_t xb_cur = gx + bsel * XBOX_BYTES;
          const int colt = (grp * RPG + rl) * 64;          // first column of this round inside the tile
          if (issuer) {
            tma_store_wait_read1();                        // the store of round r-2 has read the box round r+1 lands in
            int nt = tile, nrl = rl + 1;
            if (nrl == RPG) { nt = tile + tile_step; nrl = 0; }
            if (nt < num_tiles) {
              const int ncol = (nt % num_n) * BLOCK_N + (grp * RPG + nrl) * 64, nrow = (nt / num_n) * TILE_M + rank * BLOCK_M;
              mbar_arrive_expect_tx(xfull_bar(grp * 3 + nsel), XBOX_BYTES);
              tma_load_2d(gx + nsel * XBOX_BYTES, &map_c, xfull_bar(grp * 3 + nsel), ncol, nrow);
            }
          }
          uint32_t v[2][32];
          const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N + colt);
          tmem_ld32(taddr0, v[0]);
          mbar_wait(xfull_bar(grp * 3 + bsel), (xround / 3u) & 1u);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            tmem_ld_wait();
            if (cc == 0) tmem_ld32(taddr0 + 32u, v[1]);
            uint32_t (&w)[32] = v[cc];
            const uint32_t xrow = xb_cur + (uint32_t)r * 128u;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t addr = xrow + (uint32_t)(((cc * 4 + g) ^ (r & 7)) << 4);
              uint32_t o[4];
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(addr) : "memory");
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 xo = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&o[e]));
                __nv_bfloat162 h2 = __floats2bfloat162_rn(xo.x + __uint_as_float(w[g * 8 + 2 * e]), xo.y + __uint_as_float(w[g * 8 + 2 * e + 1]));
                if (!live_row) h2 = __floats2bfloat162_rn(0.f, 0.f);
                const float2 xn2 = __bfloat1622float2(h2);
                ssq = fmaf(xn2.x, xn2.x, ssq); ssq = fmaf(xn2.y, xn2.y, ssq);
                o[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
            }
          }
          if (rl == RPG - 1) {
            if (live_row && n_blk * (BLOCK_N >> 7) + grp < (N >> 7))   // N = 128 (mod 256): the tile's second block does not exist
              na.ss_out[(size_t)row * (N >> 7) + n_blk * (BLOCK_N >> 7) + grp] = ssq;
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
          fence_async_smem();
          group_bar();
          if (issuer) {
            const int col0 = n_blk * BLOCK_N + colt;
            if (col0 < N) tma_store_2d(&map_c, xb_cur, col0, row0);
            tma_store_commit();
          }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
#pragma unroll 1
      for (int rd = grp; rd < ROUNDS; rd += EG) {
        if (issuer) tma_store_wait_read();                 // previous bulk store has finished reading the staging tile
        group_bar();
        // software-pipelined TMEM reads: chunk cc + 1 is in flight while chunk cc is converted and staged
        uint32_t v[2][32];
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N + rd * CHUNKS_PER_ROUND * 32);
        tmem_ld32(taddr0, v[0]);
#pragma unroll
        for (int cc = 0; cc < CHUNKS_PER_ROUND; ++cc) {
          const int c = rd * CHUNKS_PER_ROUND + cc;        // 32-column chunk of the accumulator
          tmem_ld_wait();
          if (cc + 1 < CHUNKS_PER_ROUND) tmem_ld32(taddr0 + (uint32_t)((cc + 1) * 32), v[(cc + 1) & 1]);
          uint32_t (&w)[32] = v[cc & 1];
          if (kOutBf16) {
            // box (cc/2): 128 rows x 64 bf16 (128 B per row); this chunk is the 16-byte pieces (c%2)*4 .. +3
            const uint32_t box = gstage + (uint32_t)(cc >> 1) * 16384u + (uint32_t)r * 128u;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a = __uint_as_float(w[g * 8 + 2 * e]) * rs, b = __uint_as_float(w[g * 8 + 2 * e + 1]) * rs;
                if (EPI == EPI_RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                pk[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              const uint32_t piece = (uint32_t)(((c & 1) * 4 + g) ^ (r & 7));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + (piece << 4)), "r"(pk[0]), "r"(pk[1]),
                           "r"(pk[2]), "r"(pk[3]) : "memory");
            }
          } else {
            // box cc: 128 rows x 32 fp32 (128 B per row)
            const uint32_t box = gstage + (uint32_t)cc * 16384u + (uint32_t)r * 128u;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const uint32_t piece = (uint32_t)(g ^ (r & 7));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + (piece << 4)), "r"(w[g * 4]),
                           "r"(w[g * 4 + 1]), "r"(w[g * 4 + 2]), "r"(w[g * 4 + 3]) : "memory");
            }
          }
        }
        if (rd + EG >= ROUNDS) {
          // every tcgen05.ld of this group's share of the accumulator has completed: hand it back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
        }
        fence_async_smem();                                // generic-proxy writes -> visible to the async proxy
        group_bar();
        if (issuer) {
#pragma unroll
          for (int bx = 0; bx < 2; ++bx) {
            const int col0 = n_blk * BLOCK_N + (kOutBf16 ? (rd * 2 + bx) * 64 : (rd * 2 + bx) * 32);
            if (col0 < N) {
              if (EPI == EPI_RESID) tma_reduce_add_2d(&map_c, gstage + bx * 16384u, col0, row0);
              else tma_store_2d(&map_c, gstage + bx * 16384u, col0, row0);
            }
          }
          tma_store_commit();
        }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
    if (issuer) tma_store_wait_all();
  }
  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all();      // neither CTA leaves while the pair's MMAs / barrier arrivals are in flight
  else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::mutex g_mu;
EncodeTiledFn g_encode = nullptr;
std::string g_err;
std::map<std::tuple<const void*, int, long long>, CUtensorMap> g_maps;
SmemAttr g_attr[42];

bool get_encode() {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    g_err = std::string("cuTensorMapEncodeTiled entry point unavailable: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return false;
  }
  g_encode = (EncodeTiledFn)fn;
  return true;
}

// 2-D row-major [rows, cols] tensor map with a (box_rows x 128 bytes) box and 128-byte swizzle.
// kind 0: bf16 operand/output (box 64 columns), kind 1: fp32 output (box 32 columns)
bool get_map(const void* ptr, int rows, int cols, int kind, int box_rows, CUtensorMap* out) {
  auto key = std::make_tuple(ptr, rows, ((long long)cols * 2 + kind) * 4 + (box_rows >> 7));
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return true; }
  if (!get_encode()) return false;
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * (kind ? 4 : 2)};
  cuuint32_t box[2] = {(cuuint32_t)(kind ? 32 : 64), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(&m, kind ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return false;
  }
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[key] = m;
  *out = m;
  return true;
}

// resident CTA pairs the device can hold for one instantiation (cached per device)
template <typename Kern>
int max_pairs(Kern kern, size_t smem, int threads, int num_sms, int* cache) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (cache[dev] == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms & ~1);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = -1; }
    cache[dev] = n;
  }
  return cache[dev];
}

int g_pairs[42][64];

template <int EPI, int BN, int CTAS, int EG = 1>
cudaError_t launch(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mc, const CUtensorMap& mxb,
                   float2* lse_partial, int M_max, const int* m_ptr, int N, int K, const NormArgs& na, int num_sms,
                   cudaStream_t s) {
  auto kern = gemm_tc_kernel<EPI, BN, CTAS, EG>;
  using C_ = Cfg<BN, CTAS, EPI == EPI_RESID_NORM ? 1 : EPI == EPI_RESID_BF16 ? 2 : (EPI == EPI_LSE && EG == 2) ? 3 : 0, EG>;
  constexpr size_t smem = C_::SMEM_BYTES;
  constexpr int threads = C_::THREADS;
  constexpr int slot = EPI * 3 + (BN == 256) + (CTAS == 2) + 21 * (EG - 1);
  {
    cudaError_t e = g_attr[slot].ensure(kern, smem);
    if (e != cudaSuccess) return e;
  }
  const int tiles = ((M_max + BLOCK_M * CTAS - 1) / (BLOCK_M * CTAS)) * ((N + BN - 1) / BN);
  if (CTAS == 1) {
    const int grid = tiles < num_sms ? tiles : num_sms;
    kern<<<grid, threads, smem, s>>>(ma, mw, mc, mxb, lse_partial, M_max, m_ptr, N, K, na);
    return cudaGetLastError();
  }
  const int pairs = max_pairs(kern, smem, threads, num_sms, g_pairs[slot]);
  if (pairs <= 0) return cudaErrorLaunchOutOfResources;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, ma, mw, mc, mxb, lse_partial, M_max, m_ptr, N, K, na);
}

// GRAM_GEMM_EG=2 runs TWO epilogue groups on the CTA-pair tiles of the plain epilogues (A/B).  Measured (round 2, one call,
// M = 1 M rows): q|k|v 1153 vs 1229 TFLOP/s, wi 1170 vs 1241 with ONE group -- the fifth ring stage the second staging
// buffer costs and the 128 extra threads lose more than the second group gains once the TMEM reads are pipelined and the
// accumulator hand-back is a relaxed arrive; one group is the default
inline int pair_epilogue_groups() {
  static int v = 0;
  if (v == 0) { const char* e = getenv("GRAM_GEMM_EG"); v = (e && e[0] == '2') ? 2 : 1; }
  return v;
}

// Tile shape: 128 x 128 for small problems; 128 x 256 when that still gives every SM at least two tiles; a CTA pair
// (256 x 256) when pairs are allowed (max_ctas >= 2) and every SM gets at least 16 tiles -- below that the cluster
// launch + cluster barriers cost more than the pair saves (measured: M = 18880, N = 2048: 39.2 us vs 37.2 us).
struct Shape { int bn, ctas; };
inline Shape pick_shape(int M_max, int N, int num_sms, int max_ctas) {
  if (N < 256) return {128, 1};
  const long long tiles256 = (long long)((M_max + BLOCK_M - 1) / BLOCK_M) * ((N + 255) / 256);
  if (tiles256 < 2LL * num_sms) return {128, 1};
  static const long long pair_tiles = [] { const char* e = getenv("GRAM_PAIR_TILES"); return e ? atoll(e) : 16LL; }();   // A/B
  return {256, (max_ctas >= 2 && tiles256 >= pair_tiles * num_sms) ? 2 : 1};
}

}  // namespace tc

bool gemm_tc_supported(int N, int K) { return (K % tc::BLOCK_K) == 0 && (N % 16) == 0 && N >= 16 && K >= tc::BLOCK_K; }

const char* gemm_tc_last_error() { return tc::g_err.c_str(); }
namespace tc { const char* last_error() { return g_err.c_str(); } std::mutex& mutex() { return g_mu; } }

namespace tc {
// GRAM_LSE_EG=2: two epilogue groups on the 128 x 256 tiles of the vocabulary head (A/B)
inline int lse_groups() {
  static int v = 0;
  if (v == 0) { const char* e = getenv("GRAM_LSE_EG"); v = (e && e[0] == '2') ? 2 : 1; }
  return v;
}
}  // namespace tc

int gemm_tc_lse_ntiles(int M_max, int N, int num_sms) {
  const int bn = tc::pick_shape(M_max, N, num_sms, 1).bn;
  return ((N + bn - 1) / bn) * (bn == 256 ? tc::lse_groups() : 1);
}

cudaError_t gemm_tc(int epi, const void* A, const void* W, void* C, int M_max, const int* m_ptr, int N, int K,
                    int num_sms, int max_ctas, cudaStream_t s, const GemmNormAux* aux) {
  if (M_max <= 0) return cudaSuccess;
  if (!gemm_tc_supported(N, K)) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(tc::g_mu);
  CUtensorMap ma, mw, mc;
  const int ckind = (epi == EPI_STORE || epi == EPI_RELU || epi == EPI_RESID_BF16) ? 0 : 1;
  tc::NormArgs na = {};
  if (aux && aux->row_ss) {
    if ((epi != EPI_STORE && epi != EPI_RELU) || (K & 127)) return cudaErrorInvalidValue;
    na.row_ss = aux->row_ss; na.ss_blocks = K >> 7; na.inv_d = 1.0f / (float)K; na.eps = aux->eps;
  }
  if (epi == EPI_RESID_NORM) {
    // x += A W^T, plus bf16((x) * ln_w) and per-128-column sums of squares: 128-column single-CTA tiles, or CTA pairs
    if (!aux || !aux->xb || !aux->ss_out || !aux->ln_w || (N & 127)) return cudaErrorInvalidValue;
    na.ss_out = aux->ss_out; na.ln_w = aux->ln_w;
    const bool pairs = tc::pick_shape(M_max, N, num_sms, max_ctas).ctas == 2;
    CUtensorMap mxb;
    if (!tc::get_map(A, M_max, K, 0, tc::BLOCK_M, &ma) || !tc::get_map(W, N, K, 0, 128, &mw) ||
        !tc::get_map(C, M_max, N, 1, tc::BLOCK_M, &mc) || !tc::get_map(aux->xb, M_max, N, 0, tc::BLOCK_M, &mxb))
      return cudaErrorUnknown;
    return pairs ? tc::launch<EPI_RESID_NORM, 256, 2>(ma, mw, mc, mxb, nullptr, M_max, m_ptr, N, K, na, num_sms, s)
                 : tc::launch<EPI_RESID_NORM, 128, 1>(ma, mw, mc, mxb, nullptr, M_max, m_ptr, N, K, na, num_sms, s);
  }
  if (epi == EPI_RESID_BF16) {
    // bf16 residual stream updated in place (kernels.h): 128-column single-CTA tiles, or CTA pairs with two epilogue groups
    if (!aux || !aux->ss_out || (N & 127)) return cudaErrorInvalidValue;
    na.ss_out = aux->ss_out;
    const bool pairs = tc::pick_shape(M_max, N, num_sms, max_ctas).ctas == 2;
    if (!tc::get_map(A, M_max, K, 0, tc::BLOCK_M, &ma) || !tc::get_map(W, N, K, 0, 128, &mw) ||
        !tc::get_map(C, M_max, N, 0, tc::BLOCK_M, &mc))
      return cudaErrorUnknown;
    return pairs ? tc::launch<EPI_RESID_BF16, 256, 2, 2>(ma, mw, mc, mc, nullptr, M_max, m_ptr, N, K, na, num_sms, s)
                 : tc::launch<EPI_RESID_BF16, 128, 1, 1>(ma, mw, mc, mc, nullptr, M_max, m_ptr, N, K, na, num_sms, s);
  }
  // the fused log-softmax epilogue stays on single-CTA tiles: it is exp2-bound, and making the leader wait for the
  // slower of two epilogues cost 11 % on the vocabulary head (measured)
  static const bool lse_pairs = [] { const char* e = getenv("GRAM_LSE_PAIRS"); return e && e[0] == '1'; }();   // A/B
  const tc::Shape sh = tc::pick_shape(M_max, N, num_sms, (epi == EPI_LSE && !lse_pairs) ? 1 : max_ctas);
  // the W box is the rows ONE CTA loads per stage
  if (!tc::get_map(A, M_max, K, 0, tc::BLOCK_M, &ma) || !tc::get_map(W, N, K, 0, sh.bn / sh.ctas, &mw)) return cudaErrorUnknown;
#define GRAM_TC_LAUNCH(E, MC, LP)                                                                              \
  return sh.ctas == 2 ? tc::launch<E, 256, 2>(ma, mw, MC, MC, LP, M_max, m_ptr, N, K, na, num_sms, s)          \
         : sh.bn == 256 ? tc::launch<E, 256, 1>(ma, mw, MC, MC, LP, M_max, m_ptr, N, K, na, num_sms, s)        \
                        : tc::launch<E, 128, 1>(ma, mw, MC, MC, LP, M_max, m_ptr, N, K, na, num_sms, s)
#define GRAM_TC_LAUNCH_EG(E, MC)                                                                               \
  if (sh.ctas == 2 && tc::pair_epilogue_groups() == 2)                                                         \
    return tc::launch<E, 256, 2, 2>(ma, mw, MC, MC, nullptr, M_max, m_ptr, N, K, na, num_sms, s);              \
  GRAM_TC_LAUNCH(E, MC, nullptr)
  if (epi == EPI_LSE) {
    if (sh.bn == 256 && sh.ctas == 1 && tc::lse_groups() == 2)
      return tc::launch<EPI_LSE, 256, 1, 2>(ma, mw, ma, ma, (float2*)C, M_max, m_ptr, N, K, na, num_sms, s);
    GRAM_TC_LAUNCH(EPI_LSE, ma, (float2*)C);
  }
  if (!tc::get_map(C, M_max, N, ckind, tc::BLOCK_M, &mc)) return cudaErrorUnknown;
  switch (epi) {
    case EPI_STORE: { GRAM_TC_LAUNCH_EG(EPI_STORE, mc); }
    case EPI_RELU: { GRAM_TC_LAUNCH_EG(EPI_RELU, mc); }
    case EPI_RESID: { GRAM_TC_LAUNCH_EG(EPI_RESID, mc); }
    case EPI_F32: { GRAM_TC_LAUNCH_EG(EPI_F32, mc); }
    default: return cudaErrorInvalidValue;
  }
#undef GRAM_TC_LAUNCH_EG
#undef GRAM_TC_LAUNCH
}

}  // namespace gram
