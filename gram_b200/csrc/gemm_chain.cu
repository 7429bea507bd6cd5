// Row-block chain of one encoder layer's residual sublayers, bf16, sm_100a (tcgen05 / TMEM / TMA):
//
//     x  += ao W_o^T ;   xn = bf16(x * ln1) , ss = row sums of squares           (attention output projection + RMSNorm)
//     ff  = relu((xn W_i^T) * rsqrt(mean x^2 + eps))                             (T5LayerFF, first half)
//     x  += ff W_o2^T ;  xn = bf16(x * ln0') , ss                                (second half + the next layer's RMSNorm)
//
// (reference T5LayerSelfAttention / T5LayerFF / T5DenseActDense, src/model/gram_t5_modeling.py:297-310,337-352,634-667).
// As three separate GEMM launches (gemm_tc.cu) the [tokens, d_ff] activation `ff` is written to and read back from HBM
// (8 KB per token per layer at T5-small), `xn` makes the same round trip and the residual stream is streamed twice:
// 19 of the 29 KB/token/layer the encoder moves, on a path that sits exactly on the machine's FLOP/byte ridge.
//
// Here ONE persistent CTA per SM owns a 128-token row block at a time and runs the three GEMMs of that block back to
// back -- o-projection tiles, then the d_ff / 256 first-half tiles, then the second-half tiles -- handing `ff` and `xn`
// from one GEMM to the next through a per-CTA scratch (128 x d_ff bf16, 512 KiB, same addresses for every block) that
// stays resident in the 126 MB L2: the data is written by TMA stores, read back a few microseconds later by the TMA
// loads of the same CTA, and overwritten by the next block before anything evicts it.  All hand-offs are INTRA-CTA
// (a shared-memory completion counter between the epilogue's store issuer and the TMA producer warp): no CTA ever
// waits for another, so there is nothing to deadlock on.  The o-projection tiles of block i+1 are issued between the
// first-half and second-half tiles of block i, which keeps every dependency at least two tiles behind the tensor pipe.
//
// TMEM cannot hold the [128 x d_model] fp32 output tile and a hidden-chunk accumulator at once (128 x 512 fp32 is the
// whole 256 KB), so a register/TMEM-resident fusion of the two FF GEMMs is not possible at d_model = 512 without
// recomputing the first GEMM; chaining through L2 keeps both GEMMs at full tile size instead.
//
// Warp roles as in gemm_tc.cu: warp 0 = TMA producer (3-stage ring of 128x64 A + 256x64 W boxes, 128B swizzle),
// warp 1 = single-thread tcgen05.mma issuer (128 x 256 x 16, two TMEM accumulators), warps 2-5 = epilogue
// (tcgen05.ld -> registers -> swizzled shared memory -> TMA store; the residual tiles stream x through the SM by TMA,
// prefetched one 64-column round ahead, and emit the folded RMSNorm operands exactly like EPI_RESID_NORM).
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace gram {
namespace chain {

using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int STAGES = 3;
constexpr int THREADS = 192;
constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;   // 16 + 32 KiB
constexpr uint32_t XS_BYTES = 32 * 1024;          // one staging buffer: 128 rows x 64 fp32 (x) or x 128 bf16 (ff)
constexpr uint32_t XB_BYTES = 16 * 1024;          // bf16(x * w) box of a residual round
constexpr uint32_t EPI_BYTES = 2 * XS_BYTES + XB_BYTES;
constexpr int TMEM_COLS = 512;                    // two 256-column accumulators
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr int KB_PER_TILE = BN / BK;              // k-blocks of a consumer that one producer tile (256 columns) covers
constexpr int NORM_ROUNDS = BN / 64;

enum { T_O = 0, T_WI = 1, T_WO = 2 };

struct Args {
  int M_imm; const int* m_ptr;
  int D, HD, F;
  float* ss;               // [M][D/128]: written by the residual tiles, read by the first-half tiles of the same rows
  const float* ln_mid;     // [D] RMSNorm weight between attention and FF
  const float* ln_next;    // [D] first RMSNorm weight of the next layer; nullptr = the last layer (x only)
  float eps;
  int* err;                // sticky device error flag (7 = watchdog: a barrier of this kernel never completed)
  int hints;               // 1 = L2 cache-policy hints on the TMA traffic
};

struct Tile { int type, it, sub; };

// Program of one CTA over its n_iter row blocks (nd = D / 256 residual tiles, nf = F / 256 first-half tiles):
//   O(0)[0..nd) | for it: WI(it)[0..nf)  O(it+1)[0..nd) (if there is a next block)  WO(it)[0..nd)
__device__ __forceinline__ Tile decode(int seq, int n_iter, int nd, int nf) {
  if (seq < nd) return {T_O, 0, seq};
  const int per = nf + 2 * nd, s = seq - nd;
  const int it = s / per, r = s - it * per;
  if (r < nf) return {T_WI, it, r};
  if (it + 1 < n_iter) {
    if (r < nf + nd) return {T_O, it + 1, r - nf};
    return {T_WO, it, r - nf - nd};
  }
  return {T_WO, it, r - nf};
}
__device__ __forceinline__ int total_tiles(int n_iter, int nd, int nf) {
  return n_iter > 0 ? nd + (n_iter - 1) * (nf + 2 * nd) + nf + nd : 0;
}
__device__ __forceinline__ int idx_O(int it, int n, int nd, int nf) { return it == 0 ? n : nd + (it - 1) * (nf + 2 * nd) + nf + n; }
__device__ __forceinline__ int idx_WI(int it, int j, int nd, int nf) { return nd + it * (nf + 2 * nd) + j; }
// number of leading tiles of the program whose stores must have completed before k-block kb of tile t may be loaded
__device__ __forceinline__ int need_of(const Tile& t, int kb, int nd, int nf) {
  if (t.type == T_WI) return idx_O(t.it, kb / KB_PER_TILE, nd, nf) + 1;
  if (t.type == T_WO) return idx_WI(t.it, kb / KB_PER_TILE, nd, nf) + 1;
  return 0;
}

// mbarrier wait with a watchdog: a logic error must end the kernel, not hang the GPU
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, int* err) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if (++spins > (1u << 24)) { atomicExch(err, 7); __trap(); }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
enc_chain_kernel(const __grid_constant__ CUtensorMap map_ao, const __grid_constant__ CUtensorMap map_wo_attn,
                 const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_xn,
                 const __grid_constant__ CUtensorMap map_wi, const __grid_constant__ CUtensorMap map_ffs,
                 const __grid_constant__ CUtensorMap map_wo, Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t xs0 = base + STAGES * STAGE_BYTES;            // staging buffers 0 / 1, then the bf16 box
  const uint32_t xbox = xs0 + 2 * XS_BYTES;
  const uint32_t bars = xs0 + EPI_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };
  auto xfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + 4 + s); };
  uint8_t* bar_area = smem + STAGES * STAGE_BYTES + EPI_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_area + 8 * (2 * STAGES + 6));
  volatile uint32_t* done_ctr = reinterpret_cast<volatile uint32_t*>(bar_area + 8 * (2 * STAGES + 6) + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = a.m_ptr ? *a.m_ptr : a.M_imm;
  const int num_m = (M + BM - 1) / BM;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int n_iter = cta < num_m ? (num_m - cta + G - 1) / G : 0;
  const int nd = a.D / BN, nf = a.F / BN;
  const int total = total_tiles(n_iter, nd, nf);
  const int kb_o = a.HD / BK, kb_wi = a.D / BK, kb_wo = a.F / BK;
  const uint64_t pol_keep = a.hints ? kEvictLast : kEvictNormal, pol_stream = a.hints ? kEvictFirst : kEvictNormal;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ao) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wo_attn) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xn) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ffs) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wo) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); mbar_init(xfull_bar(s), 1); }
    *done_ctr = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    int seen = 0;                                             // tiles known to be complete
    for (int seq = 0; seq < total; ++seq) {
      const Tile t = decode(seq, n_iter, nd, nf);
      const int m_blk = cta + t.it * G;
      const CUtensorMap* ma = t.type == T_O ? &map_ao : (t.type == T_WI ? &map_xn : &map_ffs);
      const CUtensorMap* mw = t.type == T_O ? &map_wo_attn : (t.type == T_WI ? &map_wi : &map_wo);
      const int num_kb = t.type == T_O ? kb_o : (t.type == T_WI ? kb_wi : kb_wo);
      const int a_row = t.type == T_WO ? cta * BM : m_blk * BM;     // the ff scratch is per CTA
      const uint64_t pol_a = t.type == T_O ? pol_stream : pol_keep;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int need = need_of(t, kb, nd, nf);
        if (need > seen) {
          // the operand is produced by an earlier tile of THIS CTA: wait until its TMA stores have completed
          if (lane == 0) {
            uint32_t spins = 0;
            int v;
            while ((v = (int)*done_ctr) < need) {
              __nanosleep(64);
              if (++spins > (1u << 24)) { atomicExch(a.err, 7); __trap(); }
            }
            seen = v;
            fence_proxy_async_all();                          // generic-proxy observation -> later async-proxy (TMA) reads
          }
          seen = __shfl_sync(0xffffffffu, seen, 0);
        }
        mbar_wait_wd(empty_bar(stage), phase ^ 1u, a.err);
        if (lane == 0) {
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d_hint(sa, ma, full_bar(stage), kb * BK, a_row, pol_a);
          tma_load_2d_hint(sb, mw, full_bar(stage), kb * BK, t.sub * BN, pol_keep);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int seq = 0; seq < total; ++seq) {
      const Tile t = decode(seq, n_iter, nd, nf);
      const int num_kb = t.type == T_O ? kb_o : (t.type == T_WI ? kb_wi : kb_wo);
      mbar_wait_wd(tempty_bar(acc), acc_phase ^ 1u, a.err);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_wd(full_bar(stage), phase, a.err);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (kb | k) ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const bool issuer = (warp == 2 && lane == 0);
    int acc = 0; uint32_t acc_phase = 0;
    int published = 0;                                        // issuer: tiles announced complete to the producer warp
    uint32_t xround = 0;                                      // residual rounds consumed so far; round q lives in staging buffer q & 1
    const int ssb = a.D >> 7;
    const float inv_d = 1.0f / (float)a.D;
    auto xs = [&](uint32_t q) { return xs0 + (q & 1u) * XS_BYTES; };
    // x of residual round (seq, rd): two 32-column fp32 boxes
    auto request_x = [&](int seq, int rd, uint32_t q) {
      const Tile t = decode(seq, n_iter, nd, nf);
      const int col = t.sub * BN + rd * 64, row = (cta + t.it * G) * BM;
      mbar_arrive_expect_tx(xfull_bar(q & 1u), XS_BYTES);
      tma_load_2d_hint(xs(q), &map_x, xfull_bar(q & 1u), col, row, t.type == T_O ? pol_stream : pol_keep);
      tma_load_2d_hint(xs(q) + 16384u, &map_x, xfull_bar(q & 1u), col + 32, row, t.type == T_O ? pol_stream : pol_keep);
    };
    if (issuer && total > 0) request_x(0, 0, 0);              // the program starts with a residual tile
    for (int seq = 0; seq < total; ++seq) {
      const Tile t = decode(seq, n_iter, nd, nf);
      const int m_blk = cta + t.it * G;
      const int row0 = m_blk * BM, row = row0 + r;
      if (issuer) {
        // staging buffers: every earlier store has READ its source.  Publication to the producer warp: all stores of the
        // tiles before the previous one have COMPLETED -- only the previous tile's own bulk groups (2 for a first-half
        // tile, one per round for a residual tile) may still be in flight, so this wait is on stores that are at least a
        // whole tile old and never stalls the epilogue (waiting for everything here cost a store round trip per tile)
        tma_store_wait_read();
        if (seq >= 1) {
          if (decode(seq - 1, n_iter, nd, nf).type == T_WI) asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group 4;" ::: "memory");
          fence_proxy_async_all();
          if (seq - 1 > published) { published = seq - 1; *done_ctr = (uint32_t)published; }   // never moves backwards
        }
      }
      if (t.type == T_WI) {
        // ---- first half of the FF: relu(acc * rsqrt(mean x^2 + eps)) -> bf16 -> this CTA's ff scratch ----
        float rs = 0.f;
        if (row < M) {
          float s = 0.f;
          for (int b = 0; b < ssb; ++b) s += a.ss[(size_t)row * ssb + b];
          rs = 1.0f / sqrtf(s * inv_d + a.eps);
        }
        mbar_wait_wd(tfull_bar(acc), acc_phase, a.err);
        tcgen05_fence_after();
        const uint32_t stg = xs(xround + 1u);                 // the buffer that does NOT hold the prefetched x round
#pragma unroll 1
        for (int rd = 0; rd < 2; ++rd) {
          // TMEM reads software-pipelined as in gemm_tc.cu: chunk cc + 1 is in flight while chunk cc is converted and staged
          uint32_t vv[2][32];
          const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + rd * 128);
          tmem_ld32(taddr0, vv[0]);
          if (issuer && rd > 0) tma_store_wait_read();
          epi_bar();
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int c = rd * 4 + cc;
            tmem_ld_wait();
            if (cc + 1 < 4) tmem_ld32(taddr0 + (uint32_t)((cc + 1) * 32), vv[(cc + 1) & 1]);
            uint32_t (&v)[32] = vv[cc & 1];
            const uint32_t box = stg + (uint32_t)(cc >> 1) * 16384u + (uint32_t)r * 128u;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float p = fmaxf(__uint_as_float(v[g * 8 + 2 * e]) * rs, 0.f), q = fmaxf(__uint_as_float(v[g * 8 + 2 * e + 1]) * rs, 0.f);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(p, q);
                pk[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              const uint32_t piece = (uint32_t)(((c & 1) * 4 + g) ^ (r & 7));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + (piece << 4)), "r"(pk[0]), "r"(pk[1]),
                           "r"(pk[2]), "r"(pk[3]) : "memory");
            }
          }
          if (rd == 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          fence_async_smem();
          epi_bar();
          if (issuer) {
            const int col0 = t.sub * BN + rd * 128;
            tma_store_2d_hint(&map_ffs, stg, col0, cta * BM, pol_keep);
            tma_store_2d_hint(&map_ffs, stg + 16384u, col0 + 64, cta * BM, pol_keep);
            tma_store_commit();
          }
        }
      } else {
        // ---- residual tile: x += acc, bf16(x * w) and the row sums of squares of the next RMSNorm (as EPI_RESID_NORM) ----
        const float* lnw = t.type == T_O ? a.ln_mid : a.ln_next;
        const bool emit = lnw != nullptr;
        const bool live_row = row < M;
        float ssq = 0.f;
        mbar_wait_wd(tfull_bar(acc), acc_phase, a.err);
        tcgen05_fence_after();
#pragma unroll 1
        for (int rd = 0; rd < NORM_ROUNDS; ++rd, ++xround) {
          const uint32_t x_cur = xs(xround);
          if (issuer) {
            if (rd > 0) tma_store_wait_read();                // the other buffer's store has been read
            // prefetch the next residual round of the program (possibly several first-half tiles away)
            int ns = seq, nr = rd + 1;
            bool more = true;
            if (nr == NORM_ROUNDS) {
              nr = 0;
              do { ++ns; } while (ns < total && decode(ns, n_iter, nd, nf).type == T_WI);
              more = ns < total;
            }
            if (more) request_x(ns, nr, xround + 1u);
          }
          uint32_t vv[2][32];
          const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + rd * 64);
          tmem_ld32(taddr0, vv[0]);
          epi_bar();                                          // the bf16 box is free again
          mbar_wait_wd(xfull_bar(xround & 1u), (xround >> 1) & 1u, a.err);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = rd * 2 + cc;
            tmem_ld_wait();
            if (cc == 0) tmem_ld32(taddr0 + 32u, vv[1]);
            uint32_t (&v)[32] = vv[cc];
            const uint32_t xrow = x_cur + (uint32_t)cc * 16384u + (uint32_t)r * 128u;
            const uint32_t brow = xbox + (uint32_t)r * 128u;
            const float* gw = emit ? lnw + t.sub * BN + c * 32 : nullptr;
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
                if (emit) {
                  ssq = fmaf(xo.x, xo.x, ssq); ssq = fmaf(xo.y, xo.y, ssq); ssq = fmaf(xo.z, xo.z, ssq); ssq = fmaf(xo.w, xo.w, ssq);
                  const float4 w4 = __ldg(reinterpret_cast<const float4*>(gw + g * 4));
                  __nv_bfloat162 h0 = __floats2bfloat162_rn(xo.x * w4.x, xo.y * w4.y), h1 = __floats2bfloat162_rn(xo.z * w4.z, xo.w * w4.w);
                  pk[hh * 2] = *reinterpret_cast<uint32_t*>(&h0);
                  pk[hh * 2 + 1] = *reinterpret_cast<uint32_t*>(&h1);
                }
              }
              if (emit) {
                const uint32_t piece = (uint32_t)((cc * 4 + g2) ^ (r & 7));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(brow + (piece << 4)), "r"(pk[0]), "r"(pk[1]),
                             "r"(pk[2]), "r"(pk[3]) : "memory");
              }
            }
          }
          if ((rd & 1) == 1) {
            // one partial per (row, 128-column block), owned by this thread alone: same summation order as gemm_tc.cu
            if (emit && live_row) a.ss[(size_t)row * ssb + t.sub * (BN >> 7) + (rd >> 1)] = ssq;
            ssq = 0.f;
          }
          if (rd == NORM_ROUNDS - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          fence_async_smem();
          epi_bar();
          if (issuer) {
            const int col0 = t.sub * BN + rd * 64;
            tma_store_2d(&map_x, x_cur, col0, row0);
            tma_store_2d(&map_x, x_cur + 16384u, col0 + 32, row0);
            if (emit) tma_store_2d(&map_xn, xbox, col0, row0);
            tma_store_commit();
          }
        }
      }
      if (issuer) {
        // one of the next two tiles reads what this one wrote (first block of the CTA: o-projection -> first half; last
        // block: first half -> second half): the lazy publication above would come too late for its producer
        bool eager = false;
        for (int j = 1; j <= 2 && seq + j < total; ++j) {
          const Tile nx = decode(seq + j, n_iter, nd, nf);
          if (nx.type != T_O && need_of(nx, (nx.type == T_WI ? kb_wi : kb_wo) - 1, nd, nf) >= seq + 1) eager = true;
        }
        if (eager) {
          tma_store_wait_all();
          fence_proxy_async_all();
          published = seq + 1;
          *done_ctr = (uint32_t)published;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (issuer) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

SmemAttr g_attr;

}  // namespace chain

// D >= 512: with a single residual tile per row block the x prefetch of the second-half tile would read what the
// o-projection tile of the same block has only just committed (the program relies on a tile boundary in between)
bool enc_chain_supported(int D, int HD, int F) {
  return D >= 512 && (D % 256) == 0 && (F % 256) == 0 && F >= 256 && (HD % 64) == 0 && HD >= 64;
}

size_t enc_chain_scratch_bytes(int F, int num_sms) { return (size_t)num_sms * chain::BM * (size_t)F * 2; }

cudaError_t enc_chain(const void* ao, const void* w_o, float* x, void* xn, float* ss, const void* w_i, const void* w_o2,
                      void* scratch, const float* ln_mid, const float* ln_next, float eps, int M_max, const int* m_ptr,
                      int D, int HD, int F, int num_sms, int hints, int* err, cudaStream_t s) {
  if (M_max <= 0) return cudaSuccess;
  if (!enc_chain_supported(D, HD, F) || !err) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(tc::mutex());
  const int num_m = (M_max + chain::BM - 1) / chain::BM;
  const int grid = num_m < num_sms ? num_m : num_sms;
  CUtensorMap m_ao, m_wo_attn, m_x, m_xn, m_wi, m_ffs, m_wo;
  if (!tc::get_map(ao, M_max, HD, 0, chain::BM, &m_ao) || !tc::get_map(w_o, D, HD, 0, chain::BN, &m_wo_attn) ||
      !tc::get_map(x, M_max, D, 1, chain::BM, &m_x) || !tc::get_map(xn, M_max, D, 0, chain::BM, &m_xn) ||
      !tc::get_map(w_i, F, D, 0, chain::BN, &m_wi) || !tc::get_map(scratch, num_sms * chain::BM, F, 0, chain::BM, &m_ffs) ||
      !tc::get_map(w_o2, D, F, 0, chain::BN, &m_wo))
    return cudaErrorUnknown;
  cudaError_t e = chain::g_attr.ensure(chain::enc_chain_kernel, chain::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  chain::Args a;
  a.M_imm = M_max; a.m_ptr = m_ptr; a.D = D; a.HD = HD; a.F = F; a.ss = ss; a.ln_mid = ln_mid; a.ln_next = ln_next;
  a.eps = eps; a.err = err; a.hints = hints;
  chain::enc_chain_kernel<<<grid, chain::THREADS, chain::SMEM_BYTES, s>>>(m_ao, m_wo_attn, m_x, m_xn, m_wi, m_ffs, m_wo, a);
  return cudaGetLastError();
}

}  // namespace gram
