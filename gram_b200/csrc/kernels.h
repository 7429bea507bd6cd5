// Host-callable launchers of the gram_b200 kernels.  dtype: 0 = fp32 storage, 1 = bf16 storage
// (activations/weights); accumulation and the residual stream are always fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gram {

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remember, per device, the largest size configured for a
// kernel and raise it on demand (a process may drive several GPUs).
struct SmemAttr {
  size_t configured[64] = {0};
  template <typename Kern>
  cudaError_t ensure(Kern kern, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev &= 63;
    if (bytes > configured[dev]) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e != cudaSuccess) return e;
      configured[dev] = bytes;
    }
    return cudaSuccess;
  }
};

enum { EPI_STORE = 0, EPI_RELU = 1, EPI_RESID = 2, EPI_F32 = 3, EPI_LSE = 4, EPI_RESID_NORM = 5, EPI_RESID_BF16 = 6 };
// EPI_LSE (tcgen05 GEMM only): C is float2 [M, ceil(N/128)] of per-tile (max, sum exp(x - max)); no logits are stored
// EPI_RESID_NORM (tcgen05 GEMM only): C (fp32 residual stream) += A W^T like EPI_RESID, and the input of the RMSNorm that
// follows is produced on the way (GemmNormAux): no separate normalisation pass over the residual stream

// EPI_RESID_BF16 (tcgen05 GEMM only; bf16 mode unless GRAM_FLAG_FP32_RESID): C is a bf16 residual stream updated in place,
// C = bf16(C + A W^T), with the row sums of squares of the ROUNDED rows per 128-column block (aux->ss_out).  The stream
// itself is the next GEMM's A operand (the RMSNorm gain is folded into that GEMM's weight columns at load time), so a
// residual GEMM moves 2 + 2 B per element of the stream instead of 4 + 4 + 2

// RMSNorm folded into the tcgen05 GEMMs on either side of it (reference T5LayerNorm, src/model/gram_t5_modeling.py:
// 253-276: y = w * x * rsqrt(mean(x^2) + eps)).  The producer (EPI_RESID_NORM) writes xb = bf16(x * w) and the row's sums
// of squares per 128-column block; the consumer (EPI_STORE / EPI_RELU over A = xb) scales output row i by
// rsqrt(sum_b ss[i][b] / K + eps): (x w r) W^T = ((x w) W^T) r.
struct GemmNormAux {
  const float* row_ss = nullptr;   // consumer: [M][K/128]; nullptr = plain GEMM
  void* xb = nullptr;              // producer: [M][N] bf16
  float* ss_out = nullptr;         // producer: [M][N/128]
  const float* ln_w = nullptr;     // producer: [N] weight of the next RMSNorm
  float eps = 0.f;                 // consumer
};

// ---- gemm_simt.cu ------------------------------------------------------------------------------
// C[M,N] = A[M,K] * W[N,K]^T.  M = *m_ptr when m_ptr != nullptr (grid sized for M_max), else M_max.
cudaError_t gemm_simt(int dtype, int epi, const void* A, const void* W, void* C, int M_max, const int* m_ptr,
                      int N, int K, cudaStream_t s);

// ---- gemm_tc.cu (tcgen05 / TMEM / TMA, bf16) -----------------------------------------------------
bool gemm_tc_supported(int N, int K);
// max_ctas: 2 = large problems run on CTA pairs (tcgen05 cta_group::2, 256x256 tiles), 1 = single-CTA tiles only
cudaError_t gemm_tc(int epi, const void* A, const void* W, void* C, int M_max, const int* m_ptr, int N, int K,
                    int num_sms, int max_ctas, cudaStream_t s, const GemmNormAux* aux = nullptr);
const char* gemm_tc_last_error();
// number of column tiles (= per-row partials) the EPI_LSE epilogue writes for this problem
int gemm_tc_lse_ntiles(int M_max, int N, int num_sms);

// ---- gemm_chain.cu (tcgen05 / TMEM / TMA, bf16): one encoder layer's residual sublayers chained per 128-row block --------
// x += ao w_o^T; xn = bf16(x * ln_mid), ss;   ff = relu(RMSNorm-scaled xn w_i^T) (kept in an L2-resident per-CTA scratch);
// x += ff w_o2^T; xn = bf16(x * ln_next), ss (ln_next == nullptr: last layer, x only).  D % 256 == F % 256 == HD % 64 == 0.
bool enc_chain_supported(int D, int HD, int F);
size_t enc_chain_scratch_bytes(int F, int num_sms);
// hints: 1 = L2 cache-policy hints on the TMA traffic; err: sticky device flag (7 = watchdog)
cudaError_t enc_chain(const void* ao, const void* w_o, float* x, void* xn, float* ss, const void* w_i, const void* w_o2,
                      void* scratch, const float* ln_mid, const float* ln_next, float eps, int M_max, const int* m_ptr,
                      int D, int HD, int F, int num_sms, int hints, int* err, cudaStream_t s);

// ---- encoder_kernels.cu --------------------------------------------------------------------------
struct PackMeta {            // device-resident description of the packed (valid-token) layout
  int* plen;                 // [P]    valid length of passage p (last valid index + 1)
  int* poff;                 // [P+1]  first packed row of passage p
  int* ustart;               // [B+1]  first packed row of user b
  int* uorder;               // [B]    users ordered longest-first (launch order of the cross-attention CTAs)
  int* total;                // [1]    number of packed rows (M of every encoder GEMM)
  int* tok_id;               // [Mcap] token id per packed row
  int* tok_pos;              // [Mcap] passage index within the user (row of the position table)
  uint8_t* tok_valid;        // [Mcap] attention-mask bit per packed row
  int* row_src;              // [Mcap] flat index b*N*L + n*L + l of the packed row (for unpacking)
  int* err;                  // [1]    sticky device error flag (2 = token id outside the vocabulary)
  int vocab;                 // vocabulary size, for the id range check
  long long cap;             // rows the packed buffers hold; a batch with more valid tokens is emptied and err = 4
};
cudaError_t enc_pack(const int64_t* ids, const uint8_t* mask, int B, int N, int L, PackMeta pm, cudaStream_t s);
// per-item encoder-state cache (see encoder_kernels.cu): scatter a chunk's packed fp32 rows into the item table, and
// build a user layout + memory from prompt rows and cached item rows (4 launches)
cudaError_t cache_scatter(const float* src, const PackMeta& pm, int M_max, int D, long long row0, float* item_mem,
                          uint8_t* item_valid, cudaStream_t s);
cudaError_t cached_pack_assemble(int dtype, const PackMeta& pm, const PackMeta& prompt_pm, const float* prompt_rows,
                                 const int* items /*[B,NI], -1 = none*/, const float* item_mem, const uint8_t* item_valid,
                                 const int* item_len, int n_items, const float* pos_table, int B, int NI, int L, int D,
                                 void* mem, cudaStream_t s);
cudaError_t embed_rows(int dtype, const void* table, const int* tok_id, float* x, int M_max, const int* m_ptr,
                       int D, cudaStream_t s);
// bf16 only: embedding with the first RMSNorm folded (x fp32, xb = bf16(x * w), ss [M][D/128]); see GemmNormAux
cudaError_t embed_rows_norm(const void* table, const int* tok_id, float* x, void* xb, float* ss, const float* w, int M_max,
                            const int* m_ptr, int D, cudaStream_t s);
// y = w * (x * rsqrt(mean(x^2)+eps)) * scale  [+ pos_table[tok_pos[row]]]   (x fp32 -> y dtype)
cudaError_t rmsnorm_rows(int dtype, const float* x, const float* w, void* y, int M_max, const int* m_ptr, int D,
                         float eps, float scale, const float* pos_table, const int* tok_pos, cudaStream_t s);
// bf16 residual stream (bf16 mode unless GRAM_FLAG_FP32_RESID): the stream's first value (= the embedding row) + ss [M][D/128]; the norm over it
cudaError_t embed_rows_stream(const void* table, const int* tok_id, void* xr, float* ss, int M_max, const int* m_ptr, int D,
                              cudaStream_t s);
cudaError_t rmsnorm_rows_stream(int dtype, const void* x, const float* w, void* y, int M_max, const int* m_ptr, int D,
                                float eps, float scale, const float* pos_table, const int* tok_pos, cudaStream_t s);
// bidirectional self-attention of every (passage, head): qkv [M, 3*H*dk] -> out [M, H*dk]
cudaError_t enc_attention(int dtype, const void* qkv, void* out, const int* plen, const int* poff,
                          const uint8_t* tok_valid, const float* bias_lut /*[H, 2*Lb-1]*/, int Lb, int P, int H,
                          int dk, int Lmax, cudaStream_t s);
cudaError_t unpack_memory(int dtype, const void* mem, const int* row_src, float* out, int M_max, const int* m_ptr,
                          int D, cudaStream_t s);

// ---- decoder_kernels.cu --------------------------------------------------------------------------
// single-token causal self-attention with an ancestry-indirected KV cache.
//   qkv [R, 3*HD]; cache_k/v [Tmax][R][HD]; anc [R][Tmax] (row *within the user* holding position j);
//   dec_bias [H][n_dec]; t = current position.  Writes this step's k/v into slot t.
// Live-row compaction (slot_row != nullptr): the step decodes only *n_rows compact slots; slot s stands for beam row
// slot_row[s] (cache rows, ancestry and user membership follow the beam row, qkv/out rows follow the slot).
cudaError_t dec_self_attention(int dtype, const void* qkv, void* cache_k, void* cache_v, const int* anc, int Tmax,
                               const float* dec_bias, int n_dec, void* out, int R, int K, int H, int dk, int t,
                               const int* slot_row, const int* n_rows, cudaStream_t s);
// cross-attention over the in-place packed K/V memory (kernel (b)).
//   q [R, HD]; kv rows of `kv_stride` elements with K at column k_off and V at column v_off (+ h*dk);
//   user u owns packed rows [ustart[u], ustart[u+1]); beams of user u are rows u*K .. u*K+K-1, or -- with live-row
//   compaction (live_start != nullptr) -- the live_count[u] rows from live_start[u] (a user without live beams is skipped).
cudaError_t cross_attention(int dtype, const void* q, const void* kv, size_t kv_stride, int k_off, int v_off,
                            const int* ustart, const uint8_t* tok_valid, void* out, int users, int K, int H, int dk,
                            const int* live_start, const int* live_count, cudaStream_t s);

// ---- attention_mma.cu (bf16, tensor cores) --------------------------------------------------------------
bool cross_attention_mma_supported(int K, int H, int dk);
// kv_rows = rows of the allocation behind `kv` (TMA bound); rows past a user's range are masked, so the buffer
// must only hold finite values.
cudaError_t cross_attention_mma(const void* q, const void* kv, size_t kv_rows, size_t kv_stride, int k_off, int v_off,
                                const int* ustart, const int* uorder /* nullable: users longest-first */,
                                const uint8_t* tok_valid, void* out, int users, int K, int H,
                                const int* live_start, const int* live_count /* nullable, see cross_attention */,
                                int num_sms /* > 0: persistent kernel, one CTA per SM; <= 0: one CTA per (user, head group) */,
                                cudaStream_t s);
bool enc_attention_mma_supported(int dk, int Lmax);
cudaError_t enc_attention_mma(const void* qkv, void* out, const int* plen, const int* poff, const uint8_t* tok_valid,
                              const float* bias_lut, int Lb, int P, int H, int Lmax, cudaStream_t s);

// ---- attention_tc.cu (bf16, tcgen05/TMEM, passages of <= 256 tokens) ---------------------------------------
bool enc_attention_tc_supported(int dk, int Lmax, int Lb, int H);
// qkv_rows = rows of the allocation behind `qkv` (TMA bound); rows past a passage are masked / not stored.
// Lmax <= 128 selects the one-key-block kernel (two CTAs per SM), Lmax <= 256 the two-key-block kernel
cudaError_t enc_attention_tc(const void* qkv, size_t qkv_rows, void* out, const int* plen, const int* poff,
                             const uint8_t* tok_valid, const float* bias_lut, int Lb, int P, int H, int Lmax, cudaStream_t s);

// ---- beam_kernels.cu -------------------------------------------------------------------------------
struct TrieCSR {
  const int* child_offsets; const int* child_tokens; const int* child_nodes;
  int n_nodes; int n_edges; int root; int max_fanout;
};
struct BeamState {           // all device pointers; rows R = users*K
  int K, max_length, V, eos, pad;   // max_length = row pitch of seq/anc/hyp_tok (capacity)
  int gen_len;                       // max_length of the current generate call (<= max_length)
  float* beam_score[2];      // [R]
  int* node[2];              // [R]   trie node of the beam prefix, -1 = dead
  int* seq[2];               // [R][max_length] token prefix
  int* anc[2];               // [R][max_length] self-attention cache ancestry
  int* tok;                  // [R]   input token of the next decoder step
  // hypotheses (per user, K+1 slots)
  double* hyp_score; int* hyp_len; int* hyp_seqno; int* hyp_tok;   // [U][K+1], ..., [U][K+1][max_length]
  int* n_hyp; double* worst; int* next_seqno; int* done;            // [U]
  int* live_cnt;             // [U]   beams of the user that are still alive after the step (0 once the user is done);
                             //       they are the user's FIRST live_cnt beams (candidates are ranked best-first)
  int* err;                  // [1] sticky device-side error flag
  unsigned long long* work;  // [2] executed work of the current generate: decoder rows, K/V tokens read per layer
  const double* len_pow;     // [max_length+1]
  // optional taps
  float* tap_lse; float* tap_score; int* tap_seq;                   // [steps][R](...)
};
// n_rows (nullable, device): only the first *n_rows rows are computed
cudaError_t lse_rows(const float* logits, float* lse, int R, int V, const int* n_rows, cudaStream_t s);
// lse[row] from the per-tile (max, sumexp) partials of the EPI_LSE GEMM epilogue
cudaError_t lse_combine(const void* partial, float* lse, int R, int n_tiles, const int* n_rows, cudaStream_t s);
cudaError_t beam_init(BeamState bs, TrieCSR trie, int users, int start_tok, cudaStream_t s);
// one beam-search step: PrefixConstrainedLogitsProcessor + topk(2K) + BeamSearchScorer.process
// logits != nullptr: gather candidate logits from the materialised [R,V] matrix (fp32 parity mode);
// logits == nullptr: recompute them as dot(hidden[row], head[token]) from the bf16 decoder output `hidden` [R,D]
// and the bf16 vocabulary head `head` [V,D] (fused mode: full-vocab logits are never written)
// row_slot (nullable): live-row compaction of this step -- beam row r was decoded as row row_slot[r] of
// logits/hidden/lse (-1 = dead beam, not decoded)
cudaError_t beam_step(BeamState bs, TrieCSR trie, const float* logits, const void* hidden, const void* head, int D,
                      const float* lse, int users, int t, int cand_cap, int compact, const int* row_slot, cudaStream_t s);
// Live-row compaction for the step that follows a beam_step (two launches): an exclusive scan of bs.live_cnt gives
// every user its slot range, then slot_row / row_slot / the slots' input tokens are filled.  `cur` = index of the
// beam buffers the coming step reads.  Dead beams (-inf score: fewer finite candidates than K, typically item ids
// that ended one token earlier) and finished users cannot influence any output, so they are not decoded.
struct LiveMap {
  int* start;                // [U+1] first compact slot of user u
  int* slot_row;             // [R]   beam row of slot s
  int* row_slot;             // [R]   slot of beam row r, -1 = dead
  int* tok;                  // [R]   input token of slot s
  int* n_live;               // [1]   number of slots = M of the step's GEMMs
};
// ustart [U+1] = packed memory rows per user (for the work counters in bs.work)
cudaError_t live_compact(BeamState bs, int users, int cur, LiveMap lm, const int* ustart, cudaStream_t s);
// bs.work += (rows, tokens): accounting of a step that runs without live_compact
cudaError_t work_add(BeamState bs, long long rows, const int* tokens_ptr, cudaStream_t s);
// compact != 0 (only at t == 0): all K beams of a user are identical, so the decoder ran ONE row per user; row u of
// logits/hidden/lse serves every beam of user u and the step-0 self-attention cache row is recorded in the ancestry
cudaError_t beam_finalize(BeamState bs, int users, int t_final, int R_ret, int64_t* out_seq, float* out_scores,
                          int* out_width, cudaStream_t s);
// teacher forcing: tok[r] = ids[r*q + t], anc[r][t'] = 0
cudaError_t forced_step(BeamState bs, const int64_t* dec_ids, int q, int t, int R, cudaStream_t s);
size_t beam_step_smem(int cand_cap);

}  // namespace gram
