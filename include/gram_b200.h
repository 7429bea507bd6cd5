/*
 * gram_b200 -- C ABI of the B200-native GRAM inference/scoring hot path.
 *
 * This is the drop-in boundary.  The reference has no native layer: its
 * "FFI" for this path is the Python call `model_rec.generate(...)` made by the eval loop, which in
 * turn drives torch/ATen kernels op by op.  Each entry point below states the reference interface it
 * replaces (paths relative to the reference root).  All pointers are plain host or device pointers;
 * no torch types cross this boundary.  Integer return codes: 0 = ok, non-zero = error (message via
 * gram_last_error).  One handle per device; a handle is not re-entrant; all device work is issued on
 * the caller-provided stream and there is no host synchronisation inside the decode-step loop.
 */
#ifndef GRAM_B200_H_
#define GRAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gram_handle gram_handle;

enum { GRAM_DTYPE_F32 = 0, GRAM_DTYPE_BF16 = 1 };

enum {
  GRAM_OK = 0,
  GRAM_ERR_INVALID = 1,   /* bad argument / shape / capacity exceeded */
  GRAM_ERR_CUDA = 2,      /* a CUDA runtime call failed              */
  GRAM_ERR_STATE = 3,     /* call order violated (weights/trie/encode missing) */
  GRAM_ERR_UNSUPPORTED = 4
};

/* Model + capacity description.  Field meaning follows the reference's config object
 * (src/model/gram_t5_config.py:85-105) and the GRAM additions (src/main_generative_gram.py:67-70). */
typedef struct gram_config {
  int32_t vocab_size, d_model, d_kv, d_ff;
  int32_t num_layers, num_decoder_layers, num_heads;
  int32_t rel_buckets, rel_max_distance;
  float   ln_eps;
  int32_t pad_id, eos_id, start_id;
  int32_t tie_word_embeddings;     /* 1: scale decoder output by d_model**-0.5 (gram_t5.py:249-252) */
  int32_t n_positions;             /* rows of the passage-position table (max_item_num+1); 0 = none  */
  int32_t dtype;                   /* GRAM_DTYPE_F32 (parity mode) or GRAM_DTYPE_BF16                */
  int32_t device;                  /* CUDA device ordinal */
  /* capacities: workspaces are allocated once in gram_create, nothing is allocated per call */
  int32_t max_users;               /* B per call */
  int32_t max_passages;            /* N */
  int32_t max_seq_len;             /* L */
  int32_t max_beams;               /* K */
  int32_t max_length;              /* decoder max_length (start token included) */
  int64_t max_tokens;              /* workspace rows = cap on VALID encoder tokens per call; 0 = max_users*max_passages*
                                      max_seq_len (every passage full).  A smaller value lets a bigger user batch fit:
                                      the valid-token count is checked on the device, a batch above it is emptied and
                                      reported (gram_generate with host outputs / gram_check_errors)               */
  int32_t flags;                   /* GRAM_FLAG_* */
} gram_config;

enum {
  GRAM_FLAG_SIMT_GEMM = 1,         /* force the CUDA-core GEMM even for bf16 (debug / A-B timing) */
  GRAM_FLAG_KEEP_LOGITS = 2,       /* record per-step taps (lse, beam scores, prefixes) for parity tests; the decode path
                                      itself is unchanged (bf16: the fused log-softmax head stays on)               */
  GRAM_FLAG_SIMT_ATTN = 4,         /* force the CUDA-core attention kernels even for bf16 (A-B timing)   */
  GRAM_FLAG_MMA_ENC_ATTN = 8,      /* encoder attention through the mma.sync kernel (attention_mma.cu) instead of the
                                      persistent tcgen05/TMEM kernel (attention_tc.cu), which is the default for bf16,
                                      d_kv = 64 and passages of at most 128 tokens (A-B timing, cross-check)          */
  GRAM_FLAG_GEMM_1CTA = 16,        /* keep every tcgen05 GEMM on single-CTA tiles (no cta_group::2 pairs; A-B timing) */
  GRAM_FLAG_UNFUSED_NORM = 64,     /* bf16 encoder: run the per-layer RMSNorms as separate kernels (A-B timing).  Default: they
                                      are folded into the tcgen05 GEMMs around them (the residual GEMM emits
                                      bf16(x * ln_w) and the row's sums of squares, the consumer GEMM scales its output
                                      rows by rsqrt(mean x^2 + eps)): no normalisation pass over the residual stream.
                                      Same math, one rounding placed differently (x*w is rounded to bf16 before the
                                      row scale instead of after)                                                     */
  GRAM_FLAG_UNFUSED_HEAD = 128,    /* bf16: materialise the [rows, V] fp32 logits and reduce them with lse_rows instead of the
                                      fused log-softmax epilogue of the vocabulary GEMM (A-B timing / cross-check)     */
  GRAM_FLAG_ENC_CHAIN = 256,       /* bf16 encoder: ONE persistent launch per layer chains the o-projection, wi and wo GEMMs per
                                      128-token row block and hands ff / the normalised rows from GEMM to GEMM through a per-CTA
                                      scratch (gemm_chain.cu) instead of three launches.  Same arithmetic per element,
                                      bit-identical results.  Opt-in: measured slower than the three launches at the
                                      headline batch (the scratch does not stay L2-resident there)  */
  GRAM_FLAG_NO_L2_HINTS = 512,     /* chain kernel without L2 cache-policy hints on its TMA traffic (A-B timing)      */
  GRAM_FLAG_MMA_LONG_ATTN = 2048,  /* passages of 129-256 tokens through the mma.sync encoder attention instead of the two-key-block
                                      tcgen05 kernel (A-B timing; GRAM_FLAG_MMA_ENC_ATTN covers every length)       */
  GRAM_FLAG_CUDA_GRAPH = 4096,     /* gram_generate: capture everything between the input staging and the result copy-out (the
                                      encoder's ~70 launches and ~40 per decode step; nothing in it depends on a host value) as a
                                      CUDA graph per call shape and replay it (first call of a shape eager, second captures).
                                      Same kernels, same results; it removes the host's per-launch cost, which only shows at
                                      small batches                                             */
  GRAM_FLAG_XATTN_PER_ITEM = 8192, /* cross-attention with one CTA per (user, head group) instead of the persistent kernel that
                                      streams the K/V tiles of consecutive items without a gap (A-B timing)            */
  GRAM_FLAG_FP32_RESID = 16384,    /* bf16 encoder: keep the residual stream in fp32 (the round-1 arithmetic; A-B timing and a more
                                      conservative precision mode).  Default for dtype bf16: the encoder's residual stream is bf16,
                                      as in the reference run under model.bfloat16(): the residual GEMMs (o, wo) update it in
                                      place (x = bf16(x + A W^T), sums of squares of the rounded rows on the way) and the stream
                                      itself is the A operand of q|k|v and wi, whose weight columns carry the RMSNorm gain
                                      (folded from the fp32 weights when they are finalised): 4 B per stream element per
                                      residual GEMM instead of 10 (fp32 read + fp32 write + bf16 copy).  One more bf16
                                      rounding per sublayer; measured parity and speed in DESIGN.md section 5h.  The decoder's
                                      stream (a few thousand rows) stays fp32                                              */
  GRAM_FLAG_NO_DEC_CHAIN = 1024,   /* bf16 decoder: cross-attention output projection, wi and wo as three launches instead of
                                      one chain launch per layer (A-B timing)                                         */
  GRAM_FLAG_ALL_ROWS = 32          /* decode every beam row at every step, as the reference does (A-B timing).  Default:
                                      beams that are dead (-inf score: the user had fewer than K finite continuations,
                                      typically item ids that ended a token earlier) and the beams of users whose
                                      hypotheses are final are compacted away on the device before each step; they
                                      cannot reach any output, so results are bit-identical                          */
};

/* ---- lifetime -------------------------------------------------------------------------------- */
/* replaces: model.create_model("gram", config) + .to(device)   (src/model/__init__.py:9-24,
 * src/main_generative_gram.py:158-163) */
int gram_create(const gram_config* cfg, gram_handle** out);
void gram_destroy(gram_handle* h);
const char* gram_last_error(const gram_handle* h);   /* h may be NULL: last create error */
const char* gram_version(void);

/* ---- weights --------------------------------------------------------------------------------- */
/* replaces: GRAM.load_t5(state_dict) / load_state_dict  (src/model/gram.py:162-165).
 * `name` is a canonical tensor name (see gram_b200/weights.py for the mapping from the reference's
 * state-dict keys): "shared", "lm_head", "pos_emb", "enc.final_ln", "dec.final_ln",
 * "enc.<i>.{q,k,v,o,wi,wo,ln0,ln1,rel_bias}", "dec.<i>.{q,k,v,o,cq,ck,cv,co,wi,wo,ln0,ln1,ln2,rel_bias}".
 * `data` is a HOST pointer to contiguous fp32 in the reference's layout (nn.Linear: [out,in]). */
int gram_load_weight(gram_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim);
/* relative-position bucket tables, computed by the host with the reference's own torch fp32 ops
 * (src/model/gram_t5_modeling.py:397-450; rounding-sensitive):
 *   enc_buckets[r + (n_enc-1)/2] for r = mem - ctx in [-(L-1), L-1];  dec_buckets[dist], dist = ctx - mem >= 0 */
int gram_set_rel_buckets(gram_handle* h, const int32_t* enc_buckets, int32_t n_enc,
                         const int32_t* dec_buckets, int32_t n_dec);
/* checks that every tensor arrived, builds fused/packed device layouts.  Call once after loading. */
int gram_finalize_weights(gram_handle* h);

/* ---- item trie ------------------------------------------------------------------------------- */
/* replaces: Trie(sequences) + prefix_allowed_tokens_fn(trie) (src/utils/generation_trie.py:5-95) as
 * consumed by PrefixConstrainedLogitsProcessor.  CSR: children of node n are
 * child_tokens/child_nodes[child_offsets[n] .. child_offsets[n+1]).  `root_node` is the node reached
 * by the decoder start token (Trie.get([0]) lists its children); a trie without it (root_node < 0) is rejected
 * with GRAM_ERR_INVALID.  Calling it again replaces (and frees) the previous trie. */
int gram_set_trie(gram_handle* h, const int32_t* child_offsets, const int32_t* child_tokens,
                  const int32_t* child_nodes, int32_t n_nodes, int32_t n_edges, int32_t root_node);

/* ---- hot path -------------------------------------------------------------------------------- */
/* replaces: EncoderWrapper.forward (src/model/gram.py:200-256) + the per-layer cross-attention K/V
 * projection (src/model/gram_t5_modeling.py:531-534).  ids int64 [B,N,L], mask uint8/bool [B,N,L];
 * host or device pointers.  Leaves the fused memory and K/V resident in the handle. */
int gram_encode(gram_handle* h, const int64_t* ids, const uint8_t* mask, int32_t B, int32_t N, int32_t L,
                void* stream);

/* replaces: GRAM.generate(input_ids, attention_mask, max_length, prefix_allowed_tokens_fn=...,
 * num_beams=K, num_return_sequences=R, length_penalty=...) (src/model/gram.py:74-107; call site
 * src/runner/single_runner_gram.py:641-651) including transformers-4.26 beam_search /
 * BeamSearchScorer / PrefixConstrainedLogitsProcessor.
 *   len_pow[c] = float(c) ** length_penalty for c in [0, max_length]   (host-computed doubles)
 *   out_seq    int64 [B*R, max_length], rows best-first per user, 0-padded, EOS after the id
 *   out_scores fp32  [B*R]  (sum_logprobs / len**length_penalty)
 *   out_width  int32 [1]    min(max hypothesis length + 1, max_length) as HF pads it
 * Output pointers may be host or device.  If `ids` is NULL the batch encoded by the last
 * gram_encode call is decoded. */
int gram_generate(gram_handle* h, const int64_t* ids, const uint8_t* mask, int32_t B, int32_t N, int32_t L,
                  int32_t K, int32_t R, int32_t max_length, const double* len_pow,
                  int64_t* out_seq, int32_t* out_width, float* out_scores, void* stream);

/* ---- parity / debug taps (same kernels, results copied out) ------------------------------------ */
/* fused memory of the last gram_encode as fp32 [B, N*L, d_model]; positions that were skipped
 * (masked tail of a passage, all-masked passages) are written as 0. Host or device pointer. */
int gram_get_memory(gram_handle* h, float* out, void* stream);
/* replaces: GRAM.forward(..., decoder_input_ids=dec_ids, encoder_outputs=...) teacher-forced logits
 * (src/model/gram.py:51-69): dec_ids int64 [B,q] -> logits fp32 [B,q,V] through the cached
 * single-token decode step. */
int gram_decoder_logits(gram_handle* h, const int64_t* dec_ids, int32_t q, float* out_logits, void* stream);
/* per-step taps of the last gram_generate (requires GRAM_FLAG_KEEP_LOGITS):
 * lse fp32 [steps, B*K]; beam_scores fp32 [steps, B*K] (scores entering the step);
 * beam tokens int32 [steps, B*K, max_length] (prefix entering the step). Any pointer may be NULL. */
int gram_get_step_taps(gram_handle* h, float* lse, float* beam_scores, int32_t* beam_tokens, int32_t* n_steps);

/* ---- measurement ------------------------------------------------------------------------------- */
/* Calls whose outputs are all device pointers never synchronise, so input errors the kernels detect (token id
 * outside the vocabulary -- the reference's nn.Embedding raises IndexError, src/model/gram_t5_modeling.py:1091 --,
 * item index outside the cached table, candidate overflow) stay in a sticky device flag.  This synchronises the
 * stream, returns GRAM_ERR_INVALID with the message if the flag is set, and clears it. */
int gram_check_errors(gram_handle* h, void* stream);

typedef struct gram_stats {
  int64_t launches;          /* kernels launched by the last encode/generate call            */
  int64_t packed_tokens;     /* valid encoder tokens of the last encode (needs a sync to read) */
  int64_t kv_bytes;          /* bytes of cross-attention K/V resident for the last batch      */
  int64_t workspace_bytes;   /* device memory owned by the handle                             */
  int64_t decoded_rows;      /* decoder rows actually run by the last generate, summed over its steps (the reference
                                runs B*K rows at every step; step 0 runs one row per user here, later steps only the
                                live beams)                                                                          */
  int64_t kv_tokens_read;    /* memory tokens whose K/V one decoder layer streamed, summed over the steps (a user
                                without live beams is not read)                                                      */
} gram_stats;
int gram_get_stats(gram_handle* h, gram_stats* out);

/* kernel classes for event timing */
enum {
  GRAM_K_GEMM_ENC = 0, GRAM_K_ENC_ATTN = 1, GRAM_K_GEMM_KV = 2, GRAM_K_GEMM_DEC = 3,
  GRAM_K_CROSS_ATTN = 4, GRAM_K_LM_HEAD = 5, GRAM_K_BEAM = 6, GRAM_K_OTHER = 7, GRAM_K_SELF_ATTN = 8,
  GRAM_K_NORM_ENC = 9, GRAM_K_NORM_DEC = 10, GRAM_K_COUNT = 11
};
/* mask: bit c set = bracket every launch of class c with CUDA events on the launching stream */
int gram_profile_begin(gram_handle* h, uint32_t class_mask);
/* synchronises the stream, returns per-class total milliseconds and launch counts, clears the log */
int gram_profile_end(gram_handle* h, float* ms_per_class /*[GRAM_K_COUNT]*/, int64_t* launches_per_class);

/* ---- per-item encoder-state cache (no counterpart in the reference) ------------
 * Every passage except the user prompt depends on the ITEM only (reference src/utils/indexing.py:209-211,315-320),
 * passages are encoded independently (src/model/gram.py:206-216) and the passage-position row is added after the
 * encoder (src/model/gram.py:238-249).  gram_cache_items encodes each item passage once -- ids int64 / mask uint8
 * [n_items, L], host or device -- and keeps its final-normed encoder rows (fp32, before the position add) in the
 * handle.  gram_encode_cached then encodes only the user prompts [B, L] and assembles every user's memory from the
 * prompt rows and the cached rows of items[b][0..NI) (int32, -1 = no passage), passage index 1 + j; L must equal
 * the L of the table.  The result is bit-identical to gram_encode on the equivalent [B, 1+NI, L] input; follow it
 * with gram_generate(h, NULL, NULL, ...) to decode.  The table is dropped by gram_destroy / a new gram_cache_items. */
int gram_cache_items(gram_handle* h, const int64_t* ids, const uint8_t* mask, int32_t n_items, int32_t L, void* stream);
int gram_encode_cached(gram_handle* h, const int64_t* prompt_ids, const uint8_t* prompt_mask, const int32_t* items,
                       int32_t B, int32_t NI, int32_t L, void* stream);

/* ---- single-operator entry points (unit parity tests and roofline measurement) ------------------ */
/* C[M,N] = A[M,K] * W[N,K]^T on device pointers.  dtype as in gram_config; impl 0 = SIMT fp32-accumulate,
 * 1 = tcgen05 (bf16 only; CTA pairs on large problems), 2 = tcgen05 single-CTA tiles only.  epilogue: 0 store (dtype), 1 relu+store (dtype), 2 C_f32 += acc, 3 store fp32. */
int gram_op_gemm(int32_t device, int32_t dtype, int32_t impl, int32_t epilogue, const void* A, const void* W,
                 void* C, int32_t M, int32_t N, int32_t K, void* stream);
/* The tcgen05 GEMM with an RMSNorm (reference T5LayerNorm, src/model/gram_t5_modeling.py:253-276) folded into it, bf16.
 * epilogue 5 (producer): C fp32 [M,N] += A W^T; xb bf16 [M,N] = (C) * ln_w; ss fp32 [M, N/128] = sums of squares of the new C
 * rows per 128-column block.  epilogue 0 / 1 (consumer, row_ss != NULL): C = A W^T with output row i scaled by
 * rsqrt(sum_b row_ss[i][b] / K + eps) (then ReLU for 1).  epilogue 6 (the bf16 encoder's residual GEMMs): C is a bf16 [M,N] residual
 * stream updated in place, C = bf16(C + A W^T), ss = sums of squares of the rounded rows per 128-column block; xb / ln_w unused.
 * impl 1 = CTA pairs allowed, 2 = single-CTA tiles. */
int gram_op_gemm_norm(int32_t device, int32_t impl, int32_t epilogue, const void* A, const void* W, void* C, void* xb,
                      float* ss, const float* ln_w, const float* row_ss, float eps, int32_t M, int32_t N, int32_t K,
                      void* stream);
/* One encoder layer's residual sublayers chained per 128-row block in ONE launch (gemm_chain.cu; reference
 * T5LayerSelfAttention output projection + T5LayerFF, src/model/gram_t5_modeling.py:297-310,337-352,622,634-667), bf16:
 *   x fp32 [M,D] += ao [M,HD] w_o[D,HD]^T;  xn bf16 [M,D] = x * ln_mid;  ss fp32 [M, D/128] = row sums of squares per 128 columns;
 *   ff = relu((xn w_i[F,D]^T) * rsqrt(sum ss / D + eps))  (never leaves the L2-resident scratch);
 *   x += ff w_o2[D,F]^T;  xn = x * ln_next, ss   (ln_next NULL: x only).
 * scratch: num_SMs * 128 * F * 2 bytes; err: device int32 (7 = internal watchdog).  D % 256 == F % 256 == HD % 64 == 0. */
int gram_op_enc_chain(int32_t device, const void* ao, const void* w_o, float* x, void* xn, float* ss, const void* w_i,
                      const void* w_o2, void* scratch, int64_t scratch_bytes, const float* ln_mid, const float* ln_next, float eps,
                      int32_t M, int32_t D, int32_t HD, int32_t F, int32_t hints, int32_t* err, void* stream);
/* Kernel (c) head: lse[i] = log sum_v exp(hidden[i] . head[v]) with the log-softmax statistics fused into the epilogue of the
 * tcgen05 vocabulary GEMM (the [M, V] logits are never written) + the per-row combine -- the path gram_generate runs in
 * bf16 (reference: lm_head then log_softmax, src/model/gram_t5.py:249-254 and HF beam_search).  hidden bf16 [M, D],
 * head bf16 [V, D], lse fp32 [M], partial = scratch of M * ceil(V / 128) * 8 bytes; device pointers. */
int gram_op_lse_head(int32_t device, const void* hidden, const void* head, float* lse, void* partial, int32_t M, int32_t V,
                     int32_t D, void* stream);
/* decoder cross-attention over an in-place K/V memory: q [users*K, H*dk], kv [kv_rows, 2*H*dk] (K|V),
 * user_start int32 [users+1], tok_valid uint8 [kv_rows] or NULL, out [users*K, H*dk] (all dtype).
 * impl 0 = CUDA-core kernel (fp32 or bf16), 1 = TMA + tensor-core kernel, persistent (bf16), 2 = the same, one CTA per
 * (user, head group). */
int gram_op_cross_attention(int32_t device, int32_t dtype, int32_t impl, const void* q, const void* kv, int32_t kv_rows,
                            const int32_t* user_start, const uint8_t* tok_valid, void* out,
                            int32_t users, int32_t K, int32_t H, int32_t dk, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAM_B200_H_ */
