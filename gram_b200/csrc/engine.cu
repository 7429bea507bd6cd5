// gram_b200 engine: handle, weight/trie residency, the encode -> fused memory -> decode-step loop,
// and the C ABI declared in include/gram_b200.h.
//
// Orchestration mirrors the reference call stack:
//   GRAM.generate                      src/model/gram.py:74-107
//     EncoderWrapper.forward           src/model/gram.py:200-256          -> encode()
//     HF generate / beam_search loop   transformers 4.26 (third party)    -> generate() step loop
//       GRAM.forward one token         src/model/gram_t5.py:118-287       -> decoder_step()
//       log_softmax + trie + topk + scorer                                 -> lse_rows + beam_step
//       _reorder_cache                 src/model/gram_t5.py:320-348       -> eliminated (ancestry table; cross
//                                                                            K/V is per user, not per beam)
// Everything is enqueued on the caller's stream; the only host synchronisation is the final copy-out
// to host buffers.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <map>
#include <set>

#include "../../include/gram_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace gram;

namespace {

thread_local std::string g_create_error;

struct LayerW {
  void *qkv = nullptr, *o = nullptr, *wi = nullptr, *wo = nullptr;      // self-attn + FF
  void *cq = nullptr, *co = nullptr;                                     // decoder cross-attn q / o
  float *ln0 = nullptr, *ln1 = nullptr, *ln2 = nullptr;
  // the bf16 residual stream (encoder): fp32 copies of q|k|v and wi kept until finalize, and their bf16 versions with the
  // RMSNorm gain folded into the columns, W'[n][k] = bf16(W[n][k] * ln[k])
  float *qkv_f = nullptr, *wi_f = nullptr;
  void *qkv_g = nullptr, *wi_g = nullptr;
};

}  // namespace

struct gram_handle {
  gram_config cfg;
  int D, HD, F, V, H, dk, Le, Ld;
  size_t esz;                 // bytes per stored element
  int num_sms = 148;
  std::string err;
  std::vector<void*> allocs;
  size_t alloc_bytes = 0;

  // weights
  void* shared = nullptr;     // [V, D]
  void* lm_head = nullptr;    // [V, D] (== shared when tied weights were loaded once)
  float* pos_emb = nullptr;   // [n_positions, D] fp32
  float *enc_final_ln = nullptr, *dec_final_ln = nullptr;
  std::vector<LayerW> enc, dec;
  void* ckv_w = nullptr;      // [Ld*2*HD, D]  (ck_0 | cv_0 | ck_1 | ...)
  float *enc_rel = nullptr, *dec_rel = nullptr;     // [buckets, H] fp32
  float *enc_bias_lut = nullptr, *dec_bias_lut = nullptr;
  int n_enc_lut = 0, n_dec_lut = 0, Lb = 0;
  std::set<std::string> loaded;
  bool buckets_set = false, weights_ready = false;
  float* stage = nullptr;
  size_t stage_elems = 0;

  // trie
  TrieCSR trie{};
  int* trie_bufs[3] = {nullptr, nullptr, nullptr};   // child_offsets / child_tokens / child_nodes (freed on re-upload)
  bool trie_set = false;
  int cand_cap = 0;

  // encoder workspace
  int64_t Mcap = 0;
  PackMeta pm{};
  int64_t* d_ids = nullptr;
  uint8_t* d_mask = nullptr;
  float* x = nullptr;
  void *xn = nullptr, *qkv = nullptr, *ao = nullptr, *ff = nullptr, *mem = nullptr, *ckv = nullptr;
  void* ffs = nullptr;        // per-CTA ff scratch of the row-block chain kernel (gemm_chain.cu), L2 resident
  float* ss = nullptr;        // [Mcap][D/128] row sums of squares of the residual stream (RMSNorm folded into the GEMMs)
  int enc_B = 0, enc_N = 0, enc_L = 0;
  bool encoded = false;

  // per-item encoder-state cache: fp32 final-normed rows before the position add
  float* item_mem = nullptr;             // [n_items * item_L, D]
  uint8_t* item_valid = nullptr;         // [n_items * item_L]
  int* item_len = nullptr;               // [n_items]
  int n_items = 0, item_L = 0;
  PackMeta pm_prompt{};                  // packed layout of the user prompts alone (P = B passages)
  int* d_items = nullptr;                // staging for the item-index matrix [max_users, max_passages]

  // decoder workspace
  int Rcap = 0;
  float* dx = nullptr;
  void *dxn = nullptr, *dqkv = nullptr, *dao = nullptr, *dq = nullptr, *dff = nullptr;
  float* dss = nullptr;                  // [Rcap][D/128] row sums of squares of dx (RMSNorm folded into the decoder GEMMs)
  void *sk = nullptr, *sv = nullptr;     // [Ld][Tmax][Rcap][HD]
  float *logits = nullptr, *lse = nullptr;
  void* lse_partial = nullptr;           // float2 [Rcap][ceil(V/128)]
  BeamState bs{};
  LiveMap live{};                        // live-row compaction of the decode steps (beam_kernels.cu:live_compact)
  double* d_len_pow = nullptr;
  double* h_len_pow = nullptr;           // pinned
  cudaEvent_t len_pow_ev = nullptr;       // recorded after the H2D copy of h_len_pow; waited on before it is rewritten
  int64_t* d_out_seq = nullptr;
  float* d_out_scores = nullptr;
  int* d_out_width = nullptr;
  int* h_flags = nullptr;                // pinned: [0]=err, [1]=width, [2]=total tokens
  int* d_zero_anc = nullptr;
  int64_t* d_dec_ids = nullptr;
  int last_steps = 0, last_R = 0;

  // CUDA-graph replay of a whole generate call (GRAM_FLAG_CUDA_GRAPH): one instantiated graph per call shape
  struct GraphRec { int calls = 0; cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
  std::map<std::vector<int>, GraphRec> graphs;
  cudaStream_t gstream = nullptr;         // capture / replay stream (the caller's may be the legacy default stream)
  cudaEvent_t g_in = nullptr, g_out = nullptr;

  // measurement
  int64_t launches = 0;
  uint32_t prof_mask = 0;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct EvRec { int cls; cudaEvent_t a, b; };
  std::vector<EvRec> ev_log;
  int64_t cls_launches[GRAM_K_COUNT] = {0};
};

namespace {

// bf16 residual stream in the encoder (default for bf16; GRAM_FLAG_FP32_RESID turns it off): needs the folded-norm tcgen05 path
bool stream_mode(const gram_handle* h) {
  const gram_config& c = h->cfg;
  return c.dtype == GRAM_DTYPE_BF16 && !(c.flags & GRAM_FLAG_FP32_RESID) &&
         !(c.flags & (GRAM_FLAG_UNFUSED_NORM | GRAM_FLAG_SIMT_GEMM | GRAM_FLAG_ENC_CHAIN)) && (h->D % 128) == 0 &&
         gemm_tc_supported(3 * h->HD, h->D) && gemm_tc_supported(h->D, h->HD) && gemm_tc_supported(h->F, h->D) &&
         gemm_tc_supported(h->D, h->F);
}

#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      char buf_[512];                                                                                  \
      snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      h->err = buf_;                                                                                   \
      return GRAM_ERR_CUDA;                                                                            \
    }                                                                                                  \
  } while (0)

int fail(gram_handle* h, int code, const std::string& msg) {
  h->err = msg;
  return code;
}

void free_graphs(gram_handle* h) {
  for (auto& kv : h->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->graphs.clear();
}

void free_trie(gram_handle* h) {
  free_graphs(h);                          // captured launches hold the CSR pointers
  for (int i = 0; i < 3; ++i) {
    if (h->trie_bufs[i]) cudaFree(h->trie_bufs[i]);
    h->trie_bufs[i] = nullptr;
  }
  h->trie_set = false;
}

void free_item_cache(gram_handle* h) {
  if (h->item_mem) cudaFree(h->item_mem);
  if (h->item_valid) cudaFree(h->item_valid);
  if (h->item_len) cudaFree(h->item_len);
  h->item_mem = nullptr; h->item_valid = nullptr; h->item_len = nullptr;
  h->n_items = 0; h->item_L = 0;
}

template <typename P>
int dalloc(gram_handle* h, P** out, size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    char b[256];
    snprintf(b, sizeof b, "cudaMalloc(%zu bytes) failed: %s (handle already owns %zu bytes)", bytes,
             cudaGetErrorString(e), h->alloc_bytes);
    h->err = b;
    return GRAM_ERR_CUDA;
  }
  h->allocs.push_back(p);
  h->alloc_bytes += bytes;
  *out = reinterpret_cast<P*>(p);
  return GRAM_OK;
}
#define DA(ptr, bytes)                                         \
  do {                                                         \
    int rc_ = dalloc(h, &(ptr), (bytes));                      \
    if (rc_) return rc_;                                       \
  } while (0)

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- launch bookkeeping -----------------------------------------------------------------------
struct Scope {
  gram_handle* h; int cls; cudaStream_t s; bool timed; cudaEvent_t a{}, b{};
  Scope(gram_handle* h_, int cls_, cudaStream_t s_) : h(h_), cls(cls_), s(s_) {
    h->launches++;
    h->cls_launches[cls]++;
    timed = (h->prof_mask >> cls) & 1u;
    if (timed) {
      while (h->ev_used + 2 > h->ev_pool.size()) {
        cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e);
      }
      a = h->ev_pool[h->ev_used++]; b = h->ev_pool[h->ev_used++];
      cudaEventRecord(a, s);
    }
  }
  ~Scope() {
    if (timed) { cudaEventRecord(b, s); h->ev_log.push_back({cls, a, b}); }
  }
};

__global__ void convert_kernel_f32(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
__global__ void convert_kernel_bf16(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void fold_gain_kernel(const float* __restrict__ w, const float* __restrict__ ln, bf16* __restrict__ out, size_t n, int cols) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(w[i] * ln[i % cols]);
}
__global__ void build_lut_kernel(const float* __restrict__ rel, const int* __restrict__ buckets, int n, int H,
                                 float* __restrict__ lut) {
  // lut[h][i] = rel[buckets[i]][h]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * H) {
    const int hh = i / n, j = i % n;
    lut[i] = rel[buckets[j] * H + hh];
  }
}

int gemm(gram_handle* h, int cls, int epi, const void* A, const void* W, void* C, int M_max, const int* m_ptr, int N,
         int K, cudaStream_t s, const GemmNormAux* aux = nullptr) {
  Scope sc(h, cls, s);
  cudaError_t e;
  if (h->cfg.dtype == GRAM_DTYPE_BF16 && !(h->cfg.flags & GRAM_FLAG_SIMT_GEMM) && gemm_tc_supported(N, K)) {
    e = gemm_tc(epi, A, W, C, M_max, m_ptr, N, K, h->num_sms, (h->cfg.flags & GRAM_FLAG_GEMM_1CTA) ? 1 : 2, s, aux);
    if (e != cudaSuccess) {
      h->err = std::string("tcgen05 gemm launch failed: ") + cudaGetErrorString(e) + " / " + gemm_tc_last_error();
      return GRAM_ERR_CUDA;
    }
    return GRAM_OK;
  }
  if (epi == EPI_LSE || epi == EPI_RESID_NORM || aux) { h->err = "this epilogue needs the tcgen05 GEMM"; return GRAM_ERR_STATE; }
  e = gemm_simt(h->cfg.dtype, epi, A, W, C, M_max, m_ptr, N, K, s);
  if (e != cudaSuccess) {
    h->err = std::string("simt gemm launch failed: ") + cudaGetErrorString(e);
    return GRAM_ERR_CUDA;
  }
  return GRAM_OK;
}
#define RC(call)                 \
  do {                           \
    int rc_ = (call);            \
    if (rc_) return rc_;         \
  } while (0)
#define CKL(cls, call)                                                                      \
  do {                                                                                      \
    Scope sc_(h, (cls), s);                                                                 \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      h->err = std::string(#call) + " failed: " + cudaGetErrorString(e_);                   \
      return GRAM_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

// ---- weight name resolution ---------------------------------------------------------------------
struct WeightSlot { void* dst; int64_t rows, cols; bool f32; };

bool resolve_weight(gram_handle* h, const std::string& name, WeightSlot* ws) {
  const int D = h->D, HD = h->HD, F = h->F, V = h->V;
  const size_t esz = h->esz;
  auto T = [&](void* base, size_t row_off, int64_t rows, int64_t cols) {
    ws->dst = (char*)base + row_off * (size_t)cols * esz; ws->rows = rows; ws->cols = cols; ws->f32 = false; return true;
  };
  auto F32 = [&](float* base, int64_t rows, int64_t cols) {
    ws->dst = base; ws->rows = rows; ws->cols = cols; ws->f32 = true; return true;
  };
  if (name == "shared") return T(h->shared, 0, V, D);
  if (name == "lm_head") return T(h->lm_head, 0, V, D);
  if (name == "pos_emb") return h->pos_emb ? F32(h->pos_emb, h->cfg.n_positions, D) : false;
  if (name == "enc.final_ln") return F32(h->enc_final_ln, 1, D);
  if (name == "dec.final_ln") return F32(h->dec_final_ln, 1, D);
  char side[8]; int idx; char what[32];
  if (sscanf(name.c_str(), "%3[a-z].%d.%31s", side, &idx, what) != 3) return false;
  const bool is_enc = !strcmp(side, "enc");
  if (!is_enc && strcmp(side, "dec")) return false;
  if (idx < 0 || idx >= (is_enc ? h->Le : h->Ld)) return false;
  LayerW& L = is_enc ? h->enc[idx] : h->dec[idx];
  const std::string w = what;
  if (w == "q") return T(L.qkv, 0, HD, D);
  if (w == "k") return T(L.qkv, HD, HD, D);
  if (w == "v") return T(L.qkv, 2 * (size_t)HD, HD, D);
  if (w == "o") return T(L.o, 0, D, HD);
  if (w == "wi") return T(L.wi, 0, F, D);
  if (w == "wo") return T(L.wo, 0, D, F);
  if (w == "ln0") return F32(L.ln0, 1, D);
  if (w == "ln1") return F32(L.ln1, 1, D);
  if (w == "rel_bias") {
    if (idx != 0) return false;
    return F32(is_enc ? h->enc_rel : h->dec_rel, h->cfg.rel_buckets, h->H);
  }
  if (is_enc) return false;
  if (w == "ln2") return F32(L.ln2, 1, D);
  if (w == "cq") return T(L.cq, 0, HD, D);
  if (w == "co") return T(L.co, 0, D, HD);
  if (w == "ck") return T(h->ckv_w, (size_t)idx * 2 * HD, HD, D);
  if (w == "cv") return T(h->ckv_w, (size_t)idx * 2 * HD + HD, HD, D);
  return false;
}

std::vector<std::string> required_weights(const gram_handle* h) {
  std::vector<std::string> r = {"shared", "lm_head", "enc.final_ln", "dec.final_ln", "enc.0.rel_bias", "dec.0.rel_bias"};
  if (h->cfg.n_positions > 0) r.push_back("pos_emb");
  for (int i = 0; i < h->Le; ++i)
    for (const char* w : {"q", "k", "v", "o", "wi", "wo", "ln0", "ln1"}) r.push_back("enc." + std::to_string(i) + "." + w);
  for (int i = 0; i < h->Ld; ++i)
    for (const char* w : {"q", "k", "v", "o", "cq", "ck", "cv", "co", "wi", "wo", "ln0", "ln1", "ln2"})
      r.push_back("dec." + std::to_string(i) + "." + w);
  return r;
}

int next_pow2(int v) { int n = 1; while (n < v) n <<= 1; return n; }

// ---- encoder --------------------------------------------------------------------------------------
// embedding + all encoder blocks over the packed rows described by `pm` (P passages of at most L tokens); leaves the
// residual stream in h->x, ready for the final norm
int encoder_stack(gram_handle* h, const PackMeta& pm, int P, int L, int Mmax, cudaStream_t s) {
  const gram_config& c = h->cfg;
  const int* mp = pm.total;
  const int D = h->D, HD = h->HD, F = h->F;
  // RMSNorm folded into the GEMMs (kernels.h: GemmNormAux): the producer of the residual stream (embedding, o and wo
  // projections) emits xn = bf16(x * ln_w) and the row sums of squares, the consumer (q|k|v, wi) scales its output rows
  const bool fused = c.dtype == GRAM_DTYPE_BF16 && !(c.flags & (GRAM_FLAG_UNFUSED_NORM | GRAM_FLAG_SIMT_GEMM)) &&
                     (D % 128) == 0 && gemm_tc_supported(3 * HD, D) && gemm_tc_supported(D, HD) &&
                     gemm_tc_supported(F, D) && gemm_tc_supported(D, F);
  const bool chained = fused && (c.flags & GRAM_FLAG_ENC_CHAIN) && h->ffs != nullptr;
  GemmNormAux scaled;                  // consumer side
  scaled.row_ss = h->ss; scaled.eps = c.ln_eps;
  if (stream_mode(h)) {
    // bf16 residual stream: h->xn IS the stream (h->x is not touched); see the bf16 residual stream
    GemmNormAux upd; upd.ss_out = h->ss;
    CKL(GRAM_K_OTHER, embed_rows_stream(h->shared, pm.tok_id, h->xn, h->ss, Mmax, mp, D, s));
    for (int l = 0; l < h->Le; ++l) {
      const LayerW& W = h->enc[l];
      RC(gemm(h, GRAM_K_GEMM_ENC, EPI_STORE, h->xn, W.qkv_g, h->qkv, Mmax, mp, 3 * HD, D, s, &scaled));
      if (!(c.flags & GRAM_FLAG_MMA_ENC_ATTN) && enc_attention_tc_supported(h->dk, L, h->Lb, h->H) &&
          (L <= 128 || !(c.flags & GRAM_FLAG_MMA_LONG_ATTN))) {
        CKL(GRAM_K_ENC_ATTN, enc_attention_tc(h->qkv, (size_t)h->Mcap + 256, h->ao, pm.plen, pm.poff, pm.tok_valid,
                                              h->enc_bias_lut, h->Lb, P, h->H, L, s));
      } else if (enc_attention_mma_supported(h->dk, L)) {
        CKL(GRAM_K_ENC_ATTN, enc_attention_mma(h->qkv, h->ao, pm.plen, pm.poff, pm.tok_valid, h->enc_bias_lut,
                                               h->Lb, P, h->H, L, s));
      } else {
        CKL(GRAM_K_ENC_ATTN, enc_attention(c.dtype, h->qkv, h->ao, pm.plen, pm.poff, pm.tok_valid,
                                           h->enc_bias_lut, h->Lb, P, h->H, h->dk, L, s));
      }
      RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RESID_BF16, h->ao, W.o, h->xn, Mmax, mp, D, HD, s, &upd));
      RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RELU, h->xn, W.wi_g, h->ff, Mmax, mp, F, D, s, &scaled));
      RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RESID_BF16, h->ff, W.wo, h->xn, Mmax, mp, D, F, s, &upd));
    }
    return GRAM_OK;
  }
  auto produce = [&](const float* ln_w) { GemmNormAux a; a.xb = h->xn; a.ss_out = h->ss; a.ln_w = ln_w; return a; };
  if (fused) CKL(GRAM_K_OTHER, embed_rows_norm(h->shared, pm.tok_id, h->x, h->xn, h->ss, h->enc[0].ln0, Mmax, mp, D, s));
  else CKL(GRAM_K_OTHER, embed_rows(c.dtype, h->shared, pm.tok_id, h->x, Mmax, mp, D, s));
  for (int l = 0; l < h->Le; ++l) {
    const LayerW& W = h->enc[l];
    if (!fused) CKL(GRAM_K_NORM_ENC, rmsnorm_rows(c.dtype, h->x, W.ln0, h->xn, Mmax, mp, D, c.ln_eps, 1.f, nullptr, nullptr, s));
    RC(gemm(h, GRAM_K_GEMM_ENC, EPI_STORE, h->xn, W.qkv, h->qkv, Mmax, mp, 3 * HD, D, s, fused ? &scaled : nullptr));
    if (c.dtype == GRAM_DTYPE_BF16 && !(c.flags & (GRAM_FLAG_MMA_ENC_ATTN | GRAM_FLAG_SIMT_ATTN)) &&
        enc_attention_tc_supported(h->dk, L, h->Lb, h->H) && (L <= 128 || !(c.flags & GRAM_FLAG_MMA_LONG_ATTN))) {
      CKL(GRAM_K_ENC_ATTN, enc_attention_tc(h->qkv, (size_t)h->Mcap + 256, h->ao, pm.plen, pm.poff, pm.tok_valid,
                                            h->enc_bias_lut, h->Lb, P, h->H, L, s));
    } else if (c.dtype == GRAM_DTYPE_BF16 && !(c.flags & GRAM_FLAG_SIMT_ATTN) && enc_attention_mma_supported(h->dk, L)) {
      CKL(GRAM_K_ENC_ATTN, enc_attention_mma(h->qkv, h->ao, pm.plen, pm.poff, pm.tok_valid, h->enc_bias_lut,
                                             h->Lb, P, h->H, L, s));
    } else {
      CKL(GRAM_K_ENC_ATTN, enc_attention(c.dtype, h->qkv, h->ao, pm.plen, pm.poff, pm.tok_valid,
                                         h->enc_bias_lut, h->Lb, P, h->H, h->dk, L, s));
    }
    if (fused && chained) {
      // o-projection -> RMSNorm -> wi -> ReLU -> wo (+ the next layer's RMSNorm) as ONE persistent launch: each CTA walks
      // 128-row blocks and hands ff / xn from GEMM to GEMM through its L2-resident scratch (gemm_chain.cu)
      CKL(GRAM_K_GEMM_ENC, enc_chain(h->ao, W.o, h->x, h->xn, h->ss, W.wi, W.wo, h->ffs, W.ln1,
                                     l + 1 < h->Le ? h->enc[l + 1].ln0 : nullptr, c.ln_eps, Mmax, mp, D, HD, F, h->num_sms,
                                     (c.flags & GRAM_FLAG_NO_L2_HINTS) ? 0 : 1, h->pm.err, s));
    } else if (fused) {
      const GemmNormAux a1 = produce(W.ln1);
      RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RESID_NORM, h->ao, W.o, h->x, Mmax, mp, D, HD, s, &a1));
      RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RELU, h->xn, W.wi, h->ff, Mmax, mp, F, D, s, &scaled));
      if (l + 1 < h->Le) {
        const GemmNormAux a0 = produce(h->enc[l + 1].ln0);
        RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RESID_NORM, h->ff, W.wo, h->x, Mmax, mp, D, F, s, &a0));
      } else {
        RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RESID, h->ff, W.wo, h->x, Mmax, mp, D, F, s));   // the final norm is a kernel
      }
    } else {
      RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RESID, h->ao, W.o, h->x, Mmax, mp, D, HD, s));
      CKL(GRAM_K_NORM_ENC, rmsnorm_rows(c.dtype, h->x, W.ln1, h->xn, Mmax, mp, D, c.ln_eps, 1.f, nullptr, nullptr, s));
      RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RELU, h->xn, W.wi, h->ff, Mmax, mp, F, D, s));
      RC(gemm(h, GRAM_K_GEMM_ENC, EPI_RESID, h->ff, W.wo, h->x, Mmax, mp, D, F, s));
    }
  }
  return GRAM_OK;
}

// final encoder norm over whichever residual stream encoder_stack left behind
cudaError_t enc_final_norm(gram_handle* h, int out_dtype, void* out, int M_max, const int* m_ptr, const float* pos,
                           const int* tok_pos, cudaStream_t s) {
  if (stream_mode(h))
    return rmsnorm_rows_stream(out_dtype, h->xn, h->enc_final_ln, out, M_max, m_ptr, h->D, h->cfg.ln_eps, 1.f, pos, tok_pos, s);
  return rmsnorm_rows(out_dtype, h->x, h->enc_final_ln, out, M_max, m_ptr, h->D, h->cfg.ln_eps, 1.f, pos, tok_pos, s);
}

int check_encode_args(gram_handle* h, int B, int N, int L) {
  const gram_config& c = h->cfg;
  if (!h->weights_ready) return fail(h, GRAM_ERR_STATE, "gram_encode: weights not finalised");
  if (B <= 0 || N <= 0 || L <= 0) return fail(h, GRAM_ERR_INVALID, "gram_encode: empty batch");
  if (B > c.max_users || N > c.max_passages || L > c.max_seq_len)
    return fail(h, GRAM_ERR_INVALID, "gram_encode: B/N/L exceed the capacities given to gram_create");
  if ((int64_t)B * N * L > 0x7fffffff) return fail(h, GRAM_ERR_INVALID, "gram_encode: B*N*L exceeds 2^31");
  if (c.n_positions > 0 && N > c.n_positions)
    return fail(h, GRAM_ERR_INVALID, "gram_encode: more passages than rows in the position table");
  return GRAM_OK;
}

// device copies of the caller's ids / mask: host pointers always go through the handle's staging buffers; with `always`
// device pointers do too (a captured graph must read fixed addresses)
int stage_inputs(gram_handle* h, const int64_t* ids, const uint8_t* mask, size_t n, bool always, const int64_t** dids,
                 const uint8_t** dmask, cudaStream_t s) {
  *dids = ids; *dmask = mask;
  const bool dev_i = is_device_ptr(ids), dev_m = is_device_ptr(mask);
  if (!dev_i || always) { CK(cudaMemcpyAsync(h->d_ids, ids, n * sizeof(int64_t), dev_i ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s)); *dids = h->d_ids; }
  if (!dev_m || always) { CK(cudaMemcpyAsync(h->d_mask, mask, n, dev_m ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s)); *dmask = h->d_mask; }
  return GRAM_OK;
}

// pack + encoder + fused memory + cross-attention K/V of a batch whose ids / mask are on the device
int enqueue_encode(gram_handle* h, const int64_t* dids, const uint8_t* dmask, int B, int N, int L, cudaStream_t s) {
  const gram_config& c = h->cfg;
  const size_t n = (size_t)B * N * L;
  h->encoded = false;
  h->enc_B = B; h->enc_N = N; h->enc_L = L;
  // rows the batch can occupy: its padded size, or the workspace when max_tokens caps it below that (the VALID
  // token count is checked against the workspace on the device, by the packing scan)
  const int Mmax = (int)std::min<int64_t>((int64_t)n, h->Mcap);
  CKL(GRAM_K_OTHER, enc_pack(dids, dmask, B, N, L, h->pm, s));
  h->launches += 3;   // enc_pack issues four kernels
  RC(encoder_stack(h, h->pm, B * N, L, Mmax, s));
  CKL(GRAM_K_NORM_ENC, enc_final_norm(h, c.dtype, h->mem, Mmax, h->pm.total, h->pos_emb, h->pos_emb ? h->pm.tok_pos : nullptr, s));
  // cross-attention K/V of every decoder layer, written in place in the layout kernel (b) reads
  RC(gemm(h, GRAM_K_GEMM_KV, EPI_STORE, h->mem, h->ckv_w, h->ckv, Mmax, h->pm.total, h->Ld * 2 * h->HD, h->D, s));
  h->encoded = true;
  return GRAM_OK;
}

int run_encode(gram_handle* h, const int64_t* ids, const uint8_t* mask, int B, int N, int L, cudaStream_t s) {
  RC(check_encode_args(h, B, N, L));
  const int64_t* dids; const uint8_t* dmask;
  RC(stage_inputs(h, ids, mask, (size_t)B * N * L, false, &dids, &dmask, s));
  return enqueue_encode(h, dids, dmask, B, N, L, s);
}

// ---- one decoder step (all layers + lm_head) for R rows ---------------------------------------------
// live: the step decodes only the compact slots of h->live (their count *mp is known on the device alone; every launch
// is sized for R and bounded by *mp, like the encoder's packed token count)
int decoder_step(gram_handle* h, int R, int K, int users, int t, const int* anc, bool fused_lse, bool live, cudaStream_t s) {
  const gram_config& c = h->cfg;
  const int D = h->D, HD = h->HD, F = h->F;
  const size_t esz = h->esz;
  const size_t layer_cache = (size_t)c.max_length * h->Rcap * HD * esz;
  const LiveMap& lm = h->live;
  const int* mp = live ? lm.n_live : nullptr;
  const int* slot_row = live ? lm.slot_row : nullptr;
  const int *lstart = live ? lm.start : nullptr, *lcount = live ? h->bs.live_cnt : nullptr;
  // bf16: the three RMSNorms of every decoder layer are folded into the GEMMs around them exactly as in the encoder (the
  // producer of the residual stream emits bf16(x * w) and the row sums of squares, the consumer scales its output rows),
  // and cross-attention output projection -> wi -> wo run as one chain launch (gemm_chain.cu): 6 launches per layer
  // instead of 12.  fp32 parity mode and GRAM_FLAG_UNFUSED_NORM keep the separate kernels.
  const bool fused = c.dtype == GRAM_DTYPE_BF16 && !(c.flags & (GRAM_FLAG_UNFUSED_NORM | GRAM_FLAG_SIMT_GEMM)) &&
                     (D % 128) == 0 && gemm_tc_supported(3 * HD, D) && gemm_tc_supported(D, HD) &&
                     gemm_tc_supported(F, D) && gemm_tc_supported(D, F) && gemm_tc_supported(HD, D);
  const bool chained = fused && !(c.flags & GRAM_FLAG_NO_DEC_CHAIN) && h->ffs != nullptr;
  GemmNormAux scaled;
  scaled.row_ss = h->dss; scaled.eps = c.ln_eps;
  auto produce = [&](const float* ln_w) { GemmNormAux a; a.xb = h->dxn; a.ss_out = h->dss; a.ln_w = ln_w; return a; };
  const int* toks = live ? lm.tok : h->bs.tok;
  if (fused) CKL(GRAM_K_OTHER, embed_rows_norm(h->shared, toks, h->dx, h->dxn, h->dss, h->dec[0].ln0, R, mp, D, s));
  else CKL(GRAM_K_OTHER, embed_rows(c.dtype, h->shared, toks, h->dx, R, mp, D, s));
  for (int l = 0; l < h->Ld; ++l) {
    const LayerW& W = h->dec[l];
    if (fused) {
      const float* ln_next = l + 1 < h->Ld ? h->dec[l + 1].ln0 : nullptr;
      RC(gemm(h, GRAM_K_GEMM_DEC, EPI_STORE, h->dxn, W.qkv, h->dqkv, R, mp, 3 * HD, D, s, &scaled));
      CKL(GRAM_K_SELF_ATTN, dec_self_attention(c.dtype, h->dqkv, (char*)h->sk + l * layer_cache, (char*)h->sv + l * layer_cache,
                                           anc, c.max_length, h->dec_bias_lut, h->n_dec_lut, h->dao, R, K, h->H, h->dk, t,
                                           slot_row, mp, s));
      const GemmNormAux a1 = produce(W.ln1);
      RC(gemm(h, GRAM_K_GEMM_DEC, EPI_RESID_NORM, h->dao, W.o, h->dx, R, mp, D, HD, s, &a1));
      RC(gemm(h, GRAM_K_GEMM_DEC, EPI_STORE, h->dxn, W.cq, h->dq, R, mp, HD, D, s, &scaled));
      if (!(c.flags & GRAM_FLAG_SIMT_ATTN) && cross_attention_mma_supported(K, h->H, h->dk)) {
        CKL(GRAM_K_CROSS_ATTN, cross_attention_mma(h->dq, h->ckv, (size_t)h->Mcap + 256, (size_t)h->Ld * 2 * HD, l * 2 * HD,
                                                   l * 2 * HD + HD, h->pm.ustart, h->pm.uorder, h->pm.tok_valid, h->dao, users, K, h->H,
                                                   lstart, lcount, (c.flags & GRAM_FLAG_XATTN_PER_ITEM) ? 0 : h->num_sms, s));
      } else {
        CKL(GRAM_K_CROSS_ATTN, cross_attention(c.dtype, h->dq, h->ckv, (size_t)h->Ld * 2 * HD, l * 2 * HD, l * 2 * HD + HD,
                                               h->pm.ustart, h->pm.tok_valid, h->dao, users, K, h->H, h->dk, lstart, lcount, s));
      }
      if (chained) {
        CKL(GRAM_K_GEMM_DEC, enc_chain(h->dao, W.co, h->dx, h->dxn, h->dss, W.wi, W.wo, h->ffs, W.ln2, ln_next, c.ln_eps, R, mp,
                                       D, HD, F, h->num_sms, (c.flags & GRAM_FLAG_NO_L2_HINTS) ? 0 : 1, h->pm.err, s));
      } else {
        const GemmNormAux a2 = produce(W.ln2);
        RC(gemm(h, GRAM_K_GEMM_DEC, EPI_RESID_NORM, h->dao, W.co, h->dx, R, mp, D, HD, s, &a2));
        RC(gemm(h, GRAM_K_GEMM_DEC, EPI_RELU, h->dxn, W.wi, h->dff, R, mp, F, D, s, &scaled));
        if (ln_next) {
          const GemmNormAux a0 = produce(ln_next);
          RC(gemm(h, GRAM_K_GEMM_DEC, EPI_RESID_NORM, h->dff, W.wo, h->dx, R, mp, D, F, s, &a0));
        } else {
          RC(gemm(h, GRAM_K_GEMM_DEC, EPI_RESID, h->dff, W.wo, h->dx, R, mp, D, F, s));
        }
      }
      continue;
    }
    CKL(GRAM_K_NORM_DEC, rmsnorm_rows(c.dtype, h->dx, W.ln0, h->dxn, R, mp, D, c.ln_eps, 1.f, nullptr, nullptr, s));
    RC(gemm(h, GRAM_K_GEMM_DEC, EPI_STORE, h->dxn, W.qkv, h->dqkv, R, mp, 3 * HD, D, s));
    // cache slices are indexed [t][R][HD] with the *current* R as the row pitch
    CKL(GRAM_K_SELF_ATTN, dec_self_attention(c.dtype, h->dqkv, (char*)h->sk + l * layer_cache, (char*)h->sv + l * layer_cache,
                                         anc, c.max_length, h->dec_bias_lut, h->n_dec_lut, h->dao, R, K, h->H, h->dk, t,
                                         slot_row, mp, s));
    RC(gemm(h, GRAM_K_GEMM_DEC, EPI_RESID, h->dao, W.o, h->dx, R, mp, D, HD, s));
    CKL(GRAM_K_NORM_DEC, rmsnorm_rows(c.dtype, h->dx, W.ln1, h->dxn, R, mp, D, c.ln_eps, 1.f, nullptr, nullptr, s));
    RC(gemm(h, GRAM_K_GEMM_DEC, EPI_STORE, h->dxn, W.cq, h->dq, R, mp, HD, D, s));
    if (c.dtype == GRAM_DTYPE_BF16 && !(c.flags & GRAM_FLAG_SIMT_ATTN) && cross_attention_mma_supported(K, h->H, h->dk)) {
      CKL(GRAM_K_CROSS_ATTN, cross_attention_mma(h->dq, h->ckv, (size_t)h->Mcap + 256, (size_t)h->Ld * 2 * HD, l * 2 * HD,
                                                 l * 2 * HD + HD, h->pm.ustart, h->pm.uorder, h->pm.tok_valid, h->dao, users, K, h->H,
                                                 lstart, lcount, (c.flags & GRAM_FLAG_XATTN_PER_ITEM) ? 0 : h->num_sms, s));
    } else {
      CKL(GRAM_K_CROSS_ATTN, cross_attention(c.dtype, h->dq, h->ckv, (size_t)h->Ld * 2 * HD, l * 2 * HD, l * 2 * HD + HD,
                                             h->pm.ustart, h->pm.tok_valid, h->dao, users, K, h->H, h->dk, lstart, lcount, s));
    }
    RC(gemm(h, GRAM_K_GEMM_DEC, EPI_RESID, h->dao, W.co, h->dx, R, mp, D, HD, s));
    CKL(GRAM_K_NORM_DEC, rmsnorm_rows(c.dtype, h->dx, W.ln2, h->dxn, R, mp, D, c.ln_eps, 1.f, nullptr, nullptr, s));
    RC(gemm(h, GRAM_K_GEMM_DEC, EPI_RELU, h->dxn, W.wi, h->dff, R, mp, F, D, s));
    RC(gemm(h, GRAM_K_GEMM_DEC, EPI_RESID, h->dff, W.wo, h->dx, R, mp, D, F, s));
  }
  const float scale = c.tie_word_embeddings ? 1.0f / sqrtf((float)D) : 1.0f;
  CKL(GRAM_K_NORM_DEC, rmsnorm_rows(c.dtype, h->dx, h->dec_final_ln, h->dxn, R, mp, D, c.ln_eps, scale, nullptr, nullptr, s));
  if (fused_lse) {
    // kernel (c): vocabulary projection with the log-softmax statistics fused into the epilogue; logits never stored
    RC(gemm(h, GRAM_K_LM_HEAD, EPI_LSE, h->dxn, h->lm_head, h->lse_partial, R, mp, h->V, D, s));
    CKL(GRAM_K_LM_HEAD, lse_combine(h->lse_partial, h->lse, R, gemm_tc_lse_ntiles(R, h->V, h->num_sms), mp, s));
  } else {
    RC(gemm(h, GRAM_K_LM_HEAD, EPI_F32, h->dxn, h->lm_head, h->logits, R, mp, h->V, D, s));
  }
  return GRAM_OK;
}

// beam search over the batch that is encoded in the handle: init, T decode steps, finalize into h->d_out_*
int enqueue_decode(gram_handle* h, int users, int K, int R_ret, int max_length, cudaStream_t s) {
  const gram_config& c = h->cfg;
  const int R = users * K, T = max_length - 1;
  // the candidate buffer was sized for cfg.max_beams; K <= max_beams so it is sufficient
  BeamState bs = h->bs;
  bs.K = K; bs.max_length = c.max_length; bs.gen_len = max_length;
  CKL(GRAM_K_BEAM, beam_init(bs, h->trie, users, c.start_id, s));
  // fused head (bf16 + tcgen05 GEMM): log-softmax statistics come out of the GEMM epilogue and candidate logits are
  // recomputed from the trie children only; otherwise (fp32 parity mode) full logits are materialised
  // (GRAM_FLAG_KEEP_LOGITS only records the per-step taps: the benchmarked fused head is the one they observe)
  const bool fused = c.dtype == GRAM_DTYPE_BF16 && !(c.flags & (GRAM_FLAG_SIMT_GEMM | GRAM_FLAG_UNFUSED_HEAD)) &&
                     gemm_tc_supported(h->V, h->D) && (h->D % 8) == 0;
  for (int t = 0; t < T; ++t) {
    // step 0: the K beams of a user all hold the start token and attend to the same memory, i.e. K identical rows
    // (HF computes them K times); one row per user is decoded and shared by the user's beams
    const int compact = (t == 0 && K > 1) ? 1 : 0;
    const int Rt = compact ? users : R, Kt = compact ? 1 : K;
    // later steps: only beams that can still reach an output are decoded (dead beams and finished users are compacted
    // away on the device; the reference decodes them and discards the result)
    const bool live = t > 0 && !(c.flags & GRAM_FLAG_ALL_ROWS);
    if (live) {
      CKL(GRAM_K_BEAM, live_compact(bs, users, t & 1, h->live, h->pm.ustart, s));
      h->launches += 1;   // live_compact issues two kernels
    } else {
      CKL(GRAM_K_OTHER, work_add(bs, Rt, h->pm.total, s));   // executed-work accounting (gram_get_stats)
    }
    const int* row_slot = live ? h->live.row_slot : nullptr;
    RC(decoder_step(h, Rt, Kt, users, t, bs.anc[t & 1], fused, live, s));
    if (fused) {
      CKL(GRAM_K_BEAM, beam_step(bs, h->trie, nullptr, h->dxn, h->lm_head, h->D, h->lse, users, t, h->cand_cap, compact, row_slot, s));
    } else {
      CKL(GRAM_K_LM_HEAD, lse_rows(h->logits, h->lse, Rt, h->V, live ? h->live.n_live : nullptr, s));
      CKL(GRAM_K_BEAM, beam_step(bs, h->trie, h->logits, nullptr, nullptr, h->D, h->lse, users, t, h->cand_cap, compact, row_slot, s));
    }
  }
  CKL(GRAM_K_BEAM, beam_finalize(bs, users, T, R_ret, h->d_out_seq, h->d_out_scores, h->d_out_width, s));
  return GRAM_OK;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* gram_version(void) { return "gram_b200 0.1 (sm_100a)"; }

const char* gram_last_error(const gram_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void gram_destroy(gram_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  free_item_cache(h);
  free_trie(h);
  if (h->gstream) cudaStreamDestroy(h->gstream);
  if (h->g_in) cudaEventDestroy(h->g_in);
  if (h->g_out) cudaEventDestroy(h->g_out);
  for (void* p : h->allocs) cudaFree(p);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->len_pow_ev) cudaEventDestroy(h->len_pow_ev);
  if (h->h_len_pow) cudaFreeHost(h->h_len_pow);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  delete h;
}

int gram_create(const gram_config* cfg, gram_handle** out) {
  if (!cfg || !out) { g_create_error = "gram_create: null argument"; return GRAM_ERR_INVALID; }
  *out = nullptr;
  gram_handle* h = new gram_handle();
  h->cfg = *cfg;
  const gram_config& c = h->cfg;
  auto bail = [&](int code) { g_create_error = h->err; gram_destroy(h); return code; };
  if (c.dtype != GRAM_DTYPE_F32 && c.dtype != GRAM_DTYPE_BF16) { h->err = "gram_create: dtype must be 0 (fp32) or 1 (bf16)"; return bail(GRAM_ERR_INVALID); }
  if (c.vocab_size <= 0 || c.d_model <= 0 || c.d_kv <= 0 || c.d_ff <= 0 || c.num_layers <= 0 ||
      c.num_decoder_layers <= 0 || c.num_heads <= 0 || c.max_users <= 0 || c.max_passages <= 0 ||
      c.max_seq_len <= 0 || c.max_beams <= 0 || c.max_length < 2) {
    h->err = "gram_create: non-positive dimension"; return bail(GRAM_ERR_INVALID);
  }
  if (c.d_kv != 16 && c.d_kv != 32 && c.d_kv != 64) { h->err = "gram_create: d_kv must be 16, 32 or 64"; return bail(GRAM_ERR_UNSUPPORTED); }
  if ((c.d_model & 3) || (c.d_ff & 3) || (c.vocab_size & 3)) { h->err = "gram_create: d_model, d_ff and vocab_size must be multiples of 4"; return bail(GRAM_ERR_UNSUPPORTED); }
  if (c.max_beams > 64) { h->err = "gram_create: max_beams > 64 is not supported"; return bail(GRAM_ERR_UNSUPPORTED); }
  if (c.max_length > 64) { h->err = "gram_create: max_length > 64 is not supported"; return bail(GRAM_ERR_UNSUPPORTED); }
  if (c.max_seq_len > 256) { h->err = "gram_create: max_seq_len > 256 is not supported"; return bail(GRAM_ERR_UNSUPPORTED); }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    h->err = "gram_create: no CUDA device visible -- gram_b200 has no CPU fallback"; return bail(GRAM_ERR_CUDA);
  }
  if (c.device < 0 || c.device >= ndev) { h->err = "gram_create: bad device ordinal"; return bail(GRAM_ERR_INVALID); }
  {
    cudaError_t e = cudaSetDevice(c.device);
    if (e != cudaSuccess) { h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return bail(GRAM_ERR_CUDA); }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, c.device);
    if (prop.major != 10) {
      char b[256]; snprintf(b, sizeof b, "gram_create: device %d is sm_%d%d; this library is built for sm_100a only", c.device, prop.major, prop.minor);
      h->err = b; return bail(GRAM_ERR_UNSUPPORTED);
    }
    h->num_sms = prop.multiProcessorCount;
  }
  h->D = c.d_model; h->dk = c.d_kv; h->H = c.num_heads; h->HD = c.num_heads * c.d_kv; h->F = c.d_ff; h->V = c.vocab_size;
  h->Le = c.num_layers; h->Ld = c.num_decoder_layers;
  h->esz = c.dtype == GRAM_DTYPE_F32 ? 4 : 2;
  if (h->HD & 3) { h->err = "gram_create: num_heads*d_kv must be a multiple of 4"; return bail(GRAM_ERR_UNSUPPORTED); }
  const size_t esz = h->esz;
  const int D = h->D, HD = h->HD, F = h->F, V = h->V;
  int rc;
#define DAC(ptr, bytes) do { rc = dalloc(h, &(ptr), (bytes)); if (rc) return bail(rc); } while (0)
  // ---- weights ----
  DAC(h->shared, (size_t)V * D * esz);
  DAC(h->lm_head, (size_t)V * D * esz);
  if (c.n_positions > 0) DAC(h->pos_emb, (size_t)c.n_positions * D * 4);
  DAC(h->enc_final_ln, (size_t)D * 4);
  DAC(h->dec_final_ln, (size_t)D * 4);
  DAC(h->enc_rel, (size_t)c.rel_buckets * h->H * 4);
  DAC(h->dec_rel, (size_t)c.rel_buckets * h->H * 4);
  h->enc.resize(h->Le); h->dec.resize(h->Ld);
  for (int i = 0; i < h->Le; ++i) {
    LayerW& L = h->enc[i];
    DAC(L.qkv, (size_t)3 * HD * D * esz); DAC(L.o, (size_t)D * HD * esz);
    DAC(L.wi, (size_t)F * D * esz); DAC(L.wo, (size_t)D * F * esz);
    DAC(L.ln0, (size_t)D * 4); DAC(L.ln1, (size_t)D * 4);
    if (stream_mode(h)) {
      DAC(L.qkv_f, (size_t)3 * HD * D * 4); DAC(L.wi_f, (size_t)F * D * 4);
      DAC(L.qkv_g, (size_t)3 * HD * D * esz); DAC(L.wi_g, (size_t)F * D * esz);
    }
  }
  for (int i = 0; i < h->Ld; ++i) {
    LayerW& L = h->dec[i];
    DAC(L.qkv, (size_t)3 * HD * D * esz); DAC(L.o, (size_t)D * HD * esz);
    DAC(L.cq, (size_t)HD * D * esz); DAC(L.co, (size_t)D * HD * esz);
    DAC(L.wi, (size_t)F * D * esz); DAC(L.wo, (size_t)D * F * esz);
    DAC(L.ln0, (size_t)D * 4); DAC(L.ln1, (size_t)D * 4); DAC(L.ln2, (size_t)D * 4);
  }
  DAC(h->ckv_w, (size_t)h->Ld * 2 * HD * D * esz);
  h->Lb = c.max_seq_len;
  h->n_enc_lut = 2 * h->Lb - 1;
  h->n_dec_lut = c.max_length;
  DAC(h->enc_bias_lut, (size_t)h->H * h->n_enc_lut * 4);
  DAC(h->dec_bias_lut, (size_t)h->H * h->n_dec_lut * 4);
  // ---- encoder workspace ----
  const int64_t full = (int64_t)c.max_users * c.max_passages * c.max_seq_len;
  h->Mcap = (c.max_tokens > 0 && c.max_tokens < full) ? c.max_tokens : full;
  if (h->Mcap > 0x7fffffff) { h->err = "gram_create: token capacity exceeds 2^31"; return bail(GRAM_ERR_UNSUPPORTED); }
  const size_t Mc = (size_t)h->Mcap + 256;     // slack rows so vector loads/TMA boxes never leave the buffer
  const size_t P = (size_t)c.max_users * c.max_passages;
  DAC(h->pm.plen, P * 4); DAC(h->pm.poff, (P + 1) * 4); DAC(h->pm.ustart, ((size_t)c.max_users + 1) * 4);
  DAC(h->pm.uorder, ((size_t)c.max_users + 1) * 4);
  DAC(h->pm.total, 16);
  DAC(h->pm.tok_id, Mc * 4); DAC(h->pm.tok_pos, Mc * 4); DAC(h->pm.tok_valid, Mc); DAC(h->pm.row_src, Mc * 4);
  DAC(h->d_ids, (size_t)full * 8); DAC(h->d_mask, (size_t)full);
  if (!stream_mode(h)) DAC(h->x, Mc * D * 4);   // the bf16 stream lives in xn
  DAC(h->xn, Mc * D * esz); DAC(h->qkv, Mc * 3 * HD * esz); DAC(h->ao, Mc * HD * esz);
  DAC(h->ff, Mc * std::max<size_t>((size_t)F * esz, (size_t)D * 4));   // also an fp32 [rows, D] scratch (item cache)
  DAC(h->mem, Mc * D * esz);
  DAC(h->ckv, Mc * (size_t)h->Ld * 2 * HD * esz);
  DAC(h->ss, Mc * (size_t)((D + 127) / 128) * 4);
  if (c.dtype == GRAM_DTYPE_BF16 && enc_chain_supported(D, HD, F)) DAC(h->ffs, enc_chain_scratch_bytes(F, h->num_sms));
  // ---- decoder workspace ----
  h->Rcap = c.max_users * c.max_beams;
  const size_t R = (size_t)h->Rcap + 128;
  const int ML = c.max_length;
  DAC(h->dx, R * D * 4);
  DAC(h->dxn, R * D * esz); DAC(h->dqkv, R * 3 * HD * esz); DAC(h->dao, R * HD * esz); DAC(h->dq, R * HD * esz);
  DAC(h->dff, R * F * esz);
  DAC(h->dss, R * (size_t)((D + 127) / 128) * 4);
  DAC(h->sk, (size_t)h->Ld * ML * R * HD * esz); DAC(h->sv, (size_t)h->Ld * ML * R * HD * esz);
  DAC(h->logits, R * V * 4); DAC(h->lse, R * 4);
  DAC(h->lse_partial, R * (size_t)((V + 127) / 128 + 2) * 8);   // + 2: 256-column tiles with two partials each
  BeamState& bs = h->bs;
  bs.max_length = ML; bs.gen_len = ML; bs.V = V; bs.eos = c.eos_id; bs.pad = c.pad_id; bs.K = c.max_beams;
  for (int i = 0; i < 2; ++i) {
    DAC(bs.beam_score[i], R * 4); DAC(bs.node[i], R * 4); DAC(bs.seq[i], R * ML * 4); DAC(bs.anc[i], R * ML * 4);
  }
  DAC(bs.tok, R * 4);
  const size_t U = c.max_users, S = (size_t)c.max_beams + 1;
  DAC(bs.hyp_score, U * S * 8); DAC(bs.hyp_len, U * S * 4); DAC(bs.hyp_seqno, U * S * 4); DAC(bs.hyp_tok, U * S * ML * 4);
  DAC(bs.n_hyp, U * 4); DAC(bs.worst, U * 8); DAC(bs.next_seqno, U * 4); DAC(bs.done, U * 4);
  DAC(bs.live_cnt, U * 4);
  DAC(h->live.start, (U + 1) * 4); DAC(h->live.slot_row, R * 4); DAC(h->live.row_slot, R * 4); DAC(h->live.tok, R * 4);
  DAC(h->live.n_live, 16);
  DAC(bs.err, 16);
  DAC(bs.work, 16);
  h->pm.err = bs.err;
  h->pm.vocab = V;
  h->pm.cap = h->Mcap;
  DAC(h->d_len_pow, ((size_t)ML + 1) * 8);
  bs.len_pow = h->d_len_pow;
  bs.tap_lse = nullptr; bs.tap_score = nullptr; bs.tap_seq = nullptr;
  if (c.flags & GRAM_FLAG_KEEP_LOGITS) {
    DAC(bs.tap_lse, (size_t)ML * R * 4); DAC(bs.tap_score, (size_t)ML * R * 4); DAC(bs.tap_seq, (size_t)ML * R * ML * 4);
  }
  DAC(h->d_out_seq, U * c.max_beams * ML * 8); DAC(h->d_out_scores, U * c.max_beams * 4); DAC(h->d_out_width, 16);
  DAC(h->d_zero_anc, R * ML * 4);
  DAC(h->d_dec_ids, R * ML * 8);
#undef DAC
  // K/V rows past a user's range are read (and masked) by the TMA-fed attention kernel: they must be finite
  if (cudaMemset(h->ckv, 0, Mc * (size_t)h->Ld * 2 * HD * esz) != cudaSuccess ||
      cudaMemset(h->qkv, 0, Mc * 3 * (size_t)HD * esz) != cudaSuccess ||   // rows past a passage are read (masked) by TMA

      cudaMemset(h->d_zero_anc, 0, R * ML * 4) != cudaSuccess || cudaMemset(bs.err, 0, 16) != cudaSuccess || cudaMemset(bs.work, 0, 16) != cudaSuccess ||
      cudaMemset(bs.anc[0], 0, R * ML * 4) != cudaSuccess || cudaMemset(bs.anc[1], 0, R * ML * 4) != cudaSuccess) {
    h->err = "gram_create: cudaMemset failed"; return bail(GRAM_ERR_CUDA);
  }
  if (cudaMallocHost(&h->h_len_pow, ((size_t)ML + 1) * 8) != cudaSuccess || cudaMallocHost(&h->h_flags, 64) != cudaSuccess) {
    h->err = "gram_create: cudaMallocHost failed"; return bail(GRAM_ERR_CUDA);
  }
  *out = h;
  return GRAM_OK;
}

int gram_load_weight(gram_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  if (!h || !name || !data || !shape) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  WeightSlot ws;
  if (!resolve_weight(h, name, &ws)) return fail(h, GRAM_ERR_INVALID, std::string("gram_load_weight: unknown tensor name '") + name + "'");
  int64_t rows = 1, cols = 1;
  if (ndim == 1) { cols = shape[0]; }
  else if (ndim == 2) { rows = shape[0]; cols = shape[1]; }
  else return fail(h, GRAM_ERR_INVALID, "gram_load_weight: ndim must be 1 or 2");
  if (rows != ws.rows || cols != ws.cols) {
    char b[256];
    snprintf(b, sizeof b, "gram_load_weight: '%s' has shape [%lld,%lld], expected [%lld,%lld]", name, (long long)rows,
             (long long)cols, (long long)ws.rows, (long long)ws.cols);
    return fail(h, GRAM_ERR_INVALID, b);
  }
  const size_t n = (size_t)rows * cols;
  if (n > h->stage_elems) {
    float* p = nullptr;
    CK(cudaMalloc(&p, n * 4));
    if (h->stage) cudaFree(h->stage);
    h->stage = p; h->stage_elems = n;
  }
  CK(cudaMemcpy(h->stage, data, n * 4, cudaMemcpyHostToDevice));
  const int grid = (int)((n + 255) / 256);
  if (ws.f32 || h->cfg.dtype == GRAM_DTYPE_F32) convert_kernel_f32<<<grid, 256>>>(h->stage, (float*)ws.dst, n);
  else convert_kernel_bf16<<<grid, 256>>>(h->stage, (bf16*)ws.dst, n);
  CK(cudaGetLastError());
  {
    // the bf16 residual stream: the encoder's q|k|v and wi are folded with their RMSNorm gain from the fp32 values at finalize
    int idx; char what[32];
    if (stream_mode(h) && sscanf(name, "enc.%d.%31s", &idx, what) == 2 && idx >= 0 && idx < h->Le) {
      LayerW& L = h->enc[idx];
      const std::string w = what;
      float* keep = w == "q" ? L.qkv_f : w == "k" ? L.qkv_f + (size_t)h->HD * h->D : w == "v" ? L.qkv_f + 2 * (size_t)h->HD * h->D
                    : w == "wi" ? L.wi_f : nullptr;
      if (keep) convert_kernel_f32<<<grid, 256>>>(h->stage, keep, n);
      CK(cudaGetLastError());
    }
  }
  CK(cudaDeviceSynchronize());
  h->loaded.insert(name);
  h->weights_ready = false;
  return GRAM_OK;
}

int gram_set_rel_buckets(gram_handle* h, const int32_t* enc_buckets, int32_t n_enc, const int32_t* dec_buckets,
                         int32_t n_dec) {
  if (!h || !enc_buckets || !dec_buckets) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  if (n_enc != h->n_enc_lut || n_dec != h->n_dec_lut) {
    char b[256];
    snprintf(b, sizeof b, "gram_set_rel_buckets: expected n_enc=%d (2*max_seq_len-1) and n_dec=%d (max_length)", h->n_enc_lut, h->n_dec_lut);
    return fail(h, GRAM_ERR_INVALID, b);
  }
  for (int i = 0; i < n_enc; ++i) if (enc_buckets[i] < 0 || enc_buckets[i] >= h->cfg.rel_buckets) return fail(h, GRAM_ERR_INVALID, "bucket out of range");
  for (int i = 0; i < n_dec; ++i) if (dec_buckets[i] < 0 || dec_buckets[i] >= h->cfg.rel_buckets) return fail(h, GRAM_ERR_INVALID, "bucket out of range");
  // stash the bucket ids in the (int-sized) LUT buffers; finalize turns them into bias values
  CK(cudaMemcpy(h->enc_bias_lut, enc_buckets, (size_t)n_enc * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->dec_bias_lut, dec_buckets, (size_t)n_dec * 4, cudaMemcpyHostToDevice));
  h->buckets_set = true;
  h->weights_ready = false;
  return GRAM_OK;
}

int gram_finalize_weights(gram_handle* h) {
  if (!h) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  for (const std::string& n : required_weights(h))
    if (!h->loaded.count(n)) return fail(h, GRAM_ERR_STATE, "gram_finalize_weights: tensor '" + n + "' was never loaded");
  if (!h->buckets_set) return fail(h, GRAM_ERR_STATE, "gram_finalize_weights: gram_set_rel_buckets not called");
  // bias LUTs: lut[h][i] = rel[bucket[i]][h]
  int* tmp = nullptr;
  const int nmax = h->n_enc_lut > h->n_dec_lut ? h->n_enc_lut : h->n_dec_lut;
  CK(cudaMalloc(&tmp, (size_t)nmax * 4));
  CK(cudaMemcpy(tmp, h->enc_bias_lut, (size_t)h->n_enc_lut * 4, cudaMemcpyDeviceToDevice));
  build_lut_kernel<<<(h->n_enc_lut * h->H + 255) / 256, 256>>>(h->enc_rel, tmp, h->n_enc_lut, h->H, h->enc_bias_lut);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(tmp, h->dec_bias_lut, (size_t)h->n_dec_lut * 4, cudaMemcpyDeviceToDevice));
  build_lut_kernel<<<(h->n_dec_lut * h->H + 255) / 256, 256>>>(h->dec_rel, tmp, h->n_dec_lut, h->H, h->dec_bias_lut);
  CK(cudaDeviceSynchronize());
  cudaFree(tmp);
  if (stream_mode(h)) {
    for (int l = 0; l < h->Le; ++l) {
      LayerW& L = h->enc[l];
      const size_t nq = (size_t)3 * h->HD * h->D, nw = (size_t)h->F * h->D;
      fold_gain_kernel<<<(int)((nq + 255) / 256), 256>>>(L.qkv_f, L.ln0, (bf16*)L.qkv_g, nq, h->D);
      fold_gain_kernel<<<(int)((nw + 255) / 256), 256>>>(L.wi_f, L.ln1, (bf16*)L.wi_g, nw, h->D);
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  if (h->stage) { cudaFree(h->stage); h->stage = nullptr; h->stage_elems = 0; }
  h->buckets_set = false;    // LUT buffers now hold values; buckets must be re-sent before another finalize
  h->weights_ready = true;
  return GRAM_OK;
}

int gram_set_trie(gram_handle* h, const int32_t* child_offsets, const int32_t* child_tokens, const int32_t* child_nodes,
                  int32_t n_nodes, int32_t n_edges, int32_t root_node) {
  if (!h || !child_offsets || n_nodes <= 0 || n_edges < 0) return GRAM_ERR_INVALID;
  if (n_edges > 0 && (!child_tokens || !child_nodes)) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  // root_node < 0: no item starts with the decoder start token (Trie.to_csr) -- every beam would be dead and the
  // result all padding; the reference's HF loop raises in that situation, so refuse the trie here
  if (root_node < 0 || root_node >= n_nodes)
    return fail(h, GRAM_ERR_INVALID, "gram_set_trie: root_node out of range (no item sequence starts with the decoder start token?)");
  if (child_offsets[0] != 0 || child_offsets[n_nodes] != n_edges) return fail(h, GRAM_ERR_INVALID, "gram_set_trie: malformed CSR offsets");
  int max_fan = 0;
  for (int i = 0; i < n_nodes; ++i) {
    const int f = child_offsets[i + 1] - child_offsets[i];
    if (f < 0) return fail(h, GRAM_ERR_INVALID, "gram_set_trie: offsets must be non-decreasing");
    if (f > max_fan) max_fan = f;
  }
  for (int e = 0; e < n_edges; ++e) {
    if (child_tokens[e] < 0 || child_tokens[e] >= h->V) return fail(h, GRAM_ERR_INVALID, "gram_set_trie: token id outside the vocabulary");
    if (child_nodes[e] < 0 || child_nodes[e] >= n_nodes) return fail(h, GRAM_ERR_INVALID, "gram_set_trie: child node out of range");
  }
  const int cap = next_pow2((max_fan > 0 ? max_fan : 1) * h->cfg.max_beams);
  if (beam_step_smem(cap) > 200 * 1024)
    return fail(h, GRAM_ERR_UNSUPPORTED, "gram_set_trie: max_beams * max trie fan-out exceeds the shared-memory candidate buffer");
  // the CSR arrays live outside h->allocs: a re-upload (Trie.add bumps the version) frees the previous ones
  // (cudaFree waits for the kernels that may still read them)
  free_trie(h);
  int *d_off = nullptr, *d_tok = nullptr, *d_node = nullptr;
  if (cudaMalloc(&d_off, ((size_t)n_nodes + 1) * 4) != cudaSuccess || cudaMalloc(&d_tok, ((size_t)n_edges + 1) * 4) != cudaSuccess ||
      cudaMalloc(&d_node, ((size_t)n_edges + 1) * 4) != cudaSuccess) {
    cudaGetLastError();
    if (d_off) cudaFree(d_off);
    if (d_tok) cudaFree(d_tok);
    return fail(h, GRAM_ERR_CUDA, "gram_set_trie: cudaMalloc of the CSR arrays failed");
  }
  h->trie_bufs[0] = d_off; h->trie_bufs[1] = d_tok; h->trie_bufs[2] = d_node;
  CK(cudaMemcpy(d_off, child_offsets, ((size_t)n_nodes + 1) * 4, cudaMemcpyHostToDevice));
  if (n_edges) {
    CK(cudaMemcpy(d_tok, child_tokens, (size_t)n_edges * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_node, child_nodes, (size_t)n_edges * 4, cudaMemcpyHostToDevice));
  }
  h->trie.child_offsets = d_off; h->trie.child_tokens = d_tok; h->trie.child_nodes = d_node;
  h->trie.n_nodes = n_nodes; h->trie.n_edges = n_edges; h->trie.root = root_node; h->trie.max_fanout = max_fan;
  h->cand_cap = cap < 64 ? 64 : cap;
  h->trie_set = true;
  return GRAM_OK;
}

int gram_encode(gram_handle* h, const int64_t* ids, const uint8_t* mask, int32_t B, int32_t N, int32_t L, void* stream) {
  if (!h || !ids || !mask) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  h->launches = 0;
  return run_encode(h, ids, mask, B, N, L, (cudaStream_t)stream);
}

// ---- per-item encoder-state cache ((f)-1) ---------------------------------------------------------------------
int gram_cache_items(gram_handle* h, const int64_t* ids, const uint8_t* mask, int32_t n_items, int32_t L, void* stream) {
  if (!h || !ids || !mask) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const gram_config& c = h->cfg;
  if (!h->weights_ready) return fail(h, GRAM_ERR_STATE, "gram_cache_items: weights not finalised");
  if (n_items <= 0 || L <= 0 || L > c.max_seq_len) return fail(h, GRAM_ERR_INVALID, "gram_cache_items: bad n_items / L");
  CK(cudaStreamSynchronize(s));
  free_item_cache(h);
  h->encoded = false;
  const size_t rows = (size_t)n_items * L;
  {
    cudaError_t e1 = cudaMalloc(&h->item_mem, rows * h->D * 4), e2 = cudaMalloc(&h->item_valid, rows),
                e3 = cudaMalloc(&h->item_len, (size_t)n_items * 4);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
      cudaGetLastError();
      free_item_cache(h);
      return fail(h, GRAM_ERR_CUDA, "gram_cache_items: cudaMalloc of the item table failed");
    }
  }
  CK(cudaMemsetAsync(h->item_valid, 0, rows, s));
  if (!h->pm_prompt.plen) {
    // layout arrays of the prompt-only encoder pass (one passage per user)
    const size_t U = (size_t)c.max_users, Mp = U * c.max_seq_len + 256;
    int rc;
#define DAC2(ptr, bytes) do { rc = dalloc(h, &(ptr), (bytes)); if (rc) return rc; } while (0)
    DAC2(h->pm_prompt.plen, U * 4); DAC2(h->pm_prompt.poff, (U + 1) * 4); DAC2(h->pm_prompt.ustart, (U + 1) * 4);
    DAC2(h->pm_prompt.uorder, (U + 1) * 4); DAC2(h->pm_prompt.total, 16);
    DAC2(h->pm_prompt.tok_id, Mp * 4); DAC2(h->pm_prompt.tok_pos, Mp * 4); DAC2(h->pm_prompt.tok_valid, Mp);
    DAC2(h->pm_prompt.row_src, Mp * 4);
    DAC2(h->d_items, U * c.max_passages * 4);
#undef DAC2
    h->pm_prompt.err = h->pm.err;
    h->pm_prompt.vocab = h->V;
    // the prompt pass writes the same Mcap-row workspace as an ordinary batch: the packing scan empties a batch whose
    // prompts alone exceed it (err 4) before any kernel writes a row
    h->pm_prompt.cap = std::min<long long>((long long)U * c.max_seq_len, (long long)h->Mcap);
  }
  // encode the items in chunks shaped like an ordinary batch [Bc, N, L] (item = first + b*N + n)
  const int N = c.max_passages;
  int64_t per_chunk = (int64_t)c.max_users * N;
  if (per_chunk * L > h->Mcap) per_chunk = h->Mcap / L;
  if (per_chunk <= 0) return fail(h, GRAM_ERR_INVALID, "gram_cache_items: token capacity below one passage");
  const bool dev_in = is_device_ptr(ids), dev_mask = is_device_ptr(mask);
  for (int64_t first = 0; first < n_items; first += per_chunk) {
    const int64_t cnt = std::min<int64_t>(per_chunk, n_items - first);
    const int Bc = (int)((cnt + N - 1) / N);
    const size_t full = (size_t)Bc * N * L, have = (size_t)cnt * L;
    CK(cudaMemsetAsync(h->d_mask, 0, full, s));
    CK(cudaMemsetAsync(h->d_ids, 0, full * 8, s));
    CK(cudaMemcpyAsync(h->d_ids, ids + first * L, have * 8, dev_in ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->d_mask, mask + first * L, have, dev_mask ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    const int Mmax = (int)std::min<int64_t>((int64_t)full, h->Mcap);
    CKL(GRAM_K_OTHER, enc_pack(h->d_ids, h->d_mask, Bc, N, L, h->pm, s));
    RC(encoder_stack(h, h->pm, Bc * N, L, Mmax, s));
    // final norm in fp32, WITHOUT the position row (h->ff is free after the last block and holds >= Mcap*D floats)
    CKL(GRAM_K_NORM_ENC, enc_final_norm(h, GRAM_DTYPE_F32, h->ff, Mmax, h->pm.total, nullptr, nullptr, s));
    CKL(GRAM_K_OTHER, cache_scatter((const float*)h->ff, h->pm, Mmax, h->D, first * L, h->item_mem, h->item_valid, s));
    CK(cudaMemcpyAsync(h->item_len + first, h->pm.plen, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, s));
    if (!dev_in || !dev_mask) CK(cudaStreamSynchronize(s));     // the caller's host buffers are consumed per chunk
  }
  CK(cudaMemcpyAsync(&h->h_flags[0], h->pm.err, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (h->h_flags[0] != 0) {
    cudaMemsetAsync(h->pm.err, 0, 4, s);
    free_item_cache(h);
    return fail(h, GRAM_ERR_INVALID, "gram_cache_items: input token id outside [0, vocab_size)");
  }
  h->n_items = n_items; h->item_L = L;
  return GRAM_OK;
}

int gram_encode_cached(gram_handle* h, const int64_t* prompt_ids, const uint8_t* prompt_mask, const int32_t* items,
                       int32_t B, int32_t NI, int32_t L, void* stream) {
  if (!h || !prompt_ids || !prompt_mask || (NI > 0 && !items)) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const gram_config& c = h->cfg;
  if (!h->item_mem) return fail(h, GRAM_ERR_STATE, "gram_encode_cached: no item table (gram_cache_items)");
  if (L != h->item_L) return fail(h, GRAM_ERR_INVALID, "gram_encode_cached: L differs from the L the item table was built with");
  const int N = NI + 1;
  if (B <= 0 || NI < 0) return fail(h, GRAM_ERR_INVALID, "gram_encode_cached: empty batch");
  if (B > c.max_users || N > c.max_passages) return fail(h, GRAM_ERR_INVALID, "gram_encode_cached: B/N exceed the capacities given to gram_create");
  if ((int64_t)B * N * L > 0x7fffffff) return fail(h, GRAM_ERR_INVALID, "gram_encode_cached: B*N*L exceeds 2^31");
  if (c.n_positions > 0 && N > c.n_positions) return fail(h, GRAM_ERR_INVALID, "gram_encode_cached: more passages than rows in the position table");
  h->launches = 0;
  const size_t n = (size_t)B * L;
  const int64_t* dids = prompt_ids;
  const uint8_t* dmask = prompt_mask;
  const int* ditems = items;
  if (!is_device_ptr(prompt_ids)) { CK(cudaMemcpyAsync(h->d_ids, prompt_ids, n * 8, cudaMemcpyHostToDevice, s)); dids = h->d_ids; }
  if (!is_device_ptr(prompt_mask)) { CK(cudaMemcpyAsync(h->d_mask, prompt_mask, n, cudaMemcpyHostToDevice, s)); dmask = h->d_mask; }
  if (NI > 0 && !is_device_ptr(items)) { CK(cudaMemcpyAsync(h->d_items, items, (size_t)B * NI * 4, cudaMemcpyHostToDevice, s)); ditems = h->d_items; }
  h->encoded = false;
  h->enc_B = B; h->enc_N = N; h->enc_L = L;
  // 1. the encoder stack on the prompts alone
  CKL(GRAM_K_OTHER, enc_pack(dids, dmask, B, 1, L, h->pm_prompt, s));
  h->launches += 3;
  const int Mprompt = (int)std::min<int64_t>((int64_t)n, h->Mcap);
  RC(encoder_stack(h, h->pm_prompt, B, L, Mprompt, s));
  CKL(GRAM_K_NORM_ENC, enc_final_norm(h, GRAM_DTYPE_F32, h->ff, Mprompt, h->pm_prompt.total, nullptr, nullptr, s));
  // 2. user layout + memory = prompt rows / cached item rows + position rows
  const int Mmax = (int)std::min<int64_t>((int64_t)B * N * L, h->Mcap);
  CKL(GRAM_K_OTHER, cached_pack_assemble(c.dtype, h->pm, h->pm_prompt, (const float*)h->ff, ditems, h->item_mem, h->item_valid,
                                         h->item_len, h->n_items, h->pos_emb, B, NI, L, h->D, h->mem, s));
  h->launches += 3;
  // 3. cross-attention K/V, as in the uncached path
  RC(gemm(h, GRAM_K_GEMM_KV, EPI_STORE, h->mem, h->ckv_w, h->ckv, Mmax, h->pm.total, h->Ld * 2 * h->HD, h->D, s));
  h->encoded = true;
  return GRAM_OK;
}

int gram_generate(gram_handle* h, const int64_t* ids, const uint8_t* mask, int32_t B, int32_t N, int32_t L, int32_t K,
                  int32_t R_ret, int32_t max_length, const double* len_pow, int64_t* out_seq, int32_t* out_width,
                  float* out_scores, void* stream) {
  if (!h || !len_pow || !out_seq || !out_scores || !out_width) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const gram_config& c = h->cfg;
  if (!h->trie_set) return fail(h, GRAM_ERR_STATE, "gram_generate: no trie set (gram_set_trie)");
  if (K <= 0 || K > c.max_beams) return fail(h, GRAM_ERR_INVALID, "gram_generate: num_beams exceeds max_beams");
  if (R_ret <= 0 || R_ret > K) return fail(h, GRAM_ERR_INVALID, "`num_return_sequences` has to be smaller or equal to `num_beams`.");
  if (max_length < 2 || max_length > c.max_length) return fail(h, GRAM_ERR_INVALID, "gram_generate: max_length outside [2, cfg.max_length]");
  const bool have_ids = ids != nullptr;
  if (have_ids) {
    if (!mask) return GRAM_ERR_INVALID;
    RC(check_encode_args(h, B, N, L));
  } else {
    if (!h->encoded) return fail(h, GRAM_ERR_STATE, "gram_generate: ids == NULL but nothing has been encoded");
    B = h->enc_B; N = h->enc_N; L = h->enc_L;
  }
  const int users = B, R = B * K, T = max_length - 1;
  // GRAM_FLAG_CUDA_GRAPH: everything between the input staging and the result copy-out -- ~70 launches of the encoder and
  // ~40 per decode step, none of which depends on a host-side value -- is captured once per call shape and replayed
  cudaStream_t caller = s;
  const bool graphed = (c.flags & GRAM_FLAG_CUDA_GRAPH) && h->prof_mask == 0;
  if (graphed) {
    if (!h->gstream) {
      CK(cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&h->g_in, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->g_out, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(h->g_in, caller));
    s = h->gstream;
    CK(cudaStreamWaitEvent(s, h->g_in, 0));
  }
  const int64_t* dids = nullptr; const uint8_t* dmask = nullptr;
  if (have_ids) RC(stage_inputs(h, ids, mask, (size_t)B * N * L, graphed, &dids, &dmask, s));
  if (h->len_pow_ev) CK(cudaEventSynchronize(h->len_pow_ev));
  else CK(cudaEventCreateWithFlags(&h->len_pow_ev, cudaEventDisableTiming));
  memcpy(h->h_len_pow, len_pow, ((size_t)max_length + 1) * sizeof(double));
  for (int i = max_length + 1; i <= c.max_length; ++i) h->h_len_pow[i] = 1.0;
  CK(cudaMemcpyAsync(h->d_len_pow, h->h_len_pow, ((size_t)c.max_length + 1) * 8, cudaMemcpyHostToDevice, s));
  CK(cudaEventRecord(h->len_pow_ev, s));
  auto enqueue = [&](cudaStream_t st) -> int {
    if (have_ids) { h->launches = 0; RC(enqueue_encode(h, dids, dmask, B, N, L, st)); }
    return enqueue_decode(h, users, K, R_ret, max_length, st);
  };
  if (!graphed) {
    RC(enqueue(s));
  } else {
    gram_handle::GraphRec& g = h->graphs[{B, N, L, K, R_ret, max_length, have_ids ? 1 : 0}];
    ++g.calls;
    if (g.calls == 1) {
      RC(enqueue(s));                                   // first call of a shape runs eagerly (lazy one-time set-up inside)
    } else if (!g.exec) {
      CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      const int64_t before = have_ids ? 0 : h->launches;
      const int rc = enqueue(s);
      cudaGraph_t graph = nullptr;
      const cudaError_t e = cudaStreamEndCapture(s, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      CK(e);
      CK(cudaGraphInstantiate(&g.exec, graph, 0));
      cudaGraphDestroy(graph);
      g.launches = h->launches - before;
      CK(cudaGraphLaunch(g.exec, s));
    } else {
      if (have_ids) { h->launches = 0; h->enc_B = B; h->enc_N = N; h->enc_L = L; h->encoded = true; }
      h->launches += g.launches;
      CK(cudaGraphLaunch(g.exec, s));
    }
  }
  BeamState bs = h->bs;
  h->last_steps = T; h->last_R = R;
  // ---- copy-out: the library's result layout is [B*R_ret, cfg.max_length]; the ABI promises max_length ----
  const size_t rows = (size_t)users * R_ret;
  const bool host_out = !is_device_ptr(out_seq);
  if (max_length == c.max_length) {
    CK(cudaMemcpyAsync(out_seq, h->d_out_seq, rows * max_length * 8, host_out ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
  } else {
    CK(cudaMemcpy2DAsync(out_seq, (size_t)max_length * 8, h->d_out_seq, (size_t)c.max_length * 8, (size_t)max_length * 8, rows,
                         host_out ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
  }
  CK(cudaMemcpyAsync(out_scores, h->d_out_scores, rows * 4, is_device_ptr(out_scores) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(out_width, h->d_out_width, 4, is_device_ptr(out_width) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&h->h_flags[0], bs.err, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&h->h_flags[2], h->pm.total, 4, cudaMemcpyDeviceToHost, s));
  if (graphed) {                                        // the caller's stream continues after the replay stream
    CK(cudaEventRecord(h->g_out, s));
    CK(cudaStreamWaitEvent(caller, h->g_out, 0));
  }
  if (host_out || !is_device_ptr(out_scores) || !is_device_ptr(out_width)) {
    CK(cudaStreamSynchronize(s));
    if (h->h_flags[0] != 0) {
      const int code = h->h_flags[0];
      cudaMemsetAsync(bs.err, 0, 4, s);
      return fail(h, GRAM_ERR_INVALID, code == 2   ? "gram_generate: input token id outside [0, vocab_size)"
                                       : code == 3 ? "gram_generate: item index outside the cached item table"
                                       : code == 4 ? "gram_generate: the batch holds more valid tokens than max_tokens"
                                       : code == 5 ? "gram_generate: internal error, live beams are not a prefix of the user's beams"
                                       : code == 7 ? "gram_generate: internal error, a barrier of the chain kernel never completed (watchdog)"
                                                   : "gram_generate: candidate buffer overflow (trie fan-out larger than declared)");
    }
  }
  return GRAM_OK;
}

int gram_get_memory(gram_handle* h, float* out, void* stream) {
  if (!h || !out) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->encoded) return fail(h, GRAM_ERR_STATE, "gram_get_memory: nothing encoded");
  const size_t n = (size_t)h->enc_B * h->enc_N * h->enc_L * h->D;
  float* dst = out;
  float* tmp = nullptr;
  const bool dev = is_device_ptr(out);
  if (!dev) { CK(cudaMalloc(&tmp, n * 4)); dst = tmp; }
  CK(cudaMemsetAsync(dst, 0, n * 4, s));
  cudaError_t e = unpack_memory(h->cfg.dtype, h->mem, h->pm.row_src, dst, (int)((size_t)h->enc_B * h->enc_N * h->enc_L), h->pm.total, h->D, s);
  if (e != cudaSuccess) { if (tmp) cudaFree(tmp); return fail(h, GRAM_ERR_CUDA, cudaGetErrorString(e)); }
  if (!dev) {
    CK(cudaMemcpyAsync(out, tmp, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    cudaFree(tmp);
  }
  return GRAM_OK;
}

int gram_decoder_logits(gram_handle* h, const int64_t* dec_ids, int32_t q, float* out_logits, void* stream) {
  if (!h || !dec_ids || !out_logits) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->encoded) return fail(h, GRAM_ERR_STATE, "gram_decoder_logits: nothing encoded");
  if (q <= 0 || q > h->cfg.max_length) return fail(h, GRAM_ERR_INVALID, "gram_decoder_logits: q outside [1, max_length]");
  const int B = h->enc_B, R = B;
  const int64_t* dids = dec_ids;
  if (!is_device_ptr(dec_ids)) {
    CK(cudaMemcpyAsync(h->d_dec_ids, dec_ids, (size_t)R * q * 8, cudaMemcpyHostToDevice, s));
    dids = h->d_dec_ids;
  }
  const bool dev = is_device_ptr(out_logits);
  BeamState bs = h->bs;
  bs.K = 1;
  const size_t V = h->V;
  for (int t = 0; t < q; ++t) {
    CKL(GRAM_K_OTHER, forced_step(bs, dids, q, t, R, s));
    RC(decoder_step(h, R, 1, B, t, h->d_zero_anc, false, false, s));
    // logits [R, V] -> out[b][t][:]
    CK(cudaMemcpy2DAsync(out_logits + (size_t)t * V, (size_t)q * V * 4, h->logits, V * 4, V * 4, R,
                         dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  }
  if (!dev) CK(cudaStreamSynchronize(s));
  return GRAM_OK;
}

int gram_get_step_taps(gram_handle* h, float* lse, float* beam_scores, int32_t* beam_tokens, int32_t* n_steps) {
  if (!h) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  if (!h->bs.tap_lse) return fail(h, GRAM_ERR_STATE, "gram_get_step_taps: handle was created without GRAM_FLAG_KEEP_LOGITS");
  CK(cudaDeviceSynchronize());
  const size_t n = (size_t)h->last_steps * h->last_R;
  if (lse) CK(cudaMemcpy(lse, h->bs.tap_lse, n * 4, cudaMemcpyDefault));
  if (beam_scores) CK(cudaMemcpy(beam_scores, h->bs.tap_score, n * 4, cudaMemcpyDefault));
  if (beam_tokens) CK(cudaMemcpy(beam_tokens, h->bs.tap_seq, n * h->cfg.max_length * 4, cudaMemcpyDefault));
  if (n_steps) *n_steps = h->last_steps;
  return GRAM_OK;
}

int gram_check_errors(gram_handle* h, void* stream) {
  if (!h) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(&h->h_flags[0], h->pm.err, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const int code = h->h_flags[0];
  if (code == 0) return GRAM_OK;
  CK(cudaMemsetAsync(h->pm.err, 0, 4, s));
  return fail(h, GRAM_ERR_INVALID, code == 2   ? "input token id outside [0, vocab_size)"
                                   : code == 3 ? "item index outside the cached item table"
                                   : code == 4 ? "the batch holds more valid tokens than max_tokens"
                                   : code == 5 ? "internal error, live beams are not a prefix of the user's beams"
                                   : code == 7 ? "internal error, a barrier of the chain kernel never completed (watchdog)"
                                               : "candidate buffer overflow (trie fan-out larger than declared)");
}

int gram_get_stats(gram_handle* h, gram_stats* out) {
  if (!h || !out) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  int total = 0;
  CK(cudaMemcpy(&total, h->pm.total, 4, cudaMemcpyDeviceToHost));
  out->launches = h->launches;
  out->packed_tokens = h->encoded ? total : 0;
  out->kv_bytes = (int64_t)total * h->Ld * 2 * h->HD * (int64_t)h->esz;
  out->workspace_bytes = (int64_t)h->alloc_bytes;
  unsigned long long work[2] = {0ull, 0ull};
  CK(cudaMemcpy(work, h->bs.work, 16, cudaMemcpyDeviceToHost));
  out->decoded_rows = (int64_t)work[0];
  out->kv_tokens_read = (int64_t)work[1];
  return GRAM_OK;
}

int gram_profile_begin(gram_handle* h, uint32_t class_mask) {
  if (!h) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  h->prof_mask = class_mask;
  h->ev_used = 0;
  h->ev_log.clear();
  for (int i = 0; i < GRAM_K_COUNT; ++i) h->cls_launches[i] = 0;
  // pre-create a pool so event creation never lands inside a timed region
  while (h->ev_pool.size() < 8192) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->ev_pool.push_back(e); }
  return GRAM_OK;
}

int gram_profile_end(gram_handle* h, float* ms_per_class, int64_t* launches_per_class) {
  if (!h) return GRAM_ERR_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  float acc[GRAM_K_COUNT] = {0};
  // GRAM_PROF_DUMP=<file>: one line "class ms" per bracketed launch, in launch order (in-step per-launch times under the
  // step's real clocks; ncu's launch list runs every kernel alone at boost clocks with a cold L2)
  FILE* dump = nullptr;
  if (const char* path = getenv("GRAM_PROF_DUMP")) dump = fopen(path, "a");
  for (const auto& r : h->ev_log) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, r.a, r.b));
    acc[r.cls] += ms;
    if (dump) fprintf(dump, "%d %.4f\n", r.cls, ms);
  }
  if (dump) { fprintf(dump, "-1 0\n"); fclose(dump); }
  for (int i = 0; i < GRAM_K_COUNT; ++i) {
    if (ms_per_class) ms_per_class[i] = acc[i];
    if (launches_per_class) launches_per_class[i] = h->cls_launches[i];
  }
  h->prof_mask = 0;
  h->ev_used = 0;
  h->ev_log.clear();
  return GRAM_OK;
}

int gram_op_gemm(int32_t device, int32_t dtype, int32_t impl, int32_t epilogue, const void* A, const void* W, void* C,
                 int32_t M, int32_t N, int32_t K, void* stream) {
  if (cudaSetDevice(device) != cudaSuccess) return GRAM_ERR_CUDA;
  cudaError_t e;
  if (impl == 1 || impl == 2) {
    if (dtype != GRAM_DTYPE_BF16 || !gemm_tc_supported(N, K)) { g_create_error = "gram_op_gemm: tcgen05 path needs bf16 and a supported (N,K)"; return GRAM_ERR_UNSUPPORTED; }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    e = gemm_tc(epilogue, A, W, C, M, nullptr, N, K, sms, impl == 2 ? 1 : 2, (cudaStream_t)stream);
    if (e != cudaSuccess) { g_create_error = std::string("gemm_tc: ") + cudaGetErrorString(e) + " / " + gemm_tc_last_error(); return GRAM_ERR_CUDA; }
    return GRAM_OK;
  }
  e = gemm_simt(dtype, epilogue, A, W, C, M, nullptr, N, K, (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = std::string("gemm_simt: ") + cudaGetErrorString(e); return GRAM_ERR_CUDA; }
  return GRAM_OK;
}

int gram_op_gemm_norm(int32_t device, int32_t impl, int32_t epilogue, const void* A, const void* W, void* C, void* xb,
                      float* ss, const float* ln_w, const float* row_ss, float eps, int32_t M, int32_t N, int32_t K,
                      void* stream) {
  if (cudaSetDevice(device) != cudaSuccess) return GRAM_ERR_CUDA;
  if (!gemm_tc_supported(N, K)) { g_create_error = "gram_op_gemm_norm: unsupported (N,K)"; return GRAM_ERR_UNSUPPORTED; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  GemmNormAux aux;
  aux.row_ss = row_ss; aux.xb = xb; aux.ss_out = ss; aux.ln_w = ln_w; aux.eps = eps;
  cudaError_t e = gemm_tc(epilogue, A, W, C, M, nullptr, N, K, sms, impl == 2 ? 1 : 2, (cudaStream_t)stream, &aux);
  if (e != cudaSuccess) { g_create_error = std::string("gemm_tc: ") + cudaGetErrorString(e) + " / " + gemm_tc_last_error(); return GRAM_ERR_CUDA; }
  return GRAM_OK;
}

int gram_op_enc_chain(int32_t device, const void* ao, const void* w_o, float* x, void* xn, float* ss, const void* w_i,
                      const void* w_o2, void* scratch, int64_t scratch_bytes, const float* ln_mid, const float* ln_next, float eps,
                      int32_t M, int32_t D, int32_t HD, int32_t F, int32_t hints, int32_t* err, void* stream) {
  if (cudaSetDevice(device) != cudaSuccess) return GRAM_ERR_CUDA;
  if (!ao || !w_o || !x || !xn || !ss || !w_i || !w_o2 || !scratch || !ln_mid || !err || M <= 0) return GRAM_ERR_INVALID;
  if (!enc_chain_supported(D, HD, F)) { g_create_error = "gram_op_enc_chain: needs d_model % 256 == d_ff % 256 == inner_dim % 64 == 0"; return GRAM_ERR_UNSUPPORTED; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (scratch_bytes < (int64_t)enc_chain_scratch_bytes(F, sms)) { g_create_error = "gram_op_enc_chain: scratch smaller than num_sms * 128 * d_ff * 2 bytes"; return GRAM_ERR_INVALID; }
  cudaError_t e = enc_chain(ao, w_o, x, xn, ss, w_i, w_o2, scratch, ln_mid, ln_next, eps, M, nullptr, D, HD, F, sms, hints, err,
                            (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = std::string("enc_chain: ") + cudaGetErrorString(e) + " / " + gemm_tc_last_error(); return GRAM_ERR_CUDA; }
  return GRAM_OK;
}

int gram_op_lse_head(int32_t device, const void* hidden, const void* head, float* lse, void* partial, int32_t M, int32_t V,
                     int32_t D, void* stream) {
  if (cudaSetDevice(device) != cudaSuccess) return GRAM_ERR_CUDA;
  if (!hidden || !head || !lse || !partial || M <= 0) return GRAM_ERR_INVALID;
  if (!gemm_tc_supported(V, D)) { g_create_error = "gram_op_lse_head: unsupported (V, D)"; return GRAM_ERR_UNSUPPORTED; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaError_t e = gemm_tc(EPI_LSE, hidden, head, partial, M, nullptr, V, D, sms, 1, (cudaStream_t)stream);
  if (e == cudaSuccess) e = lse_combine(partial, lse, M, gemm_tc_lse_ntiles(M, V, sms), nullptr, (cudaStream_t)stream);
  if (e != cudaSuccess) { g_create_error = std::string("lse_head: ") + cudaGetErrorString(e) + " / " + gemm_tc_last_error(); return GRAM_ERR_CUDA; }
  return GRAM_OK;
}

int gram_op_cross_attention(int32_t device, int32_t dtype, int32_t impl, const void* q, const void* kv, int32_t kv_rows,
                            const int32_t* user_start, const uint8_t* tok_valid, void* out, int32_t users, int32_t K,
                            int32_t H, int32_t dk, void* stream) {
  if (cudaSetDevice(device) != cudaSuccess) return GRAM_ERR_CUDA;
  cudaError_t e;
  if (impl == 1 || impl == 2) {
    if (dtype != GRAM_DTYPE_BF16 || !cross_attention_mma_supported(K, H, dk)) {
      g_create_error = "gram_op_cross_attention: tensor-core path needs bf16, d_kv 64, K <= 64 and H % 4 == 0 (K <= 32) or H % 2 == 0";
      return GRAM_ERR_UNSUPPORTED;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    // impl 1 = persistent kernel (default in the engine), impl 2 = one CTA per (user, head group)
    e = cross_attention_mma(q, kv, (size_t)kv_rows, (size_t)2 * H * dk, 0, H * dk, user_start, nullptr, tok_valid, out, users, K, H,
                            nullptr, nullptr, impl == 2 ? 0 : sms, (cudaStream_t)stream);
  } else {
    e = cross_attention(dtype, q, kv, (size_t)2 * H * dk, 0, H * dk, user_start, tok_valid, out, users, K, H, dk,
                        nullptr, nullptr, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) { g_create_error = std::string("cross_attention: ") + cudaGetErrorString(e); return GRAM_ERR_CUDA; }
  return GRAM_OK;
}

}  // extern "C"
