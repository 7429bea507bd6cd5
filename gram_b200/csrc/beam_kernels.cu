// Trie-constrained beam search on the device: log-sum-exp, CSR trie gather, per-user top-2K,
// BeamSearchScorer.process / finalize.  No host round trip per step.
//
// Semantics restated from transformers==4.26.0 (the reference's pinned dependency, reference
// requirements.txt:1; reached through `super().generate(...)` at src/model/gram.py:93-99):
//   beam_search          generation/utils.py     log_softmax over the FULL vocabulary, then the
//                                                logits processor, then + beam_scores, view(B, K*V),
//                                                topk(2K, sorted)
//   PrefixConstrained... generation/logits_process.py   -inf mask outside Trie.get(prefix)
//                                                (reference src/utils/generation_trie.py:44-68,89-95);
//                                                the mask is applied AFTER the full-vocab
//                                                log-softmax, scores are not renormalised
//   BeamSearchScorer     generation/beam_search.py      process(): EOS candidates ranked < K become
//                                                hypotheses, the first K non-EOS candidates become the
//                                                next beams; BeamHypotheses.add / is_done; finalize()
// Because only trie children can survive the mask, the candidate set of a user is the union of the
// children of its K live trie nodes (<= K * max_fanout) instead of K*V masked scores.
//
// Documented divergence (SURVEY.md section 8(c) "tie/UB zones"): when a user has fewer than 2K finite
// candidates HF's topk returns -inf entries whose token ids depend on torch.topk tie order; here such
// fillers are dead beams (score -inf, token pad, never EOS).  They can only surface in the returned
// top-K when the trie holds fewer than K reachable items.
#include "common.cuh"
#include "kernels.h"

namespace gram {

// ------------------------------------------------------------------------------------------------
// lse[row] = log(sum(exp(logits[row, :])))  computed as max + log(sum(exp(x - max)))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lse_rows_kernel(const float* __restrict__ logits, float* __restrict__ lse, int V,
                                                       const int* __restrict__ n_rows) {
  __shared__ float red[8];
  __shared__ float bcast;
  if (n_rows && (int)blockIdx.x >= *n_rows) return;
  const float* x = logits + (size_t)blockIdx.x * V;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float mx = -INFINITY;
  for (int i = tid * 4; i < V; i += 1024) {
    const float4 v = load4(x + i);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  mx = warp_max(mx);
  if (lane == 0) red[wid] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int i = tid * 4; i < V; i += 1024) {
    const float4 v = load4(x + i);
    sum += expf(v.x - mx) + expf(v.y - mx) + expf(v.z - mx) + expf(v.w - mx);
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  if (tid == 0) {
    float sacc = 0.f;
    for (int i = 0; i < 8; ++i) sacc += red[i];
    lse[blockIdx.x] = mx + logf(sacc);
  }
}

cudaError_t lse_rows(const float* logits, float* lse, int R, int V, const int* n_rows, cudaStream_t s) {
  if (R <= 0) return cudaSuccess;
  if (V & 3) return cudaErrorInvalidValue;
  lse_rows_kernel<<<R, 256, 0, s>>>(logits, lse, V, n_rows);
  return cudaGetLastError();
}

__global__ void lse_combine_kernel(const float2* __restrict__ partial, float* __restrict__ lse, int R, int n_tiles,
                                   const int* __restrict__ n_rows) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= (n_rows ? *n_rows : R)) return;
  const float2* p = partial + (size_t)row * n_tiles;
  float mx = -INFINITY;
  for (int i = lane; i < n_tiles; i += 32) mx = fmaxf(mx, p[i].x);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < n_tiles; i += 32) {
    const float2 v = p[i];
    if (v.x > -INFINITY) sum += v.y * expf(v.x - mx);
  }
  sum = warp_sum(sum);
  if (lane == 0) lse[row] = mx + logf(sum);
}

cudaError_t lse_combine(const void* partial, float* lse, int R, int n_tiles, const int* n_rows, cudaStream_t s) {
  if (R <= 0) return cudaSuccess;
  lse_combine_kernel<<<(R + 7) / 8, 256, 0, s>>>((const float2*)partial, lse, R, n_tiles, n_rows);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
__global__ void beam_init_kernel(BeamState bs, int root, int users, int start_tok) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int R = users * bs.K;
  if (r < R) {
    bs.beam_score[0][r] = (r % bs.K == 0) ? 0.f : -1e9f;
    bs.node[0][r] = root;
    bs.seq[0][(size_t)r * bs.max_length] = start_tok;
    bs.tok[r] = start_tok;
  }
  if (r == 0) { bs.work[0] = 0ull; bs.work[1] = 0ull; }
  if (r < users) {
    bs.n_hyp[r] = 0;
    bs.worst[r] = 1e9;
    bs.next_seqno[r] = 0;
    bs.done[r] = 0;
    bs.live_cnt[r] = bs.K;
  }
}

cudaError_t beam_init(BeamState bs, TrieCSR trie, int users, int start_tok, cudaStream_t s) {
  const int R = users * bs.K;
  beam_init_kernel<<<(R + 255) / 256, 256, 0, s>>>(bs, trie.root, users, start_tok);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// hypothesis container (BeamHypotheses of transformers 4.26, early_stopping=False)
// ------------------------------------------------------------------------------------------------
struct HypView {
  double* score; int* len; int* seqno; int* tok;   // K+1 slots
  int* n; double* worst; int* next_seqno;
  int K, max_length;
};

__device__ inline HypView hyp_view(const BeamState& bs, int u) {
  HypView hv;
  const int S = bs.K + 1;
  hv.score = bs.hyp_score + (size_t)u * S;
  hv.len = bs.hyp_len + (size_t)u * S;
  hv.seqno = bs.hyp_seqno + (size_t)u * S;
  hv.tok = bs.hyp_tok + (size_t)u * S * bs.max_length;
  hv.n = bs.n_hyp + u;
  hv.worst = bs.worst + u;
  hv.next_seqno = bs.next_seqno + u;
  hv.K = bs.K;
  hv.max_length = bs.max_length;
  return hv;
}

// BeamHypotheses.add: score = sum_logprobs / len ** length_penalty (len counts the start token, not EOS)
__device__ void hyp_add(HypView hv, const int* tokens, int len, float sum_logprobs, const double* len_pow) {
  const double score = (double)sum_logprobs / len_pow[len];
  int n = *hv.n;
  if (n < hv.K || score > *hv.worst) {
    hv.score[n] = score;
    hv.len[n] = len;
    hv.seqno[n] = (*hv.next_seqno)++;
    for (int i = 0; i < len; ++i) hv.tok[(size_t)n * hv.max_length + i] = tokens[i];
    ++n;
    if (n > hv.K) {
      // python: sorted([(s, idx)])[0] -> lowest score, earliest inserted among equals
      int imin = 0;
      for (int i = 1; i < n; ++i)
        if (hv.score[i] < hv.score[imin] || (hv.score[i] == hv.score[imin] && hv.seqno[i] < hv.seqno[imin])) imin = i;
      const int last = n - 1;
      if (imin != last) {
        hv.score[imin] = hv.score[last];
        hv.len[imin] = hv.len[last];
        hv.seqno[imin] = hv.seqno[last];
        for (int i = 0; i < hv.len[last]; ++i)
          hv.tok[(size_t)imin * hv.max_length + i] = hv.tok[(size_t)last * hv.max_length + i];
      }
      --n;
      double w = hv.score[0];
      for (int i = 1; i < n; ++i) w = fmin(w, hv.score[i]);
      *hv.worst = w;
    } else {
      *hv.worst = fmin(score, *hv.worst);
    }
    *hv.n = n;
  }
}

// ------------------------------------------------------------------------------------------------
// one step: candidates -> top-2K -> scorer.  One CTA per user.
// ------------------------------------------------------------------------------------------------
constexpr int BEAM_THREADS = 256;
constexpr int BEAM_KMAX = 64;

size_t beam_step_smem(int cand_cap) { return (size_t)cand_cap * sizeof(unsigned long long); }

__global__ void __launch_bounds__(BEAM_THREADS)
beam_step_kernel(BeamState bs, TrieCSR trie, const float* __restrict__ logits, const bf16* __restrict__ hidden,
                 const bf16* __restrict__ head, int D, const float* __restrict__ lse, int users, int t, int cand_cap,
                 int compact, const int* __restrict__ row_slot) {
  extern __shared__ __align__(16) unsigned long long keys[];
  __shared__ int pre[BEAM_KMAX + 1];
  __shared__ int sel_parent[BEAM_KMAX], sel_tok[BEAM_KMAX], sel_node[BEAM_KMAX];
  __shared__ float sel_score[BEAM_KMAX];
  __shared__ int n_sort_s;
  __shared__ int cand_off[BEAM_KMAX], cand_cnt[BEAM_KMAX];          // first CSR edge / fan-out of each live beam
  __shared__ float r_score[2 * BEAM_KMAX];                           // the 2K best candidates, decoded in parallel
  __shared__ int r_beam[2 * BEAM_KMAX], r_tok[2 * BEAM_KMAX], r_child[2 * BEAM_KMAX];

  const int u = blockIdx.x, tid = threadIdx.x;
  const int K = bs.K, V = bs.V, ML = bs.max_length;
  const int R = users * K, base = u * K;
  const int cur = t & 1, nxt = cur ^ 1;
  const int cur_len = t + 1;
  const float* score_c = bs.beam_score[cur];
  const int* node_c = bs.node[cur];
  const int* seq_c = bs.seq[cur];
  const int* anc_c = bs.anc[cur];
  // decoder row (of logits / hidden / lse) that served beam row `row`: the user's single row at the compact step 0,
  // the beam's compact slot under live-row compaction (dead beams have none and offer no candidates), else the row
  auto dec_row = [&](int row) { return compact ? u : (row_slot ? row_slot[row] : row); };

  if (bs.tap_lse) {
    for (int b = tid; b < K; b += BEAM_THREADS) {
      const int lr = dec_row(base + b);
      bs.tap_lse[(size_t)t * R + base + b] = lr >= 0 ? lse[lr] : 0.f;
      bs.tap_score[(size_t)t * R + base + b] = score_c[base + b];
    }
    for (int i = tid; i < K * ML; i += BEAM_THREADS)
      bs.tap_seq[((size_t)t * R + base) * ML + i] = (i % ML) < cur_len ? seq_c[(size_t)base * ML + i] : 0;
  }

  const bool frozen = bs.done[u] != 0;
  if (frozen) {
    // hypotheses of this user are final (scorer pads done batches); keep rows well-formed but dead
    for (int b = tid; b < K; b += BEAM_THREADS) {
      bs.beam_score[nxt][base + b] = 0.f;
      bs.node[nxt][base + b] = -1;
      bs.tok[base + b] = bs.pad;
    }
    if (tid == 0) bs.live_cnt[u] = 0;
    for (int i = tid; i < K * ML; i += BEAM_THREADS) {
      const int b = i / ML, j = i % ML;
      bs.seq[nxt][(size_t)base * ML + i] = (j == cur_len) ? bs.pad : seq_c[(size_t)base * ML + i];
      bs.anc[nxt][(size_t)base * ML + i] = (j == t) ? (compact ? (u - base) : b) : anc_c[(size_t)base * ML + i];
    }
    return;
  }

  // ---- enumerate candidates: children of every live beam's trie node ----
  // (the per-beam fan-outs are fetched in parallel: a single thread chasing node -> offsets serially costs
  //  ~1 us of dependent global-load latency per beam)
  if (tid < K) {
    const int nd = node_c[base + tid];
    const float sc = score_c[base + tid];
    int cnt = 0;
    if (nd >= 0 && sc > -INFINITY) {
      const int o0 = trie.child_offsets[nd];
      cnt = trie.child_offsets[nd + 1] - o0;
      cand_off[tid] = o0;
    } else {
      cand_off[tid] = 0;
    }
    cand_cnt[tid] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    // step 0: beams 1..K-1 start at -1e9 on the same (root) node as beam 0.  If beam 0 alone offers >= 2K
    // candidates, theirs (all ~ -1e9) cannot enter the top-2K, so they are not enumerated: same result as HF.
    const bool first_only = (t == 0) && cand_cnt[0] >= 2 * K && score_c[base] == 0.f;
    int run = 0;
    for (int b = 0; b < K; ++b) {
      pre[b] = run;
      if (first_only && b > 0 && score_c[base + b] <= -1e8f) cand_cnt[b] = 0;
      run += cand_cnt[b];
    }
    pre[K] = run;
    int n = 2;
    while (n < run) n <<= 1;
    n_sort_s = n;
    if (run > cand_cap) { atomicExch(bs.err, 1); }
  }
  __syncthreads();
  const int C = min(pre[K], cand_cap);
  const int n_sort = min(n_sort_s, cand_cap);
  if (logits != nullptr) {
    for (int c = tid; c < n_sort; c += BEAM_THREADS) {
      unsigned long long key = 0ull;
      if (c < C) {
        int b = 0;
        while (pre[b + 1] <= c) ++b;                         // K <= 64, linear search
        const int e = cand_off[b] + (c - pre[b]);
        const int tok = trie.child_tokens[e];
        const int row = base + b, lrow = dec_row(row);
        float s = (logits[(size_t)lrow * V + tok] - lse[lrow]) + score_c[row];
        if (!(s == s)) s = -INFINITY;
        // children are stored token-ascending, so the enumeration index c orders candidates exactly like the flat
        // index beam * V + token: it is both the tie-break (smaller first) and the handle back to (beam, edge)
        key = ((unsigned long long)float_key(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)c);
      }
      keys[c] = key;
    }
  } else {
    // fused mode: a warp recomputes the logits of 4 candidates per iteration, logit = hidden[row] . head[token]
    // (bf16 operands, fp32 sum); the 4 x 2 row loads are in flight together, the chain token -> row -> reduce of a
    // one-candidate-per-iteration loop is latency-bound
    const int wid = tid >> 5, lane = tid & 31;
    constexpr int UN = 4;
    for (int c = C + tid; c < n_sort; c += BEAM_THREADS) keys[c] = 0ull;
    for (int c0 = wid * UN; c0 < C; c0 += (BEAM_THREADS / 32) * UN) {
      int bb[UN], tk[UN];
      const bf16* hr[UN];
      const bf16* er[UN];
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        const int c = min(c0 + j, C - 1);
        int b = 0;
        while (pre[b + 1] <= c) ++b;
        bb[j] = b;
        tk[j] = trie.child_tokens[cand_off[b] + (c - pre[b])];
      }
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        hr[j] = hidden + (size_t)dec_row(base + bb[j]) * D;
        er[j] = head + (size_t)tk[j] * D;
      }
      float a[UN];
#pragma unroll
      for (int j = 0; j < UN; ++j) a[j] = 0.f;
      for (int d = lane * 8; d < D; d += 256) {
        uint4 hv[UN], ev[UN];
#pragma unroll
        for (int j = 0; j < UN; ++j) {
          hv[j] = *reinterpret_cast<const uint4*>(hr[j] + d);
          ev[j] = *reinterpret_cast<const uint4*>(er[j] + d);
        }
#pragma unroll
        for (int j = 0; j < UN; ++j) {
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hv[j]);
          const __nv_bfloat162* e2 = reinterpret_cast<const __nv_bfloat162*>(&ev[j]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 x = __bfloat1622float2(h2[i]), y = __bfloat1622float2(e2[i]);
            a[j] = fmaf(x.x, y.x, a[j]);
            a[j] = fmaf(x.y, y.y, a[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < UN; ++j) a[j] = warp_sum(a[j]);
      if (lane < UN && c0 + lane < C) {
        float av = a[0];
#pragma unroll
        for (int j = 1; j < UN; ++j) av = (lane == j) ? a[j] : av;
        int b = bb[0];
#pragma unroll
        for (int j = 1; j < UN; ++j) b = (lane == j) ? bb[j] : b;
        const int row = base + b, lrow = dec_row(row);
        float sc = (av - lse[lrow]) + score_c[row];
        if (!(sc == sc)) sc = -INFINITY;
        keys[c0 + lane] = ((unsigned long long)float_key(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)(c0 + lane));
      }
    }
  }
  __syncthreads();
  // ---- bitonic sort ascending (best candidate ends at n_sort-1) ----
  for (int k = 2; k <= n_sort; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n_sort; i += BEAM_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  // ---- decode the 2K best candidates in parallel (key -> score, beam, token, child node) ----
  if (tid < 2 * K) {
    float s = -INFINITY;
    int b = 0, tok = bs.pad, child = -1;
    if (tid < C) {
      const unsigned long long key = keys[n_sort - 1 - tid];
      s = key_float((unsigned int)(key >> 32));
      const int c = (int)(0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull));
      while (pre[b + 1] <= c) ++b;
      const int e = cand_off[b] + (c - pre[b]);
      tok = trie.child_tokens[e];
      child = trie.child_nodes[e];
      if (s == -INFINITY) { tok = bs.pad; child = -1; }     // indistinguishable from a filler
    }
    r_score[tid] = s; r_beam[tid] = b; r_tok[tid] = tok; r_child[tid] = child;
  }
  __syncthreads();
  // ---- BeamSearchScorer.process for this user (inherently sequential, shared memory only) ----
  if (tid == 0) {
    HypView hv = hyp_view(bs, u);
    int slot = 0;
    const float best = r_score[0];
    for (int rank = 0; rank < 2 * K && slot < K; ++rank) {
      const float s = r_score[rank];
      const int b = r_beam[rank], tok = r_tok[rank], child = r_child[rank];
      if (tok == bs.eos && s > -INFINITY) {
        if (rank >= K) continue;
        hyp_add(hv, seq_c + (size_t)(base + b) * ML, cur_len, s, bs.len_pow);
      } else {
        sel_parent[slot] = b; sel_tok[slot] = tok; sel_node[slot] = child; sel_score[slot] = s;
        ++slot;
      }
    }
    for (; slot < K; ++slot) { sel_parent[slot] = 0; sel_tok[slot] = bs.pad; sel_node[slot] = -1; sel_score[slot] = -INFINITY; }
    // BeamHypotheses.is_done(best_sum_logprobs = max of the 2K candidate scores, cur_len)
    bool done = false;
    if (*hv.n >= K) {
      const double cur_score = (double)best / bs.len_pow[cur_len];
      if (*hv.worst >= cur_score) { bs.done[u] = 1; done = true; }
    }
    // beams the next step has to decode: candidates were taken best-first, so the live ones (finite score, a trie
    // node to continue from) are the first `live` slots
    int live = 0;
    while (live < K && sel_node[live] >= 0 && sel_score[live] > -INFINITY) ++live;
    bs.live_cnt[u] = done ? 0 : live;
  }
  __syncthreads();
  for (int b = tid; b < K; b += BEAM_THREADS) {
    bs.beam_score[nxt][base + b] = sel_score[b];
    bs.node[nxt][base + b] = sel_node[b];
    bs.tok[base + b] = sel_tok[b];
  }
  for (int i = tid; i < K * ML; i += BEAM_THREADS) {
    const int b = i / ML, j = i % ML;
    const int par = sel_parent[b];
    int tokv = 0;
    if (j < cur_len) tokv = seq_c[(size_t)(base + par) * ML + j];
    else if (j == cur_len) tokv = sel_tok[b];
    bs.seq[nxt][(size_t)base * ML + i] = tokv;
    int a = 0;
    if (j < t) a = anc_c[(size_t)(base + par) * ML + j];
    else if (j == t) a = compact ? (u - base) : par;   // compact step 0 wrote ONE cache row per user, at row u
    bs.anc[nxt][(size_t)base * ML + i] = a;
  }
}

cudaError_t beam_step(BeamState bs, TrieCSR trie, const float* logits, const void* hidden, const void* head, int D,
                      const float* lse, int users, int t, int cand_cap, int compact, const int* row_slot, cudaStream_t s) {
  if (users <= 0) return cudaSuccess;
  if (bs.K > BEAM_KMAX) return cudaErrorInvalidValue;
  if (logits == nullptr && (hidden == nullptr || head == nullptr || (D & 7))) return cudaErrorInvalidValue;
  const size_t smem = beam_step_smem(cand_cap);
  static SmemAttr attr;
  {
    cudaError_t e = attr.ensure(beam_step_kernel, smem);
    if (e != cudaSuccess) return e;
  }
  beam_step_kernel<<<users, BEAM_THREADS, smem, s>>>(bs, trie, logits, (const bf16*)hidden, (const bf16*)head, D, lse, users, t,
                                                     cand_cap, compact, row_slot);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// BeamSearchScorer.finalize: one thread per user (tiny, serial by nature)
// ------------------------------------------------------------------------------------------------
__global__ void beam_finalize_kernel(BeamState bs, int users, int t_final, int R_ret, long long* __restrict__ out_seq,
                                     float* __restrict__ out_scores, int* __restrict__ out_width) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= users) return;
  const int K = bs.K, ML = bs.max_length;
  const int cur = t_final & 1;              // buffers holding the beams after the last step
  const int cur_len = t_final + 1;
  HypView hv = hyp_view(bs, u);
  if (!bs.done[u]) {
    for (int j = 0; j < K; ++j) {
      const int row = u * K + j;
      hyp_add(hv, bs.seq[cur] + (size_t)row * ML, cur_len, bs.beam_score[cur][row], bs.len_pow);
    }
  }
  // sorted(beams, key=score) ascending, stable; pop() from the end -> (score desc, insertion desc)
  int n = *hv.n;
  int width = 1;
  for (int rnk = 0; rnk < R_ret; ++rnk) {
    long long* o = out_seq + ((size_t)u * R_ret + rnk) * ML;
    for (int i = 0; i < ML; ++i) o[i] = bs.pad;
    if (n == 0) { out_scores[(size_t)u * R_ret + rnk] = -INFINITY; continue; }
    int ib = 0;
    for (int i = 1; i < n; ++i)
      if (hv.score[i] > hv.score[ib] || (hv.score[i] == hv.score[ib] && hv.seqno[i] > hv.seqno[ib])) ib = i;
    const int len = hv.len[ib];
    for (int i = 0; i < len; ++i) o[i] = hv.tok[(size_t)ib * ML + i];
    if (len < bs.gen_len) o[len] = bs.eos;
    out_scores[(size_t)u * R_ret + rnk] = (float)hv.score[ib];
    width = max(width, min(len + 1, bs.gen_len));
    // remove ib (order of the remainder is irrelevant: seqno carries insertion order)
    const int last = n - 1;
    if (ib != last) {
      hv.score[ib] = hv.score[last]; hv.len[ib] = hv.len[last]; hv.seqno[ib] = hv.seqno[last];
      for (int i = 0; i < hv.len[last]; ++i) hv.tok[(size_t)ib * ML + i] = hv.tok[(size_t)last * ML + i];
    }
    --n;
  }
  atomicMax(out_width, width);
}

cudaError_t beam_finalize(BeamState bs, int users, int t_final, int R_ret, int64_t* out_seq, float* out_scores,
                          int* out_width, cudaStream_t s) {
  if (users <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(out_width, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  beam_finalize_kernel<<<(users + 63) / 64, 64, 0, s>>>(bs, users, t_final, R_ret, (long long*)out_seq, out_scores, out_width);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// live-row compaction: slot ranges per user (one CTA, chunked block scan), then the row <-> slot maps
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) live_scan_kernel(const int* __restrict__ live_cnt, int users, int* __restrict__ start,
                                                         int* __restrict__ n_live, const int* __restrict__ ustart,
                                                         unsigned long long* __restrict__ work) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  __shared__ unsigned long long tok_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { carry_s = 0; tok_s = 0ull; }
  __syncthreads();
  unsigned long long toks = 0ull;                          // memory tokens of the users this thread found alive
  for (int u0 = 0; u0 < users; u0 += 1024) {
    const int u = u0 + tid;
    const int c = u < users ? live_cnt[u] : 0;
    if (c > 0) toks += (unsigned long long)(ustart[u + 1] - ustart[u]);
    int x = c;                                              // inclusive scan within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;                                   // inclusive totals of warps 0..lane
    }
    __syncthreads();
    const int carry = carry_s;
    if (u < users) start[u] = carry + (wid ? warp_tot[wid - 1] : 0) + x - c;
    __syncthreads();
    if (tid == 0) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (toks) atomicAdd(&tok_s, toks);
  __syncthreads();
  if (tid == 0) {
    start[users] = carry_s;
    *n_live = carry_s;
    work[0] += (unsigned long long)carry_s;
    work[1] += tok_s;
  }
}

__global__ void work_add_kernel(unsigned long long* __restrict__ work, long long rows, const int* __restrict__ tokens_ptr) {
  work[0] += (unsigned long long)rows;
  work[1] += (unsigned long long)*tokens_ptr;
}

cudaError_t work_add(BeamState bs, long long rows, const int* tokens_ptr, cudaStream_t s) {
  work_add_kernel<<<1, 1, 0, s>>>(bs.work, rows, tokens_ptr);
  return cudaGetLastError();
}

__global__ void live_fill_kernel(BeamState bs, int users, int cur, LiveMap lm) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = bs.K;
  if (r >= users * K) return;
  const int u = r / K, b = r - u * K;
  const int c = bs.live_cnt[u];
  // cross-check of the "live beams are the first live_cnt beams" contract against the beam state itself
  const bool alive = bs.done[u] == 0 && bs.node[cur][r] >= 0 && bs.beam_score[cur][r] > -INFINITY;
  if (alive != (b < c)) atomicExch(bs.err, 5);
  int slot = -1;
  if (b < c) {
    slot = lm.start[u] + b;
    lm.slot_row[slot] = r;
    lm.tok[slot] = bs.tok[r];
  }
  lm.row_slot[r] = slot;
}

cudaError_t live_compact(BeamState bs, int users, int cur, LiveMap lm, const int* ustart, cudaStream_t s) {
  if (users <= 0) return cudaSuccess;
  live_scan_kernel<<<1, 1024, 0, s>>>(bs.live_cnt, users, lm.start, lm.n_live, ustart, bs.work);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int R = users * bs.K;
  live_fill_kernel<<<(R + 255) / 256, 256, 0, s>>>(bs, users, cur, lm);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
__global__ void forced_step_kernel(BeamState bs, const long long* __restrict__ dec_ids, int q, int t, int R) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) bs.tok[r] = (int)dec_ids[(size_t)r * q + t];
}

cudaError_t forced_step(BeamState bs, const int64_t* dec_ids, int q, int t, int R, cudaStream_t s) {
  forced_step_kernel<<<(R + 255) / 256, 256, 0, s>>>(bs, (const long long*)dec_ids, q, t, R);
  return cudaGetLastError();
}

}  // namespace gram
