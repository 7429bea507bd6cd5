// Encoder-side kernels: valid-token packing, embedding gather, T5 RMS norm (+ passage-position
// embedding), bidirectional per-passage self-attention, and the debug un-packer.
//
// Reference semantics (paths relative to the reference root):
//   packing          src/model/gram.py:206-216 views B x (N*L) as (B*N) x L and encodes every passage
//                    independently.  Masked keys receive finfo.min and contribute exactly 0 after the
//                    fp32 softmax (src/model/gram_t5_modeling.py:1130,596-610), and masked memory
//                    positions are invisible to the decoder (:1145-1147), so rows past the last valid
//                    token of a passage -- and all-masked passages -- are skipped here, exactly.
//   embedding        src/model/gram_t5_modeling.py:1091
//   RMS norm         src/model/gram_t5_modeling.py:262-276 (fp32 variance, no mean subtraction)
//   position add     src/model/gram.py:238-249
//   attention        src/model/gram_t5_modeling.py:572-621 (no 1/sqrt(d) scale, additive relative bias
//                    from layer 0 shared by all layers :1249, fp32 softmax)
#include "common.cuh"
#include "kernels.h"

namespace gram {

// ------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------
__global__ void passage_len_kernel(const uint8_t* __restrict__ mask, int P, int L, int* __restrict__ plen) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P) return;
  const uint8_t* m = mask + (size_t)warp * L;
  int last = 0;
  for (int l = lane; l < L; l += 32)
    if (m[l]) last = l + 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (lane == 0) plen[warp] = last;
}

// single-CTA exclusive scan (P is at most a few hundred thousand)
// If the batch holds more valid tokens than the workspace (`cap` rows) the whole batch is emptied (every passage
// length 0) and the sticky error flag is set: capacity is checked where the count is known, on the device.
__global__ void __launch_bounds__(1024) passage_scan_kernel(int* __restrict__ plen, int P, int N, int B,
                                                            int* __restrict__ poff, int* __restrict__ ustart,
                                                            int* __restrict__ total, long long cap, int* __restrict__ err) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (P + 1023) / 1024;
  const int beg = min(tid * per, P), end = min(beg + per, P);
  int sum = 0;
  for (int i = beg; i < end; ++i) sum += plen[i];
  // block exclusive scan of `sum`
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = warp_sums[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    warp_sums[lane] = wi - w;   // exclusive
    if (lane == 31) carry_s = wi;
  }
  __syncthreads();
  const bool over = (long long)carry_s > cap;
  int run = warp_sums[wid] + incl - sum;
  for (int i = beg; i < end; ++i) {
    poff[i] = over ? 0 : run;
    if (i % N == 0) ustart[i / N] = over ? 0 : run;
    run += plen[i];
    if (over) plen[i] = 0;
  }
  if (tid == 0) {
    const int tot = over ? 0 : carry_s;
    poff[P] = tot;
    ustart[B] = tot;
    *total = tot;
    if (over) atomicExch(err, 4);
  }
}

__global__ void passage_fill_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ mask, int N, int L,
                                    PackMeta pm) {
  const int p = blockIdx.x;
  const int len = pm.plen[p], off = pm.poff[p];
  for (int l = threadIdx.x; l < len; l += blockDim.x) {
    const size_t src = (size_t)p * L + l;
    long long id = ids[src];
    if (id < 0 || id >= pm.vocab) {              // the reference's nn.Embedding would raise IndexError
      atomicExch(pm.err, 2);
      id = 0;
    }
    pm.tok_id[off + l] = (int)id;
    pm.tok_valid[off + l] = mask[src] ? 1 : 0;
    pm.tok_pos[off + l] = p % N;
    pm.row_src[off + l] = (int)src;
  }
}

// uorder[rank] = user, ranked by packed length descending (ties by index): O(B^2 / 1024) per thread, B <= a few thousand
__global__ void __launch_bounds__(1024) user_order_kernel(const int* __restrict__ ustart, int B, int* __restrict__ uorder) {
  for (int u = threadIdx.x; u < B; u += blockDim.x) {
    const int len = ustart[u + 1] - ustart[u];
    int rank = 0;
    for (int v = 0; v < B; ++v) {
      const int lv = ustart[v + 1] - ustart[v];
      rank += (lv > len) || (lv == len && v < u);
    }
    uorder[rank] = u;
  }
}

cudaError_t enc_pack(const int64_t* ids, const uint8_t* mask, int B, int N, int L, PackMeta pm, cudaStream_t s) {
  const int P = B * N;
  passage_len_kernel<<<(P * 32 + 255) / 256, 256, 0, s>>>(mask, P, L, pm.plen);
  passage_scan_kernel<<<1, 1024, 0, s>>>(pm.plen, P, N, B, pm.poff, pm.ustart, pm.total, pm.cap, pm.err);
  passage_fill_kernel<<<P, 128, 0, s>>>(ids, mask, N, L, pm);
  user_order_kernel<<<1, 1024, 0, s>>>(pm.ustart, B, pm.uorder);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// per-item encoder-state cache.  Every passage except the user prompt is a function
// of the ITEM only (reference src/utils/indexing.py:209-211,315-320), passages are encoded independently
// (src/model/gram.py:206-216) and the passage-position row is added AFTER the encoder (src/model/gram.py:238-249):
// the final-normed encoder rows of an item passage are therefore identical for every user and are kept, as fp32
// before the position add, in a table [n_items][L][D].  A cached encode runs the encoder stack on the user prompts
// only and assembles each user's memory from prompt rows + cached item rows + position rows -- bit-identical to
// encoding all passages (same kernels per passage, same separately-rounded position add).
// ------------------------------------------------------------------------------------------------
// packed fp32 rows of a chunk of items ([B, N, L] layout, item = item0 + passage) -> table rows item*L + l
__global__ void cache_scatter_kernel(const float* __restrict__ src, const int* __restrict__ row_src,
                                     const uint8_t* __restrict__ tok_valid, const int* __restrict__ m_ptr, int D,
                                     long long row0, float* __restrict__ item_mem, uint8_t* __restrict__ item_valid) {
  const int M = *m_ptr;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const long long dst = row0 + row_src[row];
  const float* s = src + (size_t)row * D;
  float* d = item_mem + (size_t)dst * D;
  for (int c = lane * 4; c < D; c += 128) store4(d + c, load4(s + c));
  if (lane == 0) item_valid[dst] = tok_valid[row];
}

cudaError_t cache_scatter(const float* src, const PackMeta& pm, int M_max, int D, long long row0, float* item_mem,
                          uint8_t* item_valid, cudaStream_t s) {
  if (M_max <= 0) return cudaSuccess;
  cache_scatter_kernel<<<(M_max + 7) / 8, 256, 0, s>>>(src, pm.row_src, pm.tok_valid, pm.total, D, row0, item_mem, item_valid);
  return cudaGetLastError();
}

// passage lengths of the full user layout [B][1 + NI]: passage 0 = the prompt, passage 1 + j = item items[b][j]
__global__ void cached_plen_kernel(const int* __restrict__ prompt_len, const int* __restrict__ items,
                                   const int* __restrict__ item_len, int n_items, int B, int NI, int* __restrict__ plen,
                                   int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (NI + 1)) return;
  const int b = i / (NI + 1), n = i % (NI + 1);
  int len = 0;
  if (n == 0) {
    len = prompt_len[b];
  } else {
    const int it = items[b * NI + n - 1];
    if (it >= n_items) atomicExch(err, 3);
    else if (it >= 0) len = item_len[it];
  }
  plen[i] = len;
}

template <typename T>
__global__ void __launch_bounds__(128)
cached_assemble_kernel(PackMeta pm, const int* __restrict__ prompt_off, const uint8_t* __restrict__ prompt_valid,
                       const float* __restrict__ prompt_rows, const int* __restrict__ items,
                       const float* __restrict__ item_mem, const uint8_t* __restrict__ item_valid,
                       const float* __restrict__ pos_table, int NI, int L, int D, T* __restrict__ mem) {
  const int p = blockIdx.x;
  const int len = pm.plen[p];
  if (len == 0) return;
  const int off = pm.poff[p];
  const int b = p / (NI + 1), n = p % (NI + 1);
  const float* src;
  const uint8_t* vsrc;
  if (n == 0) {
    src = prompt_rows + (size_t)prompt_off[b] * D;
    vsrc = prompt_valid + prompt_off[b];
  } else {
    const size_t r0 = (size_t)items[b * NI + n - 1] * L;
    src = item_mem + r0 * D;
    vsrc = item_valid + r0;
  }
  const float* pe = pos_table ? pos_table + (size_t)n * D : nullptr;
  const int q = D / 4;
  for (int i = threadIdx.x; i < len * q; i += blockDim.x) {
    const int l = i / q, c = (i % q) * 4;
    float4 v = load4(src + (size_t)l * D + c);
    if (pe) {
      const float4 e = load4(pe + c);
      v.x = __fadd_rn(v.x, e.x); v.y = __fadd_rn(v.y, e.y); v.z = __fadd_rn(v.z, e.z); v.w = __fadd_rn(v.w, e.w);
    }
    store4(mem + (size_t)(off + l) * D + c, v);
  }
  for (int l = threadIdx.x; l < len; l += blockDim.x) {
    pm.tok_valid[off + l] = vsrc[l];
    pm.tok_pos[off + l] = n;
    pm.row_src[off + l] = p * L + l;
    pm.tok_id[off + l] = 0;
  }
}

cudaError_t cached_pack_assemble(int dtype, const PackMeta& pm, const PackMeta& prompt_pm, const float* prompt_rows,
                                 const int* items, const float* item_mem, const uint8_t* item_valid, const int* item_len,
                                 int n_items, const float* pos_table, int B, int NI, int L, int D, void* mem,
                                 cudaStream_t s) {
  const int P = B * (NI + 1);
  cached_plen_kernel<<<(P + 255) / 256, 256, 0, s>>>(prompt_pm.plen, items, item_len, n_items, B, NI, pm.plen, pm.err);
  passage_scan_kernel<<<1, 1024, 0, s>>>(pm.plen, P, NI + 1, B, pm.poff, pm.ustart, pm.total, pm.cap, pm.err);
  if (dtype == 0)
    cached_assemble_kernel<float><<<P, 128, 0, s>>>(pm, prompt_pm.poff, prompt_pm.tok_valid, prompt_rows, items, item_mem,
                                                    item_valid, pos_table, NI, L, D, (float*)mem);
  else
    cached_assemble_kernel<bf16><<<P, 128, 0, s>>>(pm, prompt_pm.poff, prompt_pm.tok_valid, prompt_rows, items, item_mem,
                                                   item_valid, pos_table, NI, L, D, (bf16*)mem);
  user_order_kernel<<<1, 1024, 0, s>>>(pm.ustart, B, pm.uorder);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// embedding gather: x[row] = table[tok_id[row]]   (fp32 residual stream)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void embed_rows_kernel(const T* __restrict__ table, const int* __restrict__ tok_id, float* __restrict__ x,
                                  int M_imm, const int* __restrict__ m_ptr, int D) {
  const int M = m_ptr ? *m_ptr : M_imm;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const T* src = table + (size_t)tok_id[row] * D;
  float* dst = x + (size_t)row * D;
  for (int c = lane * 4; c < D; c += 128) store4(dst + c, load4(src + c));
}

// embedding + the first RMSNorm's input in the folded form the tcgen05 GEMMs consume (kernels.h: GemmNormAux):
// x fp32, xb = bf16(x * w), and the row's sum of squares per 128-column block.  One warp per row; D % 128 == 0.
__global__ void embed_rows_norm_kernel(const bf16* __restrict__ table, const int* __restrict__ tok_id, float* __restrict__ x,
                                       bf16* __restrict__ xb, float* __restrict__ ss, const float* __restrict__ w,
                                       int M_imm, const int* __restrict__ m_ptr, int D) {
  const int M = m_ptr ? *m_ptr : M_imm;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const bf16* src = table + (size_t)tok_id[row] * D;
  for (int c0 = 0; c0 < D; c0 += 128) {
    const int c = c0 + lane * 4;
    const float4 v = load4(src + c);
    const float4 g = load4(w + c);
    store4(x + (size_t)row * D + c, v);
    store4(xb + (size_t)row * D + c, make_float4(v.x * g.x, v.y * g.y, v.z * g.z, v.w * g.w));
    float t = v.x * v.x;
    t = fmaf(v.y, v.y, t); t = fmaf(v.z, v.z, t); t = fmaf(v.w, v.w, t);
    t = warp_sum(t);
    if (lane == 0) ss[(size_t)row * (D >> 7) + (c0 >> 7)] = t;
  }
}

cudaError_t embed_rows_norm(const void* table, const int* tok_id, float* x, void* xb, float* ss, const float* w, int M_max,
                            const int* m_ptr, int D, cudaStream_t s) {
  if (M_max <= 0) return cudaSuccess;
  if (D & 127) return cudaErrorInvalidValue;
  embed_rows_norm_kernel<<<(M_max + 7) / 8, 256, 0, s>>>((const bf16*)table, tok_id, x, (bf16*)xb, ss, w, M_max, m_ptr, D);
  return cudaGetLastError();
}

// bf16 residual stream (bf16 mode unless GRAM_FLAG_FP32_RESID): the stream starts as the embedding row itself (exact: the table is bf16)
// with its sums of squares per 128-column block; the first RMSNorm's gain lives in the consumer GEMM's weights
__global__ void embed_rows_stream_kernel(const bf16* __restrict__ table, const int* __restrict__ tok_id, bf16* __restrict__ xr,
                                         float* __restrict__ ss, int M_imm, const int* __restrict__ m_ptr, int D) {
  const int M = m_ptr ? *m_ptr : M_imm;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const bf16* src = table + (size_t)tok_id[row] * D;
  for (int c0 = 0; c0 < D; c0 += 128) {
    const int c = c0 + lane * 4;
    const uint2 raw = *reinterpret_cast<const uint2*>(src + c);
    *reinterpret_cast<uint2*>(xr + (size_t)row * D + c) = raw;
    const float4 v = load4(src + c);
    float t = v.x * v.x;
    t = fmaf(v.y, v.y, t); t = fmaf(v.z, v.z, t); t = fmaf(v.w, v.w, t);
    t = warp_sum(t);
    if (lane == 0) ss[(size_t)row * (D >> 7) + (c0 >> 7)] = t;
  }
}

cudaError_t embed_rows_stream(const void* table, const int* tok_id, void* xr, float* ss, int M_max, const int* m_ptr, int D,
                              cudaStream_t s) {
  if (M_max <= 0) return cudaSuccess;
  if (D & 127) return cudaErrorInvalidValue;
  embed_rows_stream_kernel<<<(M_max + 7) / 8, 256, 0, s>>>((const bf16*)table, tok_id, (bf16*)xr, ss, M_max, m_ptr, D);
  return cudaGetLastError();
}

cudaError_t embed_rows(int dtype, const void* table, const int* tok_id, float* x, int M_max, const int* m_ptr, int D,
                       cudaStream_t s) {
  if (M_max <= 0) return cudaSuccess;
  const int grid = (M_max + 7) / 8;
  if (dtype == 0) embed_rows_kernel<float><<<grid, 256, 0, s>>>((const float*)table, tok_id, x, M_max, m_ptr, D);
  else embed_rows_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)table, tok_id, x, M_max, m_ptr, D);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// RMS norm: one warp per row
// ------------------------------------------------------------------------------------------------
template <typename T, typename TX = float>
__global__ void rmsnorm_rows_kernel(const TX* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                                    int M_imm, const int* __restrict__ m_ptr, int D, float eps, float scale,
                                    const float* __restrict__ pos_table, const int* __restrict__ tok_pos) {
  const int M = m_ptr ? *m_ptr : M_imm;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const TX* xr = x + (size_t)row * D;
  float ss = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = load4(xr + c);
    ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
  }
  ss = warp_sum(ss);
  const float r = 1.0f / sqrtf(ss / (float)D + eps);
  const float* pe = pos_table ? pos_table + (size_t)tok_pos[row] * D : nullptr;
  T* yr = y + (size_t)row * D;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = load4(xr + c);
    const float4 g = load4(w + c);
    float4 o;
    // separately rounded multiply / add, as the reference computes them (weight * normed, then + position row); no
    // FMA contraction, so the cached-item path (cached_assemble_kernel) reproduces these bits exactly
    o.x = __fmul_rn(g.x, __fmul_rn(v.x, r)); o.y = __fmul_rn(g.y, __fmul_rn(v.y, r));
    o.z = __fmul_rn(g.z, __fmul_rn(v.z, r)); o.w = __fmul_rn(g.w, __fmul_rn(v.w, r));
    if (scale != 1.0f) { o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale; }
    if (pe) {
      const float4 q = load4(pe + c);
      o.x = __fadd_rn(o.x, q.x); o.y = __fadd_rn(o.y, q.y); o.z = __fadd_rn(o.z, q.z); o.w = __fadd_rn(o.w, q.w);
    }
    store4(yr + c, o);
  }
}

// the same norm over a bf16 residual stream (bf16 mode unless GRAM_FLAG_FP32_RESID)
cudaError_t rmsnorm_rows_stream(int dtype, const void* x, const float* w, void* y, int M_max, const int* m_ptr, int D,
                                float eps, float scale, const float* pos_table, const int* tok_pos, cudaStream_t s) {
  if (M_max <= 0) return cudaSuccess;
  const int grid = (M_max + 7) / 8;
  if (dtype == 0)
    rmsnorm_rows_kernel<float, bf16><<<grid, 256, 0, s>>>((const bf16*)x, w, (float*)y, M_max, m_ptr, D, eps, scale, pos_table, tok_pos);
  else
    rmsnorm_rows_kernel<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)x, w, (bf16*)y, M_max, m_ptr, D, eps, scale, pos_table, tok_pos);
  return cudaGetLastError();
}

cudaError_t rmsnorm_rows(int dtype, const float* x, const float* w, void* y, int M_max, const int* m_ptr, int D,
                         float eps, float scale, const float* pos_table, const int* tok_pos, cudaStream_t s) {
  if (M_max <= 0) return cudaSuccess;
  const int grid = (M_max + 7) / 8;
  if (dtype == 0)
    rmsnorm_rows_kernel<float><<<grid, 256, 0, s>>>(x, w, (float*)y, M_max, m_ptr, D, eps, scale, pos_table, tok_pos);
  else
    rmsnorm_rows_kernel<bf16><<<grid, 256, 0, s>>>(x, w, (bf16*)y, M_max, m_ptr, D, eps, scale, pos_table, tok_pos);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// encoder self-attention, CUDA-core version: one CTA per (passage, head), one warp per query row.
// K and V of the passage/head are staged in shared memory as fp32; a lane owns keys lane, lane+32, ...
// ------------------------------------------------------------------------------------------------
template <typename T, int DK, int NI>
__global__ void __launch_bounds__(128)
enc_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, const int* __restrict__ plen,
                     const int* __restrict__ poff, const uint8_t* __restrict__ tok_valid,
                     const float* __restrict__ bias_lut, int Lb, int H) {
  constexpr int LK = NI * 32;
  constexpr int KS = DK + 1;                       // padded K row: lanes read different rows, same column
  const int p = blockIdx.x, h = blockIdx.y;
  const int len = plen[p];
  if (len == 0) return;
  const int row0 = poff[p];
  const int HD = H * DK;
  const size_t ld = (size_t)3 * HD;

  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;                                // [LK][KS]
  float* Vs = Ks + LK * KS;                        // [LK][DK]
  float* lut = Vs + LK * DK;                       // [2*Lb-1]
  float* qs = lut + (2 * Lb - 1);                  // [4][DK]
  float* ps = qs + 4 * DK;                         // [4][LK]
  uint8_t* vs = reinterpret_cast<uint8_t*>(ps + 4 * LK);   // [LK]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < len * (DK / 4); i += 128) {
    const int j = i / (DK / 4), c = (i % (DK / 4)) * 4;
    const T* base = qkv + (size_t)(row0 + j) * ld + h * DK + c;
    const float4 kf = load4(base + HD);
    const float4 vf = load4(base + 2 * HD);
    Ks[j * KS + c + 0] = kf.x; Ks[j * KS + c + 1] = kf.y; Ks[j * KS + c + 2] = kf.z; Ks[j * KS + c + 3] = kf.w;
    *reinterpret_cast<float4*>(&Vs[j * DK + c]) = vf;
  }
  for (int i = tid; i < 2 * Lb - 1; i += 128) lut[i] = bias_lut[(size_t)h * (2 * Lb - 1) + i];
  for (int i = tid; i < len; i += 128) vs[i] = tok_valid[row0 + i];
  __syncthreads();

  for (int qi = wid; qi < len; qi += 4) {
    // stage this query row
    const T* qrow = qkv + (size_t)(row0 + qi) * ld + h * DK;
    for (int d = lane; d < DK; d += 32) qs[wid * DK + d] = to_f32(qrow[d]);
    __syncwarp();
    float sc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) sc[i] = 0.f;
#pragma unroll 8
    for (int d = 0; d < DK; ++d) {
      const float qd = qs[wid * DK + d];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int j = lane + 32 * i;
        if (j < len) sc[i] = fmaf(qd, Ks[j * KS + d], sc[i]);
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int j = lane + 32 * i;
      if (j < len && vs[j]) {
        sc[i] += lut[j - qi + Lb - 1];
        mx = fmaxf(mx, sc[i]);
      } else {
        sc[i] = -INFINITY;
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      sc[i] = (sc[i] == -INFINITY) ? 0.f : expf(sc[i] - mx);
      sum += sc[i];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int j = lane + 32 * i;
      if (j < LK) ps[wid * LK + j] = sc[i] / sum;
    }
    __syncwarp();
    // out[d] = sum_j p_j V[j][d]
    float o[(DK + 31) / 32];
#pragma unroll
    for (int e = 0; e < (DK + 31) / 32; ++e) o[e] = 0.f;
    for (int j = 0; j < len; ++j) {
      const float pj = ps[wid * LK + j];
#pragma unroll
      for (int e = 0; e < (DK + 31) / 32; ++e) {
        const int d = lane + 32 * e;
        if (d < DK) o[e] = fmaf(pj, Vs[j * DK + d], o[e]);
      }
    }
    T* orow = out + (size_t)(row0 + qi) * HD + h * DK;
#pragma unroll
    for (int e = 0; e < (DK + 31) / 32; ++e) {
      const int d = lane + 32 * e;
      if (d < DK) orow[d] = from_f32<T>(o[e]);
    }
    __syncwarp();
  }
}

template <typename T, int DK, int NI>
static cudaError_t launch_enc_attn(const void* qkv, void* out, const int* plen, const int* poff,
                                   const uint8_t* tok_valid, const float* bias_lut, int Lb, int P, int H,
                                   cudaStream_t s) {
  constexpr int LK = NI * 32;
  size_t smem = sizeof(float) * ((size_t)LK * (DK + 1) + (size_t)LK * DK + (2 * Lb - 1) + 4 * DK + 4 * LK) + LK + 16;
  auto kern = enc_attention_kernel<T, DK, NI>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<dim3(P, H), 128, smem, s>>>((const T*)qkv, (T*)out, plen, poff, tok_valid, bias_lut, Lb, H);
  return cudaGetLastError();
}

template <typename T, int DK>
static cudaError_t dispatch_enc_attn_ni(const void* qkv, void* out, const int* plen, const int* poff,
                                        const uint8_t* tok_valid, const float* bias_lut, int Lb, int P, int H,
                                        int Lmax, cudaStream_t s) {
  if (Lmax <= 32) return launch_enc_attn<T, DK, 1>(qkv, out, plen, poff, tok_valid, bias_lut, Lb, P, H, s);
  if (Lmax <= 64) return launch_enc_attn<T, DK, 2>(qkv, out, plen, poff, tok_valid, bias_lut, Lb, P, H, s);
  if (Lmax <= 128) return launch_enc_attn<T, DK, 4>(qkv, out, plen, poff, tok_valid, bias_lut, Lb, P, H, s);
  if (Lmax <= 256) return launch_enc_attn<T, DK, 8>(qkv, out, plen, poff, tok_valid, bias_lut, Lb, P, H, s);
  return cudaErrorInvalidValue;
}

template <typename T>
static cudaError_t dispatch_enc_attn_dk(const void* qkv, void* out, const int* plen, const int* poff,
                                        const uint8_t* tok_valid, const float* bias_lut, int Lb, int P, int H,
                                        int dk, int Lmax, cudaStream_t s) {
  switch (dk) {
    case 16: return dispatch_enc_attn_ni<T, 16>(qkv, out, plen, poff, tok_valid, bias_lut, Lb, P, H, Lmax, s);
    case 32: return dispatch_enc_attn_ni<T, 32>(qkv, out, plen, poff, tok_valid, bias_lut, Lb, P, H, Lmax, s);
    case 64: return dispatch_enc_attn_ni<T, 64>(qkv, out, plen, poff, tok_valid, bias_lut, Lb, P, H, Lmax, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t enc_attention(int dtype, const void* qkv, void* out, const int* plen, const int* poff,
                          const uint8_t* tok_valid, const float* bias_lut, int Lb, int P, int H, int dk, int Lmax,
                          cudaStream_t s) {
  if (P <= 0) return cudaSuccess;
  if (dtype == 0)
    return dispatch_enc_attn_dk<float>(qkv, out, plen, poff, tok_valid, bias_lut, Lb, P, H, dk, Lmax, s);
  return dispatch_enc_attn_dk<bf16>(qkv, out, plen, poff, tok_valid, bias_lut, Lb, P, H, dk, Lmax, s);
}

// ------------------------------------------------------------------------------------------------
// debug tap: packed memory -> padded fp32 [B, N*L, D] (caller zero-fills first)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void unpack_memory_kernel(const T* __restrict__ mem, const int* __restrict__ row_src,
                                     float* __restrict__ out, int M_imm, const int* __restrict__ m_ptr, int D) {
  const int M = m_ptr ? *m_ptr : M_imm;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const T* src = mem + (size_t)row * D;
  float* dst = out + (size_t)row_src[row] * D;
  for (int c = lane * 4; c < D; c += 128) store4(dst + c, load4(src + c));
}

cudaError_t unpack_memory(int dtype, const void* mem, const int* row_src, float* out, int M_max, const int* m_ptr,
                          int D, cudaStream_t s) {
  if (M_max <= 0) return cudaSuccess;
  const int grid = (M_max + 7) / 8;
  if (dtype == 0) unpack_memory_kernel<float><<<grid, 256, 0, s>>>((const float*)mem, row_src, out, M_max, m_ptr, D);
  else unpack_memory_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)mem, row_src, out, M_max, m_ptr, D);
  return cudaGetLastError();
}

}  // namespace gram
