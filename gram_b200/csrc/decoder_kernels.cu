// Decoder-side attention kernels.
//
//  * dec_self_attention: single-token causal self-attention over <= max_length cached positions
//    (reference src/model/gram_t5_modeling.py:536-540 `cat` of past K/V, :572-621 attention,
//    decoder layer-0 causal relative bias :397-477).  Instead of physically re-ordering the cache
//    after every beam step (reference src/model/gram_t5.py:320-348, index_select on every tensor),
//    each beam row carries an ancestry table anc[row][j] = "row of my user that wrote position j".
//
//  * cross_attention (kernel (b) of the north star): decoder cross-attention over the fused FiD
//    memory (reference src/model/gram_t5_modeling.py:670-705, 549, 572-621).  The reference expands
//    the encoder memory K-fold per beam (HF `_expand_inputs_for_generation`) and streams K/V of shape
//    (B*K) x H x S x dk per layer per step; here the K beams of a user are the rows of ONE problem and
//    the user's K/V -- written in place by the projection GEMM in packed [token][layer][K|V][head][dk]
//    order, no concat copy -- is read once.  This file holds the CUDA-core version used by the fp32
//    parity mode; cross_attention_mma.cu holds the bf16 tensor-core version.
#include "common.cuh"
#include "kernels.h"

namespace gram {

// ------------------------------------------------------------------------------------------------
// self-attention: one warp per (row, head)
// ------------------------------------------------------------------------------------------------
template <typename T, int DK>
__global__ void __launch_bounds__(128)
dec_self_attention_kernel(const T* __restrict__ qkv, T* __restrict__ cache_k, T* __restrict__ cache_v,
                          const int* __restrict__ anc, int Tmax, const float* __restrict__ dec_bias, int n_dec,
                          T* __restrict__ out, int R, int K, int H, int t, const int* __restrict__ slot_row,
                          const int* __restrict__ n_rows) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= (n_rows ? *n_rows : R) * H) return;
  const int slot = warp / H, h = warp % H;                  // qkv / out row
  const int r = slot_row ? slot_row[slot] : slot;           // beam row: cache row, ancestry, user membership
  const int HD = H * DK;
  const int ubase = (r / K) * K;
  const T* qrow = qkv + (size_t)slot * 3 * HD + h * DK;
  // append this step's key/value (slot t, own row)
  T* kslot = cache_k + ((size_t)t * R + r) * HD + h * DK;
  T* vslot = cache_v + ((size_t)t * R + r) * HD + h * DK;
  for (int d = lane; d < DK; d += 32) {
    kslot[d] = qrow[HD + d];
    vslot[d] = qrow[2 * HD + d];
  }
  __syncwarp();
  __shared__ float qs_all[4][DK];
  __shared__ float ps_all[4][64];
  __shared__ int src_all[4][64];
  float* qs = qs_all[threadIdx.x >> 5];
  float* ps = ps_all[threadIdx.x >> 5];
  int* src = src_all[threadIdx.x >> 5];
  for (int d = lane; d < DK; d += 32) qs[d] = to_f32(qrow[d]);
  const int npos = t + 1;                                   // <= 64
  for (int j = lane; j < npos; j += 32) src[j] = (j == t) ? r : ubase + anc[(size_t)r * Tmax + j];
  __syncwarp();
  float sc[2];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = lane + 32 * i;
    sc[i] = -INFINITY;
    if (j < npos) {
      const T* kr = cache_k + ((size_t)j * R + src[j]) * HD + h * DK;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < DK; d += 4) {
        const float4 kf = load4(kr + d);
        a = fmaf(qs[d + 0], kf.x, a); a = fmaf(qs[d + 1], kf.y, a);
        a = fmaf(qs[d + 2], kf.z, a); a = fmaf(qs[d + 3], kf.w, a);
      }
      const int dist = t - j;
      sc[i] = a + dec_bias[h * n_dec + (dist < n_dec ? dist : n_dec - 1)];
      mx = fmaxf(mx, sc[i]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    sc[i] = (sc[i] == -INFINITY) ? 0.f : expf(sc[i] - mx);
    sum += sc[i];
  }
  sum = warp_sum(sum);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = lane + 32 * i;
    if (j < 64) ps[j] = sc[i] / sum;
  }
  __syncwarp();
  float o[(DK + 31) / 32];
#pragma unroll
  for (int e = 0; e < (DK + 31) / 32; ++e) o[e] = 0.f;
  for (int j = 0; j < npos; ++j) {
    const T* vr = cache_v + ((size_t)j * R + src[j]) * HD + h * DK;
    const float pj = ps[j];
#pragma unroll
    for (int e = 0; e < (DK + 31) / 32; ++e) {
      const int d = lane + 32 * e;
      if (d < DK) o[e] = fmaf(pj, to_f32(vr[d]), o[e]);
    }
  }
  T* orow = out + (size_t)slot * HD + h * DK;
#pragma unroll
  for (int e = 0; e < (DK + 31) / 32; ++e) {
    const int d = lane + 32 * e;
    if (d < DK) orow[d] = from_f32<T>(o[e]);
  }
}

// d_kv = 64 specialisation with a short dependency chain: lane = (position group pg = lane/8, 16-byte chunk
// c = lane%8).  After one round trip for q/k/v + the ancestry row, every K and V chunk the warp needs is requested
// at once (addresses depend on the ancestry only), so a step costs two memory round trips instead of five.
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// 8 consecutive elements of a cache row as they sit in memory: K/V chunks wait in registers in storage form (4
// registers per 8 bf16) and are widened only when they are consumed, so a warp keeps more positions in flight and
// more warps fit on an SM (the kernel is bound by memory latency: ncu, 127 registers, 16 warps per SM, 2.4 TB/s)
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 u;
  __device__ __forceinline__ void load(const bf16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = u; }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a; *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// One warp = (row, 4 heads): lane = (head hq = lane/8, 16-byte chunk c = lane%8), positions are walked with an
// online softmax, UN positions' K and V chunks requested per iteration.  Work is proportional to t+1.
template <typename T, int UN, int MINB>
__global__ void __launch_bounds__(128, MINB)
dec_self_attention64_kernel(const T* __restrict__ qkv, T* __restrict__ cache_k, T* __restrict__ cache_v,
                            const int* __restrict__ anc, int Tmax, const float* __restrict__ dec_bias, int n_dec,
                            T* __restrict__ out, int R, int K, int H, int t, const int* __restrict__ slot_row,
                            const int* __restrict__ n_rows) {
  constexpr int DK = 64;
  const int HQ = H >> 2;                                 // head quads per row
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= (n_rows ? *n_rows : R) * HQ) return;
  const int slot = warp / HQ, h = (warp % HQ) * 4 + (lane >> 3), c = lane & 7;
  const int r = slot_row ? slot_row[slot] : slot;        // beam row: cache row, ancestry, user membership
  const int HD = H * DK;
  const int ubase = (r / K) * K;
  const size_t col = (size_t)h * DK + c * 8;
  // ancestry of this row: lane j holds position j (and j + 32)
  const int a0 = (lane < t) ? anc[(size_t)r * Tmax + lane] : 0;
  const int a1 = (lane + 32 < t) ? anc[(size_t)r * Tmax + lane + 32] : 0;
  const T* qrow = qkv + (size_t)slot * 3 * HD + col;
  Raw8<T> qr, kc, vc;
  qr.load(qrow);
  kc.load(qrow + HD);
  vc.load(qrow + 2 * HD);
  kc.store(cache_k + ((size_t)t * R + r) * HD + col);    // append this step's key/value (slot t, own row)
  vc.store(cache_v + ((size_t)t * R + r) * HD + col);
  float qv[8];
  qr.unpack(qv);
  const float* bias = dec_bias + h * n_dec;
  float m = -INFINITY, l = 0.f;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  auto absorb = [&](const Raw8<T>& kraw, const Raw8<T>& vraw, int j) {
    float kx[8];
    kraw.unpack(kx);
    float a = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) a = fmaf(qv[e], kx[e], a);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    const int dist = t - j;
    a += bias[dist < n_dec ? dist : n_dec - 1];
    const float mn = fmaxf(m, a);
    const float corr = expf(m - mn);                     // exp(-inf) = 0 on the first position
    const float p = expf(a - mn);
    l = l * corr + p;
    float vx[8];
    vraw.unpack(vx);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vx[e], acc[e] * corr);
    m = mn;
  };
  for (int j0 = 0; j0 < t; j0 += UN) {
    Raw8<T> kx[UN], vx[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u;                              // warp-uniform
      if (j < t) {
        const int sj = ubase + __shfl_sync(0xffffffffu, j < 32 ? a0 : a1, j & 31);
        const size_t off = ((size_t)j * R + sj) * HD + col;
        kx[u].load(cache_k + off);
        vx[u].load(cache_v + off);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
      if (j0 + u < t) absorb(kx[u], vx[u], j0 + u);
  }
  absorb(kc, vc, t);
  const float inv = 1.0f / l;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] *= inv;
  store8(out + (size_t)slot * HD + col, acc);
}

template <typename T>
static cudaError_t launch_dec_self(const void* qkv, void* ck, void* cv, const int* anc, int Tmax,
                                   const float* dec_bias, int n_dec, void* out, int R, int K, int H, int dk, int t,
                                   const int* slot_row, const int* n_rows, cudaStream_t s) {
  const int warps = (dk == 64 && (H & 3) == 0) ? R * (H / 4) : R * H;
  const int grid = (warps + 3) / 4;
  if (dk == 64 && (H & 3)) {
    dec_self_attention_kernel<T, 64><<<grid, 128, 0, s>>>((const T*)qkv, (T*)ck, (T*)cv, anc, Tmax, dec_bias, n_dec, (T*)out, R, K, H, t, slot_row, n_rows);
    return cudaGetLastError();
  }
  switch (dk) {
    case 16: dec_self_attention_kernel<T, 16><<<grid, 128, 0, s>>>((const T*)qkv, (T*)ck, (T*)cv, anc, Tmax, dec_bias, n_dec, (T*)out, R, K, H, t, slot_row, n_rows); break;
    case 32: dec_self_attention_kernel<T, 32><<<grid, 128, 0, s>>>((const T*)qkv, (T*)ck, (T*)cv, anc, Tmax, dec_bias, n_dec, (T*)out, R, K, H, t, slot_row, n_rows); break;
    case 64:
      // bf16, per 1,888-user step on one B200 (A/B in one gpurun call): K/V widened to fp32 on load, 127 registers,
      // 16 warps per SM: 8.2 ms; storage-form chunks with 8 positions in flight, 117 registers, 16 warps: 8.0 ms;
      // 4 positions, 80 registers, 24 warps: 6.1 ms; 4 positions, 64 registers (32 bytes spilled), 32 warps: 5.5 ms;
      // 8-byte lanes (2 heads per warp), 8 positions, 32 warps: 8.4 ms.  Occupancy beats batch depth.  Round 2: consecutive
      // warps = consecutive beams of one head quad (shared ancestors as L1 hits): 6.1 vs 5.9-6.1 ms, no gain.
      if constexpr (sizeof(T) == 2) {
        dec_self_attention64_kernel<T, 4, 8><<<grid, 128, 0, s>>>((const T*)qkv, (T*)ck, (T*)cv, anc, Tmax, dec_bias, n_dec, (T*)out, R, K, H, t, slot_row, n_rows);
      } else {
        dec_self_attention64_kernel<T, 4, 4><<<grid, 128, 0, s>>>((const T*)qkv, (T*)ck, (T*)cv, anc, Tmax, dec_bias, n_dec, (T*)out, R, K, H, t, slot_row, n_rows);
      }
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t dec_self_attention(int dtype, const void* qkv, void* cache_k, void* cache_v, const int* anc, int Tmax,
                               const float* dec_bias, int n_dec, void* out, int R, int K, int H, int dk, int t,
                               const int* slot_row, const int* n_rows, cudaStream_t s) {
  if (R <= 0) return cudaSuccess;
  if (t + 1 > 64) return cudaErrorInvalidValue;
  if (dtype == 0)
    return launch_dec_self<float>(qkv, cache_k, cache_v, anc, Tmax, dec_bias, n_dec, out, R, K, H, dk, t, slot_row, n_rows, s);
  return launch_dec_self<bf16>(qkv, cache_k, cache_v, anc, Tmax, dec_bias, n_dec, out, R, K, H, dk, t, slot_row, n_rows, s);
}

// ------------------------------------------------------------------------------------------------
// cross-attention, CUDA-core version: one CTA per (user, head); the K beams are the query rows.
// Tiles of TS memory rows are staged through shared memory with vectorised coalesced loads;
// online (flash-style) softmax keeps one running max/sum per beam.
// ------------------------------------------------------------------------------------------------
constexpr int XA_THREADS = 128;
constexpr int XA_TS = 32;        // memory rows per tile
constexpr int XA_KMAX = 64;      // beams per user supported by this kernel

template <typename T, int DK>
__global__ void __launch_bounds__(XA_THREADS)
cross_attention_kernel(const T* __restrict__ q, const T* __restrict__ kv, size_t kv_stride, int k_off, int v_off,
                       const int* __restrict__ ustart, const uint8_t* __restrict__ tok_valid, T* __restrict__ out,
                       int K_all, int H, const int* __restrict__ live_start, const int* __restrict__ live_count) {
  constexpr int KS = DK + 4;                 // padded, keeps 16B alignment and conflict-free LDS.128
  constexpr int G = XA_THREADS / DK;         // beam groups in the PV phase
  constexpr int NBT = (XA_KMAX + G - 1) / G; // beams per thread in the PV phase
  constexpr int NBS = XA_KMAX / 4;           // beams per thread in the score phase (4 groups of 32 rows)
  const int u = blockIdx.x, h = blockIdx.y;
  // query rows of this user: its K beams, or (live-row compaction) its live beams only
  const int K = live_start ? live_count[u] : K_all;
  if (K == 0) return;
  const int qrow0 = live_start ? live_start[u] : u * K_all;
  __shared__ int s_qrow0;                    // re-read by the epilogue instead of staying live across the key loop
  if (threadIdx.x == 0) s_qrow0 = qrow0;
  const int s_beg = ustart[u], s_end = ustart[u + 1];
  const int HD = H * DK;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  extern __shared__ __align__(16) float smem[];
  float* qs = smem;                          // [K][DK]
  float* Ks = qs + XA_KMAX * DK;             // [TS][KS]
  float* Vs = Ks + XA_TS * KS;               // [TS][DK]
  float* Ps = Vs + XA_TS * DK;               // [KMAX][TS]
  float* m_s = Ps + XA_KMAX * XA_TS;         // [KMAX] running max
  float* l_s = m_s + XA_KMAX;                // [KMAX] running sum
  float* c_s = l_s + XA_KMAX;                // [KMAX] correction factor of this tile

  for (int i = tid; i < K * DK; i += XA_THREADS) {
    const int b = i / DK, d = i % DK;
    qs[b * DK + d] = to_f32(q[(size_t)(qrow0 + b) * HD + h * DK + d]);
  }
  for (int i = tid; i < XA_KMAX; i += XA_THREADS) { m_s[i] = -INFINITY; l_s[i] = 0.f; c_s[i] = 0.f; }

  float acc[NBT];
#pragma unroll
  for (int i = 0; i < NBT; ++i) acc[i] = 0.f;
  const int pv_d = tid % DK, pv_g = tid / DK;
  __syncthreads();

  for (int s0 = s_beg; s0 < s_end; s0 += XA_TS) {
    const int rows = min(XA_TS, s_end - s0);
    // ---- stage K and V rows of this head ----
    for (int i = tid; i < XA_TS * (DK / 4); i += XA_THREADS) {
      const int j = i / (DK / 4), c = (i % (DK / 4)) * 4;
      float4 kf = make_float4(0.f, 0.f, 0.f, 0.f), vf = kf;
      if (j < rows) {
        const T* base = kv + (size_t)(s0 + j) * kv_stride + h * DK + c;
        kf = load4(base + k_off);
        vf = load4(base + v_off);
      }
      *reinterpret_cast<float4*>(&Ks[j * KS + c]) = kf;
      *reinterpret_cast<float4*>(&Vs[j * DK + c]) = vf;
    }
    __syncthreads();
    // ---- scores: thread = (row j = lane, beam group = warp) ----
    {
      float sc[NBS];
#pragma unroll
      for (int i = 0; i < NBS; ++i) sc[i] = 0.f;
      const int j = lane;
#pragma unroll 4
      for (int d = 0; d < DK; d += 4) {
        const float4 kf = *reinterpret_cast<const float4*>(&Ks[j * KS + d]);
#pragma unroll
        for (int i = 0; i < NBS; ++i) {
          const int b = wid + 4 * i;
          if (b < K) {
            const float4 qf = *reinterpret_cast<const float4*>(&qs[b * DK + d]);
            sc[i] = fmaf(qf.x, kf.x, sc[i]); sc[i] = fmaf(qf.y, kf.y, sc[i]);
            sc[i] = fmaf(qf.z, kf.z, sc[i]); sc[i] = fmaf(qf.w, kf.w, sc[i]);
          }
        }
      }
      const bool ok = (j < rows) && (tok_valid == nullptr || tok_valid[s0 + j]);
      // ---- online softmax per beam (this warp owns beams wid, wid+4, ...) ----
#pragma unroll
      for (int i = 0; i < NBS; ++i) {
        const int b = wid + 4 * i;
        if (b < K) {                                        // warp-uniform
          const float v = ok ? sc[i] : -INFINITY;
          const float tmax = warp_max(v);
          const float m_old = m_s[b];
          const float m_new = fmaxf(m_old, tmax);
          const float p = (v == -INFINITY) ? 0.f : expf(v - m_new);
          const float psum = warp_sum(p);
          Ps[b * XA_TS + j] = p;
          if (lane == 0) {
            const float corr = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
            c_s[b] = corr;
            l_s[b] = l_s[b] * corr + psum;
            m_s[b] = m_new;
          }
        }
      }
    }
    __syncthreads();
    // ---- PV: thread = (d = tid % DK, beam group = tid / DK) ----
#pragma unroll
    for (int i = 0; i < NBT; ++i) {
      const int b = pv_g + G * i;
      if (b < K) {
        float a = acc[i] * c_s[b];
#pragma unroll 8
        for (int j = 0; j < XA_TS; j += 4) {
          const float4 pf = *reinterpret_cast<const float4*>(&Ps[b * XA_TS + j]);
          a = fmaf(pf.x, Vs[(j + 0) * DK + pv_d], a);
          a = fmaf(pf.y, Vs[(j + 1) * DK + pv_d], a);
          a = fmaf(pf.z, Vs[(j + 2) * DK + pv_d], a);
          a = fmaf(pf.w, Vs[(j + 3) * DK + pv_d], a);
        }
        acc[i] = a;
      }
    }
    __syncthreads();
  }
  const int out_row0 = s_qrow0;
#pragma unroll
  for (int i = 0; i < NBT; ++i) {
    const int b = pv_g + G * i;
    if (b < K) {
      const float l = l_s[b];
      out[(size_t)(out_row0 + b) * HD + h * DK + pv_d] = from_f32<T>(l > 0.f ? acc[i] / l : 0.f);
    }
  }
}

template <typename T, int DK>
static cudaError_t launch_cross(const void* q, const void* kv, size_t kv_stride, int k_off, int v_off,
                                const int* ustart, const uint8_t* tok_valid, void* out, int users, int K, int H,
                                const int* live_start, const int* live_count, cudaStream_t s) {
  constexpr int KS = DK + 4;
  const size_t smem = sizeof(float) * ((size_t)XA_KMAX * DK + XA_TS * KS + XA_TS * DK + XA_KMAX * XA_TS + 3 * XA_KMAX);
  auto kern = cross_attention_kernel<T, DK>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<dim3(users, H), XA_THREADS, smem, s>>>((const T*)q, (const T*)kv, kv_stride, k_off, v_off, ustart,
                                                 tok_valid, (T*)out, K, H, live_start, live_count);
  return cudaGetLastError();
}

cudaError_t cross_attention(int dtype, const void* q, const void* kv, size_t kv_stride, int k_off, int v_off,
                            const int* ustart, const uint8_t* tok_valid, void* out, int users, int K, int H, int dk,
                            const int* live_start, const int* live_count, cudaStream_t s) {
  if (users <= 0) return cudaSuccess;
  if (K > XA_KMAX) return cudaErrorInvalidValue;
#define GRAM_XA(TT)                                                                                              \
  switch (dk) {                                                                                                  \
    case 16: return launch_cross<TT, 16>(q, kv, kv_stride, k_off, v_off, ustart, tok_valid, out, users, K, H,  \
                                            live_start, live_count, s);                                             \
    case 32: return launch_cross<TT, 32>(q, kv, kv_stride, k_off, v_off, ustart, tok_valid, out, users, K, H,  \
                                            live_start, live_count, s);                                             \
    case 64: return launch_cross<TT, 64>(q, kv, kv_stride, k_off, v_off, ustart, tok_valid, out, users, K, H,  \
                                            live_start, live_count, s);                                             \
    default: return cudaErrorInvalidValue;                                                                       \
  }
  if (dtype == 0) { GRAM_XA(float) }
  GRAM_XA(bf16)
#undef GRAM_XA
}

}  // namespace gram
