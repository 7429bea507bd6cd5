// CUDA-core GEMM  C[M,N] = A[M,K] * W[N,K]^T  with fp32 accumulation in a fixed k order.
//
// This is the arithmetic behind every nn.Linear of the reference T5 (q/k/v/o, wi/wo, lm_head:
// reference src/model/gram_t5_modeling.py:305-310,369-372,553-569,622; src/model/gram_t5.py:254).
// It is the fp32 *parity mode* GEMM (plain TF32 tensor-core math cannot meet the 1e-4 logits
// tolerance) and the A/B reference for the tcgen05 kernel in gemm_tc.cu.
//
// Tiling: 128x128x16 per CTA, 256 threads, 8x8 outputs per thread held as four 4x4 register blocks so
// that every shared-memory read is a conflict-free 16-byte LDS; global loads are 16-byte (fp32) or
// 8-byte (bf16) vector loads along K, software-pipelined through registers.
#include "common.cuh"
#include "kernels.h"

namespace gram {

constexpr int BM = 128, BN = 128, BK = 16, GEMM_THREADS = 256;

template <typename T, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ W, void* __restrict__ Cv,
                 int M_imm, const int* __restrict__ m_ptr, int N, int K) {
  const int M = m_ptr ? *m_ptr : M_imm;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= M) return;

  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  // loader mapping: each thread moves 2 x (4 consecutive k) for A and for W
  const int lrow = tid >> 2;            // 0..63
  const int lk = (tid & 3) * 4;         // 0,4,8,12
  // compute mapping: 16x16 threads; thread owns rows {ty*4..+4, 64+ty*4..+4}, cols likewise
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + lk;
      ra[h] = (gm < M && gk < K) ? load4(A + (size_t)gm * K + gk) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[h] = (gn < N && gk < K) ? load4(W + (size_t)gn * K + gk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y;
      Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows {ty*4+i, 64+ty*4+i}, cols {tx*4.., 64+tx*4..}; N % 4 == 0 is required by the host
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int gn = n0 + jh * 64 + tx * 4;
      if (gn >= N) continue;
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      const size_t off = (size_t)gm * N + gn;
      if (EPI == EPI_STORE) {
        store4(reinterpret_cast<T*>(Cv) + off, v);
      } else if (EPI == EPI_RELU) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        store4(reinterpret_cast<T*>(Cv) + off, v);
      } else if (EPI == EPI_RESID) {
        float* c = reinterpret_cast<float*>(Cv) + off;
        float4 o = *reinterpret_cast<float4*>(c);
        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
        *reinterpret_cast<float4*>(c) = o;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(Cv) + off) = v;
      }
    }
  }
}

template <typename T>
static cudaError_t launch_simt(int epi, const T* A, const T* W, void* C, int M, const int* m_ptr, int N, int K,
                               cudaStream_t s) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  switch (epi) {
    case EPI_STORE: gemm_simt_kernel<T, EPI_STORE><<<grid, GEMM_THREADS, 0, s>>>(A, W, C, M, m_ptr, N, K); break;
    case EPI_RELU:  gemm_simt_kernel<T, EPI_RELU><<<grid, GEMM_THREADS, 0, s>>>(A, W, C, M, m_ptr, N, K); break;
    case EPI_RESID: gemm_simt_kernel<T, EPI_RESID><<<grid, GEMM_THREADS, 0, s>>>(A, W, C, M, m_ptr, N, K); break;
    case EPI_F32:   gemm_simt_kernel<T, EPI_F32><<<grid, GEMM_THREADS, 0, s>>>(A, W, C, M, m_ptr, N, K); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t gemm_simt(int dtype, int epi, const void* A, const void* W, void* C, int M_max, const int* m_ptr,
                      int N, int K, cudaStream_t s) {
  if ((N & 3) || (K & 3)) return cudaErrorInvalidValue;
  if (dtype == 0) return launch_simt<float>(epi, (const float*)A, (const float*)W, C, M_max, m_ptr, N, K, s);
  return launch_simt<bf16>(epi, (const bf16*)A, (const bf16*)W, C, M_max, m_ptr, N, K, s);
}

}  // namespace gram
