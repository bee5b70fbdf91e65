// tcgen05 / TMEM / TMA TF32 GEMM (the BF_PREC_TF32 path).  See gemm_tc.cu.
#pragma once
#include "common.cuh"

namespace bf {

// C[M,N] (row-major, ldc) = act(A @ B + bias[n]) (+ C when accumulate).
//   A: a_kmajor ? A[m*lda + k] : A[k*lda + m]        B: b_kmajor ? B[n*ldb + k] : B[k*ldb + n]
// Both operands are read straight from HBM by TMA into 128B-swizzled shared tiles; the MMA is
// tcgen05.mma.cta_group::1.kind::tf32 with the accumulator in TMEM.
bool tc_gemm_supported(int M, int N, int K, bool a_kmajor, bool b_kmajor);
int tc_gemm(const float* A, long lda, bool a_kmajor, const float* B, long ldb, bool b_kmajor, float* C, long ldc, int M,
            int N, int K, const float* bias, int act, int accumulate, cudaStream_t s);

}  // namespace bf
