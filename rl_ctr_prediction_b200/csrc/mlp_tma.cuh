// mlp_tma.cuh -- internal interface of the TMA-fed 3xTF32 GEMM (mlp_tma.cu), used by the C-ABI entry points in mlp.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rlctr {
namespace tma {

// One GEMM operand.  K-major: memory [rows = M or N][K], `pitch` floats between rows.  MN-major: memory
// [K rows][M or N], `pitch` floats between k-rows.  lo != null: (ptr, lo) are the pre-split (hi, lo) images.
struct Operand {
    const float* ptr;
    const float* lo;
    int64_t pitch;
    bool mn_major;
};

int enabled();                                            // RLCTR_GEMM_TMA != 0 and the driver exports cuTensorMapEncodeTiled
int plan_splits(int M, int N, int K, bool b_mn);          // split-K factor the planner will use (workspace sizing)
int split_weight(const float* w, float* hi, float* lo, int rows, int cols, int pitch, cudaStream_t st);
// C[(split*M + m)*ldc + n] = sum_k A[m,k] B[n,k]; RLCTR_EUNSUPPORTED when an operand is not TMA-addressable
int gemm(const Operand& A, const Operand& B, float* C, int64_t ldc, const float* bias, int M, int N, int K, int relu,
         bool allow_split, cudaStream_t st);

}  // namespace tma
}  // namespace rlctr
