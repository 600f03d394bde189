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

// optional epilogue work (all-zero = none)
struct Epilogue {
    const uint64_t* drop_state = nullptr;    // device {seed, counter}: forward dropout of C (after bias / ReLU)
    uint32_t drop_thresh = 0;                //   keep when hash >= thresh (= p * 2^32)
    float drop_scale = 1.f;                  //   1 / (1 - p)
    const float* mask_src = nullptr;         // C *= (mask_src[m*mask_ld + n] > 0 ? mask_scale : 0)
    int64_t mask_ld = 0;
    int mvec = 1;
    float mask_scale = 1.f;
};

int enabled();                                            // RLCTR_GEMM_TMA != 0 and the driver exports cuTensorMapEncodeTiled
int plan_splits(int M, int N, int K, bool b_mn);          // split-K factor the planner will use (workspace sizing)
int split_weight(const float* w, float* hi, float* lo, int rows, int cols, int pitch, cudaStream_t st);
// C[(split*M + m)*ldc + n] = sum_k A[m,k] B[n,k]; RLCTR_EUNSUPPORTED when an operand is not TMA-addressable
int gemm(const Operand& A, const Operand& B, float* C, int64_t ldc, const float* bias, int M, int N, int K, int relu,
         bool allow_split, cudaStream_t st, const Epilogue* epi = nullptr, float* colsum_part = nullptr);
// colsum_part (wgrad form only: A MN-major, no bias): [splits][M] partial sums over k of A[k, m] -- the bias gradient,
// accumulated by the converter warps while they split the A tile; reduce over splits with the split-K reduction.

}  // namespace tma
}  // namespace rlctr
