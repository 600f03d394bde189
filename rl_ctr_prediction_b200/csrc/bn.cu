// bn.cu -- BatchNorm1d in training mode fused with the ReLU that follows it: the policy nets' hidden layers
// (Linear -> BatchNorm1d -> ReLU, src/models/DDQN_model.py:32-46, DDPG_for_PG_model.py:27-40) in their learn steps.
//
// The reference runs torch's native_batch_norm + relu (and their three backward kernels) between every pair of GEMMs.  Here
// one kernel does the forward -- batch mean, biased variance, normalisation, affine map, ReLU, running statistics with torch's
// unbiased variance and momentum -- and one the backward (ReLU mask, the two column sums of batch_norm_backward, dx, dgamma,
// dbeta).  A block owns 32 columns (a warp reads 128-byte row segments) and walks ALL rows of the batch: the learn steps run on
// replay batches of a few hundred rows whose activations sit in L2, so the statistics are taken in two exact passes (mean,
// then squared deviations: no E[x^2] - E[x]^2 cancellation) in a fixed order -- bit-identical from run to run.
// The acting path (eval mode) never gets here: its BatchNorm is folded into the GEMM (mlp.Tower._forward_modules).
#include "common.cuh"

namespace rlctr {

constexpr int BN_WARPS = 8;

// sum over the block's warps of one value per (warp, lane = column): fixed order, every thread gets the total of its column
__device__ __forceinline__ float bn_block_sum(float v, float (*red)[32], int warp, int lane) {
    red[warp][lane] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < BN_WARPS; ++w) s += red[w][lane];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(BN_WARPS * 32)
bn_relu_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                   float* __restrict__ y, int64_t ldy, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                   int64_t batch, int n, int relu) {
    __shared__ float red[BN_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 32 + lane;
    const bool on = col < n;
    float s = 0.f;
    for (int64_t r = warp; r < batch; r += BN_WARPS) s += on ? __ldg(x + r * ldx + col) : 0.f;
    const float mean = bn_block_sum(s, red, warp, lane) / (float)batch;
    float q = 0.f;
    for (int64_t r = warp; r < batch; r += BN_WARPS) {
        const float d = on ? __ldg(x + r * ldx + col) - mean : 0.f;
        q = fmaf(d, d, q);
    }
    const float ssd = bn_block_sum(q, red, warp, lane);
    const float var = ssd / (float)batch;                        // biased: what the normalisation uses
    const float invstd = rsqrtf(var + eps);
    const float g = (on && gamma) ? __ldg(gamma + col) : 1.f, b = (on && beta) ? __ldg(beta + col) : 0.f;
    for (int64_t r = warp; r < batch; r += BN_WARPS) {
        if (!on) continue;
        float v = (__ldg(x + r * ldx + col) - mean) * invstd * g + b;
        if (relu) v = fmaxf(v, 0.f);
        y[r * ldy + col] = v;
    }
    if (warp == 0 && on) {
        if (save_mean) save_mean[col] = mean;
        if (save_invstd) save_invstd[col] = invstd;
        if (running_mean) running_mean[col] = (1.f - momentum) * running_mean[col] + momentum * mean;
        if (running_var) {
            const float unbiased = batch > 1 ? ssd / (float)(batch - 1) : var;
            running_var[col] = (1.f - momentum) * running_var[col] + momentum * unbiased;
        }
    }
}

// dy_r = gy * (y > 0) when relu;  dbeta = sum dy_r;  dgamma = invstd * sum dy_r (x - mean);
// dx = gamma * invstd * (dy_r - dbeta / B - (x - mean) * invstd^2 * sum dy_r (x - mean) / B)      (torch's batch_norm_backward)
__global__ void __launch_bounds__(BN_WARPS * 32)
bn_relu_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ y, int64_t ldy, const float* __restrict__ gy,
                   int64_t ldg, const float* __restrict__ gamma, const float* __restrict__ save_mean,
                   const float* __restrict__ save_invstd, float* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, int64_t batch, int n, int relu) {
    __shared__ float red[BN_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 32 + lane;
    const bool on = col < n;
    const float mean = on ? __ldg(save_mean + col) : 0.f, invstd = on ? __ldg(save_invstd + col) : 0.f;
    float s1 = 0.f, s2 = 0.f;
    for (int64_t r = warp; r < batch; r += BN_WARPS) {
        if (!on) continue;
        float d = __ldg(gy + r * ldg + col);
        if (relu && !(__ldg(y + r * ldy + col) > 0.f)) d = 0.f;
        s1 += d;
        s2 = fmaf(d, __ldg(x + r * ldx + col) - mean, s2);
    }
    const float sum_dy = bn_block_sum(s1, red, warp, lane);
    const float sum_dy_xmu = bn_block_sum(s2, red, warp, lane);
    const float g = (on && gamma) ? __ldg(gamma + col) : 1.f;
    if (dx) {
        const float inv_b = 1.f / (float)batch;
        const float k = sum_dy_xmu * invstd * invstd * inv_b, m1 = sum_dy * inv_b, scale = g * invstd;
        for (int64_t r = warp; r < batch; r += BN_WARPS) {
            if (!on) continue;
            float d = __ldg(gy + r * ldg + col);
            if (relu && !(__ldg(y + r * ldy + col) > 0.f)) d = 0.f;
            dx[r * lddx + col] = (d - m1 - (__ldg(x + r * ldx + col) - mean) * k) * scale;
        }
    }
    if (warp == 0 && on) {
        if (dgamma) dgamma[col] = sum_dy_xmu * invstd;
        if (dbeta) dbeta[col] = sum_dy;
    }
}

}  // namespace rlctr

using namespace rlctr;

extern "C" int rlctr_bn_relu_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, float momentum, float eps, float* y, int64_t ldy, float* save_mean,
                                 float* save_invstd, int64_t batch, int32_t n, int32_t relu, rlctr_stream_t stream) {
    if (!x || !y || !save_mean || !save_invstd || batch <= 0 || n <= 0 || ldx < n || ldy < n) return RLCTR_EINVAL;
    if (batch > (1 << 16)) return RLCTR_EUNSUPPORTED;            // a block walks all rows of its 32 columns: replay-batch sizes
    bn_relu_fwd_kernel<<<(n + 31) / 32, BN_WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, gamma, beta, running_mean, running_var,
                                                                                 momentum, eps, y, ldy, save_mean, save_invstd,
                                                                                 batch, n, relu ? 1 : 0);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_bn_relu_bwd(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* gy, int64_t ldg,
                                 const float* gamma, const float* save_mean, const float* save_invstd, float* dx, int64_t lddx,
                                 float* dgamma, float* dbeta, int64_t batch, int32_t n, int32_t relu, rlctr_stream_t stream) {
    if (!x || !gy || !save_mean || !save_invstd || batch <= 0 || n <= 0 || ldx < n || ldg < n) return RLCTR_EINVAL;
    if (relu && (!y || ldy < n)) return RLCTR_EINVAL;
    if (dx && lddx < n) return RLCTR_EINVAL;
    if (batch > (1 << 16)) return RLCTR_EUNSUPPORTED;
    bn_relu_bwd_kernel<<<(n + 31) / 32, BN_WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, gy, ldg, gamma, save_mean,
                                                                                 save_invstd, dx, lddx, dgamma, dbeta, batch, n,
                                                                                 relu ? 1 : 0);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}
