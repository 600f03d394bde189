// metrics.cu -- AUC and log-loss of a prediction vector on the device.
//
// The reference's test() / submission() copy every batch's predictions to the host with .tolist() and call
// sklearn.metrics.roc_auc_score (src/main/pretrain_main.py:110-139).  Here:
//     AUC = sum over tie groups g of pos_g * (neg_before_g + 0.5 * neg_g) / (n_pos * n_neg)
// (the Mann-Whitney statistic with average ranks for ties == sklearn's trapezoidal ROC area): sort the scores
// (cub radix sort of order-preserving keys, labels as values), exclusive-scan the negatives, and let every tie-group
// head find the end of its group by bisection.  Partial sums are exact integers held in doubles (< 2^53) and are
// combined in a fixed order, so the result is bit-identical from run to run and independent of the grid.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace rlctr {

constexpr int AUC_BLOCKS = 592;            // 4 per SM

// float -> uint32 whose unsigned order is the float order (-0 == +0; NaN sorts last)
__device__ __forceinline__ uint32_t order_key(float x) {
    if (x == 0.f) x = 0.f;
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256)
auc_prep_kernel(const float* __restrict__ pred, const int64_t* __restrict__ yi, const float* __restrict__ yf, int64_t n,
                uint32_t* __restrict__ keys, uint32_t* __restrict__ labs) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = order_key(__ldg(pred + i));
        const bool pos = yi ? (__ldg(yi + i) != 0) : (__ldg(yf + i) != 0.f);
        labs[i] = pos ? 0u : 1u;                      // 1 = NEGATIVE (what the scan counts)
    }
}
// cneg[i] = negatives among sorted positions < i (exclusive scan of labs); cneg[n] is not stored: total passed in
__global__ void __launch_bounds__(256)
auc_groups_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ negs, const uint32_t* __restrict__ cneg,
                  int64_t n, double* __restrict__ partial) {
    __shared__ double red[256];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = __ldg(keys + i);
        if (i > 0 && __ldg(keys + i - 1) == k) continue;               // not a tie-group head
        int64_t lo = i, hi = n;                                        // keys[lo] == k, keys[hi] > k (or hi == n)
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(keys + mid) == k) lo = mid; else hi = mid;
        }
        const int64_t ge = hi;                                         // group = [i, ge)
        const uint32_t neg_before = __ldg(cneg + i);
        const uint32_t neg_end = (ge < n) ? __ldg(cneg + ge) : (__ldg(cneg + n - 1) + __ldg(negs + n - 1));
        const double neg_g = (double)(neg_end - neg_before);
        const double pos_g = (double)(ge - i) - neg_g;
        acc += pos_g * ((double)neg_before + 0.5 * neg_g);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int half = 128; half > 0; half >>= 1) {
        if (threadIdx.x < half) red[threadIdx.x] += red[threadIdx.x + half];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
// log-loss partials: torch BCELoss per element (clamped logs), summed in double
__global__ void __launch_bounds__(256)
logloss_kernel(const float* __restrict__ pred, const int64_t* __restrict__ yi, const float* __restrict__ yf, int64_t n,
               double* __restrict__ partial) {
    __shared__ double red[256];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float p = __ldg(pred + i);
        const float y = yi ? (__ldg(yi + i) != 0 ? 1.f : 0.f) : __ldg(yf + i);
        acc += (double)((y - 1.0f) * fmaxf(log1pf(-p), -100.0f) - y * fmaxf(logf(p), -100.0f));
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int half = 128; half > 0; half >>= 1) {
        if (threadIdx.x < half) red[threadIdx.x] += red[threadIdx.x + half];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void auc_finish_kernel(const double* __restrict__ p_auc, const double* __restrict__ p_ll, int blocks,
                                  const uint32_t* __restrict__ cneg, const uint32_t* __restrict__ negs, int64_t n,
                                  float* __restrict__ out) {
    double a = 0.0, l = 0.0;
    for (int b = 0; b < blocks; ++b) { a += p_auc[b]; l += p_ll[b]; }      // fixed order
    const double n_neg = (double)(cneg[n - 1] + negs[n - 1]);
    const double n_pos = (double)n - n_neg;
    out[0] = (n_pos > 0.0 && n_neg > 0.0) ? (float)(a / (n_pos * n_neg)) : __int_as_float(0x7fc00000);   // NaN: one class only
    out[1] = (float)(l / (double)n);
}

static size_t auc_layout(int64_t n, size_t* o_keys, size_t* o_labs, size_t* o_skeys, size_t* o_slabs, size_t* o_cneg,
                         size_t* o_part, size_t* o_temp, size_t* temp_bytes) {
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t arr = up((size_t)n * sizeof(uint32_t));
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, n);
    cub::DeviceScan::ExclusiveSum(nullptr, t2, (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
    *temp_bytes = up(t1 > t2 ? t1 : t2);
    size_t off = 0;
    *o_keys = off; off += arr;
    *o_labs = off; off += arr;
    *o_skeys = off; off += arr;
    *o_slabs = off; off += arr;
    *o_cneg = off; off += arr;
    *o_part = off; off += up(2 * AUC_BLOCKS * sizeof(double));
    *o_temp = off; off += *temp_bytes;
    return off + 256;
}

}  // namespace rlctr

using namespace rlctr;

extern "C" size_t rlctr_auc_ws_bytes(int64_t n) {
    if (n <= 0) return 256;
    size_t a, b, c, d, e, f, g, t;
    return auc_layout(n, &a, &b, &c, &d, &e, &f, &g, &t);
}

extern "C" int rlctr_auc_logloss(const float* pred, const int64_t* labels_i64, const float* labels_f32, int64_t n, float* out,
                                 void* ws, size_t ws_bytes, rlctr_stream_t stream) {
    if (!pred || (!labels_i64 && !labels_f32) || !out || !ws || n <= 0) return RLCTR_EINVAL;
    if (n >= ((int64_t)1 << 31)) return RLCTR_EUNSUPPORTED;
    size_t o_keys, o_labs, o_skeys, o_slabs, o_cneg, o_part, o_temp, temp_bytes;
    if (ws_bytes < auc_layout(n, &o_keys, &o_labs, &o_skeys, &o_slabs, &o_cneg, &o_part, &o_temp, &temp_bytes)) return RLCTR_EWORKSPACE;
    if (!rlctr_aligned16(ws)) return RLCTR_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = reinterpret_cast<char*>(ws);
    uint32_t* keys = reinterpret_cast<uint32_t*>(base + o_keys);
    uint32_t* labs = reinterpret_cast<uint32_t*>(base + o_labs);
    uint32_t* skeys = reinterpret_cast<uint32_t*>(base + o_skeys);
    uint32_t* slabs = reinterpret_cast<uint32_t*>(base + o_slabs);
    uint32_t* cneg = reinterpret_cast<uint32_t*>(base + o_cneg);
    double* part = reinterpret_cast<double*>(base + o_part);
    void* temp = base + o_temp;
    int64_t blocks = (n + 255) / 256;
    const int grid = (int)(blocks < AUC_BLOCKS ? blocks : AUC_BLOCKS);
    auc_prep_kernel<<<grid, 256, 0, st>>>(pred, labels_i64, labels_f32, n, keys, labs);
    RLCTR_LAUNCH_CHECK();
    size_t tb = temp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tb, keys, skeys, labs, slabs, n, 0, 32, st);
    if (e != cudaSuccess) return (int)e;
    tb = temp_bytes;
    e = cub::DeviceScan::ExclusiveSum(temp, tb, slabs, cneg, n, st);
    if (e != cudaSuccess) return (int)e;
    RLCTR_COUNT_LAUNCH(8);
    auc_groups_kernel<<<grid, 256, 0, st>>>(skeys, slabs, cneg, n, part);
    RLCTR_LAUNCH_CHECK();
    logloss_kernel<<<grid, 256, 0, st>>>(pred, labels_i64, labels_f32, n, part + AUC_BLOCKS);
    RLCTR_LAUNCH_CHECK();
    auc_finish_kernel<<<1, 1, 0, st>>>(part, part + AUC_BLOCKS, grid, cneg, slabs, n, out);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}
