// replay.cu -- device-side replay memory of the RL agents.
//
// The reference keeps its replay memories on the device but samples them on the HOST: DDQN / DDPG draw
// random.sample(range(n), batch) in Python (DDQN_model.py:183-185), the prioritized memories copy up to 1M priorities to the
// host, run np.random.choice(n, batch, p=P, replace=False) there and copy the indices back (v10_Hybrid_TD3_model_PER.py:62-85)
// -- two PCIe round trips and a host pass over the whole memory per learn step.  Here everything stays on the device and is
// stream-ordered (graph-capturable); randomness is the counter hash of common.cuh keyed by a device (seed, counter) pair.
//
//   store      ring-buffer write with the reference's wrap-around (Memory.add :44-60, DDQN store_transition :105-120)
//   uniform    `batch` DISTINCT indices of [0, n): the first `batch` images of a keyed pseudo-random PERMUTATION of [0, n)
//              (6-round Feistel network on ceil(log2 n) bits + cycle walking) -- random.sample semantics, O(batch) work
//   PER        weighted sampling WITHOUT replacement (np.random.choice(p=P, replace=False)): exponential clocks
//              t_i = -log(u_i) / w_i, w_i = (|td_i| + eps)^alpha, keep the `batch` smallest (Efraimidis-Spirakis: the same
//              distribution as drawing one by one and renormalising); cub radix sort of the n keys
//   greedy     the `batch` largest raw priorities (Memory.greedy_sample :87-105): the same sort on -priority
//   IS weight  (p_i / min_j p_j)^(-beta) over the valid range (:82, :103)
//   gather / update  memory[idx] rows and prioritys_[idx, 0] = td (:80, :108)
#include <cub/device/device_radix_sort.cuh>
#include <math.h>

#include "common.cuh"

namespace rlctr {

static inline size_t rp_align(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int rp_grid(int64_t n, int cap) {
    int64_t b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

// dst[(start + i) % size, :] = src[i, :]
__global__ void __launch_bounds__(256)
replay_store_kernel(float* __restrict__ mem, int64_t size, int width, int64_t start, const float* __restrict__ src, int64_t n,
                    int64_t ld_src) {
    const int64_t total = n * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / width;
        const int c = (int)(i - r * width);
        mem[((start + r) % size) * width + c] = __ldg(src + r * ld_src + c);
    }
}

// out[i, :] = mem[idx[i], :]
__global__ void __launch_bounds__(256)
replay_gather_kernel(const float* __restrict__ mem, int width, const int64_t* __restrict__ idx, int64_t n, float* __restrict__ out) {
    const int64_t total = n * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / width;
        const int c = (int)(i - r * width);
        out[i] = __ldg(mem + __ldg(idx + r) * width + c);
    }
}

__global__ void __launch_bounds__(256)
replay_update_kernel(float* __restrict__ prio, int ld, const int64_t* __restrict__ idx, const float* __restrict__ td, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        prio[__ldg(idx + i) * ld] = __ldg(td + i);
}

// ---- uniform sampling without replacement: keyed permutation of [0, n) -------------------------------------------------
__device__ __forceinline__ uint32_t feistel_perm(uint32_t x, int half_bits, uint32_t base) {
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (int round = 0; round < 6; ++round) {                                    // an independent round key per round and per call
        const uint32_t f = mix32(r ^ mix32(base + 0x9e3779b9U * (uint32_t)(round + 1))) & mask;
        const uint32_t nl = r;
        r = l ^ f;
        l = nl;
    }
    return (l << half_bits) | r;
}
__global__ void __launch_bounds__(256)
replay_uniform_kernel(int64_t n, int64_t batch, const uint64_t* __restrict__ rng, int64_t* __restrict__ out) {
    const uint64_t seed = rng[0], ctr = rng[1];
    int half_bits = 1;
    while (((int64_t)1 << (2 * half_bits)) < n) ++half_bits;           // domain 2^(2*half_bits) in [n, 4n)
    const uint32_t base = dropout_key(seed ^ (ctr >> 32), (uint32_t)ctr);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < batch; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i;
        do { x = feistel_perm(x, half_bits, base); } while ((int64_t)x >= n);       // cycle walking: stays a bijection on [0, n)
        out[i] = (int64_t)x;
    }
}

// ---- prioritized sampling ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float per_priority(float td, float eps, float alpha) { return powf(fabsf(td) + eps, alpha); }

// keys[i] = exponential clock of slot i (mode 0) or -raw priority (mode 1, greedy); vals[i] = i
__global__ void __launch_bounds__(256)
replay_keys_kernel(const float* __restrict__ prio, int ld, int64_t n, float eps, float alpha, const uint64_t* __restrict__ rng,
                   int mode, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint64_t seed = 0, ctr = 0;
    if (mode == 0) { seed = rng[0]; ctr = rng[1]; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float p = __ldg(prio + i * ld);
        float key;
        if (mode == 0) {
            const uint64_t idx = ctr + (uint64_t)i;
            const uint32_t h = mix32((uint32_t)idx ^ dropout_key(seed, (uint32_t)(idx >> 32)));
            // Exp(1) clock from all 32 hash bits: E = -log1p(-u), u in (0, 1).  The draw is decided among the SMALLEST keys,
            // i.e. u near 0, where a float resolves u to 2^-32 (a 24-bit uniform fed to -log(u) had its winners at u near 1,
            // spaced 6e-8 apart: ~0.25 slots per level at memory_size 4,096,000 -- ties broken by slot index)
            const float u = fminf(((float)h + 0.5f) * (1.0f / 4294967296.0f), 0.99999994f);
            key = -log1pf(-u) / per_priority(p, eps, alpha);                            // > 0: float bits sort like the value
        } else {
            key = -p;
        }
        uint32_t b = __float_as_uint(key);
        b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);                                 // total order of floats as unsigned
        keys[i] = b;
        vals[i] = (uint32_t)i;
    }
}

// min over the valid range of the priority used for the IS weights (mode 0: get_priority(td); mode 1: raw), fixed tree
__global__ void __launch_bounds__(256)
replay_min_partial_kernel(const float* __restrict__ prio, int ld, int64_t n, float eps, float alpha, int mode, float* __restrict__ part) {
    __shared__ float red[256];
    float m = INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float p = __ldg(prio + i * ld);
        m = fminf(m, mode == 0 ? per_priority(p, eps, alpha) : p);
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fminf(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256)
replay_finish_kernel(const uint32_t* __restrict__ sorted_vals, const float* __restrict__ prio, int ld, float eps, float alpha,
                     float beta, int mode, const float* __restrict__ part, int parts, int64_t batch, int64_t* __restrict__ out_idx,
                     float* __restrict__ out_isw) {
    float mn = INFINITY;
    for (int q = 0; q < parts; ++q) mn = fminf(mn, __ldg(part + q));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < batch; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t j = __ldg(sorted_vals + i);
        out_idx[i] = (int64_t)j;
        if (out_isw) {
            const float p = __ldg(prio + (int64_t)j * ld);
            out_isw[i] = powf((mode == 0 ? per_priority(p, eps, alpha) : p) / mn, -beta);
        }
    }
}

constexpr int RP_MIN_PARTS = 296;

}  // namespace rlctr

using namespace rlctr;

extern "C" int rlctr_replay_store(float* memory, int64_t memory_size, int32_t width, int64_t counter, const float* src, int64_t n,
                                  int64_t ld_src, rlctr_stream_t stream) {
    if (!memory || !src || memory_size <= 0 || width <= 0 || counter < 0 || n < 0 || ld_src < width) return RLCTR_EINVAL;
    if (n > memory_size) return RLCTR_EUNSUPPORTED;             // the reference's slicing breaks there too
    if (n == 0) return RLCTR_OK;
    replay_store_kernel<<<rp_grid(n * width, RLCTR_SMS * 8), 256, 0, (cudaStream_t)stream>>>(memory, memory_size, width,
                                                                                           counter % memory_size, src, n, ld_src);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_replay_gather(const float* memory, int32_t width, const int64_t* idx, int64_t n, float* out, rlctr_stream_t stream) {
    if (!memory || !idx || !out || width <= 0 || n < 0) return RLCTR_EINVAL;
    if (n == 0) return RLCTR_OK;
    replay_gather_kernel<<<rp_grid(n * width, RLCTR_SMS * 8), 256, 0, (cudaStream_t)stream>>>(memory, width, idx, n, out);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_replay_update(float* priorities, int32_t ld, const int64_t* idx, const float* td, int64_t n, rlctr_stream_t stream) {
    if (!priorities || !idx || !td || ld <= 0 || n < 0) return RLCTR_EINVAL;
    if (n == 0) return RLCTR_OK;
    replay_update_kernel<<<rp_grid(n, RLCTR_SMS * 4), 256, 0, (cudaStream_t)stream>>>(priorities, ld, idx, td, n);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_replay_sample_uniform(int64_t n_valid, int64_t batch, const uint64_t* rng_state, int64_t* out_idx,
                                           rlctr_stream_t stream) {
    if (!rng_state || !out_idx || n_valid <= 0 || batch < 0) return RLCTR_EINVAL;
    if (batch > n_valid) return RLCTR_EINVAL;                   // random.sample raises ValueError
    if (n_valid > ((int64_t)1 << 30)) return RLCTR_EUNSUPPORTED;
    if (batch == 0) return RLCTR_OK;
    replay_uniform_kernel<<<rp_grid(batch, RLCTR_SMS * 4), 256, 0, (cudaStream_t)stream>>>(n_valid, batch, rng_state, out_idx);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" size_t rlctr_replay_per_ws_bytes(int64_t n_valid) {
    if (n_valid <= 0) return 256;
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, n_valid, 0, 32);
    return 4 * rp_align((size_t)n_valid * 4) + rp_align(RP_MIN_PARTS * 4) + temp + 256;
}

extern "C" int rlctr_replay_sample_per(const float* priorities, int32_t ld, int64_t n_valid, float eps, float alpha, float beta,
                                       int32_t greedy, int64_t batch, const uint64_t* rng_state, int64_t* out_idx, float* out_isw,
                                       void* ws, size_t ws_bytes, rlctr_stream_t stream) {
    if (!priorities || !out_idx || ld <= 0 || n_valid <= 0 || batch < 0 || batch > n_valid) return RLCTR_EINVAL;
    if (!greedy && !rng_state) return RLCTR_EINVAL;
    if (n_valid >= ((int64_t)1 << 31)) return RLCTR_EUNSUPPORTED;
    if (!ws || ws_bytes < rlctr_replay_per_ws_bytes(n_valid)) return RLCTR_EWORKSPACE;
    if (batch == 0) return RLCTR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t arr = rp_align((size_t)n_valid * 4);
    char* base = reinterpret_cast<char*>(ws);
    uint32_t* keys = reinterpret_cast<uint32_t*>(base);
    uint32_t* vals = reinterpret_cast<uint32_t*>(base + arr);
    uint32_t* skeys = reinterpret_cast<uint32_t*>(base + 2 * arr);
    uint32_t* svals = reinterpret_cast<uint32_t*>(base + 3 * arr);
    float* part = reinterpret_cast<float*>(base + 4 * arr);
    void* temp = base + 4 * arr + rp_align(RP_MIN_PARTS * 4);
    size_t temp_bytes = ws_bytes - (4 * arr + rp_align(RP_MIN_PARTS * 4));
    const int mode = greedy ? 1 : 0;
    replay_keys_kernel<<<rp_grid(n_valid, RLCTR_SMS * 8), 256, 0, st>>>(priorities, ld, n_valid, eps, alpha, rng_state, mode, keys, vals);
    RLCTR_LAUNCH_CHECK();
    const int parts = rp_grid(n_valid, RP_MIN_PARTS);
    replay_min_partial_kernel<<<parts, 256, 0, st>>>(priorities, ld, n_valid, eps, alpha, mode, part);
    RLCTR_LAUNCH_CHECK();
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, skeys, vals, svals, (int)n_valid, 0, 32, st);
    if (e != cudaSuccess) return (int)e;
    RLCTR_COUNT_LAUNCH(6);
    replay_finish_kernel<<<rp_grid(batch, RLCTR_SMS * 2), 256, 0, st>>>(svals, priorities, ld, eps, alpha, beta, mode, part, parts, batch,
                                                                       out_idx, out_isw);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Generalised advantage estimate of the PPO agent (Hybrid_PPO_model.py:206-212).  The reference walks the TD residuals in
// REVERSE order in a Python loop with one `.item()` host synchronisation per sample and a Python-float (fp64) accumulator:
//     adv = 0;  for i, d in enumerate(reversed(deltas)):  adv = c * adv + d;  advantages[i] = adv          (c = gamma * lambda)
// (so advantages[i] belongs to sample n-1-i: kept as written).  A first-order linear recurrence is an associative scan over
// pairs (m, v) with (m1, v1) o (m2, v2) = (m1 * m2, v1 * m2 + v2); it runs as one cub::DeviceScan in fp64.
// ------------------------------------------------------------------------------------------------------------------------
#include <cub/device/device_scan.cuh>

namespace rlctr {

struct GaePair { double m, v; };
struct GaeOp {
    __device__ __forceinline__ GaePair operator()(const GaePair& a, const GaePair& b) const { return GaePair{a.m * b.m, a.v * b.m + b.v}; }
};
__global__ void __launch_bounds__(256)
gae_pairs_kernel(const float* __restrict__ deltas, int64_t n, double c, GaePair* __restrict__ pairs) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        pairs[i] = GaePair{c, (double)__ldg(deltas + (n - 1 - i))};
}
__global__ void __launch_bounds__(256)
gae_out_kernel(const GaePair* __restrict__ pairs, int64_t n, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)pairs[i].v;
}

}  // namespace rlctr

extern "C" size_t rlctr_gae_ws_bytes(int64_t n) {
    if (n <= 0) return 256;
    size_t temp = 0;
    cub::DeviceScan::InclusiveScan(nullptr, temp, (const GaePair*)nullptr, (GaePair*)nullptr, GaeOp{}, (int)n);
    return rp_align((size_t)n * sizeof(GaePair)) + temp + 256;
}

extern "C" int rlctr_gae_scan(const float* deltas, int64_t n, double gamma_lambda, float* advantages, void* ws, size_t ws_bytes,
                              rlctr_stream_t stream) {
    if (!deltas || !advantages || n < 0) return RLCTR_EINVAL;
    if (n == 0) return RLCTR_OK;
    if (n >= ((int64_t)1 << 31)) return RLCTR_EUNSUPPORTED;
    if (!ws || ws_bytes < rlctr_gae_ws_bytes(n)) return RLCTR_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    GaePair* pairs = reinterpret_cast<GaePair*>(ws);
    const size_t arr = rp_align((size_t)n * sizeof(GaePair));
    void* temp = reinterpret_cast<char*>(ws) + arr;
    size_t temp_bytes = ws_bytes - arr;
    gae_pairs_kernel<<<rp_grid(n, RLCTR_SMS * 8), 256, 0, st>>>(deltas, n, gamma_lambda, pairs);
    RLCTR_LAUNCH_CHECK();
    cudaError_t e = cub::DeviceScan::InclusiveScan(temp, temp_bytes, pairs, pairs, GaeOp{}, (int)n, st);
    if (e != cudaSuccess) return (int)e;
    RLCTR_COUNT_LAUNCH(2);
    gae_out_kernel<<<rp_grid(n, RLCTR_SMS * 8), 256, 0, st>>>(pairs, n, advantages);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}
