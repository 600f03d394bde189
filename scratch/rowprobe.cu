// rowprobe.cu -- what bounds random row access on a B200?  (scratch; standalone: nvcc -o scratch/bin/rowprobe)
//
//   rowprobe <l2_fetch_granularity: 0 = leave default | 32 | 64 | 128> [n_rows_millions=10] [quick]
//
// Every test touches 983,040 rows (B = 65536 x F = 15) of a table of n_rows records, L2 flushed between runs,
// median of 9.  Output: one line per test:  name, us, G rows/s, logical GB/s.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ld_nc(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_na(const float* p) {          // no L1 allocation
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_ef(const float* p) {          // L2 prefetch-size hint 64B (the smallest)
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
struct f8 { float4 a, b; };
__device__ __forceinline__ f8 ld_256(const float* p) {             // Blackwell 256-bit load
    f8 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w) : "l"(p));
    return v;
}
template <int LPR, int RIF>
__global__ void __launch_bounds__(256) gather256_kernel(const uint32_t* __restrict__ ids, int64_t n, const float* __restrict__ tab, int pitch,
                                                        float* __restrict__ out) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t grp = gt / LPR, ngrp = (int64_t)gridDim.x * blockDim.x / LPR;
    const int c = (int)(gt % LPR);
    float acc = 0.f;
    for (int64_t k = grp * RIF; k < n; k += ngrp * RIF) {
        uint32_t id[RIF];
#pragma unroll
        for (int r = 0; r < RIF; ++r) id[r] = (k + r < n) ? __ldg(ids + k + r) : 0;
        f8 v[RIF];
#pragma unroll
        for (int r = 0; r < RIF; ++r) v[r] = ld_256(tab + (int64_t)id[r] * pitch + 8 * c);
#pragma unroll
        for (int r = 0; r < RIF; ++r) acc += v[r].a.x + v[r].a.w + v[r].b.y + v[r].b.w;
    }
    if (acc == 123.456f) out[gt] = acc;
}
template <int MODE> __device__ __forceinline__ float4 ldm(const float* p) {
    if (MODE == 1) return ld_na(p);
    if (MODE == 2) return ld_ef(p);
    return ld_nc(p);
}
__device__ __forceinline__ float sum4(float4 a) { return a.x + a.y + a.z + a.w; }

// ---- read-only gather: LPR lanes per row (16 B each), RIF rows in flight per lane group ---------------------------------
template <int LPR, int RIF, int MODE>
__global__ void __launch_bounds__(256) gather_kernel(const uint32_t* __restrict__ ids, int64_t n, const float* __restrict__ tab, int pitch,
                                                     float* __restrict__ out) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t grp = gt / LPR, ngrp = (int64_t)gridDim.x * blockDim.x / LPR;
    const int c = (int)(gt % LPR);
    float acc = 0.f;
    for (int64_t k = grp * RIF; k < n; k += ngrp * RIF) {
        uint32_t id[RIF];
#pragma unroll
        for (int r = 0; r < RIF; ++r) id[r] = (k + r < n) ? __ldg(ids + k + r) : 0;
        float4 v[RIF];
#pragma unroll
        for (int r = 0; r < RIF; ++r) v[r] = ldm<MODE>(tab + (int64_t)id[r] * pitch + 4 * c);
#pragma unroll
        for (int r = 0; r < RIF; ++r) acc += sum4(v[r]);
    }
    if (acc == 123.456f) out[gt] = acc;
}

// ---- read-modify-write of a record of CH chunks (16 B each) per lane; LPR lanes per record -----------------------------
// record = LPR*CH float4; lane c owns chunks c, c+LPR, ...  (so one warp-level request covers contiguous 16 B * LPR)
template <int LPR, int CH, int RIF>
__global__ void __launch_bounds__(256) rmw_kernel(const uint32_t* __restrict__ ids, int64_t n, float* __restrict__ tab, int pitch) {
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t grp = gt / LPR, ngrp = (int64_t)gridDim.x * blockDim.x / LPR;
    const int c = (int)(gt % LPR);
    for (int64_t k = grp * RIF; k < n; k += ngrp * RIF) {
        uint32_t id[RIF];
#pragma unroll
        for (int r = 0; r < RIF; ++r) id[r] = (k + r < n) ? __ldg(ids + k + r) : 0xffffffffu;
        float4 v[RIF][CH];
#pragma unroll
        for (int r = 0; r < RIF; ++r)
#pragma unroll
            for (int h = 0; h < CH; ++h)
                if (id[r] != 0xffffffffu) v[r][h] = *reinterpret_cast<const float4*>(tab + (int64_t)id[r] * pitch + 4 * (c + LPR * h));
#pragma unroll
        for (int r = 0; r < RIF; ++r)
#pragma unroll
            for (int h = 0; h < CH; ++h)
                if (id[r] != 0xffffffffu) {
                    float4 x = v[r][h];
                    x.x = x.x * 0.999f + 1e-3f; x.y = x.y * 0.999f + 1e-3f; x.z = x.z * 0.999f + 1e-3f; x.w = x.w * 0.999f + 1e-3f;
                    *reinterpret_cast<float4*>(tab + (int64_t)id[r] * pitch + 4 * (c + LPR * h)) = x;
                }
    }
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

struct Ctx {
    float* flush; int64_t flush_n;
    cudaEvent_t e0, e1;
};
static int g_reps = 9;
template <class F> static float timed(Ctx& c, F&& launch) {
    const int reps = g_reps;
    std::vector<float> ts;
    for (int i = 0; i < reps + 2; ++i) {
        fill_kernel<<<148 * 8, 256>>>(c.flush, c.flush_n, (float)i);
        CK(cudaEventRecord(c.e0));
        launch();
        CK(cudaEventRecord(c.e1));
        CK(cudaEventSynchronize(c.e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
        if (i >= 2) ts.push_back(ms * 1e3f);
    }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}
static void report(const char* name, float us, int64_t n, int logical_bytes) {
    printf("%-44s %8.1f us  %6.2f G rows/s  %8.1f GB/s logical\n", name, us, n / us * 1e-3, (double)n * logical_bytes / us * 1e-3);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const int gran = argc > 1 ? atoi(argv[1]) : 0;
    const int64_t n_rows = (int64_t)((argc > 2 ? atof(argv[2]) : 10.0) * 1e6);
    const bool quick = argc > 3;
    if (quick) g_reps = 1;
    if (gran > 0) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
        printf("cudaDeviceSetLimit(MaxL2FetchGranularity, %d) -> %s\n", gran, cudaGetErrorString(e));
    }
    CK(cudaFree(0));
    size_t lim = 0;
    CK(cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity));
    printf("MaxL2FetchGranularity = %zu, n_rows = %lld\n", lim, (long long)n_rows);
    const int64_t n = 983040;
    const bool rec384 = argc > 3 && std::string(argv[3]) == "rec384";   // the co-located record: 3 lines per id
    const int pitch = rec384 ? 96 : 64;                              // floats: 256 B per record slot (all layouts fit inside)
    float* tab; CK(cudaMalloc(&tab, (size_t)n_rows * pitch * 4));
    fill_kernel<<<148 * 8, 256>>>(tab, n_rows * pitch, 1.0f);
    Ctx c; c.flush_n = 96 << 20; CK(cudaMalloc(&c.flush, c.flush_n * 4));
    CK(cudaEventCreate(&c.e0)); CK(cudaEventCreate(&c.e1));
    float* out; CK(cudaMalloc(&out, 64 << 20));
    std::mt19937_64 rng(1);
    std::vector<uint32_t> h(n);
    for (auto& x : h) x = (uint32_t)(rng() % (uint64_t)n_rows);
    uint32_t *ids_r, *ids_s;
    CK(cudaMalloc(&ids_r, n * 4)); CK(cudaMalloc(&ids_s, n * 4));
    CK(cudaMemcpy(ids_r, h.data(), n * 4, cudaMemcpyHostToDevice));
    std::sort(h.begin(), h.end());
    h.erase(std::unique(h.begin(), h.end()), h.end());
    const int64_t nu = (int64_t)h.size();
    CK(cudaMemcpy(ids_s, h.data(), nu * 4, cudaMemcpyHostToDevice));
    printf("unique sorted ids: %lld\n", (long long)nu);

    auto grid_for = [](int64_t threads) { int64_t b = (threads + 255) / 256; return (int)std::min<int64_t>(b, 148 * 64); };
    char name[128];
#define GATHER(LPR, RIF, MODE, IDS, NN, P, TAG)                                                                       \
    { float us = timed(c, [&] { gather_kernel<LPR, RIF, MODE><<<grid_for((NN) * LPR / RIF), 256>>>(IDS, NN, tab, P, out); }); \
      snprintf(name, sizeof name, "gather %3dB pitch %3dB rif%d mode%d %s", LPR * 16, P * 4, RIF, MODE, TAG); report(name, us, NN, LPR * 16); }
    if (rec384) {
        // what the co-located step does to its table, stripped of everything else: the gather reads the first line of a 384-byte
        // record per occurrence; catch-up and update each read and write the record's three lines once per distinct id
        GATHER(8, 1, 0, ids_r, n, 96, "random (group gather pattern)");
        GATHER(8, 2, 0, ids_r, n, 96, "random (group gather pattern)");
        GATHER(8, 4, 0, ids_r, n, 96, "random (group gather pattern)");
        { float us = timed(c, [&] { rmw_kernel<8, 3, 1><<<grid_for(nu * 8), 256>>>(ids_s, nu, tab, 96); });
          report("rmw 384B pitch 384B lpr8 ch3 rif1 sorted (catch-up / update pattern)", us, nu, 2 * 384); }
        { float us = timed(c, [&] { rmw_kernel<8, 3, 2><<<grid_for(nu * 8 / 2), 256>>>(ids_s, nu, tab, 96); });
          report("rmw 384B pitch 384B lpr8 ch3 rif2 sorted (catch-up / update pattern)", us, nu, 2 * 384); }
        { float us = timed(c, [&] { rmw_kernel<8, 3, 4><<<grid_for(nu * 8 / 4), 256>>>(ids_s, nu, tab, 96); });
          report("rmw 384B pitch 384B lpr8 ch3 rif4 sorted (catch-up / update pattern)", us, nu, 2 * 384); }
        printf("done\n");
        return 0;
    }
    // 64 B rows at 64 B / 192 B / 256 B pitch (the table slot is 256 B: smaller pitches use a prefix of the allocation)
    GATHER(4, 1, 0, ids_r, n, 16, "random");
    GATHER(4, 1, 0, ids_r, n, 48, "random");
    GATHER(4, 1, 0, ids_r, n, 64, "random");
    GATHER(4, 2, 0, ids_r, n, 48, "random");
    GATHER(4, 4, 0, ids_r, n, 48, "random");
    GATHER(4, 2, 1, ids_r, n, 48, "random");
    GATHER(4, 2, 2, ids_r, n, 48, "random");
    GATHER(4, 2, 0, ids_s, nu, 48, "sorted");
    GATHER(4, 4, 0, ids_s, nu, 48, "sorted");
    GATHER(2, 2, 0, ids_r, n, 48, "random");           // 32 B rows
    GATHER(2, 4, 0, ids_r, n, 8, "random");            // 32 B rows, 32 B pitch
    GATHER(1, 4, 0, ids_r, n, 4, "random");            // 16 B records, 16 B pitch (LR-like [w|m|v|stamp])
    GATHER(8, 1, 0, ids_r, n, 32, "random");           // 128 B rows at 128 B pitch (one line)
    GATHER(8, 2, 0, ids_r, n, 32, "random");
    GATHER(8, 2, 0, ids_s, nu, 32, "sorted");
#define GATHER256(LPR, RIF, IDS, NN, P, TAG)                                                                          \
    { float us = timed(c, [&] { gather256_kernel<LPR, RIF><<<grid_for((NN) * LPR / RIF), 256>>>(IDS, NN, tab, P, out); }); \
      snprintf(name, sizeof name, "gather256 %3dB pitch %3dB rif%d %s", LPR * 32, P * 4, RIF, TAG); report(name, us, NN, LPR * 32); }
    GATHER256(2, 2, ids_r, n, 48, "random");           // 64 B rows by two 256-bit loads
    GATHER256(2, 4, ids_r, n, 48, "random");
    GATHER256(4, 2, ids_r, n, 32, "random");           // 128 B rows
    if (!quick) {
        GATHER(8, 2, 0, ids_r, n, 48, "random");       // 128 B of a 192 B record (straddles lines)
        GATHER(16, 1, 0, ids_r, n, 64, "random");      // 256 B rows
    }
#define RMW(LPR, CH, RIF, IDS, NN, P, TAG)                                                                             \
    { float us = timed(c, [&] { rmw_kernel<LPR, CH, RIF><<<grid_for((NN) * LPR / RIF), 256>>>(IDS, NN, tab, P); });     \
      snprintf(name, sizeof name, "rmw %3dB pitch %3dB lpr%d ch%d rif%d %s", LPR * CH * 16, P * 4, LPR, CH, RIF, TAG); \
      report(name, us, NN, 2 * LPR * CH * 16); }
    // 192 B records [p|m|v] (today's layout): 4 lanes x 3 chunks
    RMW(4, 3, 1, ids_s, nu, 48, "sorted");
    RMW(4, 3, 2, ids_s, nu, 48, "sorted");
    RMW(4, 3, 1, ids_r, n, 48, "random(dups race; timing only)");
    // 128 B single-line records: 8 lanes x 1 chunk, or 4 lanes x 2, or 2 lanes x 4
    RMW(8, 1, 1, ids_s, nu, 32, "sorted");
    RMW(8, 1, 2, ids_s, nu, 32, "sorted");
    RMW(8, 1, 4, ids_s, nu, 32, "sorted");
    RMW(4, 2, 2, ids_s, nu, 32, "sorted");
    RMW(2, 4, 2, ids_s, nu, 32, "sorted");
    RMW(8, 1, 2, ids_r, n, 32, "random(dups race; timing only)");
    // 16 B records (LR [w|m|v|stamp])
    RMW(1, 1, 4, ids_s, nu, 4, "sorted");
    RMW(1, 1, 4, ids_s, nu, 8, "sorted");              // 16 B records at 32 B pitch (one sector each)
    if (!quick) {
        RMW(4, 4, 1, ids_s, nu, 64, "sorted");         // 256 B records
        RMW(4, 1, 2, ids_s, nu, 16, "sorted");         // 64 B rows only
    }
    printf("done\n");
    return 0;
}
