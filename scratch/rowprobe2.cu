// rowprobe2.cu -- where is the ~31 G line-requests/s ceiling of random row access?  (scratch; standalone)
//   (1) the same gather on 37 / 74 / 111 / 148 SMs (one persistent 1024-thread CTA per SM): SM-side or memory-side limit?
//   (2) load flavours: ld.global.nc / .cg / .cv(volatile) / no_allocate + L2::64B / 256-bit
//   (3) TMA: cp.async.bulk.shared.global of RB bytes per row (one instruction per row), mbarrier completion
//   (4) larger problem (4x rows) to take the launch ramp out of the rate
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ float4 ldm(const float* p) {
    float4 v;
    if (MODE == 0) asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (MODE == 3) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (MODE == 4) asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (MODE == 5) asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// persistent: gridDim.x CTAs of 1024 threads (dynamic smem forces one CTA per SM)
template <int LPR, int RIF, int MODE>
__global__ void __launch_bounds__(1024, 1) gather_p_kernel(const uint32_t* __restrict__ ids, int64_t n, const float* __restrict__ tab, int pitch,
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
        for (int r = 0; r < RIF; ++r) acc += v[r].x + v[r].y + v[r].z + v[r].w;
    }
    if (acc == 123.456f) out[gt] = acc;
}

// ---- TMA bulk gather: each lane copies its row (RB bytes) global -> shared, one mbarrier per (warp, stage) -----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
template <int RB, int STAGES>
__global__ void __launch_bounds__(256) tma_gather_kernel(const uint32_t* __restrict__ ids, int64_t n, const char* __restrict__ tab, int pitch_bytes,
                                                         float* __restrict__ out) {
    extern __shared__ __align__(128) char smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    char* buf = smem + (size_t)wib * STAGES * 32 * RB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nw * STAGES * 32 * RB) + wib * STAGES;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int64_t warp0 = (int64_t)blockIdx.x * nw + wib, nwarps = (int64_t)gridDim.x * nw;
    const int64_t nbatch = (n + 31) / 32;
    float acc = 0.f;
    int64_t issue = warp0;
    int s_issue = 0;
    auto issue_batch = [&](int64_t bt, int s) {
        const int64_t k = bt * 32 + lane;
        const uint32_t id = k < n ? __ldg(ids + k) : 0;
        if (lane == 0) mbar_expect(smem_u32(&bars[s]), 32 * RB);
        __syncwarp();
        bulk_g2s(smem_u32(buf + ((size_t)s * 32 + lane) * RB), tab + (int64_t)id * pitch_bytes, RB, smem_u32(&bars[s]));
    };
    for (int s = 0; s < STAGES - 1 && issue < nbatch; ++s) { issue_batch(issue, s_issue); issue += nwarps; s_issue = (s_issue + 1) % STAGES; }
    int s_wait = 0;
    uint32_t phase = 0;
    for (int64_t bt = warp0; bt < nbatch; bt += nwarps) {
        if (issue < nbatch) { issue_batch(issue, s_issue); issue += nwarps; s_issue = (s_issue + 1) % STAGES; }
        mbar_wait(smem_u32(&bars[s_wait]), phase);
        const float4* r = reinterpret_cast<const float4*>(buf + ((size_t)s_wait * 32 + lane) * RB);
#pragma unroll
        for (int q = 0; q < RB / 16; ++q) { float4 v = r[q]; acc += v.x + v.w; }
        __syncwarp();
        if (++s_wait == STAGES) { s_wait = 0; phase ^= 1; }
    }
    if (acc == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
struct Ctx { float* flush; int64_t flush_n; cudaEvent_t e0, e1; };
static int g_reps = 9;
template <class F> static float timed(Ctx& c, F&& launch) {
    std::vector<float> ts;
    for (int i = 0; i < g_reps + 2; ++i) {
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
    printf("%-52s %8.1f us  %6.2f G rows/s  %8.1f GB/s logical\n", name, us, n / us * 1e-3, (double)n * logical_bytes / us * 1e-3);
    fflush(stdout);
}

int main(int argc, char** argv) {
    if (argc > 1) g_reps = atoi(argv[1]);
    CK(cudaFree(0));
    const int64_t n_rows = 10000000, n = 983040 * 4;
    const int pitch = 64;
    float* tab; CK(cudaMalloc(&tab, (size_t)n_rows * pitch * 4));
    fill_kernel<<<148 * 8, 256>>>(tab, n_rows * pitch, 1.0f);
    Ctx c; c.flush_n = 96 << 20; CK(cudaMalloc(&c.flush, c.flush_n * 4));
    CK(cudaEventCreate(&c.e0)); CK(cudaEventCreate(&c.e1));
    float* out; CK(cudaMalloc(&out, 64 << 20));
    std::mt19937_64 rng(1);
    std::vector<uint32_t> h(n);
    for (auto& x : h) x = (uint32_t)(rng() % (uint64_t)n_rows);
    uint32_t* ids; CK(cudaMalloc(&ids, n * 4));
    CK(cudaMemcpy(ids, h.data(), n * 4, cudaMemcpyHostToDevice));
    char name[160];
    const int64_t n1 = 983040;
#define GP(LPR, RIF, MODE, SMS, NN, P)                                                                                 \
    { CK(cudaFuncSetAttribute(gather_p_kernel<LPR, RIF, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));  \
      float us = timed(c, [&] { gather_p_kernel<LPR, RIF, MODE><<<SMS, 1024, 120 * 1024>>>(ids, NN, tab, P, out); });     \
      snprintf(name, sizeof name, "persistent gather %3dB pitch %3dB rif%d mode%d SMs %3d n %lld", LPR * 16, P * 4, RIF, MODE, SMS, (long long)(NN)); \
      report(name, us, NN, LPR * 16); }
    // (1) SM scaling, 64 B rows and 128 B rows, 4x problem
    GP(4, 2, 1, 37, n, 48); GP(4, 2, 1, 74, n, 48); GP(4, 2, 1, 111, n, 48); GP(4, 2, 1, 148, n, 48);
    GP(8, 2, 1, 37, n, 32); GP(8, 2, 1, 74, n, 32); GP(8, 2, 1, 111, n, 32); GP(8, 2, 1, 148, n, 32);
    GP(4, 4, 1, 148, n, 48); GP(8, 4, 1, 148, n, 32); GP(4, 8, 1, 148, n, 48);
    GP(4, 2, 1, 148, n1, 48); GP(8, 2, 1, 148, n1, 32);
    // (2) flavours at 148 SMs
    GP(4, 2, 0, 148, n, 48); GP(4, 2, 2, 148, n, 48); GP(4, 2, 3, 148, n, 48); GP(4, 2, 4, 148, n, 48); GP(4, 2, 5, 148, n, 48);
    GP(8, 2, 0, 148, n, 32); GP(8, 2, 3, 148, n, 32); GP(8, 2, 5, 148, n, 32);
    // (3) TMA bulk gather
#define TG(RB, STAGES, BLOCKS_PER_SM, NN, PB)                                                                          \
    { const size_t sm = (size_t)8 * STAGES * 32 * RB + 8 * STAGES * 8;                                                   \
      CK(cudaFuncSetAttribute(tma_gather_kernel<RB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
      float us = timed(c, [&] { tma_gather_kernel<RB, STAGES><<<148 * BLOCKS_PER_SM, 256, sm>>>(ids, NN, (const char*)tab, PB, out); }); \
      snprintf(name, sizeof name, "tma bulk gather %3dB pitch %3dB stages %d ctas/SM %d n %lld", RB, PB, STAGES, BLOCKS_PER_SM, (long long)(NN)); \
      report(name, us, NN, RB); }
    TG(64, 2, 4, n, 192); TG(64, 4, 4, n, 192); TG(128, 2, 4, n, 128); TG(128, 4, 2, n, 128); TG(192, 2, 2, n, 192); TG(192, 4, 1, n, 192);
    TG(128, 2, 4, n1, 128); TG(192, 2, 2, n1, 192); TG(256, 2, 2, n, 256);
    printf("done\n");
    return 0;
}
