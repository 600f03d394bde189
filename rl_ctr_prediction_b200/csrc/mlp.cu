// mlp.cu -- K4: the dense layers of the DeepFM tower and of the policy networks on the 5th-gen tensor
// cores (tcgen05 + TMEM), sm_100a only.
//
//     C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (ReLU)        fp32 in, fp32 out
//
// fp32 parity on tf32 tensor cores: 3xTF32 error-compensated split.  tcgen05 kind::tf32 reads fp32
// words from shared memory and ignores the low 13 mantissa bits, so each operand tile is staged twice:
//     hi = rn_tf32(x)       (round to nearest, done by the loader: exact for the tensor core)
//     lo = x - hi           (exact in fp32, |lo| <= 2^-11 |x|; its own truncation is a 2^-21 residual)
// and every K-slice issues three MMAs into the same TMEM accumulator: lo*hi + hi*lo + hi*hi.  The
// dropped lo*lo term and the residuals are ~2^-21 relative per product with random signs (SURVEY H2:
// plain TF32 is 1e-3 and fails the 1e-5 bar; the reference runs fp32 SGEMM, torch allow_tf32=False).
//
// One kernel serves the three GEMM forms of a Linear layer by taking element strides for both operands:
//     forward   Y  = X  W^T      A = X  [M=B , K=in ] k-contiguous   B = W  [N=out, K=in ] k-contiguous
//     dgrad     dX = dY W        A = dY [M=B , K=out] k-contiguous   B = W  [N=in , K=out] n-contiguous
//     wgrad     dW = dY^T X      A = dY [M=out, K=B ] m-contiguous   B = X  [N=in , K=B  ] n-contiguous  (split-K)
// Operands are staged by software loader warps (global -> registers -> hi/lo split -> swizzled smem),
// because the split has to be computed anyway and the strided forms are not TMA-box shaped.
//
// CTA = 4 loader warps | 1 MMA warp (one elected lane issues tcgen05.mma) | 4 epilogue warps
// (tcgen05.ld -> bias/ReLU -> global).  Persistent over output tiles; smem ring of NSTAGE k-blocks
// (full/empty mbarriers), two TMEM accumulator stages (tmem_full/tmem_empty mbarriers) so the epilogue
// of tile i overlaps the main loop of tile i+1.
//
// Shared-memory operand layout = the canonical UMMA K-major SWIZZLE_128B layout: a k-block is 32 fp32
// = 128 B per row; rows are 128 B apart; inside each group of 8 rows (1024 B) the 16-byte chunk index
// is XOR-ed with (row & 7).  Descriptor: start address, SBO = 1024 B (8-row group pitch), layout type
// SWIZZLE_128B, version 1; advancing K by one MMA (8 tf32 = 32 B) adds 32 B to the start address.
#include "common.cuh"

namespace rlctr {

constexpr int GEMM_BM = 128;              // UMMA M
constexpr int GEMM_BK = 32;               // fp32 per k-block row = 128 B = one swizzle atom
constexpr int GEMM_UK = 8;                // K per tcgen05.mma kind::tf32
constexpr int GEMM_LOADER_WARPS = 4;
constexpr int GEMM_EPI_WARPS = 4;
constexpr int GEMM_THREADS = 32 * (GEMM_LOADER_WARPS + 1 + GEMM_EPI_WARPS);
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_TMEM_COLS = 512;

struct GemmArgs {
    const float* A; int64_t sam, sak;     // A[m*sam + k*sak]   (one of the two strides is 1)
    const float* B; int64_t sbn, sbk;     // B[n*sbn + k*sbk]
    int vec_a, vec_b;                     // elements per global vector load along the contiguous dimension (1, 2, 4)
    float* C; int64_t ldc;                // C[(split*M + m)*ldc + n]
    const float* bias;                    // [N] or null (ignored when splits > 1)
    int M, N, K;
    int n_tile;                           // UMMA N (multiple of 16, <= 256)
    int m_tiles, n_tiles, splits, kb_per_split;
    int stages;
    int relu;
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);           // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                            // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8-row group pitch
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
    return d;
}
// byte offset of element (row, k) inside a [rows x 32 fp32] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
    return (uint32_t)(row * 128 + ((((k >> 2) ^ (row & 7)) << 4) | ((k & 3) << 2)));
}

__device__ __forceinline__ void split_store(char* hi, char* lo, uint32_t off, float x) {
    // hi = x rounded to nearest tf32 (10 explicit mantissa bits): the tensor core then reads it exactly;
    // lo = x - hi is exact in fp32, |lo| <= 2^-11 |x| with a random sign, and loses only 2^-21 |x| to the
    // tensor core's own truncation.
    const float h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    *reinterpret_cast<float*>(hi + off) = h;
    *reinterpret_cast<float*>(lo + off) = x - h;
}

// Stage one operand tile: rows [r0, r0+rows) x k [k0, k0+32) of G[r*sr + k*sk], zero-filled outside (R, K).
// VEC contiguous elements per global load (along k when KCONTIG, along rows otherwise); LOADS_IN_FLIGHT
// independent vector loads are issued before any is consumed (latency hiding: the loaders are the
// producers of a tensor-core pipeline and see full DRAM/L2 latency).
template <int VEC, bool KCONTIG>
__device__ __forceinline__ void load_tile_v(const float* __restrict__ G, int64_t sr, int64_t sk, int R, int K, int r0,
                                            int k0, int rows, char* hi, char* lo, int tid, int nthreads) {
    constexpr int U = 8;
    const int per = KCONTIG ? (GEMM_BK / VEC) : (rows / VEC);     // vectors along the contiguous dimension
    const int nvec = rows * GEMM_BK / VEC;
    for (int base = tid; base < nvec; base += nthreads * U) {
        float buf[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = base + u * nthreads;
#pragma unroll
            for (int e = 0; e < VEC; ++e) buf[u][e] = 0.f;
            if (v < nvec) {
                const int c = v % per, o = v / per;
                const int r = KCONTIG ? o : c * VEC, k = KCONTIG ? c * VEC : o;
                const int gr = r0 + r, gk = k0 + k;
                const float* src = G + (int64_t)gr * sr + (int64_t)gk * sk;
                const bool full = KCONTIG ? (gr < R && gk + VEC <= K) : (gk < K && gr + VEC <= R);
                if (full) {
                    if (VEC == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(src)); buf[u][0] = t.x; buf[u][1] = t.y; buf[u][2 % VEC] = t.z; buf[u][3 % VEC] = t.w; }
                    else if (VEC == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(src)); buf[u][0] = t.x; buf[u][1 % VEC] = t.y; }
                    else buf[u][0] = __ldg(src);
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const bool ok = KCONTIG ? (gr < R && gk + e < K) : (gk < K && gr + e < R);
                        if (ok) buf[u][e] = __ldg(src + (KCONTIG ? (int64_t)e * sk : (int64_t)e * sr));
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = base + u * nthreads;
            if (v < nvec) {
                const int c = v % per, o = v / per;
                const int r = KCONTIG ? o : c * VEC, k = KCONTIG ? c * VEC : o;
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    split_store(hi, lo, KCONTIG ? sw128_off(r, k + e) : sw128_off(r + e, k), buf[u][e]);
            }
        }
    }
}

__device__ __forceinline__ void load_tile(const float* __restrict__ G, int64_t sr, int64_t sk, int vec, int R, int K,
                                          int r0, int k0, int rows, char* hi, char* lo, int tid, int nthreads) {
    if (sk == 1) {
        if (vec == 4) load_tile_v<4, true>(G, sr, sk, R, K, r0, k0, rows, hi, lo, tid, nthreads);
        else if (vec == 2) load_tile_v<2, true>(G, sr, sk, R, K, r0, k0, rows, hi, lo, tid, nthreads);
        else load_tile_v<1, true>(G, sr, sk, R, K, r0, k0, rows, hi, lo, tid, nthreads);
    } else {
        if (vec == 4) load_tile_v<4, false>(G, sr, sk, R, K, r0, k0, rows, hi, lo, tid, nthreads);
        else if (vec == 2) load_tile_v<2, false>(G, sr, sk, R, K, r0, k0, rows, hi, lo, tid, nthreads);
        else load_tile_v<1, false>(G, sr, sk, R, K, r0, k0, rows, hi, lo, tid, nthreads);
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm3x_tf32_kernel(const GemmArgs g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[GEMM_MAX_STAGES], empty_bar[GEMM_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    // dynamic smem is only guaranteed 16 B aligned: round up to the 1024 B the swizzle atom needs
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_bytes = GEMM_BM * 128, b_bytes = (uint32_t)g.n_tile * 128;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const int total_tiles = g.m_tiles * g.n_tiles * g.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), GEMM_LOADER_WARPS * 32);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tmem_full_bar[s]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[s]), GEMM_EPI_WARPS * 32);
        }
        fence_barrier_init();
    }
    if (warp == GEMM_LOADER_WARPS) tmem_alloc(smem_u32(&tmem_base_smem), GEMM_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < GEMM_LOADER_WARPS) {
        // ================= loaders =================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int split = tile / (g.m_tiles * g.n_tiles);
            const int mn = tile - split * (g.m_tiles * g.n_tiles);
            const int mt = mn / g.n_tiles, nt = mn - mt * g.n_tiles;
            const int kb0 = split * g.kb_per_split;
            const int kb_total = (g.K + GEMM_BK - 1) / GEMM_BK;
            const int kb1 = min(kb0 + g.kb_per_split, kb_total);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                unsigned char* st = smem + (size_t)stage * stage_bytes;
                load_tile(g.A, g.sam, g.sak, g.vec_a, g.M, g.K, mt * GEMM_BM, kb * GEMM_BK, GEMM_BM, (char*)st,
                          (char*)st + a_bytes, threadIdx.x, GEMM_LOADER_WARPS * 32);
                load_tile(g.B, g.sbn, g.sbk, g.vec_b, g.N, g.K, nt * g.n_tile, kb * GEMM_BK, g.n_tile, (char*)st + 2 * a_bytes,
                          (char*)st + 2 * a_bytes + b_bytes, threadIdx.x, GEMM_LOADER_WARPS * 32);
                fence_proxy_async();                   // generic-proxy stores -> visible to the tensor core (async proxy)
                mbar_arrive(smem_u32(&full_bar[stage]));
                if (++stage == g.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == GEMM_LOADER_WARPS) {
        // ================= MMA issuer =================
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major, M=128, N=n_tile
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.n_tile >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int split = tile / (g.m_tiles * g.n_tiles);
            const int kb0 = split * g.kb_per_split;
            const int kb_total = (g.K + GEMM_BK - 1) / GEMM_BK;
            const int kb1 = min(kb0 + g.kb_per_split, kb_total);
            mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);       // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t a_hi = sa, a_lo = sa + a_bytes, b_hi = sa + 2 * a_bytes, b_lo = b_hi + b_bytes;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / GEMM_UK; ++k) {
                        const uint32_t ko = (uint32_t)k * GEMM_UK * 4;       // 32 B per MMA along K inside the swizzle atom
                        const uint32_t first = (kb > kb0 || k > 0) ? 1u : 0u;
                        umma_tf32(d_tmem, make_smem_desc(a_lo + ko), make_smem_desc(b_hi + ko), idesc, first);
                        umma_tf32(d_tmem, make_smem_desc(a_hi + ko), make_smem_desc(b_lo + ko), idesc, 1u);
                        umma_tf32(d_tmem, make_smem_desc(a_hi + ko), make_smem_desc(b_hi + ko), idesc, 1u);
                    }
                    umma_commit(smem_u32(&empty_bar[stage]));                // frees the smem slot when these MMAs retire
                    if (kb == kb1 - 1) umma_commit(smem_u32(&tmem_full_bar[acc]));
                }
                __syncwarp();
                if (++stage == g.stages) { stage = 0; phase ^= 1; }
            }
            if (kb1 <= kb0 && lane == 0) umma_commit(smem_u32(&tmem_full_bar[acc]));   // empty split: nothing accumulated
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;                        // TMEM lane quadrant this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int split = tile / (g.m_tiles * g.n_tiles);
            const int mn = tile - split * (g.m_tiles * g.n_tiles);
            const int mt = mn / g.n_tiles, nt = mn - mt * g.n_tiles;
            const int kb0 = split * g.kb_per_split;
            const int kb_total = (g.K + GEMM_BK - 1) / GEMM_BK;
            const bool empty_split = min(kb0 + g.kb_per_split, kb_total) <= kb0;
            mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
            tc_fence_after();
            const int m = mt * GEMM_BM + q * 32 + lane;
            float* crow = g.C + ((int64_t)split * g.M + m) * g.ldc;
            const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(q * 32) << 16);
            for (int c0 = 0; c0 < g.n_tile; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + (uint32_t)c0, r);
                tmem_ld_wait();
                if (m < g.M) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = nt * g.n_tile + c0 + j;
                        if (n < g.N) {
                            float v = empty_split ? 0.f : __uint_as_float(r[j]);
                            if (g.bias && g.splits == 1) v += __ldg(g.bias + n);
                            if (g.relu && g.splits == 1) v = fmaxf(v, 0.f);
                            crow[n] = v;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == GEMM_LOADER_WARPS) tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
}

// split-K reduction (fixed order over splits => bit-identical run to run), + bias/ReLU never needed here
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t mn, int splits) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < mn; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += __ldg(part + (int64_t)k * mn + i);
        out[i] = s;
    }
}

// column sums of dY [B, N] (the bias gradient), fixed-shape: each block owns 32 columns, 8 warps stride the
// rows, smem tree over the 8 partials.
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ Y, float* __restrict__ part, int64_t rows, int N, int rows_per_block) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float s = 0.f;
    if (n < N)
        for (int64_t r = r0 + w; r < r1; r += 8) s += __ldg(Y + r * N + n);
    red[w][lane] = s;
    __syncthreads();
    if (w == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][lane];
        if (n < N) part[(int64_t)blockIdx.y * N + n] = t;
    }
}

// dY = g * (out > 0)  (ReLU backward on the saved post-activation), elementwise
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float* g, const float* __restrict__ out, float* dy, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dy[i] = __ldg(out + i) > 0.f ? g[i] : 0.f;
}

static int round16(int n) { return (n + 15) / 16 * 16; }

struct GemmPlan {
    int n_tile, n_tiles, m_tiles, splits, kb_per_split, stages;
    size_t smem;
};
static GemmPlan plan_gemm(int M, int N, int K, bool allow_split) {
    GemmPlan p;
    p.n_tiles = (N + 255) / 256;
    p.n_tile = round16((N + p.n_tiles - 1) / p.n_tiles);
    if (p.n_tile < 16) p.n_tile = 16;
    p.m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    const int kb_total = (K + GEMM_BK - 1) / GEMM_BK;
    p.splits = 1;
    if (allow_split) {
        int tiles = p.m_tiles * p.n_tiles;
        int s = RLCTR_SMS / (tiles > 0 ? tiles : 1);
        if (s < 1) s = 1;
        int max_s = kb_total / 8;                       // keep >= 8 k-blocks per split
        if (max_s < 1) max_s = 1;
        p.splits = s < max_s ? s : max_s;
    }
    p.kb_per_split = (kb_total + p.splits - 1) / p.splits;
    if (p.kb_per_split < 1) p.kb_per_split = 1;
    p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
    if (p.splits < 1) p.splits = 1;
    const size_t stage_bytes = 2 * (size_t)GEMM_BM * 128 + 2 * (size_t)p.n_tile * 128;
    int st = (int)((size_t)(220 * 1024) / stage_bytes);
    if (st > GEMM_MAX_STAGES) st = GEMM_MAX_STAGES;
    if (st < 1) st = 1;
    p.stages = st;
    p.smem = stage_bytes * st + 1024;
    return p;
}

// widest aligned vector load for an operand whose non-contiguous stride is `pitch` floats
static int vec_of(const float* p, int64_t pitch) {
    const uintptr_t a = (uintptr_t)p;
    if (pitch % 4 == 0 && a % 16 == 0) return 4;
    if (pitch % 2 == 0 && a % 8 == 0) return 2;
    return 1;
}

static int launch_gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C,
                       int64_t ldc, const float* bias, int M, int N, int K, int relu, const GemmPlan& p, cudaStream_t st) {
    GemmArgs g;
    g.A = A; g.sam = sam; g.sak = sak;
    g.B = B; g.sbn = sbn; g.sbk = sbk;
    g.C = C; g.ldc = ldc; g.bias = bias;
    g.M = M; g.N = N; g.K = K;
    g.n_tile = p.n_tile; g.m_tiles = p.m_tiles; g.n_tiles = p.n_tiles; g.splits = p.splits; g.kb_per_split = p.kb_per_split;
    g.stages = p.stages; g.relu = relu;
    g.vec_a = vec_of(A, sak == 1 ? sam : sak);
    g.vec_b = vec_of(B, sbk == 1 ? sbn : sbk);
    RLCTR_CUDA(cudaFuncSetAttribute(gemm3x_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    const int total = p.m_tiles * p.n_tiles * p.splits;
    const int grid = total < RLCTR_SMS ? total : RLCTR_SMS;
    gemm3x_tf32_kernel<<<grid, GEMM_THREADS, p.smem, st>>>(g);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

constexpr int COLSUM_ROWS_PER_BLOCK = 4096;

}  // namespace rlctr

using namespace rlctr;

extern "C" size_t rlctr_mlp_ws_bytes(int64_t batch, int32_t in_dim, int32_t out_dim) {
    if (batch <= 0 || in_dim <= 0 || out_dim <= 0) return 256;
    GemmPlan p = plan_gemm(out_dim, in_dim, (int)batch, true);
    size_t wgrad = (size_t)p.splits * out_dim * in_dim * sizeof(float);
    size_t colsum = (size_t)((batch + COLSUM_ROWS_PER_BLOCK - 1) / COLSUM_ROWS_PER_BLOCK) * out_dim * sizeof(float);
    return wgrad + colsum + 512;
}

extern "C" int rlctr_linear_fwd(const float* x, const float* w, const float* bias, float* y, int64_t batch,
                                int32_t in_dim, int32_t out_dim, int32_t flags, void* ws, size_t ws_bytes,
                                rlctr_stream_t stream) {
    (void)ws; (void)ws_bytes;
    if (!x || !w || !y || batch < 0 || in_dim <= 0 || out_dim <= 0) return RLCTR_EINVAL;
    if (batch == 0) return RLCTR_OK;
    if (batch > 0x7fffffff) return RLCTR_EUNSUPPORTED;
    GemmPlan p = plan_gemm((int)batch, out_dim, in_dim, false);
    return launch_gemm(x, in_dim, 1, w, in_dim, 1, y, out_dim, bias, (int)batch, out_dim, in_dim,
                       (flags & RLCTR_MLP_RELU) ? 1 : 0, p, (cudaStream_t)stream);
}

extern "C" int rlctr_linear_bwd(const float* x, const float* w, const float* y, float* gy, float* dx, float* dw,
                                float* db, int64_t batch, int32_t in_dim, int32_t out_dim, int32_t flags, void* ws,
                                size_t ws_bytes, rlctr_stream_t stream) {
    if (!x || !w || !gy || batch <= 0 || in_dim <= 0 || out_dim <= 0) return RLCTR_EINVAL;
    if (batch > 0x7fffffff) return RLCTR_EUNSUPPORTED;
    if ((flags & RLCTR_MLP_RELU) && !y) return RLCTR_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const float* dy = gy;
    if (flags & RLCTR_MLP_RELU) {                       // ReLU backward on the saved output, in place on gy
        const int64_t n = batch * out_dim;
        int64_t blocks = (n + 255) / 256;
        int grid = (int)(blocks < RLCTR_SMS * 8 ? blocks : RLCTR_SMS * 8);
        relu_bwd_kernel<<<grid, 256, 0, st>>>(gy, y, gy, n);
        RLCTR_LAUNCH_CHECK();
    }
    if (dx) {   // dX[B,in] = dY[B,out] * W[out,in]:  B operand = W^T, n-contiguous
        GemmPlan p = plan_gemm((int)batch, in_dim, out_dim, false);
        int rc = launch_gemm(dy, out_dim, 1, w, 1, in_dim, dx, in_dim, nullptr, (int)batch, in_dim, out_dim, 0, p, st);
        if (rc) return rc;
    }
    if (dw || db) {
        if (!ws || ws_bytes < rlctr_mlp_ws_bytes(batch, in_dim, out_dim)) return RLCTR_EWORKSPACE;
    }
    if (dw) {   // dW[out,in] = dY^T[out,B] * X[B,in]: both operands strided along K = batch; split-K
        GemmPlan p = plan_gemm(out_dim, in_dim, (int)batch, true);
        float* part = reinterpret_cast<float*>(ws);
        if (p.splits == 1) {
            int rc = launch_gemm(dy, 1, out_dim, x, 1, in_dim, dw, in_dim, nullptr, out_dim, in_dim, (int)batch, 0, p, st);
            if (rc) return rc;
        } else {
            int rc = launch_gemm(dy, 1, out_dim, x, 1, in_dim, part, in_dim, nullptr, out_dim, in_dim, (int)batch, 0, p, st);
            if (rc) return rc;
            const int64_t mn = (int64_t)out_dim * in_dim;
            int64_t blocks = (mn + 255) / 256;
            splitk_reduce_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, st>>>(part, dw, mn, p.splits);
            RLCTR_LAUNCH_CHECK();
        }
    }
    if (db) {
        GemmPlan p = plan_gemm(out_dim, in_dim, (int)batch, true);
        float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) +
                                               (((size_t)p.splits * out_dim * in_dim * sizeof(float) + 255) & ~(size_t)255));
        const int yb = (int)((batch + COLSUM_ROWS_PER_BLOCK - 1) / COLSUM_ROWS_PER_BLOCK);
        dim3 grid((out_dim + 31) / 32, yb);
        colsum_partial_kernel<<<grid, 256, 0, st>>>(dy, part, batch, out_dim, COLSUM_ROWS_PER_BLOCK);
        RLCTR_LAUNCH_CHECK();
        splitk_reduce_kernel<<<(out_dim + 255) / 256, 256, 0, st>>>(part, db, out_dim, yb);
        RLCTR_LAUNCH_CHECK();
    }
    return RLCTR_OK;
}
