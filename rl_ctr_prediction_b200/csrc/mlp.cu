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
#include "mlp_tma.cuh"

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
    int mode_a, mode_b;                   // staging mode: 4/2/1 = k-contiguous, that many fp32 per copy; 0 = row-contiguous
    float* C; int64_t ldc;                // C[(split*M + m)*ldc + n]
    int cvec;                             // widest aligned vector store for a C row segment (1, 2, 4)
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

// hi = x rounded to nearest tf32 (10 explicit mantissa bits): the tensor core then reads it exactly;
// lo = x - hi is exact in fp32, |lo| <= 2^-11 |x| with a random sign, and loses only 2^-21 |x| to the
// tensor core's own truncation.
__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst), "l"(src), "n"(BYTES), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Operand staging, two phases per k-block:
//   issue  : cp.async (LDGSTS) global -> smem `raw` tile in the swizzled layout, zero-filled outside (R, K);
//            asynchronous, so PREFETCH k-blocks of copies are in flight per thread with no register cost.
//            k-contiguous operands: VEC fp32 per copy, a thread keeps its k-chunk and walks rows with constant
//            strides (no index arithmetic in the loop); row-contiguous operands (dgrad's W^T, wgrad's dY^T and
//            X): 4-byte copies, lanes along rows, k unrolled.
//   convert: after cp.async.wait_group + a named barrier over the loader warps, raw -> (hi in place, lo)
//            with 128-bit shared-memory accesses.
constexpr int LOADER_THREADS = GEMM_LOADER_WARPS * 32;

template <int VEC>
__device__ __forceinline__ void tile_issue_k(const float* __restrict__ G, int64_t sr, int R, int K, int r0, int k0,
                                             int rows, uint32_t raw, int tid) {
    constexpr int PER = GEMM_BK / VEC;                 // vectors per row
    constexpr int RSTEP = LOADER_THREADS / PER;        // rows advanced per iteration
    const int c = tid % PER, rr = tid / PER;
    const int gk = k0 + c * VEC;
    const int kvalid = gk < K ? min(K - gk, VEC) : 0;
    const float* src = G + (int64_t)(r0 + rr) * sr + gk;
    const int64_t src_step = (int64_t)RSTEP * sr;
    if (VEC >= 2) {                                    // RSTEP % 8 == 0: the swizzle phase (row & 7) is loop-invariant
        uint32_t dst = raw + sw128_off(rr, c * VEC);
        for (int r = rr; r < rows; r += RSTEP) {
            const int valid = (r0 + r < R) ? kvalid : 0;
            cp_async<4 * VEC>(dst, valid ? src : G, 4 * valid);
            dst += RSTEP * 128;
            src += src_step;
        }
    } else {
        for (int r = rr; r < rows; r += RSTEP) {
            const int valid = (r0 + r < R) ? kvalid : 0;
            cp_async<4>(raw + sw128_off(r, c), valid ? src : G, 4 * valid);
            src += src_step;
        }
    }
}
__device__ __forceinline__ void tile_issue_r(const float* __restrict__ G, int64_t sk, int R, int K, int r0, int k0,
                                             int rows, uint32_t raw, int tid) {
    for (int r = tid; r < rows; r += LOADER_THREADS) {
        const bool rok = r0 + r < R;
        const float* src = G + (r0 + r) + (int64_t)k0 * sk;
        const uint32_t dst = raw + r * 128;
        const int r7 = r & 7;
#pragma unroll
        for (int k = 0; k < GEMM_BK; ++k) {
            const bool ok = rok && (k0 + k < K);
            cp_async<4>(dst + ((((k >> 2) ^ r7) << 4) | ((k & 3) << 2)), ok ? src + (int64_t)k * sk : G, ok ? 4 : 0);
        }
    }
}
// mode: 4 / 2 / 1 = k-contiguous with that vector width; 0 = row-contiguous
__device__ __forceinline__ void tile_issue_any(int mode, const float* __restrict__ G, int64_t sr, int64_t sk, int R, int K,
                                               int r0, int k0, int rows, uint32_t raw, int tid) {
    if (mode == 4) tile_issue_k<4>(G, sr, R, K, r0, k0, rows, raw, tid);
    else if (mode == 2) tile_issue_k<2>(G, sr, R, K, r0, k0, rows, raw, tid);
    else if (mode == 1) tile_issue_k<1>(G, sr, R, K, r0, k0, rows, raw, tid);
    else tile_issue_r(G, sk, R, K, r0, k0, rows, raw, tid);
}
// rows is a multiple of 16: thread tid owns 16-byte chunk (tid % 8) of rows tid/8, tid/8 + 16, ...
__device__ __forceinline__ void tile_convert(int rows, char* hi, char* lo, int tid) {
    const int c = tid & 7, rr = tid >> 3;
    uint32_t off = (uint32_t)(rr * 128 + ((c ^ (rr & 7)) << 4));
    for (int r = rr; r < rows; r += LOADER_THREADS / 8) {
        const float4 x = *reinterpret_cast<const float4*>(hi + off);
        const float4 h = make_float4(tf32_rn(x.x), tf32_rn(x.y), tf32_rn(x.z), tf32_rn(x.w));
        *reinterpret_cast<float4*>(hi + off) = h;
        *reinterpret_cast<float4*>(lo + off) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        off += (LOADER_THREADS / 8) * 128;
    }
}

// walks this CTA's (tile, k-block) work items in order; used twice by the loaders (issue / convert cursors)
struct ItemCursor {
    int tile, kb, kb1, mt, nt, stage;
    uint32_t phase;
    bool valid;
};
__device__ __forceinline__ void cursor_set_tile(ItemCursor& c, const GemmArgs& g, int total_tiles) {
    const int kb_total = (g.K + GEMM_BK - 1) / GEMM_BK;
    while (c.tile < total_tiles) {
        const int split = c.tile / (g.m_tiles * g.n_tiles);
        const int mn = c.tile - split * (g.m_tiles * g.n_tiles);
        c.mt = mn / g.n_tiles;
        c.nt = mn - c.mt * g.n_tiles;
        c.kb = split * g.kb_per_split;
        c.kb1 = min(c.kb + g.kb_per_split, kb_total);
        if (c.kb < c.kb1) { c.valid = true; return; }
        c.tile += gridDim.x;                                   // empty split: no k-blocks to stage
    }
    c.valid = false;
}
__device__ __forceinline__ void cursor_next(ItemCursor& c, const GemmArgs& g, int total_tiles) {
    if (++c.stage == g.stages) { c.stage = 0; c.phase ^= 1; }
    if (++c.kb >= c.kb1) { c.tile += gridDim.x; cursor_set_tile(c, g, total_tiles); }
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
        const int tid = threadIdx.x;
        ItemCursor ci{(int)blockIdx.x, 0, 0, 0, 0, 0, 0u, false}, cc = ci;
        cursor_set_tile(ci, g, total_tiles);
        cursor_set_tile(cc, g, total_tiles);
        const int prefetch = g.stages >= 3 ? 2 : 1;            // k-blocks of copies in flight per thread
        int lead = 0;                                          // issued - converted
        while (cc.valid) {
            if (ci.valid && lead <= prefetch) {
                mbar_wait(smem_u32(&empty_bar[ci.stage]), ci.phase ^ 1);
                const uint32_t st = smem_u32(smem + (size_t)ci.stage * stage_bytes);
                tile_issue_any(g.mode_a, g.A, g.sam, g.sak, g.M, g.K, ci.mt * GEMM_BM, ci.kb * GEMM_BK, GEMM_BM, st, tid);
                tile_issue_any(g.mode_b, g.B, g.sbn, g.sbk, g.N, g.K, ci.nt * g.n_tile, ci.kb * GEMM_BK, g.n_tile,
                               st + 2 * a_bytes, tid);
                cp_async_commit();
                cursor_next(ci, g, total_tiles);
                ++lead;
                if (ci.valid && lead <= prefetch) continue;    // fill the pipeline first
            }
            // oldest outstanding group of THIS thread has landed once at most (lead-1) newer groups remain
            if (lead >= 3) cp_async_wait<2>(); else if (lead == 2) cp_async_wait<1>(); else cp_async_wait<0>();
            asm volatile("bar.sync 1, %0;" ::"n"(LOADER_THREADS) : "memory");     // every loader's copies of this k-block landed
            char* st = reinterpret_cast<char*>(smem + (size_t)cc.stage * stage_bytes);
            tile_convert(GEMM_BM, st, st + a_bytes, tid);
            tile_convert(g.n_tile, st + 2 * a_bytes, st + 2 * a_bytes + b_bytes, tid);
            fence_proxy_async();                               // generic-proxy stores -> visible to the tensor core
            mbar_arrive(smem_u32(&full_bar[cc.stage]));
            cursor_next(cc, g, total_tiles);
            --lead;
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
            const bool fuse = g.splits == 1;
            for (int c0 = 0; c0 < g.n_tile; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + (uint32_t)c0, r);
                tmem_ld_wait();
                const int n0 = nt * g.n_tile + c0;
                if (m < g.M && n0 < g.N) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float x = empty_split ? 0.f : __uint_as_float(r[j]);
                        if (fuse && g.bias && n0 + j < g.N) x += __ldg(g.bias + n0 + j);
                        if (fuse && g.relu) x = fmaxf(x, 0.f);
                        v[j] = x;
                    }
                    float* dst = crow + n0;
                    if (n0 + 16 <= g.N && g.cvec == 4) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else if (n0 + 16 <= g.N && g.cvec == 2) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) *reinterpret_cast<float2*>(dst + j) = make_float2(v[j], v[j + 1]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (n0 + j < g.N) dst[j] = v[j];
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

// the same for two partial arrays in one launch (a layer's dW and db: one graph node instead of two; the same sums in the same order)
__global__ void __launch_bounds__(256)
splitk_reduce2_kernel(const float* __restrict__ part_a, float* __restrict__ out_a, int64_t n_a, const float* __restrict__ part_b,
                      float* __restrict__ out_b, int64_t n_b, int splits) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_a + n_b; i += (int64_t)gridDim.x * blockDim.x) {
        const bool a = i < n_a;
        const float* src = a ? part_a + i : part_b + (i - n_a);
        const int64_t stride = a ? n_a : n_b;
        float s = 0.f;
        int k = 0;
        for (; k + 8 <= splits; k += 8) {                // eight independent loads in flight, added in split order
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (int64_t)(k + u) * stride);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; k < splits; ++k) s += __ldg(src + (int64_t)k * stride);
        if (a) out_a[i] = s; else out_b[i - n_a] = s;
    }
}

// column sums of dY [B, N] (the bias gradient), fixed-shape: each block owns 32 columns, 8 warps stride the
// rows, smem tree over the 8 partials.
// wt != null: rows are scaled by wt[r] first (dW of a single-output layer: sum_b gy[b] * x[b, :])
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ Y, int64_t ldy, const float* __restrict__ wt, float* __restrict__ part,
                      int64_t rows, int N, int rows_per_block) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float s = 0.f;
    if (n < N)
#pragma unroll 4
        for (int64_t r = r0 + w; r < r1; r += 8) s += wt ? __ldg(wt + r) * __ldg(Y + r * ldy + n) : __ldg(Y + r * ldy + n);
    red[w][lane] = s;
    __syncthreads();
    if (w == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][lane];
        if (n < N) part[(int64_t)blockIdx.y * N + n] = t;
    }
}

// dY = g * (out > 0 ? scale : 0)  (ReLU [+ dropout: scale = 1/(1-p)] backward on the saved output), elementwise
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float* g, const float* __restrict__ out, float* dy, int64_t n, float scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dy[i] = __ldg(out + i) > 0.f ? g[i] * scale : 0.f;
}
// dx[r, k] *= (x[r*ldx + k] > 0 ? scale : 0): the mask of the layer below applied to a dense dx (fallback of the fused
// dgrad epilogue of mlp_tma.cu)
__global__ void __launch_bounds__(256)
mask_rows_kernel(float* __restrict__ dx, const float* __restrict__ x, int64_t ldx, int64_t rows, int K, float scale) {
    const int64_t total = rows * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / K;
        const int k = (int)(i - r * K);
        dx[i] = __ldg(x + r * ldx + k) > 0.f ? dx[i] * scale : 0.f;
    }
}
// y[i] = keep(i) ? y[i] / (1-p) : 0 with the mask of common.cuh (fallback of the fused forward epilogue)
__global__ void __launch_bounds__(256)
dropout_kernel(float* __restrict__ y, int64_t n, const uint64_t* __restrict__ state, uint32_t thresh, float scale) {
    const uint64_t seed = state[0], ctr = state[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = dropout_keep(seed, ctr + (uint64_t)i, thresh) ? y[i] * scale : 0.f;
}
__global__ void rng_advance_kernel(uint64_t* state, uint64_t delta) { state[1] += delta; }

// ---- single-output layer (the tower's last Linear(200,1), p_model.py:290): GEMV-shaped, CUDA cores, HBM-bound
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ y, int64_t rows, int K, int relu) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float b0 = bias ? __ldg(bias) : 0.f;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s = fmaf(__ldg(x + r * ldx + k), __ldg(w + k), s);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(RLCTR_FULL, s, off);
        if (lane == 0) {
            float v = s + b0;
            if (relu) v = fmaxf(v, 0.f);
            y[r] = v;
        }
    }
}
// the same layer for 16-byte-aligned rows of at most 256 floats: a lane holds its (at most two) float4 of w, a warp takes four
// rows per trip with all eight 16-byte loads issued before the first use
__global__ void __launch_bounds__(256)
gemv_rows4_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int64_t rows, int K4, int relu) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const bool on0 = lane < K4, on1 = lane + 32 < K4;
    const float4 w0 = on0 ? __ldg(w4 + lane) : zero, w1 = on1 ? __ldg(w4 + lane + 32) : zero;
    const float b0 = bias ? __ldg(bias) : 0.f;
    for (int64_t r = warp0 * 4; r < rows; r += nwarps * 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4* xr = reinterpret_cast<const float4*>(x + (r + u) * ldx);
            const bool ok = r + u < rows;
            a[u] = (ok && on0) ? __ldg(xr + lane) : zero;
            b[u] = (ok && on1) ? __ldg(xr + lane + 32) : zero;
        }
        float s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float t = a[u].x * w0.x;
            t = fmaf(a[u].y, w0.y, t); t = fmaf(a[u].z, w0.z, t); t = fmaf(a[u].w, w0.w, t);
            t = fmaf(b[u].x, w1.x, t); t = fmaf(b[u].y, w1.y, t); t = fmaf(b[u].z, w1.z, t); t = fmaf(b[u].w, w1.w, t);
            s[u] = t;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(RLCTR_FULL, s[u], off);
        }
        if (lane < 4 && r + lane < rows) {
            float v = (lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3]) + b0;
            if (relu) v = fmaxf(v, 0.f);
            y[r + lane] = v;
        }
    }
}
__global__ void __launch_bounds__(256)
outer_rows_kernel(const float* __restrict__ gy, const float* __restrict__ w, float* __restrict__ dx, int64_t rows, int K,
                  const float* __restrict__ x, int64_t ldx, float scale) {
    const int64_t total = rows * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / K;
        const int k = (int)(i - r * K);
        float v = __ldg(gy + r) * __ldg(w + k);
        if (x) v = __ldg(x + r * ldx + k) > 0.f ? v * scale : 0.f;      // ReLU(+dropout) mask of the layer below, fused
        dx[i] = v;
    }
}


// ---- backward of a single-output layer in ONE pass over x: dx[r, :] = gy[r] * w (masked by the layer below), dw = sum_r gy[r] x[r, :],
// db = sum_r gy[r].  Warp per row, lane owns columns lane + 32c; block b owns rows [b*rpb, (b+1)*rpb); per-block partials
// [dw (K) | db] summed over the 8 warps in a fixed order; splitk_reduce_kernel adds the blocks.
template <int CMAX>
__global__ void __launch_bounds__(256)
gemv_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gy, const float* __restrict__ w,
                float* __restrict__ dx, float* __restrict__ part, int64_t rows, int K, int rows_per_block, int dx_mask,
                float dx_scale) {
    __shared__ float red[8][CMAX * 32 + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float wk[CMAX], acc[CMAX], gsum = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        const int k = lane + 32 * c;
        wk[c] = k < K ? __ldg(w + k) : 0.f;
        acc[c] = 0.f;
    }
    // four rows per trip: all of their loads are issued before the first use (the loop is latency-bound otherwise)
    for (int64_t rb = r0 + wib; rb < r1; rb += 32) {
        float g[4], xv[4][CMAX];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = rb + 8 * u;
            g[u] = r < r1 ? __ldg(gy + r) : 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                const int k = lane + 32 * c;
                xv[u][c] = (r < r1 && k < K) ? __ldg(x + r * ldx + k) : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = rb + 8 * u;
            gsum += g[u];
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                const int k = lane + 32 * c;
                acc[c] = fmaf(g[u], xv[u][c], acc[c]);
                if (dx && r < r1 && k < K) {
                    float v = g[u] * wk[c];
                    if (dx_mask) v = xv[u][c] > 0.f ? v * dx_scale : 0.f;
                    __stcs(dx + r * (int64_t)K + k, v);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) red[wib][lane + 32 * c] = acc[c];
    if (lane == 0) red[wib][CMAX * 32] = gsum;
    __syncthreads();
    for (int i = threadIdx.x; i <= K; i += blockDim.x) {
        const int col = i < K ? i : CMAX * 32;
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][col];
        part[(int64_t)blockIdx.x * (K + 1) + i] = t;
    }
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

// ---- exact fp32 GEMM on the CUDA cores (RLCTR_MLP_FP32) --------------------------------------------------------------------
// C[M,N] = sum_k A(m,k) * B(n,k) (+ bias[n]) (ReLU), strides in floats: A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk].
// One FFMA per product, k in ascending order: the reference's fp32 SGEMM arithmetic (error ~1e-7 sqrt(K) of sum |a b|), where
// the 3xTF32 tensor-core kernels carry ~2e-6 of the output SCALE.  That difference is invisible in a CTR tower but not in the
// learn steps of the BatchNorm policy nets: BatchNorm's backward subtracts the batch mean of a gradient whose dominant part is
// one constant (d mean(Q) / dQ), so a 2e-6-of-scale error becomes 5e-4 of what is left (measured against the reference's own
// modules in float64: tests/test_gpu_parity_scale.py).  Those steps run on replay batches of 32-256 rows -- a few MFLOP, where
// tensor cores buy nothing -- so small batches take this kernel.  64 x 64 tile, 256 threads, 4 x 4 outputs per thread.
namespace simt {
constexpr int TK = 16;
// TM x TN output tile, (TM/4) x (TN/4) threads with a 4 x 4 micro-tile each; every output element is ONE fma chain over k in
// ascending order whatever the tile shape, so the tile only decides how many blocks there are.  The learn steps of the policy
// nets are 256 x 300 x 300 problems: 64 x 64 tiles give 20 blocks on 148 SMs (48 us per GEMM, 3 ms of the C5 step over its 63
// launches); 32 x 32 tiles give 80.
template <int TM, int TN>
__global__ void __launch_bounds__((TM / 4) * (TN / 4))
sgemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbn, int64_t sbk,
             float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int M, int N, int K, int relu) {
    constexpr int NT = (TM / 4) * (TN / 4);
    constexpr int LA = TK * TM / NT, LB = TK * TN / NT;       // operand elements each thread fetches per k-tile
    __shared__ float As[TK][TM + 4], Bs[TK][TN + 4];
    const int tx = threadIdx.x % (TN / 4), ty = threadIdx.x / (TN / 4);
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // the next k-tile travels global -> registers while the current one is multiplied out of shared memory (a small problem puts
    // one block of two warps on an SM: nothing else would hide the load latency)
    float ra[LA], rb[LB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < LA; ++u) {
            const int e = threadIdx.x + u * NT;
            // the faster-varying index follows the operand's unit stride so that the global loads coalesce
            const int kk = sak == 1 ? e % TK : e / TM, mm = sak == 1 ? e / TK : e % TM;
            const int m = m0 + mm, k = k0 + kk;
            ra[u] = (m < M && k < K) ? __ldg(A + (int64_t)m * sam + (int64_t)k * sak) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < LB; ++u) {
            const int e = threadIdx.x + u * NT;
            const int kk = sbk == 1 ? e % TK : e / TN, nn = sbk == 1 ? e / TK : e % TN;
            const int n = n0 + nn, k = k0 + kk;
            rb[u] = (n < N && k < K) ? __ldg(B + (int64_t)n * sbn + (int64_t)k * sbk) : 0.f;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
        for (int u = 0; u < LA; ++u) {
            const int e = threadIdx.x + u * NT;
            const int kk = sak == 1 ? e % TK : e / TM, mm = sak == 1 ? e / TK : e % TM;
            As[kk][mm] = ra[u];
        }
#pragma unroll
        for (int u = 0; u < LB; ++u) {
            const int e = threadIdx.x + u * NT;
            const int kk = sbk == 1 ? e % TK : e / TN, nn = sbk == 1 ? e / TK : e % TN;
            Bs[kk][nn] = rb[u];
        }
        __syncthreads();
        if (k0 + TK < K) fetch(k0 + TK);
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            C[(int64_t)m * ldc + n] = v;
        }
    }
}
static int gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C, int64_t ldc,
                const float* bias, int M, int N, int K, int relu, cudaStream_t st) {
    const int64_t big = (int64_t)((N + 63) / 64) * ((M + 63) / 64);
    if (big >= 2 * RLCTR_SMS) {
        dim3 grid((N + 63) / 64, (M + 63) / 64);
        sgemm_kernel<64, 64><<<grid, 256, 0, st>>>(A, sam, sak, B, sbn, sbk, C, ldc, bias, M, N, K, relu);
    } else {
        dim3 grid((N + 31) / 32, (M + 31) / 32);
        sgemm_kernel<32, 32><<<grid, 64, 0, st>>>(A, sam, sak, B, sbn, sbk, C, ldc, bias, M, N, K, relu);
    }
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}
}  // namespace simt

static int launch_gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C,
                       int64_t ldc, const float* bias, int M, int N, int K, int relu, const GemmPlan& p, cudaStream_t st) {
    GemmArgs g;
    g.A = A; g.sam = sam; g.sak = sak;
    g.B = B; g.sbn = sbn; g.sbk = sbk;
    g.C = C; g.ldc = ldc; g.bias = bias;
    g.M = M; g.N = N; g.K = K;
    g.n_tile = p.n_tile; g.m_tiles = p.m_tiles; g.n_tiles = p.n_tiles; g.splits = p.splits; g.kb_per_split = p.kb_per_split;
    g.stages = p.stages; g.relu = relu;
    g.cvec = vec_of(C, ldc);
    g.mode_a = sak == 1 ? vec_of(A, sam) : 0;
    g.mode_b = sbk == 1 ? vec_of(B, sbn) : 0;
    RLCTR_CUDA(cudaFuncSetAttribute(gemm3x_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    const int total = p.m_tiles * p.n_tiles * p.splits;
    const int grid = total < RLCTR_SMS ? total : RLCTR_SMS;
    gemm3x_tf32_kernel<<<grid, GEMM_THREADS, p.smem, st>>>(g);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

constexpr int COLSUM_ROWS_PER_BLOCK = 128;

static inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
static inline int round4(int n) { return (n + 3) / 4 * 4; }

// workspace of one Linear call: [split-K partials of dW | column-sum partials | W_hi | W_lo]
struct MlpWs {
    size_t wgrad, colsum, wsplit, total;
    int in_pitch;
};
static MlpWs mlp_ws(int64_t batch, int in_dim, int out_dim) {
    MlpWs w;
    const GemmPlan p = plan_gemm(out_dim, in_dim, (int)batch, true);
    int splits = p.splits;
    const int s2 = tma::plan_splits(out_dim, in_dim, (int)batch, true);
    if (s2 > splits) splits = s2;
    w.wgrad = align256((size_t)splits * out_dim * in_dim * sizeof(float));
    const size_t yb = (size_t)((batch + COLSUM_ROWS_PER_BLOCK - 1) / COLSUM_ROWS_PER_BLOCK);
    w.colsum = align256(yb * ((size_t)out_dim + (out_dim == 1 ? (size_t)in_dim : 0)) * sizeof(float));
    w.in_pitch = round4(in_dim);
    w.wsplit = align256((size_t)out_dim * w.in_pitch * sizeof(float));
    w.total = w.wgrad + w.colsum + 2 * w.wsplit + 256;
    return w;
}

}  // namespace rlctr

using namespace rlctr;

extern "C" size_t rlctr_mlp_ws_bytes(int64_t batch, int32_t in_dim, int32_t out_dim) {
    if (batch <= 0 || in_dim <= 0 || out_dim <= 0) return 256;
    return mlp_ws(batch, in_dim, out_dim).total;
}

// (W_hi, W_lo) images of the weight at the end of the workspace; null when the workspace is absent / too small
// reuse: the images are already there (RLCTR_MLP_W_PRESPLIT: this workspace served the layer's forward call with the same weights)
static bool split_weights_into(void* ws, size_t ws_bytes, const MlpWs& l, const float* w, int in_dim, int out_dim,
                               cudaStream_t st, float** whi, float** wlo, int* rc, bool reuse = false) {
    *rc = RLCTR_OK;
    if (!tma::enabled() || !ws || ws_bytes < l.total || !rlctr_aligned16(ws) || out_dim == 1) return false;
    char* base = reinterpret_cast<char*>(ws) + l.wgrad + l.colsum;
    *whi = reinterpret_cast<float*>(base);
    *wlo = reinterpret_cast<float*>(base + l.wsplit);
    if (!reuse) *rc = tma::split_weight(w, *whi, *wlo, out_dim, in_dim, l.in_pitch, st);
    return *rc == RLCTR_OK;
}

static int grid_elems(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    return (int)(blocks < RLCTR_SMS * 8 ? (blocks < 1 ? 1 : blocks) : RLCTR_SMS * 8);
}

extern "C" int rlctr_rng_advance(uint64_t* state, uint64_t delta, rlctr_stream_t stream) {
    if (!state) return RLCTR_EINVAL;
    rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, delta);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_linear_fwd(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t batch,
                                int32_t in_dim, int32_t out_dim, int32_t flags, float dropout_p, const uint64_t* rng_state,
                                void* ws, size_t ws_bytes, rlctr_stream_t stream) {
    if (!x || !w || !y || batch < 0 || in_dim <= 0 || out_dim <= 0) return RLCTR_EINVAL;
    if (ldx == 0) ldx = in_dim;
    if (ldx < in_dim) return RLCTR_EINVAL;
    const bool drop = (flags & RLCTR_MLP_DROPOUT) != 0 && dropout_p > 0.f;
    if (drop && (!rng_state || !(dropout_p < 1.f))) return RLCTR_EINVAL;
    if (batch == 0) return RLCTR_OK;
    if (batch > 0x7fffffff) return RLCTR_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int relu = (flags & RLCTR_MLP_RELU) ? 1 : 0;
    const uint32_t thresh = dropout_thresh(dropout_p);
    const float dscale = drop ? 1.0f / (1.0f - dropout_p) : 1.f;
    bool fused_drop = false;
    int rc = RLCTR_OK;
    if (out_dim == 1) {
        if (in_dim % 4 == 0 && in_dim <= 256 && ldx % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0) {
            const int64_t blocks = (batch + 31) / 32;                      // 8 warps x 4 rows
            gemv_rows4_kernel<<<(unsigned)(blocks < RLCTR_SMS * 8 ? blocks : RLCTR_SMS * 8), 256, 0, st>>>(
                x, ldx, w, bias, y, batch, in_dim / 4, relu);
        } else {
            const int64_t blocks = (batch + 7) / 8;
            gemv_rows_kernel<<<(unsigned)(blocks < RLCTR_SMS * 8 ? blocks : RLCTR_SMS * 8), 256, 0, st>>>(
                x, ldx, w, bias, y, batch, in_dim, relu);
        }
        RLCTR_LAUNCH_CHECK();
    } else if (flags & RLCTR_MLP_FP32) {
        rc = simt::gemm(x, ldx, 1, w, in_dim, 1, y, out_dim, bias, (int)batch, out_dim, in_dim, relu, st);
        if (rc) return rc;
    } else {
        const MlpWs l = mlp_ws(batch, in_dim, out_dim);
        float *whi = nullptr, *wlo = nullptr;
        rc = RLCTR_EUNSUPPORTED;
        int src = RLCTR_OK;
        if (split_weights_into(ws, ws_bytes, l, w, in_dim, out_dim, st, &whi, &wlo, &src)) {
            tma::Epilogue epi;
            if (drop) { epi.drop_state = rng_state; epi.drop_thresh = thresh; epi.drop_scale = dscale; }
            rc = tma::gemm(tma::Operand{x, nullptr, ldx, false}, tma::Operand{whi, wlo, l.in_pitch, false}, y, out_dim, bias,
                           (int)batch, out_dim, in_dim, relu, false, st, &epi);
            if (rc == RLCTR_OK) fused_drop = drop;
        } else if (src) {
            return src;
        }
        if (rc == RLCTR_EUNSUPPORTED) {
            GemmPlan p = plan_gemm((int)batch, out_dim, in_dim, false);
            rc = launch_gemm(x, ldx, 1, w, in_dim, 1, y, out_dim, bias, (int)batch, out_dim, in_dim, relu, p, st);
        }
        if (rc) return rc;
    }
    if (drop && !fused_drop) {
        const int64_t n = batch * out_dim;
        dropout_kernel<<<grid_elems(n), 256, 0, st>>>(y, n, rng_state, thresh, dscale);
        RLCTR_LAUNCH_CHECK();
    }
    return RLCTR_OK;
}

extern "C" int rlctr_linear_bwd(const float* x, int64_t ldx, const float* w, const float* y, float* gy, float* dx, float* dw,
                                float* db, int64_t batch, int32_t in_dim, int32_t out_dim, int32_t flags, float gy_scale,
                                float dx_scale, void* ws, size_t ws_bytes, rlctr_stream_t stream) {
    if (!x || !w || !gy || batch <= 0 || in_dim <= 0 || out_dim <= 0) return RLCTR_EINVAL;
    if (ldx == 0) ldx = in_dim;
    if (ldx < in_dim) return RLCTR_EINVAL;
    if (batch > 0x7fffffff) return RLCTR_EUNSUPPORTED;
    if ((flags & RLCTR_MLP_RELU) && !y) return RLCTR_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const float* dy = gy;
    if (flags & RLCTR_MLP_RELU) {                       // ReLU backward on the saved output, in place on gy
        const int64_t n = batch * out_dim;
        int64_t blocks = (n + 255) / 256;
        int grid = (int)(blocks < RLCTR_SMS * 8 ? blocks : RLCTR_SMS * 8);
        relu_bwd_kernel<<<grid, 256, 0, st>>>(gy, y, gy, n, gy_scale);
        RLCTR_LAUNCH_CHECK();
    }
    const bool dx_mask = (flags & RLCTR_MLP_DX_MASK) != 0;
    const MlpWs l = mlp_ws(batch, in_dim, out_dim);
    if ((dw || db) && (!ws || ws_bytes < l.total)) return RLCTR_EWORKSPACE;
    float* part = reinterpret_cast<float*>(ws);
    float* cpart = ws ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + l.wgrad) : nullptr;
    const int yb = (int)((batch + COLSUM_ROWS_PER_BLOCK - 1) / COLSUM_ROWS_PER_BLOCK);
    if (out_dim == 1 && in_dim <= 512) {
        // one pass over x: dx, dw and db together (the ReLU/dropout mask of the layer below is read off the same x)
        if (!ws || ws_bytes < l.total) return RLCTR_EWORKSPACE;
        if (in_dim <= 256)
            gemv_bwd_kernel<8><<<yb, 256, 0, st>>>(x, ldx, dy, w, dx, cpart, batch, in_dim, COLSUM_ROWS_PER_BLOCK, dx_mask ? 1 : 0,
                                                   dx_scale);
        else
            gemv_bwd_kernel<16><<<yb, 256, 0, st>>>(x, ldx, dy, w, dx, cpart, batch, in_dim, COLSUM_ROWS_PER_BLOCK, dx_mask ? 1 : 0,
                                                    dx_scale);
        RLCTR_LAUNCH_CHECK();
        if (dw || db) {
            colsum_parts_kernel<<<colsum_parts_grid(in_dim + 1), 256, 0, st>>>(cpart, dw, db, in_dim, 1, in_dim + 1, yb);
            RLCTR_LAUNCH_CHECK();
        }
        return RLCTR_OK;
    }
    if (out_dim == 1) {
        if (dx) {
            const int64_t n = batch * in_dim;
            int64_t blocks = (n + 255) / 256;
            outer_rows_kernel<<<(unsigned)(blocks < RLCTR_SMS * 8 ? blocks : RLCTR_SMS * 8), 256, 0, st>>>(
                dy, w, dx, batch, in_dim, dx_mask ? x : nullptr, ldx, dx_scale);
            RLCTR_LAUNCH_CHECK();
        }
        if (dw) {
            dim3 grid((in_dim + 31) / 32, yb);
            colsum_partial_kernel<<<grid, 256, 0, st>>>(x, ldx, dy, cpart, batch, in_dim, COLSUM_ROWS_PER_BLOCK);
            RLCTR_LAUNCH_CHECK();
            colsum_parts_kernel<<<colsum_parts_grid(in_dim), 256, 0, st>>>(cpart, dw, nullptr, in_dim, 0, in_dim, yb);
            RLCTR_LAUNCH_CHECK();
        }
        if (db) {
            float* part2 = cpart + (size_t)yb * in_dim;
            dim3 grid(1, yb);
            colsum_partial_kernel<<<grid, 256, 0, st>>>(dy, 1, nullptr, part2, batch, 1, COLSUM_ROWS_PER_BLOCK);
            RLCTR_LAUNCH_CHECK();
            colsum_parts_kernel<<<1, 256, 0, st>>>(part2, db, nullptr, 1, 0, 1, yb);
            RLCTR_LAUNCH_CHECK();
        }
        return RLCTR_OK;
    }
    const bool fp32 = (flags & RLCTR_MLP_FP32) != 0;
    if (dx && fp32) {
        int rc = simt::gemm(dy, out_dim, 1, w, 1, in_dim, dx, in_dim, nullptr, (int)batch, in_dim, out_dim, 0, st);
        if (rc) return rc;
        if (dx_mask) {
            mask_rows_kernel<<<grid_elems(batch * in_dim), 256, 0, st>>>(dx, x, ldx, batch, in_dim, dx_scale);
            RLCTR_LAUNCH_CHECK();
        }
    } else if (dx) {   // dX[B,in] = dY[B,out] * W[out,in]:  B operand = W, contiguous along N
        float *whi = nullptr, *wlo = nullptr;
        int rc = RLCTR_OK;
        bool done = false;
        if (split_weights_into(ws, ws_bytes, l, w, in_dim, out_dim, st, &whi, &wlo, &rc, (flags & RLCTR_MLP_W_PRESPLIT) != 0)) {
            tma::Epilogue epi;
            if (dx_mask) { epi.mask_src = x; epi.mask_ld = ldx; epi.mask_scale = dx_scale; }
            rc = tma::gemm(tma::Operand{dy, nullptr, out_dim, false}, tma::Operand{whi, wlo, l.in_pitch, true}, dx, in_dim,
                           nullptr, (int)batch, in_dim, out_dim, 0, false, st, &epi);
            if (rc == RLCTR_OK) done = true;
            else if (rc != RLCTR_EUNSUPPORTED) return rc;
        } else if (rc) {
            return rc;
        }
        if (!done) {
            GemmPlan p = plan_gemm((int)batch, in_dim, out_dim, false);
            rc = launch_gemm(dy, out_dim, 1, w, 1, in_dim, dx, in_dim, nullptr, (int)batch, in_dim, out_dim, 0, p, st);
            if (rc) return rc;
            if (dx_mask) {
                mask_rows_kernel<<<grid_elems(batch * in_dim), 256, 0, st>>>(dx, x, ldx, batch, in_dim, dx_scale);
                RLCTR_LAUNCH_CHECK();
            }
        }
    }
    bool db_done = false;
    const float* dbp_pending = nullptr;
    if (dw && fp32) {   // one pass, k = batch rows in ascending order
        int rc = simt::gemm(dy, 1, out_dim, x, 1, ldx, dw, in_dim, nullptr, out_dim, in_dim, (int)batch, 0, st);
        if (rc) return rc;
    } else if (dw) {   // dW[out,in] = dY^T[out,B] * X[B,in]: both operands contiguous along M / N, K = batch; split-K
        const int64_t mn = (int64_t)out_dim * in_dim;
        int splits = tma::enabled() ? tma::plan_splits(out_dim, in_dim, (int)batch, true) : 0;
        int rc = RLCTR_EUNSUPPORTED;
        if (splits >= 1) {
            // the converter warps of the TMA kernel sum dY's columns while they split the tile: db comes for free
            float* dbp = (db && (size_t)splits * out_dim * sizeof(float) <= l.colsum) ? cpart : nullptr;
            rc = tma::gemm(tma::Operand{dy, nullptr, out_dim, true}, tma::Operand{x, nullptr, ldx, true}, splits == 1 ? dw : part,
                           in_dim, nullptr, out_dim, in_dim, (int)batch, 0, true, st, nullptr, dbp);
            if (rc == RLCTR_OK && dbp) {
                if (splits > 1) {
                    dbp_pending = dbp;                   // reduced together with dW below
                } else {
                    splitk_reduce_kernel<<<(out_dim + 255) / 256, 256, 0, st>>>(dbp, db, out_dim, splits);
                    RLCTR_LAUNCH_CHECK();
                }
                db_done = true;
            }
        }
        if (rc == RLCTR_EUNSUPPORTED) {
            GemmPlan p = plan_gemm(out_dim, in_dim, (int)batch, true);
            splits = p.splits;
            rc = launch_gemm(dy, 1, out_dim, x, 1, ldx, splits == 1 ? dw : part, in_dim, nullptr, out_dim, in_dim, (int)batch,
                             0, p, st);
        }
        if (rc) return rc;
        if (splits > 1 && dbp_pending) {
            int64_t blocks = (mn + out_dim + 255) / 256;
            splitk_reduce2_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, st>>>(part, dw, mn, dbp_pending, db, out_dim, splits);
            RLCTR_LAUNCH_CHECK();
        } else if (splits > 1) {
            int64_t blocks = (mn + 255) / 256;
            splitk_reduce_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, st>>>(part, dw, mn, splits);
            RLCTR_LAUNCH_CHECK();
        }
    }
    if (db && !db_done) {
        dim3 grid((out_dim + 31) / 32, yb);
        colsum_partial_kernel<<<grid, 256, 0, st>>>(dy, out_dim, nullptr, cpart, batch, out_dim, COLSUM_ROWS_PER_BLOCK);
        RLCTR_LAUNCH_CHECK();
        colsum_parts_kernel<<<colsum_parts_grid(out_dim), 256, 0, st>>>(cpart, db, nullptr, out_dim, 0, out_dim, yb);
        RLCTR_LAUNCH_CHECK();
    }
    return RLCTR_OK;
}
