// embed.cu -- K1 fused gather + first-order + FM second-order forward, the plain row gather,
// and K5 the Feature_Embedding RL state encoder.  sm_100a.
//
// Thread mapping (all three kernels): one warp per sample, LPR lanes per gathered row, each
// lane reading one aligned float4 (128-bit) chunk of the fused row.  For the default
// latent_dims=10 the row is [w, v0..v9, pad] = 12 floats = 3 chunks -> LPR=4, 8 rows per warp
// request, the 15 fields of a sample in two requests.  Lanes of one row hit one or two 32 B
// sectors in a single request (1.25 L1 wavefronts/row instead of 3 for a lane-per-row map).
#include "common.cuh"

namespace rlctr {

// id of gathered element e = b * fields + f; ids == NULL: the rows are already in sample order (rlctr_rows_lookup's `gathered`)
__device__ __forceinline__ int64_t id_at(const int64_t* __restrict__ ids, int64_t e) { return ids ? __ldg(ids + e) : e; }

// ------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256, 5)
embed_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ tab, const __grid_constant__ ShardView sv, int64_t n_rows, int pitch,
                 int rs, int lin_col, int emb_col, int dim, const float* __restrict__ bias,
                 float* __restrict__ logit, float* __restrict__ pctr, int64_t pctr_stride,
                 float* __restrict__ sums, float* __restrict__ rows_out, int64_t rows_pitch,
                 int64_t batch, int fields, int flags) {
    constexpr int RPW = 32 / LPR;                   // rows per warp request
    const int lane = threadIdx.x & 31;
    const int c = lane % LPR;                       // chunk of the row this lane owns
    const int g = lane / LPR;                       // row slot within the request
    const int col0 = 4 * c;
    const bool chunk_on = col0 < rs;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const bool fm = (flags & RLCTR_FM_TERM) != 0;

    bool is_emb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) is_emb[k] = (col0 + k >= emb_col) && (col0 + k < emb_col + dim);
    const float b0 = bias ? __ldg(bias) : 0.f;

    // software pipeline: the ids of this warp's NEXT sample are fetched while the rows of the
    // current one are in flight (two dependent DRAM latencies -> one on the critical path).
    int64_t b = warp0;
    int64_t id_a = -1, id_b = -1;                   // first two row slots of the next sample
    if (b < batch && chunk_on) {
        if (g < fields) id_a = id_at(ids, b * fields + g);
        if (g + RPW < fields) id_b = id_at(ids, b * fields + g + RPW);
    }
    for (; b < batch; b += nwarps) {
        const int64_t cur_a = id_a, cur_b = id_b;
        const int64_t nb = b + nwarps;
        id_a = -1; id_b = -1;
        if (nb < batch && chunk_on) {
            if (g < fields) id_a = id_at(ids, nb * fields + g);
            if (g + RPW < fields) id_b = id_at(ids, nb * fields + g + RPW);
        }
        float4 s = f4zero();
        float q = 0.f;
        for (int f0 = 0; f0 < fields; f0 += 2 * RPW) {
            const int fa = f0 + g, fb = f0 + RPW + g;
            int64_t ia = -1, ib = -1;
            if (f0 == 0) { ia = cur_a; ib = cur_b; }
            else if (chunk_on) {
                if (fa < fields) ia = id_at(ids, b * fields + fa);
                if (fb < fields) ib = id_at(ids, b * fields + fb);
            }
            float4 ra = f4zero(), rb = f4zero();
            const bool oka = (uint64_t)ia < (uint64_t)n_rows, okb = (uint64_t)ib < (uint64_t)n_rows;
            if (oka) ra = ldg4_once(row_ptr(tab, sv, ia, pitch) + col0);
            if (okb) rb = ldg4_once(row_ptr(tab, sv, ib, pitch) + col0);
            s = f4add(s, f4add(ra, rb));
            if (fm) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (is_emb[k]) { float x = f4get(ra, k), y = f4get(rb, k); q = fmaf(x, x, q); q = fmaf(y, y, q); }
            }
            if (rows_out) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!is_emb[k]) continue;
                    const int d = col0 + k - emb_col;
                    if (chunk_on && fa < fields) rows_out[b * rows_pitch + fa * dim + d] = f4get(ra, k);
                    if (chunk_on && fb < fields) rows_out[b * rows_pitch + fb * dim + d] = f4get(rb, k);
                }
            }
        }
        // sum over the F rows: lanes owning the same chunk sit LPR apart
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) {
            s = f4add(s, shfl_xor4(s, off));
            q += __shfl_xor_sync(RLCTR_FULL, q, off);
        }
        float t = 0.f;
        if (fm) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (is_emb[k]) { float x = f4get(s, k); t = fmaf(x, x, t); }
            t -= q;                                  // sum over this chunk of (S_d^2 - sum_f v_d^2)
#pragma unroll
            for (int off = 1; off < LPR; off <<= 1) t += __shfl_xor_sync(RLCTR_FULL, t, off);
        }
        float first = 0.f;
        if (lin_col >= 0) {
            const int src = lin_col >> 2;            // lane (group 0) holding the linear column
            const float4 sv = make_float4(__shfl_sync(RLCTR_FULL, s.x, src), __shfl_sync(RLCTR_FULL, s.y, src),
                                          __shfl_sync(RLCTR_FULL, s.z, src), __shfl_sync(RLCTR_FULL, s.w, src));
            first = f4get(sv, lin_col & 3);
        }
        if (lane == 0) {
            const float z = b0 + first + 0.5f * t;
            if (logit) logit[b] = z;
            if (pctr) pctr[b * pctr_stride] = sigmoidf_ref(z);
        }
        if (sums && g == 0 && chunk_on) st4(sums + b * rs + col0, s);
    }
}

// LR table: one float per row (row_stride == 1).  Lane f gathers w[x_f]; warp-sum.
__global__ void __launch_bounds__(256)
embed_fwd_scalar_kernel(const int64_t* __restrict__ ids, const float* __restrict__ tab, const __grid_constant__ ShardView sv, int64_t n_rows, int pitch,
                        const float* __restrict__ bias, float* __restrict__ logit,
                        float* __restrict__ pctr, int64_t pctr_stride, float* __restrict__ sums,
                        int64_t batch, int fields) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float b0 = bias ? __ldg(bias) : 0.f;
    for (int64_t b = warp0; b < batch; b += nwarps) {
        float s = 0.f;
        for (int f = lane; f < fields; f += 32) {
            const int64_t id = id_at(ids, b * fields + f);
            if ((uint64_t)id < (uint64_t)n_rows) s += ldg1_once(row_ptr(tab, sv, id, pitch));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(RLCTR_FULL, s, off);
        if (lane == 0) {
            const float z = b0 + s;
            if (logit) logit[b] = z;
            if (pctr) pctr[b * pctr_stride] = sigmoidf_ref(z);
            if (sums) sums[b] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------
// plain gather
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_rows_kernel(const int64_t* __restrict__ ids, int64_t n, const float* __restrict__ tab,
                   int64_t n_rows, int pitch, int rs, float* __restrict__ out) {
    const int chunks = rs >> 2;
    const int64_t total = n * chunks;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / chunks;
        const int c = (int)(t - k * chunks);
        const int64_t id = __ldg(ids + k);
        float4 r = f4zero();
        if ((uint64_t)id < (uint64_t)n_rows) r = ldg4(tab + id * pitch + 4 * c);
        st4_stream(out + k * rs + 4 * c, r);
    }
}
__global__ void __launch_bounds__(256)
gather_rows_scalar_kernel(const int64_t* __restrict__ ids, int64_t n, const float* __restrict__ tab,
                          int64_t n_rows, int pitch, float* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = __ldg(ids + t);
        out[t] = ((uint64_t)id < (uint64_t)n_rows) ? __ldg(tab + id * pitch) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// K5  Feature_Embedding: gather, stage the F latent vectors of the sample in shared memory
// (F*D contiguous floats == the second block of the output row), P pairwise dots from smem.
// ------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(128)
featemb_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ tab, int64_t n_rows, int pitch,
                   int rs, int emb_col, int dim, float* __restrict__ out, int64_t out_stride,
                   int64_t batch, int fields) {
    constexpr int RPW = 32 / LPR;
    extern __shared__ float smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;
    const int fd = fields * dim;
    const int npair = fields * (fields - 1) / 2;
    float* stage = smem + wib * fd;
    unsigned char* pi = reinterpret_cast<unsigned char*>(smem + nw * fd);
    unsigned char* pj = pi + npair;
    for (int i = threadIdx.x; i < fields - 1; i += blockDim.x) {
        int base = i * fields - i * (i + 1) / 2;     // pairs (i, i+1..F-1), Feature_embedding.py:40-43
        for (int j = i + 1; j < fields; ++j) { pi[base + j - i - 1] = (unsigned char)i; pj[base + j - i - 1] = (unsigned char)j; }
    }
    __syncthreads();
    const int c = lane % LPR, g = lane / LPR, col0 = 4 * c;
    const bool chunk_on = col0 < rs;
    const int64_t warp0 = (int64_t)blockIdx.x * nw + wib;
    const int64_t nwarps = (int64_t)gridDim.x * nw;
    for (int64_t b = warp0; b < batch; b += nwarps) {
        for (int f0 = 0; f0 < fields; f0 += 2 * RPW) {
            const int fa = f0 + g, fb = f0 + RPW + g;
            int64_t ia = -1, ib = -1;
            if (chunk_on && fa < fields) ia = __ldg(ids + b * fields + fa);
            if (chunk_on && fb < fields) ib = __ldg(ids + b * fields + fb);
            float4 ra = f4zero(), rb = f4zero();
            if ((uint64_t)ia < (uint64_t)n_rows) ra = ldg4_once(tab + ia * pitch + col0);
            if ((uint64_t)ib < (uint64_t)n_rows) rb = ldg4_once(tab + ib * pitch + col0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int d = col0 + k - emb_col;
                if (d < 0 || d >= dim || !chunk_on) continue;
                if (fa < fields) stage[fa * dim + d] = f4get(ra, k);
                if (fb < fields) stage[fb * dim + d] = f4get(rb, k);
            }
        }
        __syncwarp();
        float* o = out + b * out_stride;
        for (int p = lane; p < npair; p += 32) {
            const float* vi = stage + pi[p] * dim;
            const float* vj = stage + pj[p] * dim;
            float acc = 0.f;
            for (int d = 0; d < dim; ++d) acc += vi[d] * vj[d];    // mul then sum over d, in d order
            __stcs(o + p, acc);
        }
        for (int i = lane; i < fd; i += 32) __stcs(o + npair + i, stage[i]);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Pairwise inner products over already-gathered rows (InnerPNN, p_model.py:178-198): the tower input
// [E (F*D) | ip (P)] (or [ip | E], the Feature_Embedding order) and its backward
//     d E[i] = g_E[i] + sum_{j != i} g_ip[pair(i,j)] * E[j]
// Warp per sample; E and g_ip are staged in shared memory; pair index by a byte table.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_tables(unsigned char* pi, unsigned char* pj, unsigned char* pidx, int fields) {
    for (int i = threadIdx.x; i < fields; i += blockDim.x) {
        const int base = i * fields - i * (i + 1) / 2;       // pairs (i, i+1..F-1) in the reference's order
        for (int j = i + 1; j < fields; ++j) {
            const int p = base + j - i - 1;
            if (pi) { pi[p] = (unsigned char)i; pj[p] = (unsigned char)j; }
            if (pidx) { pidx[i * fields + j] = (unsigned char)p; pidx[j * fields + i] = (unsigned char)p; }
        }
    }
}
__global__ void __launch_bounds__(128)
pairdots_fwd_kernel(const float* __restrict__ rows, int64_t ld_rows, float* __restrict__ out, int64_t ld_out, int64_t batch,
                    int fields, int dim, int ip_first) {
    extern __shared__ float smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int fd = fields * dim, npair = fields * (fields - 1) / 2;
    float* stage = smem + wib * fd;
    unsigned char* pi = reinterpret_cast<unsigned char*>(smem + nw * fd);
    unsigned char* pj = pi + npair;
    pair_tables(pi, pj, nullptr, fields);
    __syncthreads();
    const int e_off = ip_first ? npair : 0, ip_off = ip_first ? 0 : fd;
    for (int64_t b = (int64_t)blockIdx.x * nw + wib; b < batch; b += (int64_t)gridDim.x * nw) {
        const float* r = rows + b * ld_rows;
        float* o = out + b * ld_out;
        for (int i = lane; i < fd; i += 32) { const float v = __ldg(r + i); stage[i] = v; o[e_off + i] = v; }
        __syncwarp();
        for (int p = lane; p < npair; p += 32) {
            const float* vi = stage + pi[p] * dim;
            const float* vj = stage + pj[p] * dim;
            float acc = 0.f;
            for (int d = 0; d < dim; ++d) acc += vi[d] * vj[d];      // mul then sum over d, in d order
            o[ip_off + p] = acc;
        }
        __syncwarp();
    }
}
__global__ void __launch_bounds__(128)
pairdots_bwd_kernel(const float* __restrict__ rows, int64_t ld_rows, const float* __restrict__ gout, int64_t ld_g,
                    float* __restrict__ grows, int64_t ld_grows, int64_t batch, int fields, int dim, int ip_first) {
    extern __shared__ float smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int fd = fields * dim, npair = fields * (fields - 1) / 2;
    float* stage = smem + wib * (fd + npair);
    float* gip = stage + fd;
    unsigned char* pidx = reinterpret_cast<unsigned char*>(smem + nw * (fd + npair));
    pair_tables(nullptr, nullptr, pidx, fields);
    __syncthreads();
    const int e_off = ip_first ? npair : 0, ip_off = ip_first ? 0 : fd;
    for (int64_t b = (int64_t)blockIdx.x * nw + wib; b < batch; b += (int64_t)gridDim.x * nw) {
        const float* r = rows + b * ld_rows;
        const float* g = gout + b * ld_g;
        for (int i = lane; i < fd; i += 32) stage[i] = __ldg(r + i);
        for (int p = lane; p < npair; p += 32) gip[p] = __ldg(g + ip_off + p);
        __syncwarp();
        for (int t = lane; t < fd; t += 32) {
            const int i = t / dim, d = t - i * dim;
            const unsigned char* prow = pidx + i * fields;
            float a0 = __ldg(g + e_off + t), a1 = 0.f;                  // two chains: a string of dependent LDS -> FMA otherwise
            int j = 0;
            for (; j + 1 < fields; j += 2) {
                if (j != i) a0 = fmaf(gip[prow[j]], stage[j * dim + d], a0);
                if (j + 1 != i) a1 = fmaf(gip[prow[j + 1]], stage[(j + 1) * dim + d], a1);
            }
            if (j < fields && j != i) a0 = fmaf(gip[prow[j]], stage[j * dim + d], a0);
            grows[b * ld_grows + t] = a0 + a1;
        }
        __syncwarp();
    }
}

static int grid_for_warps(int64_t items, int warps_per_block, int blocks_per_sm) {
    int64_t want = (items + warps_per_block - 1) / warps_per_block;
    int64_t cap = (int64_t)RLCTR_SMS * blocks_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace rlctr

using namespace rlctr;

static inline int pitch_of(const rlctr_table* t) { return t->row_pitch > 0 ? t->row_pitch : t->row_stride; }

static int check_table(const rlctr_table* t) {
    if (!t || !t->data || t->n_rows <= 0) return RLCTR_EINVAL;
    if (pitch_of(t) < t->row_stride || (t->row_stride != 1 && pitch_of(t) % 4 != 0)) return RLCTR_EUNSUPPORTED;
    if (t->row_stride != 1 && (t->row_stride % 4 != 0 || t->row_stride <= 0)) return RLCTR_EUNSUPPORTED;
    if (t->row_stride != 1 && !rlctr_aligned16(t->data)) return RLCTR_EALIGN;
    if (t->dim < 0 || t->emb_col < 0 || t->emb_col + t->dim > t->row_stride) return RLCTR_EINVAL;
    if (t->lin_col >= t->row_stride) return RLCTR_EINVAL;
    return RLCTR_OK;
}

extern "C" int rlctr_embed_fwd(const int64_t* ids, const rlctr_table* table, const float* bias,
                               float* logit, float* pctr, int64_t pctr_stride, float* sums,
                               float* rows_out, int64_t rows_pitch, int64_t batch, int32_t fields, int32_t flags,
                               rlctr_stream_t stream) {
    int rc = check_table(table);
    if (rc) return rc;
    if (batch < 0 || fields <= 0) return RLCTR_EINVAL;
    if (!ids && (table->world > 1 || table->n_rows < batch * fields)) return RLCTR_EINVAL;   // sample-ordered rows: local
    if (batch == 0) return RLCTR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int rs = table->row_stride;
    ShardView sv;
    if (!shard_view_of(table, &sv)) return RLCTR_EUNSUPPORTED;
    if (pctr && pctr_stride < 1) return RLCTR_EINVAL;
    if (rows_pitch == 0) rows_pitch = (int64_t)fields * table->dim;
    if (rows_out && rows_pitch < (int64_t)fields * table->dim) return RLCTR_EINVAL;
    if (rs == 1) {
        if (rows_out || (flags & RLCTR_FM_TERM) || table->lin_col != 0) return RLCTR_EUNSUPPORTED;
        int grid = grid_for_warps(batch, 8, 8);
        embed_fwd_scalar_kernel<<<grid, 256, 0, st>>>(ids, table->data, sv, table->n_rows, pitch_of(table), bias, logit, pctr,
                                                     pctr_stride, sums, batch, fields);
        RLCTR_LAUNCH_CHECK();
        return RLCTR_OK;
    }
    if (rs > 32) return RLCTR_EUNSUPPORTED;
    if (sums && !rlctr_aligned16(sums)) return RLCTR_EALIGN;
    int grid = grid_for_warps(batch, 8, 5);      // 5 resident CTAs of 8 warps per SM (48 registers)
#define LAUNCH_EMBED(L)                                                                              \
    embed_fwd_kernel<L><<<grid, 256, 0, st>>>(ids, table->data, sv, table->n_rows, pitch_of(table), rs, table->lin_col,   \
                                              table->emb_col, table->dim, bias, logit, pctr,         \
                                              pctr_stride, sums, rows_out, rows_pitch, batch, fields, flags)
    switch (rlctr_lanes_per_row(rs)) {
        case 1: LAUNCH_EMBED(1); break;
        case 2: LAUNCH_EMBED(2); break;
        case 4: LAUNCH_EMBED(4); break;
        default: LAUNCH_EMBED(8); break;
    }
#undef LAUNCH_EMBED
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_gather_rows(const int64_t* ids, int64_t n, const rlctr_table* table, float* out,
                                 rlctr_stream_t stream) {
    int rc = check_table(table);
    if (rc) return rc;
    if (!ids || !out || n < 0) return RLCTR_EINVAL;
    if (n == 0) return RLCTR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int rs = table->row_stride;
    if (rs == 1) {
        int64_t blocks = (n + 255) / 256;
        int grid = (int)(blocks < RLCTR_SMS * 8 ? blocks : RLCTR_SMS * 8);
        gather_rows_scalar_kernel<<<grid, 256, 0, st>>>(ids, n, table->data, table->n_rows, pitch_of(table), out);
    } else {
        if (!rlctr_aligned16(out)) return RLCTR_EALIGN;
        int64_t blocks = (n * (rs / 4) + 255) / 256;
        int grid = (int)(blocks < RLCTR_SMS * 8 ? blocks : RLCTR_SMS * 8);
        gather_rows_kernel<<<grid, 256, 0, st>>>(ids, n, table->data, table->n_rows, pitch_of(table), rs, out);
    }
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_featemb_fwd(const int64_t* ids, const rlctr_table* table, float* out,
                                 int64_t out_stride, int64_t batch, int32_t fields,
                                 rlctr_stream_t stream) {
    int rc = check_table(table);
    if (rc) return rc;
    if (!ids || !out || batch < 0 || fields < 2) return RLCTR_EINVAL;
    const int rs = table->row_stride, dim = table->dim;
    if (rs == 1 || rs > 32 || dim <= 0 || fields > 255) return RLCTR_EUNSUPPORTED;
    const int npair = fields * (fields - 1) / 2;
    if (out_stride < npair + fields * dim) return RLCTR_EINVAL;
    if (batch == 0) return RLCTR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int nw = 4;
    size_t smem = (size_t)nw * fields * dim * sizeof(float) + 2 * (size_t)npair;
    if (smem > 48 * 1024) return RLCTR_EUNSUPPORTED;
    int grid = grid_for_warps(batch, nw, 16);
#define LAUNCH_FE(L)                                                                               \
    featemb_fwd_kernel<L><<<grid, nw * 32, smem, st>>>(ids, table->data, table->n_rows, pitch_of(table), rs, \
                                                       table->emb_col, dim, out, out_stride, batch, fields)
    switch (rlctr_lanes_per_row(rs)) {
        case 1: LAUNCH_FE(1); break;
        case 2: LAUNCH_FE(2); break;
        case 4: LAUNCH_FE(4); break;
        default: LAUNCH_FE(8); break;
    }
#undef LAUNCH_FE
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}


extern "C" int rlctr_pairdots_fwd(const float* rows, int64_t ld_rows, float* out, int64_t ld_out, int64_t batch,
                                  int32_t fields, int32_t dim, int32_t ip_first, rlctr_stream_t stream) {
    if (!rows || !out || batch < 0 || fields < 2 || dim <= 0) return RLCTR_EINVAL;
    if (fields > 23) return RLCTR_EUNSUPPORTED;                         // pair indices are bytes: F <= 23 (P <= 253)
    const int fd = fields * dim, npair = fields * (fields - 1) / 2;
    if (ld_rows < fd || ld_out < fd + npair) return RLCTR_EINVAL;
    if (batch == 0) return RLCTR_OK;
    const int nw = 4;
    const size_t smem = (size_t)nw * fd * sizeof(float) + 2 * (size_t)npair;
    if (smem > 48 * 1024) return RLCTR_EUNSUPPORTED;
    pairdots_fwd_kernel<<<grid_for_warps(batch, nw, 16), nw * 32, smem, (cudaStream_t)stream>>>(rows, ld_rows, out, ld_out, batch,
                                                                                               fields, dim, ip_first ? 1 : 0);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_pairdots_bwd(const float* rows, int64_t ld_rows, const float* gout, int64_t ld_g, float* grows,
                                  int64_t ld_grows, int64_t batch, int32_t fields, int32_t dim, int32_t ip_first,
                                  rlctr_stream_t stream) {
    if (!rows || !gout || !grows || batch < 0 || fields < 2 || dim <= 0) return RLCTR_EINVAL;
    if (fields > 23) return RLCTR_EUNSUPPORTED;
    const int fd = fields * dim, npair = fields * (fields - 1) / 2;
    if (ld_rows < fd || ld_g < fd + npair || ld_grows < fd) return RLCTR_EINVAL;
    if (batch == 0) return RLCTR_OK;
    const int nw = 4;
    const size_t smem = (size_t)nw * (fd + npair) * sizeof(float) + (size_t)fields * fields;
    if (smem > 48 * 1024) return RLCTR_EUNSUPPORTED;
    pairdots_bwd_kernel<<<grid_for_warps(batch, nw, 16), nw * 32, smem, (cudaStream_t)stream>>>(rows, ld_rows, gout, ld_g, grows,
                                                                                               ld_grows, batch, fields, dim,
                                                                                               ip_first ? 1 : 0);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}
