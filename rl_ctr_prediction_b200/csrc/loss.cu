// loss.cu -- the loss head (sigmoid + BCELoss and their autograd), K6 generate_preds, and the
// REINFORCE head.  All per-sample, HBM-trivial; the point is to replace dozens of tiny ATen
// launches (and the host-synchronising nonzero() calls of generate_preds) by one kernel each,
// with reductions shaped as fixed trees so results are bit-identical from run to run.
#include "common.cuh"

namespace rlctr {

constexpr int RED_MAX_BLOCKS = 1024;
// ws layout: [0] u32 arrival counter | [64 floats in] K arrays of RED_MAX_BLOCKS block partials
struct RedWs {
    unsigned int* counter;
    float* partial;                                   // [K][RED_MAX_BLOCKS]
};
__host__ __device__ inline RedWs red_ws(void* ws) {
    return RedWs{reinterpret_cast<unsigned int*>(ws), reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 256)};
}

// fixed-shape block sum (256 threads): warp shuffles then 8 partials through smem
__device__ __forceinline__ float block_sum_256(float x, float* smem8) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(RLCTR_FULL, x, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem8[threadIdx.x >> 5] = x;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 8) t = smem8[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) t += __shfl_xor_sync(RLCTR_FULL, t, off);
    }
    return t;                                          // valid in thread 0
}

// true in every thread of the LAST block to arrive (its view of all partials is complete)
__device__ __forceinline__ bool last_block_arrives(unsigned int* counter) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    return is_last;
}

__global__ void __launch_bounds__(256)
bce_kernel(const float* __restrict__ logit, const int64_t* __restrict__ yi, const float* __restrict__ yf,
           float* __restrict__ pctr, float* __restrict__ loss, float* __restrict__ dlogit, float* __restrict__ dbias,
           void* ws, int64_t batch) {
    __shared__ float sm[8];
    const float ginv = 1.0f / (float)batch;            // mean-reduction backward: grad / numel
    float acc = 0.f, dsum = 0.f;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += (int64_t)gridDim.x * blockDim.x) {
        const float z = __ldg(logit + b);
        const float y = yi ? (float)__ldg(yi + b) : __ldg(yf + b);
        const float p = sigmoidf_ref(z);
        // aten binary_cross_entropy: (y-1)*max(log1p(-p),-100) - y*max(log(p),-100)
        acc += (y - 1.0f) * fmaxf(log1pf(-p), -100.0f) - y * fmaxf(logf(p), -100.0f);
        // backward: grad*(p-y)/max((1-p)*p, 1e-12), then sigmoid_backward: g*(1-p)*p
        const float dp = ginv * (p - y) / fmaxf((1.0f - p) * p, 1e-12f);
        const float dz = dp * (1.0f - p) * p;
        if (pctr) pctr[b] = p;
        if (dlogit) dlogit[b] = dz;
        dsum += dz;
    }
    RedWs w = red_ws(ws);
    const float bs = block_sum_256(acc, sm), ds = block_sum_256(dsum, sm);
    if (threadIdx.x == 0) { w.partial[blockIdx.x] = bs; w.partial[RED_MAX_BLOCKS + blockIdx.x] = ds; }
    if (last_block_arrives(w.counter)) {
        float t = 0.f, d = 0.f;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) {
            t += __ldcg(w.partial + i);
            d += __ldcg(w.partial + RED_MAX_BLOCKS + i);
        }
        t = block_sum_256(t, sm);
        d = block_sum_256(d, sm);
        if (threadIdx.x == 0) {
            if (loss) loss[0] = t / (float)batch;
            if (dbias) dbias[0] = d;                   // d L / d bias = sum_b dlogit[b]
            *w.counter = 0;                            // leave the workspace re-usable
        }
    }
}

// torch sigmoid_backward after a user-side loss (e.g. nn.BCELoss): dlogit = g * (1 - p) * p, and
// the bias gradient sum_b dlogit[b] as a fixed-shape tree.  grad_p == NULL: dlogit is given, only sum it.
__global__ void __launch_bounds__(256)
sigmoid_bwd_kernel(const float* __restrict__ grad_p, const float* __restrict__ pctr, float* __restrict__ dlogit,
                   float* __restrict__ dbias, void* ws, int64_t batch) {
    __shared__ float sm[8];
    float dsum = 0.f;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += (int64_t)gridDim.x * blockDim.x) {
        float dz;
        if (grad_p) {
            const float p = __ldg(pctr + b);
            dz = __ldg(grad_p + b) * (1.0f - p) * p;
            dlogit[b] = dz;
        } else {
            dz = dlogit[b];
        }
        dsum += dz;
    }
    if (!dbias) return;
    RedWs w = red_ws(ws);
    const float ds = block_sum_256(dsum, sm);
    if (threadIdx.x == 0) w.partial[blockIdx.x] = ds;
    if (last_block_arrives(w.counter)) {
        float d = 0.f;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) d += __ldcg(w.partial + i);
        d = block_sum_256(d, sm);
        if (threadIdx.x == 0) { dbias[0] = d; *w.counter = 0; }
    }
}

// ------------------------------------------------------------------------------------------
// K6 generate_preds: one thread per sample, everything in registers (M <= 8)
// ------------------------------------------------------------------------------------------
constexpr int GP_MAX = 8;

__global__ void __launch_bounds__(256)
generate_preds_kernel(const float* __restrict__ pctr, const float* __restrict__ w, const int64_t* __restrict__ action,
                      const int64_t* __restrict__ label, float* __restrict__ y, float* __restrict__ w_out,
                      float* __restrict__ reward, int64_t batch, int M, int variant) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    // everything is statically indexed after unrolling: the tuples (weight, pctr, model) are
    // moved by the sort instead of being addressed through a permutation (no local memory)
    float p[GP_MAX], wt[GP_MAX], ow[GP_MAX], sw[GP_MAX], sp[GP_MAX];
    int si[GP_MAX];
#pragma unroll
    for (int m = 0; m < GP_MAX; ++m) {
        p[m] = m < M ? __ldg(pctr + b * M + m) : 0.f;
        wt[m] = m < M ? __ldg(w + b * M + m) : -INFINITY;
        ow[m] = 0.f;
        sw[m] = wt[m]; sp[m] = p[m]; si[m] = m;
    }
    // stable insertion sort by descending weight (torch.sort(-w), main.py:188); the -inf padding stays last
#pragma unroll
    for (int i = 1; i < GP_MAX; ++i) {
#pragma unroll
        for (int j = i; j > 0; --j) {
            if (sw[j] > sw[j - 1]) {
                float tw = sw[j]; sw[j] = sw[j - 1]; sw[j - 1] = tw;
                float tp = sp[j]; sp[j] = sp[j - 1]; sp[j - 1] = tp;
                int ti = si[j]; si[j] = si[j - 1]; si[j - 1] = ti;
            }
        }
    }
    const int k = (int)__ldg(action + b);
    const int lab = (int)__ldg(label + b);
    float psum = 0.f;
#pragma unroll
    for (int m = 0; m < GP_MAX; ++m) if (m < M) psum += p[m];
    const float mean_all = psum / (float)M;
    const int k_lo = variant == 0 ? 2 : 1;
    float yv = 1.0f, rv = 1.0f;                        // main.py:185-186 defaults
    if (k >= k_lo && k <= M) {
        float base;
        if (k == M) {                                  // main.py:208-231
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < GP_MAX; ++m) if (m < M) { acc += wt[m] * p[m]; ow[m] = wt[m]; }
            yv = acc;
            base = mean_all;
        } else if (k == 1) {                           // hybrid_td3_main_per.py:86-91
            yv = sp[0];
#pragma unroll
            for (int m = 0; m < GP_MAX; ++m) if (si[0] == m) ow[m] = 1.0f;
            base = mean_all;
        } else {                                       // main.py:232-267
            float e[GP_MAX], esum = 0.f;
#pragma unroll
            for (int j = 0; j < GP_MAX; ++j) { e[j] = j < k ? expf(sw[j] - sw[0]) : 0.f; esum += e[j]; }
            float acc = 0.f, ps = 0.f;
#pragma unroll
            for (int j = 0; j < GP_MAX; ++j) {
                if (j < k) {
                    const float sj = e[j] / esum;
                    acc += sj * sp[j];
                    ps += sp[j];
#pragma unroll
                    for (int m = 0; m < GP_MAX; ++m) if (si[j] == m) ow[m] = sj;
                }
            }
            yv = acc;
            base = variant == 0 ? ps / (float)k : mean_all;
        }
        const bool good = lab == 1 ? (yv >= base) : (yv <= base);
        rv = good ? 1.0f : (variant == 0 ? -1.0f : 0.0f);
    }
    y[b] = yv;
    reward[b] = rv;
    if (w_out) {
#pragma unroll
        for (int m = 0; m < GP_MAX; ++m) if (m < M) w_out[b * M + m] = ow[m];
    }
}

// ------------------------------------------------------------------------------------------
// K6, v10 form (src/all_main/hybrid_td3_main_per_v10.py:54-164): models chosen by descending prob_weights, softmax over the k
// largest c_actions in their own order, rewards 1 / 0 on strict comparisons, and return_c_actions -- whose partial-ensemble rows
// read sort_c_actions at the sample's RANK within its action group (:117 indexes the whole-batch tensor with subset-relative
// positions).  The rank is an exclusive count of earlier samples with the same action: per-block counts, a scan over the blocks,
// then match_any inside the block.
// ------------------------------------------------------------------------------------------
constexpr int GP10_THREADS = 256;

__global__ void __launch_bounds__(GP10_THREADS)
gp10_count_kernel(const int64_t* __restrict__ action, int64_t batch, int M, int* __restrict__ counts) {
    __shared__ int s[GP_MAX];
    if (threadIdx.x < GP_MAX) s[threadIdx.x] = 0;
    __syncthreads();
    const int64_t b = (int64_t)blockIdx.x * GP10_THREADS + threadIdx.x;
    const int64_t k = b < batch ? __ldg(action + b) : 0;
    if (k >= 1 && k <= M) atomicAdd(&s[k - 1], 1);                   // integer counts: order-free
    __syncthreads();
    if (threadIdx.x < GP_MAX) counts[(int64_t)blockIdx.x * GP_MAX + threadIdx.x] = s[threadIdx.x];
}
// counts[block][a] -> number of samples with action a + 1 in EARLIER blocks.  One warp per action; a lane owns a contiguous run of
// blocks (sum, warp scan of the run totals, second pass writes the exclusive prefixes).
__global__ void __launch_bounds__(GP_MAX * 32)
gp10_scan_kernel(int* __restrict__ counts, int nblocks) {
    const int a = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (nblocks + 31) / 32;
    const int i0 = lane * per, i1 = min(i0 + per, nblocks);
    int total = 0;
    for (int i = i0; i < i1; ++i) total += counts[(int64_t)i * GP_MAX + a];
    int incl = total;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(RLCTR_FULL, incl, off);
        if (lane >= off) incl += t;
    }
    int run = incl - total;
    for (int i = i0; i < i1; ++i) {
        const int v = counts[(int64_t)i * GP_MAX + a];
        counts[(int64_t)i * GP_MAX + a] = run;
        run += v;
    }
}
// descending stable insertion sort of GP_MAX values (the -inf padding stays last)
__device__ __forceinline__ void gp_sort_desc(float (&v)[GP_MAX]) {
#pragma unroll
    for (int i = 1; i < GP_MAX; ++i) {
#pragma unroll
        for (int j = i; j > 0; --j) {
            if (v[j] > v[j - 1]) { const float t = v[j]; v[j] = v[j - 1]; v[j - 1] = t; }
        }
    }
}
__global__ void __launch_bounds__(GP10_THREADS)
generate_preds_v10_kernel(const float* __restrict__ pctr, const float* __restrict__ w, const float* __restrict__ c_actions,
                          const int64_t* __restrict__ action, const int64_t* __restrict__ label, float* __restrict__ y,
                          float* __restrict__ c_out, float* __restrict__ reward, int64_t batch, int M,
                          const int* __restrict__ prefix) {
    __shared__ int s_cnt[GP10_THREADS / 32][GP_MAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * GP10_THREADS + threadIdx.x;
    const bool in = b < batch;
    const int64_t k64 = in ? __ldg(action + b) : 0;
    const int k = (k64 >= 1 && k64 <= M) ? (int)k64 : 0;              // 0: outside 1..M (defaults)
    // rank of this sample among the samples of the batch with the same action
    if (lane < GP_MAX) s_cnt[warp][lane] = 0;
    __syncwarp();
    const unsigned same = __match_any_sync(RLCTR_FULL, k);
    if (k > 0 && lane == __ffs(same) - 1) s_cnt[warp][k - 1] = __popc(same);
    __syncthreads();
    int64_t rank = 0;
    if (k > 0) {
        rank = prefix[(int64_t)blockIdx.x * GP_MAX + k - 1] + __popc(same & ((1u << lane) - 1u));
        for (int q = 0; q < warp; ++q) rank += s_cnt[q][k - 1];
    }
    if (!in) return;
    float p[GP_MAX], wt[GP_MAX], sw[GP_MAX], sp[GP_MAX], cv[GP_MAX], cs[GP_MAX], oc[GP_MAX];
    int si[GP_MAX];
#pragma unroll
    for (int m = 0; m < GP_MAX; ++m) {
        p[m] = m < M ? __ldg(pctr + b * M + m) : 0.f;
        wt[m] = m < M ? __ldg(w + b * M + m) : -INFINITY;
        cv[m] = m < M ? __ldg(c_actions + b * M + m) : -INFINITY;
        oc[m] = 0.f;
        sw[m] = wt[m]; sp[m] = p[m]; si[m] = m; cs[m] = cv[m];
    }
    // models by descending prob_weights (:62), tuples moved by the sort
#pragma unroll
    for (int i = 1; i < GP_MAX; ++i) {
#pragma unroll
        for (int j = i; j > 0; --j) {
            if (sw[j] > sw[j - 1]) {
                float tw = sw[j]; sw[j] = sw[j - 1]; sw[j - 1] = tw;
                float tp = sp[j]; sp[j] = sp[j - 1]; sp[j - 1] = tp;
                int ti = si[j]; si[j] = si[j - 1]; si[j - 1] = ti;
            }
        }
    }
    gp_sort_desc(cs);                                                 // this row's c_actions, descending (:63)
    const int lab = (int)__ldg(label + b);
    float psum = 0.f;
#pragma unroll
    for (int m = 0; m < GP_MAX; ++m) if (m < M) psum += p[m];
    const float mean_all = psum / (float)M;
    float yv = 1.0f, rv = 1.0f;
    if (k > 0) {
        if (k == M) {                                                 // :88-96
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < GP_MAX; ++m) if (m < M) { acc += wt[m] * p[m]; oc[m] = cv[m]; }
            yv = acc;
        } else {                                                      // :97-123
            float cr[GP_MAX];                                         // sort_c_actions of row `rank` of the whole batch (:117)
#pragma unroll
            for (int m = 0; m < GP_MAX; ++m) cr[m] = m < M ? __ldg(c_actions + rank * M + m) : -INFINITY;
            gp_sort_desc(cr);
            float e[GP_MAX], esum = 0.f;
#pragma unroll
            for (int j = 0; j < GP_MAX; ++j) { e[j] = j < k ? expf(cs[j] - cs[0]) : 0.f; esum += e[j]; }
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < GP_MAX; ++j) {
                if (j < k) {
                    acc += (e[j] / esum) * sp[j];
#pragma unroll
                    for (int m = 0; m < GP_MAX; ++m) if (si[j] == m) oc[m] = cr[j];
                }
            }
            yv = acc;
        }
        const bool good = lab == 1 ? (yv > mean_all) : (yv < mean_all);     // strict (:127-146)
        rv = good ? 1.0f : 0.0f;
    }
    y[b] = yv;
    reward[b] = rv;
#pragma unroll
    for (int m = 0; m < GP_MAX; ++m) if (m < M) c_out[b * M + m] = oc[m];
}

// ------------------------------------------------------------------------------------------
// REINFORCE head
// ------------------------------------------------------------------------------------------
constexpr int RF_MAX = 32;

// pass 1: logp, and the three global sums the loss / its gradient need
__global__ void __launch_bounds__(256)
reinforce_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ act, const float* __restrict__ vt,
                     float* __restrict__ logp, float* __restrict__ loss, void* ws, int64_t batch, int A, int variant) {
    __shared__ float sm[8];
    float s_nl = 0.f, s_vt = 0.f, s_nlvt = 0.f;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += (int64_t)gridDim.x * blockDim.x) {
        const float* z = logits + b * A;
        float mx = -INFINITY;
        for (int a = 0; a < A; ++a) mx = fmaxf(mx, __ldg(z + a));
        float sum = 0.f;
        for (int a = 0; a < A; ++a) sum += expf(__ldg(z + a) - mx);
        const int ai = (int)__ldg(act + b) - 1;
        const float pa = expf(__ldg(z + ai) - mx) / sum;        // softmax then gather (PG_model.py:56,105)
        const float lp = logf(pa);
        const float v = __ldg(vt + b);
        if (logp) logp[b] = lp;
        s_nl += -lp; s_vt += v; s_nlvt += -lp * v;
    }
    RedWs w = red_ws(ws);
    float r0 = block_sum_256(s_nl, sm), r1 = block_sum_256(s_vt, sm), r2 = block_sum_256(s_nlvt, sm);
    if (threadIdx.x == 0) {
        w.partial[blockIdx.x] = r0;
        w.partial[RED_MAX_BLOCKS + blockIdx.x] = r1;
        w.partial[2 * RED_MAX_BLOCKS + blockIdx.x] = r2;
    }
    if (last_block_arrives(w.counter)) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) {
            t0 += __ldcg(w.partial + i);
            t1 += __ldcg(w.partial + RED_MAX_BLOCKS + i);
            t2 += __ldcg(w.partial + 2 * RED_MAX_BLOCKS + i);
        }
        t0 = block_sum_256(t0, sm); t1 = block_sum_256(t1, sm); t2 = block_sum_256(t2, sm);
        if (threadIdx.x == 0) {
            const float mean_vt = t1 / (float)batch;
            if (loss) loss[0] = variant == 0 ? t0 * mean_vt : t2 / (float)batch;
            w.partial[3 * RED_MAX_BLOCKS] = mean_vt;              // read by pass 2
            *w.counter = 0;
        }
    }
}
// pass 2: dlogits = coef_b * (pi - onehot);  coef = mean(vt) (literal) or vt_b / B (per-sample)
__global__ void __launch_bounds__(256)
reinforce_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ act, const float* __restrict__ vt,
                     float* __restrict__ dlogits, const void* ws, int64_t batch, int A, int variant) {
    const float mean_vt = __ldcg(reinterpret_cast<const float*>(reinterpret_cast<const char*>(ws) + 256) + 3 * RED_MAX_BLOCKS);
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += (int64_t)gridDim.x * blockDim.x) {
        const float* z = logits + b * A;
        float mx = -INFINITY;
        for (int a = 0; a < A; ++a) mx = fmaxf(mx, __ldg(z + a));
        float sum = 0.f;
        for (int a = 0; a < A; ++a) sum += expf(__ldg(z + a) - mx);
        const int ai = (int)__ldg(act + b) - 1;
        const float coef = variant == 0 ? mean_vt : __ldg(vt + b) / (float)batch;
        for (int a = 0; a < A; ++a) {
            const float pi = expf(__ldg(z + a) - mx) / sum;
            dlogits[b * A + a] = coef * (pi - (a == ai ? 1.0f : 0.0f));
        }
    }
}

static inline int red_grid(int64_t batch) {
    int64_t blocks = (batch + 255) / 256;
    if (blocks < 1) blocks = 1;
    return (int)(blocks < RED_MAX_BLOCKS ? blocks : RED_MAX_BLOCKS);
}

}  // namespace rlctr

using namespace rlctr;

extern "C" int rlctr_bce_fwd_bwd(const float* logit, const int64_t* labels_i64, const float* labels_f32, float* pctr,
                                 float* loss, float* dlogit, float* dbias, void* ws, int64_t batch,
                                 rlctr_stream_t stream) {
    if (!logit || !ws || batch <= 0) return RLCTR_EINVAL;
    if ((labels_i64 == nullptr) == (labels_f32 == nullptr)) return RLCTR_EINVAL;
    bce_kernel<<<red_grid(batch), 256, 0, (cudaStream_t)stream>>>(logit, labels_i64, labels_f32, pctr, loss, dlogit,
                                                                  dbias, ws, batch);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_sigmoid_bwd(const float* grad_p, const float* pctr, float* dlogit, float* dbias, void* ws,
                                 int64_t batch, rlctr_stream_t stream) {
    if (!dlogit || batch <= 0 || (grad_p && !pctr) || (dbias && !ws)) return RLCTR_EINVAL;
    sigmoid_bwd_kernel<<<red_grid(batch), 256, 0, (cudaStream_t)stream>>>(grad_p, pctr, dlogit, dbias, ws, batch);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_generate_preds(const float* pctr, const float* w, const int64_t* action, const int64_t* label,
                                    float* y, float* w_out, float* reward, int64_t batch, int32_t models,
                                    int32_t variant, rlctr_stream_t stream) {
    if (!pctr || !w || !action || !label || !y || !reward || batch < 0) return RLCTR_EINVAL;
    if (models < 1 || models > GP_MAX || (variant != 0 && variant != 1)) return RLCTR_EUNSUPPORTED;
    if (batch == 0) return RLCTR_OK;
    generate_preds_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pctr, w, action, label, y, w_out, reward, batch, models, variant);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" size_t rlctr_generate_preds_v10_ws_bytes(int64_t batch) {
    if (batch < 0) return 0;
    return (size_t)((batch + GP10_THREADS - 1) / GP10_THREADS + 1) * GP_MAX * sizeof(int);
}
extern "C" int rlctr_generate_preds_v10(const float* pctr, const float* w, const float* c_actions, const int64_t* action,
                                        const int64_t* label, float* y, float* c_out, float* reward, int64_t batch,
                                        int32_t models, void* ws, size_t ws_bytes, rlctr_stream_t stream) {
    if (!pctr || !w || !c_actions || !action || !label || !y || !c_out || !reward || batch < 0) return RLCTR_EINVAL;
    if (models < 1 || models > GP_MAX || batch > ((int64_t)1 << 31)) return RLCTR_EUNSUPPORTED;
    if (batch == 0) return RLCTR_OK;
    if (!ws || ws_bytes < rlctr_generate_preds_v10_ws_bytes(batch)) return RLCTR_EWORKSPACE;
    const int nblocks = (int)((batch + GP10_THREADS - 1) / GP10_THREADS);
    int* counts = reinterpret_cast<int*>(ws);
    cudaStream_t st = (cudaStream_t)stream;
    gp10_count_kernel<<<nblocks, GP10_THREADS, 0, st>>>(action, batch, models, counts);
    gp10_scan_kernel<<<1, GP_MAX * 32, 0, st>>>(counts, nblocks);
    generate_preds_v10_kernel<<<nblocks, GP10_THREADS, 0, st>>>(pctr, w, c_actions, action, label, y, c_out, reward, batch,
                                                                 models, counts);
    RLCTR_LAUNCH_CHECK();
    return RLCTR_OK;
}

extern "C" int rlctr_reinforce_loss_bwd(const float* logits, const int64_t* act, const float* vt, float* logp,
                                        float* loss, float* dlogits, void* ws, int64_t batch, int32_t actions,
                                        int32_t variant, rlctr_stream_t stream) {
    if (!logits || !act || !vt || !ws || batch <= 0) return RLCTR_EINVAL;
    if (actions < 1 || actions > RF_MAX || (variant != 0 && variant != 1)) return RLCTR_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    reinforce_fwd_kernel<<<red_grid(batch), 256, 0, st>>>(logits, act, vt, logp, loss, ws, batch, actions, variant);
    RLCTR_LAUNCH_CHECK();
    if (dlogits) {
        reinforce_bwd_kernel<<<red_grid(batch), 256, 0, st>>>(logits, act, vt, dlogits, ws, batch, actions, variant);
        RLCTR_LAUNCH_CHECK();
    }
    return RLCTR_OK;
}

unsigned long long g_rlctr_launches = 0;
extern "C" unsigned long long rlctr_launch_count(void) { return __atomic_load_n(&g_rlctr_launches, __ATOMIC_RELAXED); }

extern "C" int rlctr_version(void) { return RLCTR_VERSION; }

extern "C" const char* rlctr_strerror(int code) {
    switch (code) {
        case RLCTR_OK: return "ok";
        case RLCTR_EINVAL: return "invalid argument";
        case RLCTR_EUNSUPPORTED: return "unsupported shape";
        case RLCTR_EWORKSPACE: return "workspace too small";
        case RLCTR_EALIGN: return "pointer not 16-byte aligned";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}
