// Fused PPO loss + gradient over the five action heads and the value head.
//
// Replaces, for one stacked batch of N tokens (optimizer.py line numbers, TimZaman/dotaclient):
//   :587-589  advantage normalisation (mean, unbiased std, +eps) over ALL tokens incl. padding
//   :621-646  per head: actions_step, masked log-softmax (policy.py:169-178: no max-subtraction,
//             normalised over the mask), selected log-prob, ratio, clipped surrogate, entropy
//   :649-665  policy loss = mean over the 5 heads (skipped heads count as 0), entropy loss,
//             value loss over ALL tokens, total
//   :672      loss.backward() down to d loss / d logits and d loss / d value
//
// Two launches.  (1) ppo_stats_kernel: per-head action-row counts and sum / sum-of-squares of
// the raw advantage (float64) -- the counts are needed as divisors by every gradient.
// (2) ppo_loss_kernel: one thread per token; a 128-token tile of every head's logits / masks /
// actions is staged through shared memory with coalesced 16-byte loads (rows are 12..160 bytes,
// so per-thread row reads from global would waste most of every sector), the tile is overwritten
// in place with d loss / d logits and written back with coalesced 16-byte stores.
//
// Algorithmic HBM bytes per token: logits 260 + masks 65 + actions 65 + old 20 + adv/ret/value 12
// read, dlogits 260 + dvalue 4 written = 686 (+ 69 for the statistics pass).
#include "dc_common.cuh"

namespace {

constexpr int kHeads = DC_NUM_HEADS;
constexpr int kTile = 128;  // tokens per CTA == threads per CTA
__host__ __device__ constexpr int head_n(int h) { return h == 0 ? 4 : h == 1 ? 9 : h == 2 ? 9 : h == 3 ? 40 : 3; }
// odd row pitch in shared memory -> conflict-free per-thread row access
__host__ __device__ constexpr int head_pitch(int h) { return h == 0 ? 5 : h == 1 ? 9 : h == 2 ? 9 : h == 3 ? 41 : 3; }
__host__ __device__ constexpr int logit_off(int h) {  // float offset of head h's tile in smem
    int o = 0;
    for (int i = 0; i < h; ++i) o += head_pitch(i) * kTile;
    return o;
}
__host__ __device__ constexpr int byte_off(int h) {
    int o = 0;
    for (int i = 0; i < h; ++i) o += head_n(i) * kTile;
    return o;
}
constexpr int kLogitFloats = logit_off(kHeads);      // 67 * 128
constexpr int kOldFloats = 5 * kTile;
constexpr int kByteTile = byte_off(kHeads);          // 65 * 128
constexpr size_t kSmemBytes = (size_t)(kLogitFloats + kOldFloats) * 4 + 2 * (size_t)kByteTile;

struct HeadPtrs {
    const float *logits[kHeads];
    const uint8_t *masks[kHeads];
    const uint8_t *actions[kHeads];
    float *dlogits[kHeads];
    long long ld_l[kHeads];     // row pitch (floats) of logits[h]; == n_h when contiguous.  A pitch of 128 lets the four
    long long ld_d[kHeads];     // small heads + value live as column ranges of ONE packed [N,128] GEMM output / gradient
    long long ld_v, ld_dv;      // pitches of value / dvalue
};

// Workspace layout (DC_PPO_WORKSPACE_BYTES, zeroed per call)
struct Workspace {
    double pol[kHeads];   // sum over action rows of min(surr1, surr2)
    double ent[kHeads];   // sum over masked entries of -p*logp
    double vl;            // sum (ret - v)^2
    double adv_sum, adv_sq;
    int cnt[kHeads];
    unsigned ticket_stats, ticket_loss;
    float adv_mean, adv_std;
};
static_assert(sizeof(Workspace) <= DC_PPO_WORKSPACE_BYTES, "workspace too small");

// Cooperative copy of `count` rows of N floats (contiguous in global) into smem rows of pitch P.
template <int N, int P>
__device__ __forceinline__ void stage_rows_f32(float *dst, const float *__restrict__ base, long long ld, int64_t t0, int count) {
    const int total = count * N;
    if (ld != N) {                                   // strided rows (column range of a wider matrix)
        for (int idx = threadIdx.x; idx < total; idx += kTile) dst[(idx / N) * P + (idx % N)] = base[(t0 + idx / N) * ld + (idx % N)];
        return;
    }
    const float *src = base + t0 * N;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int nvec = total >> 2;
        for (int i = threadIdx.x; i < nvec; i += kTile) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + i);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = i * 4 + q;
                dst[(idx / N) * P + (idx % N)] = e[q];
            }
        }
        for (int idx = (nvec << 2) + threadIdx.x; idx < total; idx += kTile) dst[(idx / N) * P + (idx % N)] = src[idx];
    } else {
        for (int idx = threadIdx.x; idx < total; idx += kTile) dst[(idx / N) * P + (idx % N)] = src[idx];
    }
}

template <int N, int P>
__device__ __forceinline__ void unstage_rows_f32(float *__restrict__ base, long long ld, int64_t t0, const float *src, int count) {
    const int total = count * N;
    if (ld != N) {
        for (int idx = threadIdx.x; idx < total; idx += kTile) base[(t0 + idx / N) * ld + (idx % N)] = src[(idx / N) * P + (idx % N)];
        return;
    }
    float *dst = base + t0 * N;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int nvec = total >> 2;
        for (int i = threadIdx.x; i < nvec; i += kTile) {
            float e[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = i * 4 + q;
                e[q] = src[(idx / N) * P + (idx % N)];
            }
            reinterpret_cast<float4 *>(dst)[i] = make_float4(e[0], e[1], e[2], e[3]);
        }
        for (int idx = (nvec << 2) + threadIdx.x; idx < total; idx += kTile) dst[idx] = src[(idx / N) * P + (idx % N)];
    } else {
        for (int idx = threadIdx.x; idx < total; idx += kTile) dst[idx] = src[(idx / N) * P + (idx % N)];
    }
}

__device__ __forceinline__ void stage_bytes(uint8_t *dst, const uint8_t *__restrict__ src, int total) {
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int nvec = total >> 4;
        for (int i = threadIdx.x; i < nvec; i += kTile)
            reinterpret_cast<uint4 *>(dst)[i] = __ldg(reinterpret_cast<const uint4 *>(src) + i);
        for (int idx = (nvec << 4) + threadIdx.x; idx < total; idx += kTile) dst[idx] = src[idx];
    } else {
        for (int idx = threadIdx.x; idx < total; idx += kTile) dst[idx] = src[idx];
    }
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T *scratch) {
    v = dc_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = 0;
    for (int w = 0; w < kTile / 32; ++w) r += scratch[w];
    return r;
}

// ---- pass 1: counts + advantage statistics ------------------------------------------------
__global__ void __launch_bounds__(kTile) ppo_stats_kernel(HeadPtrs hp, const float *__restrict__ adv, int64_t N,
                                                           Workspace *ws, int32_t *n_actions_out) {
    __shared__ __align__(16) uint8_t s_act[kByteTile];
    __shared__ double s_red[kTile / 32];
    __shared__ bool s_last;
    const int64_t t0 = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, N - t0);
#pragma unroll
    for (int h = 0; h < kHeads; ++h) stage_bytes(s_act + byte_off(h), hp.actions[h] + t0 * head_n(h), count * head_n(h));
    __syncthreads();
    const bool live = threadIdx.x < count;
    double a = 0.0;
    if (live) a = (double)adv[t0 + threadIdx.x];
    int has[kHeads];
#pragma unroll
    for (int h = 0; h < kHeads; ++h) {
        int any = 0;
        if (live) {
            const uint8_t *row = s_act + byte_off(h) + threadIdx.x * head_n(h);
            for (int j = 0; j < head_n(h); ++j) any |= row[j];
        }
        has[h] = any != 0;
    }
    const double sa = block_sum(a, s_red);
    const double sq = block_sum(a * a, s_red);
    int tot[kHeads];
#pragma unroll
    for (int h = 0; h < kHeads; ++h) tot[h] = __syncthreads_count(has[h]);
    if (threadIdx.x == 0) {
        atomicAdd(&ws->adv_sum, sa);
        atomicAdd(&ws->adv_sq, sq);
#pragma unroll
        for (int h = 0; h < kHeads; ++h)
            if (tot[h]) atomicAdd(&ws->cnt[h], tot[h]);
        __threadfence();
        s_last = atomicAdd(&ws->ticket_stats, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const double n = (double)N;
        const double sum = *((volatile double *)&ws->adv_sum), sq2 = *((volatile double *)&ws->adv_sq);
        const double mean = sum / n;
        // torch.std: unbiased (N-1); NaN for N == 1 like torch.
        const double var = (sq2 - n * mean * mean) / (n - 1.0);
        ws->adv_mean = (float)mean;
        ws->adv_std = (float)sqrt(var > 0.0 ? var : (var == var ? 0.0 : var));
        for (int h = 0; h < kHeads; ++h) n_actions_out[h] = *((volatile int *)&ws->cnt[h]);
    }
}

// ---- pass 2: loss + gradient --------------------------------------------------------------
template <int H, bool kGrad>
__device__ __forceinline__ void head_token(float *lrow, const uint8_t *mrow, const uint8_t *arow, float old_lp,
                                           float adv_n, int n_h, float e_clip, float entropy_coef, float &pol_acc,
                                           float &ent_acc, float *logp_out) {
    constexpr int N = head_n(H);
    float l[N], e[N];
    int mask_any = 0, a_idx = -1;
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        l[j] = lrow[j];
        const int m = mrow[j];
        mask_any |= m;
        e[j] = m ? __expf(l[j]) : 0.f;   // masked_exp[~mask] = 0 (policy.py:172-174)
        se += e[j];
        if (arow[j]) a_idx = j;
    }
    if (logp_out) {  // no-grad selected log-prob (optimizer.py:387-390)
        float lp = 0.f;
        if (a_idx >= 0) {
            float la = l[0];
#pragma unroll
            for (int j = 1; j < N; ++j) la = (j == a_idx) ? l[j] : la;
            lp = la - logf(se);
        }
        *logp_out = lp;
        return;
    }
    // A head nobody used this batch is skipped entirely (optimizer.py:627-630); a row with an
    // empty mask and no action contributes exactly zero (its NaN is erased by the index_put backward).
    if (n_h == 0 || (!mask_any && a_idx < 0)) {
        if (kGrad) {
#pragma unroll
            for (int j = 0; j < N; ++j) lrow[j] = 0.f;
        }
        return;
    }
    const float lse = logf(se);          // policy.py:175-177
    const float inv_se = 1.0f / se;
    float ent_row = 0.f;
    float p[N], lp[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        lp[j] = l[j] - lse;
        p[j] = e[j] * inv_se;            // exp(log_prob) over the mask, 0 outside
        if (mrow[j]) ent_row -= p[j] * lp[j];
    }
    ent_acc += ent_row;                   // optimizer.py:644-646 (divided by n_actions at the end)
    float g_lp = 0.f;                     // d loss / d logp[a]
    if (a_idx >= 0) {
        float lpa = lp[0];
#pragma unroll
        for (int j = 1; j < N; ++j) lpa = (j == a_idx) ? lp[j] : lpa;
        const float ratio = __expf(lpa - old_lp);                     // optimizer.py:638
        const float lo = 1.0f - e_clip, hi = 1.0f + e_clip;
        const float s1 = ratio * adv_n;                                // :639
        const float s2 = fminf(fmaxf(ratio, lo), hi) * adv_n;          // :640
        pol_acc += fminf(s1, s2);                                      // :641 (negated, averaged at the end)
        // autograd of torch.min(a, b): ties split the gradient in half; clamp passes it inside [lo, hi].
        const float g1 = s1 < s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f);
        const float g2 = s2 < s1 ? 1.f : (s1 == s2 ? 0.5f : 0.f);
        const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
        g_lp = -(1.0f / kHeads) / (float)n_h * adv_n * (g1 + g2 * in_range) * ratio;
    }
    if (kGrad) {
        const float ce = entropy_coef > 0.f ? entropy_coef / (float)n_h : 0.f;   // optimizer.py:652-656
#pragma unroll
        for (int j = 0; j < N; ++j) {
            float g = -g_lp * p[j];                       // through logsumexp (masked entries only: p = 0 outside)
            if (j == a_idx) g += g_lp;
            if (mrow[j]) g += ce * p[j] * (lp[j] + ent_row);   // d(-coef * entropy)/d logit
            lrow[j] = g;
        }
    }
}

template <bool kSelectOnly>
__global__ void __launch_bounds__(kTile) ppo_loss_kernel(HeadPtrs hp, const float *__restrict__ old_logp,
                                                          const float *__restrict__ adv_raw,
                                                          const float *__restrict__ ret,
                                                          const float *__restrict__ value, int64_t N, float e_clip,
                                                          float entropy_coef, float vf_coef,
                                                          float *__restrict__ dvalue, float *__restrict__ out,
                                                          Workspace *ws, float *__restrict__ logp_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_logits = reinterpret_cast<float *>(smem_raw);
    float *s_old = s_logits + kLogitFloats;
    uint8_t *s_mask = reinterpret_cast<uint8_t *>(s_old + kOldFloats);
    uint8_t *s_act = s_mask + kByteTile;
    __shared__ float s_red[kTile / 32];
    __shared__ bool s_last;

    const int64_t t0 = (int64_t)blockIdx.x * kTile;
    const int count = (int)min((int64_t)kTile, N - t0);
    stage_rows_f32<4, 5>(s_logits + logit_off(0), hp.logits[0], hp.ld_l[0], t0, count);
    stage_rows_f32<9, 9>(s_logits + logit_off(1), hp.logits[1], hp.ld_l[1], t0, count);
    stage_rows_f32<9, 9>(s_logits + logit_off(2), hp.logits[2], hp.ld_l[2], t0, count);
    stage_rows_f32<40, 41>(s_logits + logit_off(3), hp.logits[3], hp.ld_l[3], t0, count);
    stage_rows_f32<3, 3>(s_logits + logit_off(4), hp.logits[4], hp.ld_l[4], t0, count);
    if (!kSelectOnly) stage_rows_f32<5, 5>(s_old, old_logp, 5, t0, count);
#pragma unroll
    for (int h = 0; h < kHeads; ++h) {
        stage_bytes(s_mask + byte_off(h), hp.masks[h] + t0 * head_n(h), count * head_n(h));
        stage_bytes(s_act + byte_off(h), hp.actions[h] + t0 * head_n(h), count * head_n(h));
    }
    __syncthreads();

    const int t = threadIdx.x;
    const bool live = t < count;
    float pol[kHeads] = {0, 0, 0, 0, 0}, ent[kHeads] = {0, 0, 0, 0, 0};
    float vl = 0.f;
    int cnt[kHeads] = {0, 0, 0, 0, 0};
    float adv_n = 0.f;
    if (!kSelectOnly) {
#pragma unroll
        for (int h = 0; h < kHeads; ++h) cnt[h] = ws->cnt[h];
        if (live) {
            // (advantage - mean) / (std + eps), fp32 like optimizer.py:588
            adv_n = __fdiv_rn(__fsub_rn(adv_raw[t0 + t], ws->adv_mean), __fadd_rn(ws->adv_std, 1.1920928955078125e-07f));
        }
    }
    if (live) {
        float lp_sel[kHeads];
#define DC_HEAD(H)                                                                                              \
        head_token<H, true>(s_logits + logit_off(H) + t * head_pitch(H), s_mask + byte_off(H) + t * head_n(H),  \
                            s_act + byte_off(H) + t * head_n(H), kSelectOnly ? 0.f : s_old[t * 5 + H], adv_n,   \
                            cnt[H], e_clip, entropy_coef, pol[H], ent[H], kSelectOnly ? &lp_sel[H] : nullptr);
        DC_HEAD(0) DC_HEAD(1) DC_HEAD(2) DC_HEAD(3) DC_HEAD(4)
#undef DC_HEAD
        if (kSelectOnly) {
#pragma unroll
            for (int h = 0; h < kHeads; ++h) logp_out[(t0 + t) * 5 + h] = lp_sel[h];
        } else {
            const float v = value[(t0 + t) * hp.ld_v], r = ret[t0 + t];
            const float d = r - v;
            vl = d * d;                                                     // optimizer.py:660
            dvalue[(t0 + t) * hp.ld_dv] = vf_coef > 0.f ? vf_coef * (v - r) / (float)N : 0.f;
        }
    }
    if (kSelectOnly) return;
    __syncthreads();
    unstage_rows_f32<4, 5>(hp.dlogits[0], hp.ld_d[0], t0, s_logits + logit_off(0), count);
    unstage_rows_f32<9, 9>(hp.dlogits[1], hp.ld_d[1], t0, s_logits + logit_off(1), count);
    unstage_rows_f32<9, 9>(hp.dlogits[2], hp.ld_d[2], t0, s_logits + logit_off(2), count);
    unstage_rows_f32<40, 41>(hp.dlogits[3], hp.ld_d[3], t0, s_logits + logit_off(3), count);
    unstage_rows_f32<3, 3>(hp.dlogits[4], hp.ld_d[4], t0, s_logits + logit_off(4), count);

    float sums[2 * kHeads + 1];
#pragma unroll
    for (int h = 0; h < kHeads; ++h) {
        sums[h] = block_sum(pol[h], s_red);
        sums[kHeads + h] = block_sum(ent[h], s_red);
    }
    sums[2 * kHeads] = block_sum(vl, s_red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int h = 0; h < kHeads; ++h) {
            if (sums[h] != 0.f) atomicAdd(&ws->pol[h], (double)sums[h]);
            if (sums[kHeads + h] != 0.f) atomicAdd(&ws->ent[h], (double)sums[kHeads + h]);
        }
        atomicAdd(&ws->vl, (double)sums[2 * kHeads]);
        __threadfence();
        s_last = atomicAdd(&ws->ticket_loss, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        volatile Workspace *w = ws;
        float policy = 0.f, entropy = 0.f;
        for (int h = 0; h < kHeads; ++h) {
            const int n = w->cnt[h];
            const float pl = n ? (float)(-w->pol[h] / (double)n) : 0.f;     // optimizer.py:641 / :628
            const float en = n ? (float)(w->ent[h] / (double)n) : 0.f;      // optimizer.py:646 / :629
            out[9 + h] = pl;
            out[4 + h] = en;
            policy += pl;
            entropy += en;
        }
        policy /= (float)kHeads;                                            // optimizer.py:650
        const float e_loss = entropy_coef > 0.f ? -entropy_coef * entropy : 0.f;
        const float v_loss = vf_coef > 0.f ? vf_coef * (0.5f * (float)(w->vl / (double)N)) : 0.f;
        out[0] = policy + e_loss + v_loss;                                  // optimizer.py:665
        out[1] = policy;
        out[2] = e_loss;
        out[3] = v_loss;
        out[14] = w->adv_mean;
        out[15] = w->adv_std;
    }
}

int check_heads(const float *const logits[], const uint8_t *const masks[], const uint8_t *const actions[]) {
    for (int h = 0; h < kHeads; ++h)
        if (!logits[h] || !masks[h] || !actions[h]) return 0;
    return 1;
}

}  // namespace

extern "C" int dc_ppo_loss_fwd_bwd_strided(const float *const logits[DC_NUM_HEADS], const int64_t ld_logits[DC_NUM_HEADS],
                                           const uint8_t *const masks[DC_NUM_HEADS],
                                           const uint8_t *const actions[DC_NUM_HEADS], const float *old_logp,
                                           const float *adv_raw, const float *ret, const float *value, int64_t ld_value,
                                           int64_t N, float e_clip, float entropy_coef, float vf_coef,
                                           float *const dlogits[DC_NUM_HEADS], const int64_t ld_dlogits[DC_NUM_HEADS],
                                           float *dvalue, int64_t ld_dvalue, float *out, int32_t *n_actions, void *workspace,
                                           dc_stream_t stream) {
    DC_REQUIRE(N > 0, DC_EINVAL, "dc_ppo_loss_fwd_bwd: N=%lld", (long long)N);
    DC_REQUIRE(check_heads(logits, masks, actions) && old_logp && adv_raw && ret && value && dvalue && out &&
                   n_actions && workspace && ld_logits && ld_dlogits, DC_EINVAL, "dc_ppo_loss_fwd_bwd: null pointer");
    HeadPtrs hp;
    for (int h = 0; h < kHeads; ++h) {
        DC_REQUIRE(dlogits[h], DC_EINVAL, "dc_ppo_loss_fwd_bwd: null dlogits[%d]", h);
        DC_REQUIRE(ld_logits[h] >= head_n(h) && ld_dlogits[h] >= head_n(h), DC_EINVAL, "dc_ppo_loss_fwd_bwd: row pitch of head %d", h);
        hp.logits[h] = logits[h]; hp.masks[h] = masks[h]; hp.actions[h] = actions[h]; hp.dlogits[h] = dlogits[h];
        hp.ld_l[h] = ld_logits[h]; hp.ld_d[h] = ld_dlogits[h];
    }
    DC_REQUIRE(ld_value >= 1 && ld_dvalue >= 1, DC_EINVAL, "dc_ppo_loss_fwd_bwd: value pitch");
    hp.ld_v = ld_value; hp.ld_dv = ld_dvalue;
    cudaStream_t st = dc_cu_stream(stream);
    Workspace *ws = reinterpret_cast<Workspace *>(workspace);
    DC_CUDA(cudaMemsetAsync(ws, 0, sizeof(Workspace), st));
    const unsigned blocks = (unsigned)((N + kTile - 1) / kTile);
    ppo_stats_kernel<<<blocks, kTile, 0, st>>>(hp, adv_raw, N, ws, n_actions);
    DC_LAUNCH_OK();
    // per-device attribute: set on every call (a process-wide "done" flag breaks the second GPU of a process)
    DC_CUDA(cudaFuncSetAttribute(ppo_loss_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    DC_CUDA(cudaFuncSetAttribute(ppo_loss_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    ppo_loss_kernel<false><<<blocks, kTile, kSmemBytes, st>>>(hp, old_logp, adv_raw, ret, value, N, e_clip,
                                                              entropy_coef, vf_coef, dvalue, out, ws, nullptr);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_ppo_loss_fwd_bwd(const float *const logits[DC_NUM_HEADS], const uint8_t *const masks[DC_NUM_HEADS],
                                   const uint8_t *const actions[DC_NUM_HEADS], const float *old_logp,
                                   const float *adv_raw, const float *ret, const float *value, int64_t N,
                                   float e_clip, float entropy_coef, float vf_coef,
                                   float *const dlogits[DC_NUM_HEADS], float *dvalue, float *out,
                                   int32_t *n_actions, void *workspace, dc_stream_t stream) {
    const int64_t ld[DC_NUM_HEADS] = {4, 9, 9, 40, 3};
    return dc_ppo_loss_fwd_bwd_strided(logits, ld, masks, actions, old_logp, adv_raw, ret, value, 1, N, e_clip, entropy_coef,
                                       vf_coef, dlogits, ld, dvalue, 1, out, n_actions, workspace, stream);
}

extern "C" int dc_selected_logp(const float *const logits[DC_NUM_HEADS], const uint8_t *const masks[DC_NUM_HEADS],
                                const uint8_t *const actions[DC_NUM_HEADS], int64_t N, float *logp_out,
                                dc_stream_t stream) {
    DC_REQUIRE(N > 0, DC_EINVAL, "dc_selected_logp: N=%lld", (long long)N);
    DC_REQUIRE(check_heads(logits, masks, actions) && logp_out, DC_EINVAL, "dc_selected_logp: null pointer");
    HeadPtrs hp;
    for (int h = 0; h < kHeads; ++h) {
        hp.logits[h] = logits[h]; hp.masks[h] = masks[h]; hp.actions[h] = actions[h]; hp.dlogits[h] = nullptr;
        hp.ld_l[h] = head_n(h); hp.ld_d[h] = head_n(h);
    }
    hp.ld_v = 1; hp.ld_dv = 1;
    // per-device attribute: set on every call (a process-wide "done" flag breaks the second GPU of a process)
    DC_CUDA(cudaFuncSetAttribute(ppo_loss_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    const unsigned blocks = (unsigned)((N + kTile - 1) / kTile);
    ppo_loss_kernel<true><<<blocks, kTile, kSmemBytes, dc_cu_stream(stream)>>>(
        hp, nullptr, nullptr, nullptr, nullptr, N, 0.f, 0.f, 0.f, nullptr, nullptr, nullptr, logp_out);
    DC_LAUNCH_OK();
    return DC_OK;
}
