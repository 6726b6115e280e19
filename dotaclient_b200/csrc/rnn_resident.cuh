// Weight-resident recurrence kernels for H = 128 (GRU and LSTM), forward and backward.
//
// The time loop of nn.GRU / nn.LSTM (policy.py:66,141) is strictly sequential in t, so a launch is
// bound by the latency of ONE step, 2*S times per optimizer step.  Design:
//
//   * one CTA owns kBT sequences for all S steps -- no inter-CTA traffic, no grid sync;
//   * W_hh never leaves the SM: every thread owns a 32 x 4 slab of the [M, Nout] weight matrix,
//     16 rows in REGISTERS (64 regs) and 16 rows in SHARED MEMORY (128 KB LSTM / 96 KB GRU), read
//     as conflict-free 16-byte loads.  (fp32 W_hh is 256 KB for the LSTM: neither the register
//     file nor shared memory alone can hold it, together they can.)
//   * the step's batched mat-vec is FFMA on the fp32 pipe (tensor cores are deliberately not used
//     here: with kBT = 2..4 rows per CTA it is not a dense contraction -- north_star);
//     each warp handles one 32-row slice of the contraction (its input slice is a broadcast
//     shared-memory read) and 128 output columns; the 4 (fwd) / 12-16 (bwd) partial sums per
//     output are combined through shared memory by the gate threads;
//   * the per-step global inputs (i2h pre-activations forward; saved gates, dy, c / hn, h_{t-1}
//     backward) are contiguous [kBT, ...] tiles in the time-major layout and are prefetched
//     kStages steps ahead with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//   * h (fwd) / the gate gradients (bwd) -- the mat-vec input of the next step -- stay in shared
//     memory; c (fwd) and dc (bwd) stay in registers of the thread that owns the (b, unit) pair.
//
// Forward and backward are the same skeleton around one templated mat-vec:
//     fwd: out[b][j] = sum_k h[b][k]   * W_hh^T[k][j]     M = H,   Nout = G*H
//     bwd: out[b][k] = sum_j dg[b][j]  * W_hh[j][k]       M = G*H, Nout = H
//
// Algorithmic HBM bytes per token: fwd reads G*H*4 (gi) and writes (G+2)*H*4 (gates, h, c|hn);
// bwd reads (G+3)*H*4 and writes G*H*4 (+H*4 GRU).
#pragma once
#include "dc_common.cuh"

namespace dc_rnn {

constexpr int kH = 128;
constexpr int kStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

// Thread layout of the stationary mat-vec: tid = ms * NCG + cg; the thread owns rows [ms*32, ms*32+32) and
// columns [cg*4, cg*4+4) of A [M, NOUT] (row-major in global memory).
template <int NT, int NOUT>
struct Tiling {
    static constexpr int NCG = NOUT / 4;
    static constexpr int NMS = NT / NCG;
    static constexpr int M = NMS * 32;
};

// Weights are held as (row 2j, row 2j+1) PAIRS so that the inner product runs on the packed fp32x2 FMA of
// sm_100 (FFMA2, __ffma2_rn): one instruction = two FMAs on 64-bit register pairs.  A plain 3-register FFMA issues
// at half rate on this SM (register-bank limited); FFMA2 restores the full 128 FMA/clk/SM.  The (h[2j], h[2j+1])
// operand pairs fall out of the 16-byte shared-memory loads for free; even and odd rows accumulate separately and
// are added once at the end.
template <int NT, int NOUT>
__device__ __forceinline__ void load_weights(const float *__restrict__ A, float2 (&wr)[8][4], float4 *w_s) {
    using T = Tiling<NT, NOUT>;
    const int cg = threadIdx.x % T::NCG, ms = threadIdx.x / T::NCG;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 r0 = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(ms * 32 + 2 * j) * NOUT) + cg);
        const float4 r1 = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(ms * 32 + 2 * j + 1) * NOUT) + cg);
        wr[j][0] = make_float2(r0.x, r1.x); wr[j][1] = make_float2(r0.y, r1.y);
        wr[j][2] = make_float2(r0.z, r1.z); wr[j][3] = make_float2(r0.w, r1.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 r0 = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(ms * 32 + 16 + 2 * j) * NOUT) + cg);
        const float4 r1 = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(ms * 32 + 17 + 2 * j) * NOUT) + cg);
        w_s[((ms * 8 + j) * 2 + 0) * T::NCG + cg] = make_float4(r0.x, r1.x, r0.y, r1.y);
        w_s[((ms * 8 + j) * 2 + 1) * T::NCG + cg] = make_float4(r0.z, r1.z, r0.w, r1.w);
    }
}

// part_s[ms][b][n] = sum over this thread's 32 rows of in_s[b][m] * A[m][n]
template <int NT, int NOUT, int BT>
__device__ __forceinline__ void matvec_partial(const float2 (&wr)[8][4], const float4 *w_s, const float *in_s, float *part_s) {
    using T = Tiling<NT, NOUT>;
    const int cg = threadIdx.x % T::NCG, ms = threadIdx.x / T::NCG;
    float2 acc[BT][4];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[b][c] = make_float2(0.f, 0.f);
    const float *in0 = in_s + ms * 32;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {            // register-resident rows 0..15
        float4 hv[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) hv[b] = *reinterpret_cast<const float4 *>(in0 + b * T::M + 2 * j);
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            const float2 h0 = make_float2(hv[b].x, hv[b].y), h1 = make_float2(hv[b].z, hv[b].w);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[b][c] = __ffma2_rn(h0, wr[j][c], acc[b][c]);
                acc[b][c] = __ffma2_rn(h1, wr[j + 1][c], acc[b][c]);
            }
        }
    }
    const float4 *wp = w_s + (ms * 16) * T::NCG + cg;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {            // shared-memory-resident rows 16..31
        float4 hv[BT];
#pragma unroll
        for (int b = 0; b < BT; ++b) hv[b] = *reinterpret_cast<const float4 *>(in0 + b * T::M + 16 + 2 * j);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const float4 wa = wp[((j + p) * 2 + 0) * T::NCG];
            const float4 wb = wp[((j + p) * 2 + 1) * T::NCG];
            const float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w);
            const float2 w2 = make_float2(wb.x, wb.y), w3 = make_float2(wb.z, wb.w);
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                const float2 h = p == 0 ? make_float2(hv[b].x, hv[b].y) : make_float2(hv[b].z, hv[b].w);
                acc[b][0] = __ffma2_rn(h, w0, acc[b][0]);
                acc[b][1] = __ffma2_rn(h, w1, acc[b][1]);
                acc[b][2] = __ffma2_rn(h, w2, acc[b][2]);
                acc[b][3] = __ffma2_rn(h, w3, acc[b][3]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < BT; ++b)
        *reinterpret_cast<float4 *>(part_s + ((ms * BT + b) * NOUT) + cg * 4) =
            make_float4(acc[b][0].x + acc[b][0].y, acc[b][1].x + acc[b][1].y, acc[b][2].x + acc[b][2].y, acc[b][3].x + acc[b][3].y);
}

template <int G, int BT>
struct FwdSmem {
    static constexpr int NT = G * kH;
    static constexpr int GH = G * kH;
    static constexpr size_t w_bytes = (size_t)NT * 16 * sizeof(float4);          // smem half of W_hh^T
    static constexpr size_t in_bytes = (size_t)BT * kH * 4;                       // h
    static constexpr size_t part_bytes = (size_t)4 * BT * GH * 4;                 // NMS = 4 partials
    static constexpr size_t stage_bytes = (size_t)BT * GH * 4;                    // gi tile
    static constexpr size_t total = w_bytes + in_bytes + part_bytes + kStages * stage_bytes + kStages * 8 + 16;
};

// ---- forward --------------------------------------------------------------------------------
template <int G, int BT>
__global__ void __launch_bounds__(G *kH, 1) fwd_resident_kernel(float *__restrict__ gates, const float *__restrict__ wT,
                                                                 const float *__restrict__ b_hh, float *__restrict__ ybuf,
                                                                 float *__restrict__ cbuf, int B, int S) {
    using SM = FwdSmem<G, BT>;
    constexpr int NT = SM::NT, GH = SM::GH, H = kH;
    static_assert(BT * kH <= G * kH, "one gate thread per (sequence, unit) pair");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *w_s = reinterpret_cast<float4 *>(smem_raw);
    float *in_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes);
    float *part_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes + SM::in_bytes);
    float *stage_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes + SM::in_bytes + SM::part_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + SM::w_bytes + SM::in_bytes + SM::part_bytes + kStages * SM::stage_bytes);

    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * BT;
    const int nb = min(BT, B - b0);
    const uint32_t tile_bytes = (uint32_t)nb * GH * 4;

    float2 wr[8][4];
    load_weights<NT, GH>(wT, wr, w_s);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // unit threads: (b, u) pairs
    const bool unit = tid < BT * H;
    const int ub = tid / H, uu = tid % H;
    const bool live = unit && ub < nb;
    float c_reg = 0.f, bias[G];
    if (unit) {
        in_s[ub * H + uu] = live ? ybuf[(size_t)(b0 + ub) * H + uu] : 0.f;      // h_0
        if (G == 4 && live) c_reg = cbuf[(size_t)(b0 + ub) * H + uu];           // c_0
#pragma unroll
        for (int g = 0; g < G; ++g) bias[g] = b_hh[g * H + uu];
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < kStages && s < S; ++s) {
            mbar_expect_tx(&bars[s], tile_bytes);
            bulk_g2s(stage_s + (size_t)s * BT * GH, gates + ((size_t)s * B + b0) * GH, tile_bytes, &bars[s]);
        }
    }
    for (int t = 0; t < S; ++t) {
        matvec_partial<NT, GH, BT>(wr, w_s, in_s, part_s);
        __syncthreads();
        const int st = t % kStages;
        if (unit) {
            mbar_wait(&bars[st], (t / kStages) & 1);
            if (live) {
                const float *gi = stage_s + (size_t)st * BT * GH + ub * GH;
                float pre[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float a = bias[g];
#pragma unroll
                    for (int ms = 0; ms < 4; ++ms) a += part_s[(ms * BT + ub) * GH + g * H + uu];
                    pre[g] = a;
                }
                const size_t tok = (size_t)t * B + b0 + ub;
                float *gout = gates + tok * GH;
                float hnew;
                if (G == 3) {
                    const float r = dc_sigmoid(gi[uu] + pre[0]);
                    const float z = dc_sigmoid(gi[H + uu] + pre[1]);
                    const float n = dc_tanh(gi[2 * H + uu] + r * pre[2]);
                    hnew = (1.0f - z) * n + z * in_s[ub * H + uu];
                    gout[uu] = r; gout[H + uu] = z; gout[2 * H + uu] = n;
                    cbuf[(tok + B) * H + uu] = pre[2];                             // W_hn h + b_hn, slot t+1
                } else {
                    const float ig = dc_sigmoid(gi[uu] + pre[0]);
                    const float fg = dc_sigmoid(gi[H + uu] + pre[1]);
                    const float gg = dc_tanh(gi[2 * H + uu] + pre[2]);
                    const float og = dc_sigmoid(gi[3 * H + uu] + pre[G - 1]);
                    c_reg = fg * c_reg + ig * gg;
                    hnew = og * dc_tanh(c_reg);
                    gout[uu] = ig; gout[H + uu] = fg; gout[2 * H + uu] = gg; gout[3 * H + uu] = og;
                    cbuf[(tok + B) * H + uu] = c_reg;
                }
                ybuf[(tok + B) * H + uu] = hnew;
                in_s[ub * H + uu] = hnew;
            }
        }
        __syncthreads();
        if (tid == 0 && t + kStages < S) {      // every reader of this stage is behind the barrier: refill it
            mbar_expect_tx(&bars[st], tile_bytes);
            bulk_g2s(stage_s + (size_t)st * BT * GH, gates + ((size_t)(t + kStages) * B + b0) * GH, tile_bytes, &bars[st]);
        }
    }
}

// ---- backward -------------------------------------------------------------------------------
template <int G, int BT>
struct BwdSmem {
    static constexpr int NT = G * kH;
    static constexpr int GH = G * kH;
    static constexpr int NMS = NT / (kH / 4);                                     // 16 (LSTM) / 12 (GRU)
    static constexpr size_t w_bytes = (size_t)NT * 16 * sizeof(float4);
    static constexpr size_t in_bytes = (size_t)BT * GH * 4;                       // gate gradients (mat-vec input)
    static constexpr size_t part_bytes = (size_t)NMS * BT * kH * 4;
    // stage: gates tile [BT, GH] + dy [BT, H] + aux1 [BT, H] (LSTM c_{t-1} | GRU hn) + aux2 [BT, H] (GRU h_{t-1})
    static constexpr size_t stage_floats = (size_t)BT * (GH + 3 * kH);
    static constexpr size_t total = w_bytes + in_bytes + part_bytes + kStages * stage_floats * 4 + kStages * 8 + 16;
};

template <int G, int BT>
__global__ void __launch_bounds__(G *kH, 1) bwd_resident_kernel(float *__restrict__ gates, const float *__restrict__ w,
                                                                 const float *__restrict__ ybuf, float *__restrict__ cbuf,
                                                                 const float *__restrict__ dy, const float *__restrict__ dhn,
                                                                 const float *__restrict__ dcn, float *__restrict__ dh0,
                                                                 float *__restrict__ dc0, int B, int S) {
    using SM = BwdSmem<G, BT>;
    constexpr int NT = SM::NT, GH = SM::GH, H = kH, NMS = SM::NMS;
    static_assert(BT * kH <= G * kH, "one gate thread per (sequence, unit) pair");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *w_s = reinterpret_cast<float4 *>(smem_raw);
    float *in_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes);
    float *part_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes + SM::in_bytes);
    float *stage_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes + SM::in_bytes + SM::part_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + SM::w_bytes + SM::in_bytes + SM::part_bytes + kStages * SM::stage_floats * 4);

    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * BT;
    const int nb = min(BT, B - b0);
    const uint32_t gate_bytes = (uint32_t)nb * GH * 4, row_bytes = (uint32_t)nb * H * 4;

    float2 wr[8][4];
    load_weights<NT, H>(w, wr, w_s);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool unit = tid < BT * H;
    const int ub = tid / H, uu = tid % H;
    const bool live = unit && ub < nb;
    // recurrent gradients owned by the (b, u) thread
    float dh_carry = 0.f;      // GRU: direct path dh * z;  both: initial dh_n
    float dc_carry = 0.f;      // LSTM: dL/dc flowing to step t-1
    float c_cur = 0.f;         // LSTM: c_t of the step being processed (slot t+1)
    if (live) {
        if (dhn) dh_carry = dhn[(size_t)(b0 + ub) * H + uu];
        if (G == 4) {
            if (dcn) dc_carry = dcn[(size_t)(b0 + ub) * H + uu];
            c_cur = cbuf[((size_t)S * B + b0 + ub) * H + uu];
        }
    }
    for (int i = tid; i < NMS * BT * H; i += NT) part_s[i] = 0.f;
    for (int i = tid; i < BT * GH; i += NT) in_s[i] = 0.f;
    __syncthreads();

    auto issue = [&](int t, int st) {
        float *dst = stage_s + (size_t)st * SM::stage_floats;
        const size_t tok = (size_t)t * B + b0;
        const uint32_t total = gate_bytes + row_bytes * (G == 3 ? 3u : 2u);
        mbar_expect_tx(&bars[st], total);
        bulk_g2s(dst, gates + tok * GH, gate_bytes, &bars[st]);
        bulk_g2s(dst + BT * GH, dy + tok * H, row_bytes, &bars[st]);
        if (G == 4) {
            bulk_g2s(dst + BT * (GH + H), cbuf + tok * H, row_bytes, &bars[st]);            // c_{t-1} (slot t)
        } else {
            bulk_g2s(dst + BT * (GH + H), cbuf + (tok + B) * H, row_bytes, &bars[st]);      // hn (slot t+1)
            bulk_g2s(dst + BT * (GH + 2 * H), ybuf + tok * H, row_bytes, &bars[st]);        // h_{t-1} (slot t)
        }
    };
    if (tid == 0)
        for (int s = 0; s < kStages && s < S; ++s) issue(S - 1 - s, s);

    for (int it = 0; it < S; ++it) {
        const int t = S - 1 - it;
        const int st = it % kStages;
        if (unit) {
            mbar_wait(&bars[st], (it / kStages) & 1);
            if (live) {
                const float *sg = stage_s + (size_t)st * SM::stage_floats;
                const float *g = sg + ub * GH;
                float dh = sg[BT * GH + ub * H + uu] + dh_carry;
#pragma unroll
                for (int ms = 0; ms < NMS; ++ms) dh += part_s[(ms * BT + ub) * H + uu];
                const size_t tok = (size_t)t * B + b0 + ub;
                float *gout = gates + tok * GH;
                float *dg = in_s + ub * GH;
                if (G == 3) {
                    const float r = g[uu], z = g[H + uu], n = g[2 * H + uu];
                    const float hn = sg[BT * (GH + H) + ub * H + uu];
                    const float hprev = sg[BT * (GH + 2 * H) + ub * H + uu];
                    const float dpn = dh * (1.0f - z) * (1.0f - n * n);
                    const float dpz = dh * (hprev - n) * z * (1.0f - z);
                    const float dpr = dpn * hn * r * (1.0f - r);
                    const float dghn = dpn * r;
                    gout[uu] = dpr; gout[H + uu] = dpz; gout[2 * H + uu] = dpn;
                    cbuf[(tok + B) * H + uu] = dghn;
                    dg[uu] = dpr; dg[H + uu] = dpz; dg[2 * H + uu] = dghn;
                    dh_carry = dh * z;
                } else {
                    const float ig = g[uu], fg = g[H + uu], gg = g[2 * H + uu], og = g[3 * H + uu];
                    const float cprev = sg[BT * (GH + H) + ub * H + uu];
                    const float tc = dc_tanh(c_cur);
                    const float dc = dc_carry + dh * og * (1.0f - tc * tc);
                    const float dpi = dc * gg * ig * (1.0f - ig);
                    const float dpf = dc * cprev * fg * (1.0f - fg);
                    const float dpg = dc * ig * (1.0f - gg * gg);
                    const float dpo = dh * tc * og * (1.0f - og);
                    gout[uu] = dpi; gout[H + uu] = dpf; gout[2 * H + uu] = dpg; gout[3 * H + uu] = dpo;
                    dg[uu] = dpi; dg[H + uu] = dpf; dg[2 * H + uu] = dpg; dg[3 * H + uu] = dpo;
                    dc_carry = dc * fg;
                    c_cur = cprev;
                    dh_carry = 0.f;
                }
            }
        }
        __syncthreads();
        if (tid == 0 && it + kStages < S) issue(S - 1 - (it + kStages), st);
        matvec_partial<NT, H, BT>(wr, w_s, in_s, part_s);
        __syncthreads();
    }
    if (live) {
        float dh = dh_carry;
#pragma unroll
        for (int ms = 0; ms < NMS; ++ms) dh += part_s[(ms * BT + ub) * H + uu];
        if (dh0) dh0[(size_t)(b0 + ub) * H + uu] = dh;
        if (G == 4 && dc0) dc0[(size_t)(b0 + ub) * H + uu] = dc_carry;
    }
}

inline bool resident_supported(int cell, int H) { (void)cell; return H == kH; }

template <int G, int BT>
int launch_fwd_t(float *gates, const float *wT, const float *b_hh, float *ybuf, float *cbuf, int B, int S, cudaStream_t st) {
    const size_t smem = FwdSmem<G, BT>::total;
    DC_CUDA(cudaFuncSetAttribute(fwd_resident_kernel<G, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fwd_resident_kernel<G, BT><<<(B + BT - 1) / BT, G * kH, smem, st>>>(gates, wT, b_hh, ybuf, cbuf, B, S);
    DC_LAUNCH_OK();
    return DC_OK;
}
template <int G, int BT>
int launch_bwd_t(float *gates, const float *w, const float *ybuf, float *cbuf, const float *dy, const float *dhn,
                 const float *dcn, float *dh0, float *dc0, int B, int S, cudaStream_t st) {
    const size_t smem = BwdSmem<G, BT>::total;
    DC_CUDA(cudaFuncSetAttribute(bwd_resident_kernel<G, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_resident_kernel<G, BT><<<(B + BT - 1) / BT, G * kH, smem, st>>>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S);
    DC_LAUNCH_OK();
    return DC_OK;
}

// kBT: the smallest batch tile that still fits the batch in one wave of CTAs (1 CTA / SM).  The gate phase maps one
// thread to one (sequence, unit) pair, so kBT * H must not exceed the CTA size G * H: kBT <= 3 (GRU) / 4 (LSTM).
inline bool small_tile(int B) { return (B + 1) / 2 <= dc_sm_count(); }

inline int launch_fwd_resident(int cell, float *gates, const float *wT, const float *b_hh, float *ybuf, float *cbuf, int B,
                               int S, int H, cudaStream_t st) {
    (void)H;
    if (cell == DC_CELL_GRU)
        return small_tile(B) ? launch_fwd_t<3, 2>(gates, wT, b_hh, ybuf, cbuf, B, S, st)
                             : launch_fwd_t<3, 3>(gates, wT, b_hh, ybuf, cbuf, B, S, st);
    return small_tile(B) ? launch_fwd_t<4, 2>(gates, wT, b_hh, ybuf, cbuf, B, S, st)
                         : launch_fwd_t<4, 4>(gates, wT, b_hh, ybuf, cbuf, B, S, st);
}
inline int launch_bwd_resident(int cell, float *gates, const float *w, const float *ybuf, float *cbuf, const float *dy,
                               const float *dhn, const float *dcn, float *dh0, float *dc0, int B, int S, int H, cudaStream_t st) {
    (void)H;
    if (cell == DC_CELL_GRU)
        return small_tile(B) ? launch_bwd_t<3, 2>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st)
                             : launch_bwd_t<3, 3>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st);
    return small_tile(B) ? launch_bwd_t<4, 2>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st)
                         : launch_bwd_t<4, 4>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st);
}

}  // namespace dc_rnn
