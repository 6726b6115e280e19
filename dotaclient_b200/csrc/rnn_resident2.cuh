// Weight-resident recurrence kernels for H = 128, second generation: ONE block barrier per time step.
//
// rnn_resident.cuh needs two barriers per step (mat-vec partials -> shared memory -> gate threads) and leaves half the
// CTA idle during the gate phase; ncu shows 21 % barrier stalls and 25 % short-scoreboard stalls (profiles/
// r1_recurrence_ncu.md).  Here the partial sums of the mat-vec are reduced INSIDE the warp with shuffles, the gate math
// runs on the mat-vec warps themselves, and the step's only barrier publishes the new h (fwd) / gate gradients (bwd),
// which are double-buffered in shared memory.
//
// Forward: thread (unit u, k-slice ks) owns the G gate columns {g*H + u} x rows [32 ks, 32 ks + 32) of W_hh^T -- i.e.
// 32 CONTIGUOUS floats of each of the G rows (g*H+u) of W_hh as stored, so no transposed copy is needed.  Half of
// the slab sits in registers as (row 2j, row 2j+1) pairs for the packed FFMA2, half in shared memory.  The 4 k-slices
// of a unit are the 4 lanes of a quad: after the mat-vec the quad reduce-scatters its BT x G partial sums with
// xor-shuffles so that lane b of the quad holds the complete G pre-activations of sequence b, applies the gates, keeps
// c in a register and writes h_t into the next shared-memory h buffer.
//
// Backward: thread (column group cg of 4 hidden units, j-slice js of 32 gate rows) as in rnn_resident.cuh, but the 16
// j-slices of a column group are 16 lanes of one warp: an xor-shuffle reduce-scatter leaves dh_{t-1}[b][unit] in the
// register of the lane that owns (b, unit), which then does the element-wise gate backward of the NEXT step and writes
// the gate gradients (the next mat-vec's input) to the other shared-memory buffer.
#pragma once
#include "dc_common.cuh"
#include "rnn_resident.cuh"

namespace dc_rnn2 {

using dc_rnn::bulk_g2s;
using dc_rnn::kH;
using dc_rnn::kStages;
using dc_rnn::mbar_expect_tx;
using dc_rnn::mbar_init;
using dc_rnn::mbar_wait;

constexpr int kNT = 512;
constexpr unsigned kFull = 0xffffffffu;
__host__ __device__ constexpr int pad32(int k) { return k + 4 * (k >> 5); }   // 4 floats of padding per 32: slices hit distinct banks

// ---------------------------------------------------------------------------------------------- forward
template <int G, int BT>
struct Fwd2Smem {
    static constexpr int GH = G * kH;
    static constexpr int HP = pad32(kH);                                    // 144
    static constexpr size_t w_bytes = (size_t)kNT * 8 * G * sizeof(float2);   // shared-memory half of the slab
    static constexpr size_t in_bytes = (size_t)2 * BT * HP * 4;             // double-buffered h
    static constexpr size_t stage_bytes = (size_t)BT * GH * 4;
    static constexpr size_t total = w_bytes + in_bytes + kStages * stage_bytes + kStages * 8 + 16;
};

template <int G, int BT>
__global__ void __launch_bounds__(kNT, 1) fwd2_kernel(float *__restrict__ gates, const float *__restrict__ w_hh,
                                                      const float *__restrict__ b_hh, float *__restrict__ ybuf,
                                                      float *__restrict__ cbuf, int B, int S) {
    using SM = Fwd2Smem<G, BT>;
    constexpr int GH = SM::GH, H = kH, HP = SM::HP;
    static_assert(BT == 1 || BT == 2 || BT == 4, "a quad of lanes serves up to 4 sequences");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *w_s = reinterpret_cast<float2 *>(smem_raw);
    float *in_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes);
    float *stage_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes + SM::in_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + SM::w_bytes + SM::in_bytes + kStages * SM::stage_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ks = lane & 3, u = warp * 8 + (lane >> 2);
    const int b0 = blockIdx.x * BT;
    const int nb = min(BT, B - b0);
    const uint32_t tile_bytes = (uint32_t)nb * GH * 4;

    // ---- weights: rows [32 ks, 32 ks + 32) of column g*H + u of W_hh^T == 32 contiguous floats of row g*H + u of W_hh
    float2 wr[8][G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float4 *src = reinterpret_cast<const float4 *>(w_hh + (size_t)(g * H + u) * H + ks * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(src + q);
            wr[2 * q][g] = make_float2(v.x, v.y);
            wr[2 * q + 1][g] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(src + 4 + q);
            w_s[((2 * q) * G + g) * kNT + tid] = make_float2(v.x, v.y);
            w_s[((2 * q + 1) * G + g) * kNT + tid] = make_float2(v.z, v.w);
        }
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the lane of the quad that finishes sequence b
    const bool owner = BT == 4 ? true : BT == 2 ? (ks & 1) == 0 : ks == 0;
    const int ob = BT == 4 ? ks : BT == 2 ? (ks >> 1) : 0;
    const bool live = owner && ob < nb;
    float c_reg = 0.f, bias[G];
#pragma unroll
    for (int g = 0; g < G; ++g) bias[g] = b_hh[g * H + u];
    for (int i = tid; i < 2 * BT * HP; i += kNT) in_s[i] = 0.f;
    __syncthreads();
    if (live) {
        in_s[ob * HP + pad32(u)] = ybuf[(size_t)(b0 + ob) * H + u];          // h_0 into buffer 0
        if (G == 4) c_reg = cbuf[(size_t)(b0 + ob) * H + u];                  // c_0
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < kStages && s < S; ++s) {
            mbar_expect_tx(&bars[s], tile_bytes);
            bulk_g2s(stage_s + (size_t)s * BT * GH, gates + ((size_t)s * B + b0) * GH, tile_bytes, &bars[s]);
        }
    }
    const float2 *wp = w_s + tid;
    for (int t = 0; t < S; ++t) {
        const float *hin = in_s + (t & 1) * BT * HP + ks * 36;               // pad32(32 ks) = 36 ks
        float2 acc[BT][G];
#pragma unroll
        for (int b = 0; b < BT; ++b)
#pragma unroll
            for (int g = 0; g < G; ++g) acc[b][g] = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                         // register-resident rows 0..15
            float4 hv[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) hv[b] = *reinterpret_cast<const float4 *>(hin + b * HP + 4 * q);
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                const float2 h0 = make_float2(hv[b].x, hv[b].y), h1 = make_float2(hv[b].z, hv[b].w);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    acc[b][g] = __ffma2_rn(h0, wr[2 * q][g], acc[b][g]);
                    acc[b][g] = __ffma2_rn(h1, wr[2 * q + 1][g], acc[b][g]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                         // shared-memory-resident rows 16..31
            float4 hv[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) hv[b] = *reinterpret_cast<const float4 *>(hin + b * HP + 16 + 4 * q);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float2 w = wp[((2 * q + p) * G + g) * kNT];
#pragma unroll
                    for (int b = 0; b < BT; ++b) {
                        const float2 h = p == 0 ? make_float2(hv[b].x, hv[b].y) : make_float2(hv[b].z, hv[b].w);
                        acc[b][g] = __ffma2_rn(h, w, acc[b][g]);
                    }
                }
            }
        }
        // ---- quad reduce-scatter: lane `ob` of the quad ends with the G complete sums of sequence `ob`
        float pre[G];
        if (BT == 1) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float v = acc[0][g].x + acc[0][g].y;
                v += __shfl_xor_sync(kFull, v, 1);
                v += __shfl_xor_sync(kFull, v, 2);
                pre[g] = v;
            }
        } else if (BT == 2) {
            const bool hi = (ks & 2) != 0;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float s0 = acc[0][g].x + acc[0][g].y, s1 = acc[BT - 1][g].x + acc[BT - 1][g].y;
                float v = (hi ? s1 : s0) + __shfl_xor_sync(kFull, hi ? s0 : s1, 2);
                v += __shfl_xor_sync(kFull, v, 1);
                pre[g] = v;
            }
        } else {
            const bool hi = (ks & 2) != 0, odd = (ks & 1) != 0;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float s[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) s[b] = acc[b < BT ? b : 0][g].x + acc[b < BT ? b : 0][g].y;
                const float k0 = (hi ? s[2] : s[0]) + __shfl_xor_sync(kFull, hi ? s[0] : s[2], 2);
                const float k1 = (hi ? s[3] : s[1]) + __shfl_xor_sync(kFull, hi ? s[1] : s[3], 2);
                pre[g] = (odd ? k1 : k0) + __shfl_xor_sync(kFull, odd ? k0 : k1, 1);
            }
        }
        const int st = t % kStages;
        if (owner) {
            mbar_wait(&bars[st], (t / kStages) & 1);
            if (live) {
                const float *gi = stage_s + (size_t)st * BT * GH + ob * GH;
                const size_t tok = (size_t)t * B + b0 + ob;
                float *gout = gates + tok * GH;
                float hnew;
                if (G == 3) {
                    const float r = dc_sigmoid(gi[u] + pre[0] + bias[0]);
                    const float z = dc_sigmoid(gi[H + u] + pre[1] + bias[1]);
                    const float hn = pre[2] + bias[2];
                    const float n = dc_tanh(gi[2 * H + u] + r * hn);
                    hnew = (1.0f - z) * n + z * in_s[(t & 1) * BT * HP + ob * HP + pad32(u)];
                    gout[u] = r; gout[H + u] = z; gout[2 * H + u] = n;
                    cbuf[(tok + B) * H + u] = hn;
                } else {
                    const float ig = dc_sigmoid(gi[u] + pre[0] + bias[0]);
                    const float fg = dc_sigmoid(gi[H + u] + pre[1] + bias[1]);
                    const float gg = dc_tanh(gi[2 * H + u] + pre[2] + bias[2]);
                    const float og = dc_sigmoid(gi[3 * H + u] + pre[G - 1] + bias[G - 1]);
                    c_reg = fg * c_reg + ig * gg;
                    hnew = og * dc_tanh(c_reg);
                    gout[u] = ig; gout[H + u] = fg; gout[2 * H + u] = gg; gout[3 * H + u] = og;
                    cbuf[(tok + B) * H + u] = c_reg;
                }
                ybuf[(tok + B) * H + u] = hnew;
                in_s[((t + 1) & 1) * BT * HP + ob * HP + pad32(u)] = hnew;
            }
        }
        __syncthreads();                                                      // the step's only barrier
        if (tid == 0 && t + kStages < S) {
            mbar_expect_tx(&bars[st], tile_bytes);
            bulk_g2s(stage_s + (size_t)st * BT * GH, gates + ((size_t)(t + kStages) * B + b0) * GH, tile_bytes, &bars[st]);
        }
    }
}

// ---------------------------------------------------------------------------------------------- backward
template <int G, int BT>
struct Bwd2Smem {
    static constexpr int GH = G * kH;
    static constexpr int GP = pad32(16 * 32);                               // padded gate-gradient row (16 slices)
    static constexpr size_t w_bytes = (size_t)kNT * 16 * sizeof(float4);
    static constexpr size_t in_bytes = (size_t)2 * BT * GP * 4;             // double-buffered gate gradients
    static constexpr size_t stage_floats = (size_t)BT * (GH + 3 * kH);
    static constexpr size_t total = w_bytes + in_bytes + kStages * stage_floats * 4 + kStages * 8 + 16;
};

template <int G, int BT>
__global__ void __launch_bounds__(kNT, 1) bwd2_kernel(float *__restrict__ gates, const float *__restrict__ w,
                                                      const float *__restrict__ ybuf, float *__restrict__ cbuf,
                                                      const float *__restrict__ dy, const float *__restrict__ dhn,
                                                      const float *__restrict__ dcn, float *__restrict__ dh0,
                                                      float *__restrict__ dc0, int B, int S) {
    using SM = Bwd2Smem<G, BT>;
    constexpr int GH = SM::GH, H = kH, GP = SM::GP;
    constexpr int NJS = GH / 32;                                              // 16 (LSTM) / 12 (GRU) real j-slices
    static_assert(BT == 1 || BT == 2 || BT == 4, "16 lanes x 4 columns serve up to 4 sequences");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *w_s = reinterpret_cast<float4 *>(smem_raw);
    float *in_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes);
    float *stage_s = reinterpret_cast<float *>(smem_raw + SM::w_bytes + SM::in_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + SM::w_bytes + SM::in_bytes + kStages * SM::stage_floats * 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int js = lane & 15, cg = warp * 2 + (lane >> 4);
    const int b0 = blockIdx.x * BT;
    const int nb = min(BT, B - b0);
    const uint32_t gate_bytes = (uint32_t)nb * GH * 4, row_bytes = (uint32_t)nb * H * 4;

    // ---- weights: rows [32 js, 32 js + 32) x columns [4 cg, 4 cg + 4) of W_hh [G*H, H]; slices beyond G*H are zero
    float2 wr[8][4];
    {
        const bool real = js < NJS;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
            if (real) {
                r0 = __ldg(reinterpret_cast<const float4 *>(w + (size_t)(js * 32 + 2 * j) * H) + cg);
                r1 = __ldg(reinterpret_cast<const float4 *>(w + (size_t)(js * 32 + 2 * j + 1) * H) + cg);
            }
            wr[j][0] = make_float2(r0.x, r1.x); wr[j][1] = make_float2(r0.y, r1.y);
            wr[j][2] = make_float2(r0.z, r1.z); wr[j][3] = make_float2(r0.w, r1.w);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
            if (real) {
                r0 = __ldg(reinterpret_cast<const float4 *>(w + (size_t)(js * 32 + 16 + 2 * j) * H) + cg);
                r1 = __ldg(reinterpret_cast<const float4 *>(w + (size_t)(js * 32 + 17 + 2 * j) * H) + cg);
            }
            w_s[(j * 2 + 0) * kNT + tid] = make_float4(r0.x, r1.x, r0.y, r1.y);
            w_s[(j * 2 + 1) * kNT + tid] = make_float4(r0.z, r1.z, r0.w, r1.w);
        }
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // after the 16-lane reduce-scatter, lane js owns (sequence ob, column oc) -- see the shuffle network below
    bool owner;
    int ob, oc;
    if (BT == 4) { owner = true; ob = js >> 2; oc = js & 3; }
    else if (BT == 2) { owner = (js & 1) == 0; ob = js >> 3; oc = (js >> 1) & 3; }
    else { owner = (js & 3) == 0; ob = 0; oc = js >> 2; }
    const int uu = cg * 4 + oc;                                               // hidden unit of the owned element
    const bool live = owner && ob < nb;
    float dh_rec = 0.f;      // dh flowing into the step being processed (recurrent mat-vec result + direct paths)
    float dc_carry = 0.f, c_cur = 0.f;
    if (live) {
        if (dhn) dh_rec = dhn[(size_t)(b0 + ob) * H + uu];
        if (G == 4) {
            if (dcn) dc_carry = dcn[(size_t)(b0 + ob) * H + uu];
            c_cur = cbuf[((size_t)S * B + b0 + ob) * H + uu];
        }
    }
    for (int i = tid; i < 2 * BT * GP; i += kNT) in_s[i] = 0.f;
    __syncthreads();

    auto issue = [&](int t, int st) {
        float *dst = stage_s + (size_t)st * SM::stage_floats;
        const size_t tok = (size_t)t * B + b0;
        mbar_expect_tx(&bars[st], gate_bytes + row_bytes * (G == 3 ? 3u : 2u));
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

    const float4 *wp = w_s + tid;
    for (int it = 0; it < S; ++it) {
        const int t = S - 1 - it;
        const int st = it % kStages;
        float *dg_buf = in_s + (it & 1) * BT * GP;
        // ---- element-wise gate backward for the owned (sequence, unit); writes this step's mat-vec input
        if (owner) {
            mbar_wait(&bars[st], (it / kStages) & 1);
            if (live) {
                const float *sg = stage_s + (size_t)st * SM::stage_floats;
                const float *g = sg + ob * GH;
                const float dh = sg[BT * GH + ob * H + uu] + dh_rec;
                const size_t tok = (size_t)t * B + b0 + ob;
                float *gout = gates + tok * GH;
                float *dg = dg_buf + ob * GP;
                if (G == 3) {
                    const float r = g[uu], z = g[H + uu], n = g[2 * H + uu];
                    const float hn = sg[BT * (GH + H) + ob * H + uu];
                    const float hprev = sg[BT * (GH + 2 * H) + ob * H + uu];
                    const float dpn = dh * (1.0f - z) * (1.0f - n * n);
                    const float dpz = dh * (hprev - n) * z * (1.0f - z);
                    const float dpr = dpn * hn * r * (1.0f - r);
                    const float dghn = dpn * r;
                    gout[uu] = dpr; gout[H + uu] = dpz; gout[2 * H + uu] = dpn;
                    cbuf[(tok + B) * H + uu] = dghn;
                    dg[pad32(uu)] = dpr; dg[pad32(H + uu)] = dpz; dg[pad32(2 * H + uu)] = dghn;
                    dh_rec = dh * z;                                          // direct path; the mat-vec result is added below
                } else {
                    const float ig = g[uu], fg = g[H + uu], gg = g[2 * H + uu], og = g[3 * H + uu];
                    const float cprev = sg[BT * (GH + H) + ob * H + uu];
                    const float tc = dc_tanh(c_cur);
                    const float dc = dc_carry + dh * og * (1.0f - tc * tc);
                    const float dpi = dc * gg * ig * (1.0f - ig);
                    const float dpf = dc * cprev * fg * (1.0f - fg);
                    const float dpg = dc * ig * (1.0f - gg * gg);
                    const float dpo = dh * tc * og * (1.0f - og);
                    gout[uu] = dpi; gout[H + uu] = dpf; gout[2 * H + uu] = dpg; gout[3 * H + uu] = dpo;
                    dg[pad32(uu)] = dpi; dg[pad32(H + uu)] = dpf; dg[pad32(2 * H + uu)] = dpg; dg[pad32((G - 1) * H + uu)] = dpo;
                    dc_carry = dc * fg;
                    c_cur = cprev;
                    dh_rec = 0.f;
                }
            }
        }
        __syncthreads();                                                      // the step's only barrier
        if (tid == 0 && it + kStages < S) issue(S - 1 - (it + kStages), st);
        // ---- mat-vec: partial dh_{t-1}[b][4 cg .. 4 cg + 3] over rows [32 js, 32 js + 32)
        const float *din = dg_buf + js * 36;
        float2 acc[BT][4];
#pragma unroll
        for (int b = 0; b < BT; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[b][c] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            float4 hv[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) hv[b] = *reinterpret_cast<const float4 *>(din + b * GP + 2 * j);
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
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            float4 hv[BT];
#pragma unroll
            for (int b = 0; b < BT; ++b) hv[b] = *reinterpret_cast<const float4 *>(din + b * GP + 16 + 2 * j);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const float4 wa = wp[((j + p) * 2 + 0) * kNT];
                const float4 wb = wp[((j + p) * 2 + 1) * kNT];
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
        // ---- 16-lane reduce-scatter (xor 8, 4, 2, 1): BT*4 sums -> one per owner lane
        float v[BT * 4];
#pragma unroll
        for (int b = 0; b < BT; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) v[b * 4 + c] = acc[b][c].x + acc[b][c].y;
        float res;
        if (BT == 4) {
            float k8[8], k4[4], k2[2];
            const bool b3 = (js & 8) != 0, b2 = (js & 4) != 0, b1 = (js & 2) != 0, b0_ = (js & 1) != 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) k8[i] = (b3 ? v[8 + i] : v[i]) + __shfl_xor_sync(kFull, b3 ? v[i] : v[8 + i], 8);
#pragma unroll
            for (int i = 0; i < 4; ++i) k4[i] = (b2 ? k8[4 + i] : k8[i]) + __shfl_xor_sync(kFull, b2 ? k8[i] : k8[4 + i], 4);
#pragma unroll
            for (int i = 0; i < 2; ++i) k2[i] = (b1 ? k4[2 + i] : k4[i]) + __shfl_xor_sync(kFull, b1 ? k4[i] : k4[2 + i], 2);
            res = (b0_ ? k2[1] : k2[0]) + __shfl_xor_sync(kFull, b0_ ? k2[0] : k2[1], 1);
        } else if (BT == 2) {
            float k4[4], k2[2];
            const bool b3 = (js & 8) != 0, b2 = (js & 4) != 0, b1 = (js & 2) != 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) k4[i] = (b3 ? v[(BT * 4 - 4) + i] : v[i]) + __shfl_xor_sync(kFull, b3 ? v[i] : v[(BT * 4 - 4) + i], 8);
#pragma unroll
            for (int i = 0; i < 2; ++i) k2[i] = (b2 ? k4[2 + i] : k4[i]) + __shfl_xor_sync(kFull, b2 ? k4[i] : k4[2 + i], 4);
            res = (b1 ? k2[1] : k2[0]) + __shfl_xor_sync(kFull, b1 ? k2[0] : k2[1], 2);
            res += __shfl_xor_sync(kFull, res, 1);
        } else {
            float k2[2];
            const bool b3 = (js & 8) != 0, b2 = (js & 4) != 0;
#pragma unroll
            for (int i = 0; i < 2; ++i) k2[i] = (b3 ? v[2 + i] : v[i]) + __shfl_xor_sync(kFull, b3 ? v[i] : v[2 + i], 8);
            res = (b2 ? k2[1] : k2[0]) + __shfl_xor_sync(kFull, b2 ? k2[0] : k2[1], 4);
            res += __shfl_xor_sync(kFull, res, 2);
            res += __shfl_xor_sync(kFull, res, 1);
        }
        if (owner) dh_rec += res;
    }
    if (live) {
        if (dh0) dh0[(size_t)(b0 + ob) * H + uu] = dh_rec;
        if (G == 4 && dc0) dc0[(size_t)(b0 + ob) * H + uu] = dc_carry;
    }
}

template <int G, int BT>
int launch_fwd2(float *gates, const float *w_hh, const float *b_hh, float *ybuf, float *cbuf, int B, int S, cudaStream_t st) {
    const size_t smem = Fwd2Smem<G, BT>::total;
    DC_CUDA(cudaFuncSetAttribute(fwd2_kernel<G, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fwd2_kernel<G, BT><<<(B + BT - 1) / BT, kNT, smem, st>>>(gates, w_hh, b_hh, ybuf, cbuf, B, S);
    DC_LAUNCH_OK();
    return DC_OK;
}
template <int G, int BT>
int launch_bwd2(float *gates, const float *w, const float *ybuf, float *cbuf, const float *dy, const float *dhn,
                const float *dcn, float *dh0, float *dc0, int B, int S, cudaStream_t st) {
    const size_t smem = Bwd2Smem<G, BT>::total;
    DC_CUDA(cudaFuncSetAttribute(bwd2_kernel<G, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd2_kernel<G, BT><<<(B + BT - 1) / BT, kNT, smem, st>>>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S);
    DC_LAUNCH_OK();
    return DC_OK;
}

// batch tile: 1 for a single sequence (experience prep of one rollout), 2 while the batch fits one wave, else 4
inline int pick_bt(int B) { return B == 1 ? 1 : ((B + 1) / 2 <= dc_sm_count() ? 2 : 4); }

inline int launch_fwd(int cell, float *gates, const float *w_hh, const float *b_hh, float *ybuf, float *cbuf, int B, int S,
                      cudaStream_t st) {
    const int bt = pick_bt(B);
    if (cell == DC_CELL_GRU)
        return bt == 1 ? launch_fwd2<3, 1>(gates, w_hh, b_hh, ybuf, cbuf, B, S, st)
             : bt == 2 ? launch_fwd2<3, 2>(gates, w_hh, b_hh, ybuf, cbuf, B, S, st)
                       : launch_fwd2<3, 4>(gates, w_hh, b_hh, ybuf, cbuf, B, S, st);
    return bt == 1 ? launch_fwd2<4, 1>(gates, w_hh, b_hh, ybuf, cbuf, B, S, st)
         : bt == 2 ? launch_fwd2<4, 2>(gates, w_hh, b_hh, ybuf, cbuf, B, S, st)
                   : launch_fwd2<4, 4>(gates, w_hh, b_hh, ybuf, cbuf, B, S, st);
}
inline int launch_bwd(int cell, float *gates, const float *w, const float *ybuf, float *cbuf, const float *dy,
                      const float *dhn, const float *dcn, float *dh0, float *dc0, int B, int S, cudaStream_t st) {
    const int bt = pick_bt(B);
    if (cell == DC_CELL_GRU)
        return bt == 1 ? launch_bwd2<3, 1>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st)
             : bt == 2 ? launch_bwd2<3, 2>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st)
                       : launch_bwd2<3, 4>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st);
    return bt == 1 ? launch_bwd2<4, 1>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st)
         : bt == 2 ? launch_bwd2<4, 2>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st)
                   : launch_bwd2<4, 4>(gates, w, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, B, S, st);
}

}  // namespace dc_rnn2
