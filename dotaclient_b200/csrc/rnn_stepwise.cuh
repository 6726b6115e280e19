// Step-wise recurrence for the wide layers (H a multiple of 128 other than 128 / 256, e.g. BASELINE's H = 512).
//
// At H = 512 the fp32 W_hh (4 MB LSTM; 8 MB as tf32 hi + lo) fits neither one SM nor a 16-CTA cluster, so the weights cannot
// stay on chip across steps the way rnn_resident.cuh / rnn_cluster.cuh keep them.  The step is a genuine dense contraction
// ([B, H] x [H, G*H], 1.07 GFLOP per step at B = 512), so every step runs as
//     (1) the split-K tcgen05 3xTF32 GEMM of csrc/gemm_tf32x3.cu over all SMs (W_hh streams from L2, where it stays
//         resident: 4 MB of 126 MB), writing `ksplit` partial products, and
//     (2) one elementwise gate kernel that sums the partials in a fixed order (deterministic) and applies the cell.
// 2 launches per time step; same buffers, same saved tensors and the same in-place reuse of the gate buffer as the other
// recurrence kernels (include/dotaclient_b200.h).  fp32-level accuracy like every other dense layer of the step.
#pragma once
#include "dc_common.cuh"
#include "rnn_generic.cuh"   // transpose_kernel

namespace dc_rnns {

inline bool stepwise_supported(int H) { return H % 128 == 0 && H != 128 && H != 256; }

// split-K factor: fill the SMs (tiles * ksplit <= #SM) with K ranges that stay multiples of 32 (one swizzle row)
inline int pick_ksplit(int64_t M, int N, int K) {
    const int tiles = (int)((M + 127) / 128) * (N / 128);
    int ks = 1;
    while (ks < 16 && tiles * ks * 2 <= dc_sm_count() && (K / 32) % (ks * 2) == 0) ks *= 2;
    return ks;
}

struct Workspace {
    float *wT, *part_f, *part_b, *dgbuf, *dh_carry, *dc_carry;
    int ksf, ksb;
    size_t total_floats;
};
inline Workspace carve(void *base, int cell, int B, int H) {
    const int G = cell == DC_CELL_GRU ? 3 : 4;
    Workspace w;
    w.ksf = pick_ksplit(B, G * H, H);
    w.ksb = pick_ksplit(B, H, G * H);
    float *p = reinterpret_cast<float *>(base);
    auto take = [&](size_t n) { float *q = p; p += (n + 63) / 64 * 64; return q; };
    w.wT = take((size_t)G * H * H);
    w.part_f = take((size_t)w.ksf * B * G * H);
    w.part_b = take((size_t)w.ksb * B * H);
    w.dgbuf = take((size_t)B * G * H);
    w.dh_carry = take((size_t)B * H);
    w.dc_carry = take((size_t)B * H);
    w.total_floats = (size_t)(p - reinterpret_cast<float *>(base));
    return w;
}
inline size_t workspace_bytes(int cell, int B, int H) { return carve(nullptr, cell, B, H).total_floats * sizeof(float); }

// ---- forward gate kernel: one thread per (sequence, unit) ---------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256) fwd_gate_kernel(float *__restrict__ gates_t, const float *__restrict__ part, int ksplit,
                                                       const float *__restrict__ b_hh, const float *__restrict__ h_prev,
                                                       const float *__restrict__ c_prev, float *__restrict__ h_next,
                                                       float *__restrict__ aux_next, int B, int H) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, u = idx - b * H, GH = G * H;
    const size_t stride = (size_t)B * GH;
    float pre[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float a = __ldg(b_hh + g * H + u);
        const float *p = part + (size_t)b * GH + g * H + u;
        for (int s = 0; s < ksplit; ++s) a += p[(size_t)s * stride];
        pre[g] = a;
    }
    float *g = gates_t + (size_t)b * GH + u;
    float hnew;
    if (G == 3) {
        const float r = dc_sigmoid(g[0] + pre[0]);
        const float z = dc_sigmoid(g[H] + pre[1]);
        const float n = dc_tanh(g[2 * H] + r * pre[2]);
        hnew = (1.0f - z) * n + z * h_prev[idx];
        g[0] = r; g[H] = z; g[2 * H] = n;
        aux_next[idx] = pre[2];                              // W_hn h + b_hn (cbuf slot t+1)
    } else {
        const float ig = dc_sigmoid(g[0] + pre[0]);
        const float fg = dc_sigmoid(g[H] + pre[1]);
        const float gg = dc_tanh(g[2 * H] + pre[2]);
        const float og = dc_sigmoid(g[(G - 1) * H] + pre[G - 1]);
        const float c = fg * c_prev[idx] + ig * gg;
        hnew = og * dc_tanh(c);
        g[0] = ig; g[H] = fg; g[2 * H] = gg; g[(G - 1) * H] = og;
        aux_next[idx] = c;                                   // c_t (cbuf slot t+1)
    }
    h_next[idx] = hnew;
}

// ---- backward gate kernel -------------------------------------------------------------------------------------------------
// dh = dy_t + carry + sum of the previous step's partial products; writes dgi (in place of the saved gates), the h2h gate
// gradients (GEMM operand of this step) and the carries.  first != 0: the carries are initialised from dhn / dcn.
template <int G>
__global__ void __launch_bounds__(256) bwd_gate_kernel(float *__restrict__ gates_t, const float *__restrict__ dy_t,
                                                       const float *__restrict__ part, int ksplit, float *__restrict__ dh_carry,
                                                       float *__restrict__ dc_carry, const float *__restrict__ dhn,
                                                       const float *__restrict__ dcn, const float *__restrict__ h_prev,
                                                       const float *__restrict__ c_prev, float *aux_cur, float *__restrict__ dgbuf,
                                                       int first, int B, int H) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, u = idx - b * H, GH = G * H;
    float dh = dy_t[idx];
    if (first) {
        if (dhn) dh += dhn[idx];
    } else {
        dh += dh_carry[idx];
        const size_t stride = (size_t)B * H;
        for (int s = 0; s < ksplit; ++s) dh += part[(size_t)s * stride + idx];
    }
    float *g = gates_t + (size_t)b * GH + u;
    float *dg = dgbuf + (size_t)b * GH + u;
    if (G == 3) {
        const float r = g[0], z = g[H], n = g[2 * H];
        const float hn = aux_cur[idx], hprev = h_prev[idx];
        const float dpn = dh * (1.0f - z) * (1.0f - n * n);
        const float dpz = dh * (hprev - n) * z * (1.0f - z);
        const float dpr = dpn * hn * r * (1.0f - r);
        const float dghn = dpn * r;
        g[0] = dpr; g[H] = dpz; g[2 * H] = dpn;
        aux_cur[idx] = dghn;
        dg[0] = dpr; dg[H] = dpz; dg[2 * H] = dghn;
        dh_carry[idx] = dh * z;
    } else {
        const float ig = g[0], fg = g[H], gg = g[2 * H], og = g[(G - 1) * H];
        const float c = aux_cur[idx], cprev = c_prev[idx];
        const float tc = dc_tanh(c);
        const float dcin = first ? (dcn ? dcn[idx] : 0.f) : dc_carry[idx];
        const float dc = dcin + dh * og * (1.0f - tc * tc);
        const float dpi = dc * gg * ig * (1.0f - ig);
        const float dpf = dc * cprev * fg * (1.0f - fg);
        const float dpg = dc * ig * (1.0f - gg * gg);
        const float dpo = dh * tc * og * (1.0f - og);
        g[0] = dpi; g[H] = dpf; g[2 * H] = dpg; g[(G - 1) * H] = dpo;
        dg[0] = dpi; dg[H] = dpf; dg[2 * H] = dpg; dg[(G - 1) * H] = dpo;
        dc_carry[idx] = dc * fg;
        dh_carry[idx] = 0.f;
    }
}

__global__ void __launch_bounds__(256) bwd_final_kernel(const float *__restrict__ part, int ksplit, const float *__restrict__ dh_carry,
                                                        const float *__restrict__ dc_carry, float *__restrict__ dh0,
                                                        float *__restrict__ dc0, int n) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= n) return;
    float dh = dh_carry[idx];
    for (int s = 0; s < ksplit; ++s) dh += part[(size_t)s * n + idx];
    if (dh0) dh0[idx] = dh;
    if (dc0) dc0[idx] = dc_carry[idx];
}

inline int launch_fwd(int cell, float *gates, const float *w_hh, const float *b_hh, float *ybuf, float *cbuf, int B, int S, int H,
                      void *workspace, cudaStream_t st) {
    const int G = cell == DC_CELL_GRU ? 3 : 4, GH = G * H;
    const Workspace w = carve(workspace, cell, B, H);
    const int blocks = (B * H + 255) / 256;
    const size_t BH = (size_t)B * H;
    for (int t = 0; t < S; ++t) {
        int rc = dc_gemm_tf32x3_splitk(ybuf + t * BH, H, w_hh, H, w.part_f, B, GH, H, w.ksf, t == 0, st);
        if (rc) return rc;
        float *gt = gates + (size_t)t * B * GH;
        if (G == 3)
            fwd_gate_kernel<3><<<blocks, 256, 0, st>>>(gt, w.part_f, w.ksf, b_hh, ybuf + t * BH, nullptr, ybuf + (t + 1) * BH, cbuf + (t + 1) * BH, B, H);
        else
            fwd_gate_kernel<4><<<blocks, 256, 0, st>>>(gt, w.part_f, w.ksf, b_hh, ybuf + t * BH, cbuf + t * BH, ybuf + (t + 1) * BH, cbuf + (t + 1) * BH, B, H);
    }
    DC_LAUNCH_OK();
    return DC_OK;
}

inline int launch_bwd(int cell, float *gates, const float *w_hh, const float *ybuf, float *cbuf, const float *dy, const float *dhn,
                      const float *dcn, float *dh0, float *dc0, int B, int S, int H, void *workspace, cudaStream_t st) {
    const int G = cell == DC_CELL_GRU ? 3 : 4, GH = G * H;
    const Workspace w = carve(workspace, cell, B, H);
    const int blocks = (B * H + 255) / 256;
    const size_t BH = (size_t)B * H;
    dim3 tb(32, 8), tg((H + 31) / 32, (GH + 31) / 32);
    dc_rnn::transpose_kernel<<<tg, tb, 0, st>>>(w_hh, w.wT, GH, H);            // W_hh^T [H, G*H]: the GEMM's [N, K] operand
    for (int it = 0; it < S; ++it) {
        const int t = S - 1 - it;
        float *gt = gates + (size_t)t * B * GH;
        if (G == 3)
            bwd_gate_kernel<3><<<blocks, 256, 0, st>>>(gt, dy + t * BH, w.part_b, w.ksb, w.dh_carry, w.dc_carry, dhn, nullptr, ybuf + t * BH,
                                                       nullptr, cbuf + (t + 1) * BH, w.dgbuf, it == 0, B, H);
        else
            bwd_gate_kernel<4><<<blocks, 256, 0, st>>>(gt, dy + t * BH, w.part_b, w.ksb, w.dh_carry, w.dc_carry, dhn, dcn, nullptr,
                                                       cbuf + t * BH, cbuf + (t + 1) * BH, w.dgbuf, it == 0, B, H);
        int rc = dc_gemm_tf32x3_splitk(w.dgbuf, GH, w.wT, GH, w.part_b, B, H, GH, w.ksb, it == 0, st);
        if (rc) return rc;
    }
    bwd_final_kernel<<<blocks, 256, 0, st>>>(w.part_b, w.ksb, w.dh_carry, G == 4 ? w.dc_carry : nullptr, dh0, G == 4 ? dc0 : nullptr, B * H);
    DC_LAUNCH_OK();
    return DC_OK;
}

}  // namespace dc_rnns
