// Shape-generic recurrence kernels (any H that is a multiple of 4, GRU or LSTM).
//
// Fallback for widths the register/shared-memory-resident kernels in rnn_seq.cu do not cover.
// One CTA owns kBT sequences for all S steps (no inter-CTA communication); h lives in shared
// memory; W_hh is streamed from L2 every step (it is <= 4 MB and stays L2-resident), read with
// fully coalesced loads (forward reads the [H, G*H] transpose held in the workspace, backward
// reads W_hh [G*H, H] as stored).  Correct everywhere, FMA/L2-bound at large H.
#pragma once
#include "dc_common.cuh"

namespace dc_rnn {

constexpr int kBT = 4;         // sequences per CTA
constexpr int kThreads = 256;

__global__ void transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
}

// Forward.  gates [S,B,G,H] (in: x W_ih^T + b_ih; out: activated gates), wT [H, G*H].
template <int G>
__global__ void __launch_bounds__(kThreads) fwd_generic_kernel(float *__restrict__ gates, const float *__restrict__ wT,
                                                               const float *__restrict__ b_hh, float *__restrict__ ybuf,
                                                               float *__restrict__ cbuf, int B, int S, int H) {
    extern __shared__ __align__(16) float smem[];
    float *h_s = smem;                 // [kBT][H]
    float *pre_s = smem + kBT * H;     // [kBT][G*H]
    const int GH = G * H;
    const int b0 = blockIdx.x * kBT;
    const int nb = min(kBT, B - b0);
    for (int i = threadIdx.x; i < kBT * H; i += kThreads) {
        const int b = i / H, u = i % H;
        h_s[i] = b < nb ? ybuf[(size_t)(b0 + b) * H + u] : 0.f;
    }
    __syncthreads();
    for (int t = 0; t < S; ++t) {
        for (int j = threadIdx.x; j < GH; j += kThreads) {
            float acc[kBT];
            const float bj = b_hh[j];
#pragma unroll
            for (int b = 0; b < kBT; ++b) acc[b] = bj;
            for (int k = 0; k < H; k += 4) {
                float w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) w[q] = __ldg(wT + (size_t)(k + q) * GH + j);
#pragma unroll
                for (int b = 0; b < kBT; ++b) {
                    const float4 hv = *reinterpret_cast<const float4 *>(h_s + b * H + k);
                    acc[b] = fmaf(hv.x, w[0], acc[b]);
                    acc[b] = fmaf(hv.y, w[1], acc[b]);
                    acc[b] = fmaf(hv.z, w[2], acc[b]);
                    acc[b] = fmaf(hv.w, w[3], acc[b]);
                }
            }
#pragma unroll
            for (int b = 0; b < kBT; ++b) pre_s[b * GH + j] = acc[b];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nb * H; i += kThreads) {
            const int b = i / H, u = i % H;
            const size_t tok = (size_t)t * B + b0 + b;
            float *g = gates + tok * GH;
            const float *pre = pre_s + b * GH;
            float hnew;
            if (G == 3) {   // GRU: r, z, n (torch.nn.GRU)
                const float r = dc_sigmoid(g[u] + pre[u]);
                const float z = dc_sigmoid(g[H + u] + pre[H + u]);
                const float hn = pre[2 * H + u];
                const float n = dc_tanh(g[2 * H + u] + r * hn);
                hnew = (1.0f - z) * n + z * h_s[b * H + u];
                g[u] = r; g[H + u] = z; g[2 * H + u] = n;
                cbuf[((size_t)(t + 1) * B + b0 + b) * H + u] = hn;
            } else {        // LSTM: i, f, g, o (torch.nn.LSTM)
                const float ig = dc_sigmoid(g[u] + pre[u]);
                const float fg = dc_sigmoid(g[H + u] + pre[H + u]);
                const float gg = dc_tanh(g[2 * H + u] + pre[2 * H + u]);
                const float og = dc_sigmoid(g[3 * H + u] + pre[3 * H + u]);
                const float c = fg * cbuf[((size_t)t * B + b0 + b) * H + u] + ig * gg;
                hnew = og * dc_tanh(c);
                g[u] = ig; g[H + u] = fg; g[2 * H + u] = gg; g[3 * H + u] = og;
                cbuf[((size_t)(t + 1) * B + b0 + b) * H + u] = c;
            }
            ybuf[((size_t)(t + 1) * B + b0 + b) * H + u] = hnew;
            h_s[b * H + u] = hnew;   // each (b,u) is owned by one thread; matvec readers are behind the barrier
        }
        __syncthreads();
    }
}

// Backward.  gates in: activated gates, out: dgi.  w [G*H, H] as stored.
template <int G>
__global__ void __launch_bounds__(kThreads) bwd_generic_kernel(float *__restrict__ gates, const float *__restrict__ w,
                                                               const float *__restrict__ ybuf, float *__restrict__ cbuf,
                                                               const float *__restrict__ dy, const float *__restrict__ dhn,
                                                               const float *__restrict__ dcn, float *__restrict__ dh0,
                                                               float *__restrict__ dc0, int B, int S, int H) {
    extern __shared__ __align__(16) float smem[];
    float *dh_s = smem;                    // [kBT][H] recurrent gradient wrt h
    float *dc_s = smem + kBT * H;          // [kBT][H] (LSTM) recurrent gradient wrt c
    float *dg_s = smem + 2 * kBT * H;      // [kBT][G*H] gradient wrt the hidden-to-hidden pre-activations
    const int GH = G * H;
    const int b0 = blockIdx.x * kBT;
    const int nb = min(kBT, B - b0);
    for (int i = threadIdx.x; i < kBT * H; i += kThreads) {
        const int b = i / H, u = i % H;
        dh_s[i] = (b < nb && dhn) ? dhn[(size_t)(b0 + b) * H + u] : 0.f;
        dc_s[i] = (b < nb && dcn) ? dcn[(size_t)(b0 + b) * H + u] : 0.f;
    }
    for (int i = threadIdx.x; i < kBT * GH; i += kThreads) dg_s[i] = 0.f;
    __syncthreads();
    const int nslice = (kThreads >= H) ? kThreads / H : 1;   // split the contraction when H < kThreads
    for (int t = S - 1; t >= 0; --t) {
        for (int i = threadIdx.x; i < nb * H; i += kThreads) {
            const int b = i / H, u = i % H;
            const size_t tok = (size_t)t * B + b0 + b;
            float *g = gates + tok * GH;
            float *dg = dg_s + b * GH;
            const float dh = dy[tok * H + u] + dh_s[b * H + u];
            if (G == 3) {
                const float r = g[u], z = g[H + u], n = g[2 * H + u];
                const size_t ci = ((size_t)(t + 1) * B + b0 + b) * H + u;
                const float hn = cbuf[ci];
                const float hprev = ybuf[tok * H + u];       // slot t == h_{t-1}
                const float dpn = dh * (1.0f - z) * (1.0f - n * n);
                const float dpz = dh * (hprev - n) * z * (1.0f - z);
                const float dpr = dpn * hn * r * (1.0f - r);
                g[u] = dpr; g[H + u] = dpz; g[2 * H + u] = dpn;          // dgi
                const float dghn = dpn * r;
                cbuf[ci] = dghn;                                          // n-gate part of dgh
                dg[u] = dpr; dg[H + u] = dpz; dg[2 * H + u] = dghn;
                dh_s[b * H + u] = dh * z;                                 // direct path; matvec adds on top
            } else {
                const float ig = g[u], fg = g[H + u], gg = g[2 * H + u], og = g[3 * H + u];
                const float c = cbuf[((size_t)(t + 1) * B + b0 + b) * H + u];
                const float cprev = cbuf[tok * H + u];
                const float tc = dc_tanh(c);
                const float dc = dc_s[b * H + u] + dh * og * (1.0f - tc * tc);
                const float dpi = dc * gg * ig * (1.0f - ig);
                const float dpf = dc * cprev * fg * (1.0f - fg);
                const float dpg = dc * ig * (1.0f - gg * gg);
                const float dpo = dh * tc * og * (1.0f - og);
                g[u] = dpi; g[H + u] = dpf; g[2 * H + u] = dpg; g[3 * H + u] = dpo;
                dg[u] = dpi; dg[H + u] = dpf; dg[2 * H + u] = dpg; dg[3 * H + u] = dpo;
                dc_s[b * H + u] = dc * fg;
                dh_s[b * H + u] = 0.f;
            }
        }
        __syncthreads();
        // dh_{t-1}[b][k] += sum_j dgh[b][j] * W[j][k]
        for (int item = threadIdx.x; item < H * nslice; item += kThreads) {
            const int k = item % H, sl = item / H;
            const int jlo = (int)((long long)GH * sl / nslice), jhi = (int)((long long)GH * (sl + 1) / nslice);
            float acc[kBT];
#pragma unroll
            for (int b = 0; b < kBT; ++b) acc[b] = 0.f;
            for (int j = jlo; j < jhi; ++j) {
                const float wv = __ldg(w + (size_t)j * H + k);
#pragma unroll
                for (int b = 0; b < kBT; ++b) acc[b] = fmaf(dg_s[b * GH + j], wv, acc[b]);
            }
#pragma unroll
            for (int b = 0; b < kBT; ++b) atomicAdd(&dh_s[b * H + k], acc[b]);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < nb * H; i += kThreads) {
        const int b = i / H, u = i % H;
        if (dh0) dh0[(size_t)(b0 + b) * H + u] = dh_s[b * H + u];
        if (dc0 && G == 4) dc0[(size_t)(b0 + b) * H + u] = dc_s[b * H + u];
    }
}

}  // namespace dc_rnn
