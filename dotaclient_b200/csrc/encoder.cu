// Memory-bound pieces of the unit encoder and the target-unit attention head (policy.py:99-136,144-153).
//
// The 128x128 unit-embedding GEMMs run on the tensor cores (gemm_tf32x3.cu); everything around them is
// bandwidth work on ~20 KB/token of activations and is written here so that each tensor crosses HBM once:
//
//   env_fwd / env_bwd   relu(env W_e^T + b_e), 3 -> 128, written into / read from columns [0,128) of the concatenated
//                       [N, 896] pre-rnn input row (no torch.cat), policy.py:55,97
//   unit_basic_fwd      relu(units W_b^T + b_b)            [R,12] -> [R,128]     (K = 12: not a tensor-core shape)
//   unit_basic_bwd      dW_b, db_b from d_basic, the ReLU mask and the raw units  (the inputs need no gradient); the three
//                       input streams come through a 3-stage ring of 1-D TMA bulk copies
//   unit_max_fwd        max over the units of a group + argmax (uint8), written straight into the concatenated
//                       pre-rnn input row, policy.py:102-136
//   unit_grad_assemble  d(unit embedding) in ONE dense pass: rank-1 target-unit part + max-pool routing to the arg-max unit
//                       (the training path; unit_max_bwd is the in-place scatter used when an explicit gradient arrives)
//   target_unit_fwd     logits[n,u] = <attention[n,:], unit_embedding[n,u,:]>      (policy.py:152-153)
//   target_unit_bwd     d_attention (and, outside the training path, the rank-1 d(unit embedding))
//
// Thread mapping everywhere: one warp per row of 128 channels, lane l owns channels 4l..4l+3 -> every global access
// is a fully coalesced 512-byte row segment (16 bytes per lane).
#include "dc_common.cuh"

namespace {

constexpr int kC = 128;          // embedding width (policy.py:56-63)
constexpr int kIn = 12;          // unit feature count (policy.py:56)
constexpr int kWarps = 8;
constexpr int kThreadsE = kWarps * 32;
constexpr int kMaxUnits = 40;

// ---- relu(units W_b^T + b_b) -------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsE) unit_basic_fwd_kernel(const float *__restrict__ units,
                                                                   const float *__restrict__ w_b,
                                                                   const float *__restrict__ b_b,
                                                                   float *__restrict__ basic, int64_t R) {
    __shared__ float s_u[kWarps][32][kIn];                        // 32 rows of raw features per warp per iteration
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // weights as channel PAIRS for the packed fp32x2 FMA (FFMA2): w2[p][k] = (W_b[4l+2p][k], W_b[4l+2p+1][k])
    float2 w2[2][kIn], b2[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        b2[p] = make_float2(b_b[lane * 4 + 2 * p], b_b[lane * 4 + 2 * p + 1]);
#pragma unroll
        for (int k = 0; k < kIn; ++k) w2[p][k] = make_float2(w_b[(lane * 4 + 2 * p) * kIn + k], w_b[(lane * 4 + 2 * p + 1) * kIn + k]);
    }
    const int64_t rows_per_iter = (int64_t)gridDim.x * kWarps * 32;
    for (int64_t base = ((int64_t)blockIdx.x * kWarps + warp) * 32; base < R; base += rows_per_iter) {
        const int nrows = (int)min((int64_t)32, R - base);
        // stage 32 x 12 contiguous floats (coalesced), then every lane reads each row as a broadcast
        float *su = &s_u[warp][0][0];
        for (int i = lane; i < nrows * kIn; i += 32) su[i] = units[base * kIn + i];
        __syncwarp();
        for (int r = 0; r < nrows; ++r) {
            float2 a0 = b2[0], a1 = b2[1];
#pragma unroll
            for (int k = 0; k < kIn; ++k) {
                const float u = s_u[warp][r][k];
                const float2 uu = make_float2(u, u);
                a0 = __ffma2_rn(uu, w2[0][k], a0);
                a1 = __ffma2_rn(uu, w2[1][k], a1);
            }
            *reinterpret_cast<float4 *>(basic + (base + r) * kC + lane * 4) =
                make_float4(fmaxf(a0.x, 0.f), fmaxf(a0.y, 0.f), fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f));
        }
        __syncwarp();
    }
}

// ---- dW_b, db_b ----------------------------------------------------------------------------------
// partial[block][128][13]: 12 weight-gradient columns + the bias gradient, reduced by a second tiny kernel.
__global__ void __launch_bounds__(kThreadsE, 2) unit_basic_bwd_kernel(const float *__restrict__ d_basic,
                                                                   const float *__restrict__ basic,
                                                                   const float *__restrict__ units, int64_t R,
                                                                   float *__restrict__ partial) {
    __shared__ float s_u[kWarps][32][kIn];
    __shared__ float s_red[kC][kIn + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 acc2[4][kIn / 2];       // (dW[c][2kk], dW[c][2kk+1]) pairs for FFMA2
    float accb[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        accb[c] = 0.f;
#pragma unroll
        for (int kk = 0; kk < kIn / 2; ++kk) acc2[c][kk] = make_float2(0.f, 0.f);
    }
    const int64_t rows_per_iter = (int64_t)gridDim.x * kWarps * 32;
    for (int64_t base = ((int64_t)blockIdx.x * kWarps + warp) * 32; base < R; base += rows_per_iter) {
        const int nrows = (int)min((int64_t)32, R - base);
        float *su = &s_u[warp][0][0];
        for (int i = lane; i < nrows * kIn; i += 32) su[i] = units[base * kIn + i];
        __syncwarp();
        for (int r0 = 0; r0 < nrows; r0 += 4) {      // 4 rows per trip: 8 independent 16-byte loads in flight per lane
            float4 g4[4], y4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = min(r0 + j, nrows - 1);
                g4[j] = __ldg(reinterpret_cast<const float4 *>(d_basic + (base + r) * kC) + lane);
                y4[j] = __ldg(reinterpret_cast<const float4 *>(basic + (base + r) * kC) + lane);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (r0 + j >= nrows) break;
                const float g[4] = {y4[j].x > 0.f ? g4[j].x : 0.f, y4[j].y > 0.f ? g4[j].y : 0.f,
                                    y4[j].z > 0.f ? g4[j].z : 0.f, y4[j].w > 0.f ? g4[j].w : 0.f};
#pragma unroll
                for (int kk = 0; kk < kIn / 2; ++kk) {
                    const float2 u2 = *reinterpret_cast<const float2 *>(&s_u[warp][r0 + j][2 * kk]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc2[c][kk] = __ffma2_rn(make_float2(g[c], g[c]), u2, acc2[c][kk]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) accb[c] += g[c];
            }
        }
        __syncwarp();
    }
    // fold the 8 warps into one [128][13] tile in a fixed order (deterministic), then one partial per block
    for (int w = 0; w < kWarps; ++w) {
        if (warp == w) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int k = 0; k <= kIn; ++k) {
                    float *dst = &s_red[lane * 4 + c][k];
                    const float mine = k == kIn ? accb[c] : ((k & 1) ? acc2[c][k >> 1].y : acc2[c][k >> 1].x);
                    *dst = (w == 0 ? 0.f : *dst) + mine;
                }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < kC * (kIn + 1); i += kThreadsE)
        partial[(size_t)blockIdx.x * kC * (kIn + 1) + i] = (&s_red[0][0])[i];
}

// TMA-staged variant: the three input streams (d_basic, basic, units) are contiguous row ranges, so a whole 64-row tile
// is three 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) into a 3-stage shared-memory ring.  The ring, not the
// register file, holds the bytes in flight (up to 134 KB per SM): the register-staged kernel above cannot keep more than
// ~64 KB in flight next to its 52 accumulators and stalls on HBM latency (1.7 ms for 5.4 GB = 3.2 TB/s).
constexpr int kBwdTile = 64;                                   // rows per stage
constexpr int kBwdStages = 3;
constexpr size_t kBwdStageBytes = (size_t)kBwdTile * (2 * kC + kIn) * 4;   // 68,608 B
constexpr size_t kBwdSmem = kBwdStages * kBwdStageBytes + 64;

__device__ __forceinline__ uint32_t e_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void e_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(e_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void e_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(e_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void e_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(e_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void e_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(e_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void e_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(e_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(e_smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kThreadsE, 1) unit_basic_bwd_tma_kernel(const float *__restrict__ d_basic,
                                                                          const float *__restrict__ basic,
                                                                          const float *__restrict__ units, int64_t R,
                                                                          float *__restrict__ partial) {
    extern __shared__ __align__(128) unsigned char e_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(e_smem + kBwdStages * kBwdStageBytes);
    uint64_t *empty = full + kBwdStages;
    __shared__ float s_red[kC][kIn + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_tiles = (R + kBwdTile - 1) / kBwdTile;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kBwdStages; ++s) { e_mbar_init(&full[s], 1); e_mbar_init(&empty[s], kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t i) {                              // tile number i of this CTA -> stage i % kBwdStages
        const int64_t row0 = (blockIdx.x + i * gridDim.x) * kBwdTile;
        const uint32_t rows = (uint32_t)min((int64_t)kBwdTile, R - row0);
        unsigned char *st = e_smem + (i % kBwdStages) * kBwdStageBytes;
        uint64_t *bar = &full[i % kBwdStages];
        e_mbar_expect_tx(bar, rows * (2 * kC + kIn) * 4);
        e_bulk_g2s(st, d_basic + row0 * kC, rows * kC * 4, bar);
        e_bulk_g2s(st + kBwdTile * kC * 4, basic + row0 * kC, rows * kC * 4, bar);
        e_bulk_g2s(st + 2 * kBwdTile * kC * 4, units + row0 * kIn, rows * kIn * 4, bar);
    };
    if (threadIdx.x == 0)
        for (int64_t i = 0; i < kBwdStages && i < my_tiles; ++i) issue(i);

    float2 acc2[4][kIn / 2];
    float accb[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        accb[c] = 0.f;
#pragma unroll
        for (int kk = 0; kk < kIn / 2; ++kk) acc2[c][kk] = make_float2(0.f, 0.f);
    }
    for (int64_t i = 0; i < my_tiles; ++i) {
        const int s = (int)(i % kBwdStages);
        const uint32_t ph = (uint32_t)((i / kBwdStages) & 1);
        e_mbar_wait(&full[s], ph);
        const int64_t row0 = (blockIdx.x + i * gridDim.x) * kBwdTile;
        const int rows = (int)min((int64_t)kBwdTile, R - row0);
        const float *sg = reinterpret_cast<const float *>(e_smem + s * kBwdStageBytes);
        const float *sy = sg + kBwdTile * kC;
        const float *su = sy + kBwdTile * kC;
        for (int r = warp; r < rows; r += kWarps) {           // 8 rows per warp per tile
            const float4 g4 = *reinterpret_cast<const float4 *>(sg + r * kC + lane * 4);
            const float4 y4 = *reinterpret_cast<const float4 *>(sy + r * kC + lane * 4);
            const float g[4] = {y4.x > 0.f ? g4.x : 0.f, y4.y > 0.f ? g4.y : 0.f, y4.z > 0.f ? g4.z : 0.f, y4.w > 0.f ? g4.w : 0.f};
#pragma unroll
            for (int kk = 0; kk < kIn / 2; ++kk) {
                const float2 u2 = *reinterpret_cast<const float2 *>(su + r * kIn + 2 * kk);
#pragma unroll
                for (int c = 0; c < 4; ++c) acc2[c][kk] = __ffma2_rn(make_float2(g[c], g[c]), u2, acc2[c][kk]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) accb[c] += g[c];
        }
        __syncwarp();
        if (lane == 0) e_mbar_arrive(&empty[s]);              // this warp is done with the stage
        if (threadIdx.x == 0 && i + kBwdStages < my_tiles) {   // refill once all 8 warps have released it
            e_mbar_wait(&empty[s], ph);
            issue(i + kBwdStages);
        }
    }
    for (int w = 0; w < kWarps; ++w) {
        if (warp == w) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int k = 0; k <= kIn; ++k) {
                    float *dst = &s_red[lane * 4 + c][k];
                    const float mine = k == kIn ? accb[c] : ((k & 1) ? acc2[c][k >> 1].y : acc2[c][k >> 1].x);
                    *dst = (w == 0 ? 0.f : *dst) + mine;
                }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < kC * (kIn + 1); i += kThreadsE)
        partial[(size_t)blockIdx.x * kC * (kIn + 1) + i] = (&s_red[0][0])[i];
}

__global__ void unit_basic_bwd_reduce_kernel(const float *__restrict__ partial, int nblocks, float *__restrict__ dw_b,
                                             float *__restrict__ db_b, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // over 128 x 13
    if (i >= kC * (kIn + 1)) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * kC * (kIn + 1) + i];
    const int o = i / (kIn + 1), k = i % (kIn + 1);
    float *dst = k < kIn ? dw_b + o * kIn + k : db_b + o;
    *dst = accumulate ? *dst + s : s;
}

// ---- environment encoder: relu(env W_e^T + b_e), 3 -> 128 (policy.py:55,97) ---------------------------------------
// Written straight into columns [0,128) of the concatenated pre-rnn input row (row pitch ld), next to the group maxima
// that unit_max_fwd puts in columns [128,896): the reference's torch.cat (policy.py:129-136) never materialises.
constexpr int kEnvIn = 3;
__global__ void __launch_bounds__(kThreadsE) env_fwd_kernel(const float *__restrict__ env, const float *__restrict__ w_e,
                                                            const float *__restrict__ b_e, float *__restrict__ out, int ld,
                                                            int64_t N) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float w[4][kEnvIn], b[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        b[c] = b_e[lane * 4 + c];
#pragma unroll
        for (int k = 0; k < kEnvIn; ++k) w[c][k] = w_e[(lane * 4 + c) * kEnvIn + k];
    }
    const int64_t rows_per_iter = (int64_t)gridDim.x * kWarps * 32;
    for (int64_t base = ((int64_t)blockIdx.x * kWarps + warp) * 32; base < N; base += rows_per_iter) {
        const int nrows = (int)min((int64_t)32, N - base);
        float e[kEnvIn];                                           // lane r holds row base+r, rows are broadcast by shuffle
#pragma unroll
        for (int k = 0; k < kEnvIn; ++k) e[k] = lane < nrows ? env[(base + lane) * kEnvIn + k] : 0.f;
        for (int r = 0; r < nrows; ++r) {
            float a[4] = {b[0], b[1], b[2], b[3]};
#pragma unroll
            for (int k = 0; k < kEnvIn; ++k) {
                const float u = __shfl_sync(0xffffffffu, e[k], r);
#pragma unroll
                for (int c = 0; c < 4; ++c) a[c] = fmaf(u, w[c][k], a[c]);
            }
            *reinterpret_cast<float4 *>(out + (base + r) * ld + lane * 4) =
                make_float4(fmaxf(a[0], 0.f), fmaxf(a[1], 0.f), fmaxf(a[2], 0.f), fmaxf(a[3], 0.f));
        }
    }
}

// dW_e, db_e from d_out, the ReLU mask (out > 0) and env: partial[block][128][4] (3 weight columns + bias), then reduced.
__global__ void __launch_bounds__(kThreadsE) env_bwd_kernel(const float *__restrict__ d_out, const float *__restrict__ out, int ld,
                                                            const float *__restrict__ env, int64_t N,
                                                            float *__restrict__ partial) {
    __shared__ float s_red[kC][kEnvIn + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[4][kEnvIn + 1];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k <= kEnvIn; ++k) acc[c][k] = 0.f;
    const int64_t rows_per_iter = (int64_t)gridDim.x * kWarps * 32;
    for (int64_t base = ((int64_t)blockIdx.x * kWarps + warp) * 32; base < N; base += rows_per_iter) {
        const int nrows = (int)min((int64_t)32, N - base);
        float e[kEnvIn];
#pragma unroll
        for (int k = 0; k < kEnvIn; ++k) e[k] = lane < nrows ? env[(base + lane) * kEnvIn + k] : 0.f;
        for (int r0 = 0; r0 < nrows; r0 += 8) {                    // 8 rows per trip: 16 independent 16-byte loads per lane
            float4 g4[8], y4[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = min(r0 + j, nrows - 1);
                g4[j] = __ldg(reinterpret_cast<const float4 *>(d_out + (base + r) * ld) + lane);
                y4[j] = __ldg(reinterpret_cast<const float4 *>(out + (base + r) * ld) + lane);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool live = r0 + j < nrows;                  // uniform across the warp
                const float g[4] = {live && y4[j].x > 0.f ? g4[j].x : 0.f, live && y4[j].y > 0.f ? g4[j].y : 0.f,
                                    live && y4[j].z > 0.f ? g4[j].z : 0.f, live && y4[j].w > 0.f ? g4[j].w : 0.f};
#pragma unroll
                for (int k = 0; k < kEnvIn; ++k) {
                    const float u = __shfl_sync(0xffffffffu, e[k], min(r0 + j, 31));
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[c][k] = fmaf(g[c], u, acc[c][k]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c][kEnvIn] += g[c];
            }
        }
    }
    for (int w = 0; w < kWarps; ++w) {                             // fixed fold order: deterministic
        if (warp == w) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int k = 0; k <= kEnvIn; ++k) {
                    float *dst = &s_red[lane * 4 + c][k];
                    *dst = (w == 0 ? 0.f : *dst) + acc[c][k];
                }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < kC * (kEnvIn + 1); i += kThreadsE)
        partial[(size_t)blockIdx.x * kC * (kEnvIn + 1) + i] = (&s_red[0][0])[i];
}

// 32 outputs x 8 block-groups per CTA: group y sums blocks y, y+8, ...; the groups are then added in order.
__global__ void __launch_bounds__(256) env_bwd_reduce_kernel(const float *__restrict__ partial, int nblocks, float *__restrict__ dw_e,
                                                             float *__restrict__ db_e) {
    __shared__ float sh[8][32];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + x;                             // over 128 x 4
    float s = 0.f;
    if (i < kC * (kEnvIn + 1))
        for (int b = y; b < nblocks; b += 8) s += partial[(size_t)b * kC * (kEnvIn + 1) + i];
    sh[y][x] = s;
    __syncthreads();
    if (y == 0 && i < kC * (kEnvIn + 1)) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += sh[r][x];
        const int o = i / (kEnvIn + 1), k = i % (kEnvIn + 1);
        if (k < kEnvIn) dw_e[o * kEnvIn + k] = t; else db_e[o] = t;
    }
}

// ---- max over the units of one group --------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsE) unit_max_fwd_kernel(const float *__restrict__ emb, int64_t tok_stride,
                                                                 int units, float *__restrict__ xmax, int ld_x,
                                                                 float *__restrict__ xmax2, uint8_t *__restrict__ argmax,
                                                                 int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const float4 *row = reinterpret_cast<const float4 *>(emb + n * tok_stride) + lane;
    float4 best = __ldg(row);
    uchar4 idx = make_uchar4(0, 0, 0, 0);
    for (int u = 1; u < units; ++u) {
        const float4 v = __ldg(row + u * (kC / 4));
        if (v.x > best.x) { best.x = v.x; idx.x = u; }            // strict >: the first maximum wins, like torch.max
        if (v.y > best.y) { best.y = v.y; idx.y = u; }
        if (v.z > best.z) { best.z = v.z; idx.z = u; }
        if (v.w > best.w) { best.w = v.w; idx.w = u; }
    }
    *reinterpret_cast<float4 *>(xmax + n * ld_x + lane * 4) = best;
    if (xmax2) *reinterpret_cast<float4 *>(xmax2 + n * ld_x + lane * 4) = best;     // policy.py:127: eth slot <- enh max
    *reinterpret_cast<uchar4 *>(argmax + n * kC + lane * 4) = idx;
}

__global__ void __launch_bounds__(kThreadsE) unit_max_bwd_kernel(float *__restrict__ d_emb, int64_t tok_stride,
                                                                 const float *__restrict__ d_xmax,
                                                                 const float *__restrict__ d_xmax2, int ld_dx,
                                                                 const uint8_t *__restrict__ argmax, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    float4 g = __ldg(reinterpret_cast<const float4 *>(d_xmax + n * ld_dx) + lane);
    if (d_xmax2) {
        const float4 g2 = __ldg(reinterpret_cast<const float4 *>(d_xmax2 + n * ld_dx) + lane);
        g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
    }
    const uchar4 idx = *reinterpret_cast<const uchar4 *>(argmax + n * kC + lane * 4);
    float *base = d_emb + n * tok_stride + lane * 4;
    base[idx.x * kC + 0] += g.x;
    base[idx.y * kC + 1] += g.y;
    base[idx.z * kC + 2] += g.z;
    base[idx.w * kC + 3] += g.w;
}

// ---- target-unit attention head ---------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsE) target_unit_fwd_kernel(const float *__restrict__ att,
                                                                    const float *__restrict__ ue,
                                                                    float *__restrict__ logits, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const float4 a = __ldg(reinterpret_cast<const float4 *>(att + n * kC) + lane);
    const float4 *row = reinterpret_cast<const float4 *>(ue + n * kMaxUnits * kC) + lane;
    float mine = 0.f;                                             // lane u keeps logit u (u < 32), lanes 0..7 also u+32
    float mine_hi = 0.f;
#pragma unroll 8
    for (int u = 0; u < kMaxUnits; ++u) {
        const float4 v = __ldg(row + u * (kC / 4));
        float d = v.x * a.x + v.y * a.y + v.z * a.z + v.w * a.w;
        d = dc_warp_sum(d);
        if (u < 32) { if (lane == u) mine = d; } else { if (lane == u - 32) mine_hi = d; }
    }
    logits[n * kMaxUnits + lane] = mine;
    if (lane < kMaxUnits - 32) logits[n * kMaxUnits + 32 + lane] = mine_hi;
}

__global__ void __launch_bounds__(kThreadsE) target_unit_bwd_kernel(const float *__restrict__ dlogits,
                                                                    const float *__restrict__ att,
                                                                    const float *__restrict__ ue,
                                                                    float *__restrict__ d_att, float *__restrict__ d_ue,
                                                                    int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const float g_lo = dlogits[n * kMaxUnits + lane];
    const float g_hi = lane < kMaxUnits - 32 ? dlogits[n * kMaxUnits + 32 + lane] : 0.f;
    const bool any = __any_sync(0xffffffffu, g_lo != 0.f || g_hi != 0.f);
    float4 *drow = d_ue ? reinterpret_cast<float4 *>(d_ue + n * kMaxUnits * kC) + lane : nullptr;   // NULL: d_att only
    float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!any) {                                                   // head unused on this token: exact zeros, no reads
        if (drow) {
#pragma unroll 8
            for (int u = 0; u < kMaxUnits; ++u) drow[u * (kC / 4)] = da;
        }
    } else {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(att + n * kC) + lane);
        const float4 *row = reinterpret_cast<const float4 *>(ue + n * kMaxUnits * kC) + lane;
#pragma unroll 8
        for (int u = 0; u < kMaxUnits; ++u) {
            const float g = __shfl_sync(0xffffffffu, u < 32 ? g_lo : g_hi, u & 31);
            const float4 v = __ldg(row + u * (kC / 4));
            da.x = fmaf(g, v.x, da.x); da.y = fmaf(g, v.y, da.y); da.z = fmaf(g, v.z, da.z); da.w = fmaf(g, v.w, da.w);
            if (drow) drow[u * (kC / 4)] = make_float4(g * a.x, g * a.y, g * a.z, g * a.w);
        }
    }
    *(reinterpret_cast<float4 *>(d_att + n * kC) + lane) = da;
}

// ---- d(unit embedding) assembled in ONE dense pass ---------------------------------------------------
// d_ue[n,u,c] = dlogits[n,u] * att[n,c]                      (target-unit head, rank 1; only where the head was used)
//             + (u == argmax_g[n,c]) ? d_xmax[n,g,c] : 0     (max-pool of group g routes to its arg-max unit)
// The scattered in-place version (unit_max_bwd) costs a 32-byte sector read+write per 4-byte update -- as much
// traffic as a dense pass (ncu: 935 MB read / 394 MB written for the 16-unit group) on top of target_unit_bwd's own
// dense write; fusing the two writes [N,40,128] once.
__global__ void __launch_bounds__(kThreadsE) unit_grad_assemble_kernel(const float *__restrict__ dlogits,
                                                                       const float *__restrict__ att,
                                                                       const float *__restrict__ d_xm, int ld_dx,
                                                                       const uint8_t *__restrict__ argmax,
                                                                       float *__restrict__ d_ue, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    float g_lo = 0.f, g_hi = 0.f;
    if (dlogits) {
        g_lo = dlogits[n * kMaxUnits + lane];
        g_hi = lane < kMaxUnits - 32 ? dlogits[n * kMaxUnits + 32 + lane] : 0.f;
    }
    const bool any = __any_sync(0xffffffffu, g_lo != 0.f || g_hi != 0.f);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (any) a = __ldg(reinterpret_cast<const float4 *>(att + n * kC) + lane);
    float4 *drow = reinterpret_cast<float4 *>(d_ue + n * kMaxUnits * kC) + lane;
    constexpr int units_of[6] = {1, 5, 16, 16, 1, 1};
    int u = 0;
#pragma unroll
    for (int grp = 0; grp < 6; ++grp) {
        float4 dm = make_float4(0.f, 0.f, 0.f, 0.f);
        uchar4 idx = make_uchar4(255, 255, 255, 255);
        if (d_xm && grp < 5) {                                   // group 5 (enemy towers) has no max path (policy.py:127)
            dm = __ldg(reinterpret_cast<const float4 *>(d_xm + n * ld_dx + grp * kC) + lane);
            if (grp == 3) {                                       // ... its slot was fed from the enemy non-hero maximum
                const float4 d2 = __ldg(reinterpret_cast<const float4 *>(d_xm + n * ld_dx + 5 * kC) + lane);
                dm.x += d2.x; dm.y += d2.y; dm.z += d2.z; dm.w += d2.w;
            }
            idx = *reinterpret_cast<const uchar4 *>(argmax + ((int64_t)grp * N + n) * kC + lane * 4);
        }
#pragma unroll
        for (int j = 0; j < units_of[grp]; ++j, ++u) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (any) {
                const float g = __shfl_sync(0xffffffffu, u < 32 ? g_lo : g_hi, u & 31);
                o = make_float4(g * a.x, g * a.y, g * a.z, g * a.w);
            }
            if (idx.x == j) o.x += dm.x;
            if (idx.y == j) o.y += dm.y;
            if (idx.z == j) o.z += dm.z;
            if (idx.w == j) o.w += dm.w;
            drow[u * (kC / 4)] = o;
        }
    }
}

int grid_rows(int64_t rows_per_block_iter_unused) { (void)rows_per_block_iter_unused; return 4 * dc_sm_count(); }


// ---- target-unit head without the unit embedding ------------------------------------------------------------------------
// logits[n,u] = <att[n], W_g basic[n,u] + b_g> = <att[n] W_g, basic[n,u]> + <att[n], b_g> (policy.py:144-153; algebra pinned in
// tests/test_oracle.py): with q[n, g*128 + j] = (att W_g)[n, j] and q[n, 768 + g] = <att[n], b_g> from ONE small GEMM over tokens,
// the head reads the stored `basic` rows and the [N, 40, 128] embedding is never materialised.
struct BasicPtrs { const float *p[6]; };
__constant__ int kGroupUnits[6] = {1, 5, 16, 16, 1, 1};
__constant__ int kGroupOffset[6] = {0, 1, 6, 22, 38, 39};

__global__ void __launch_bounds__(kThreadsE) target_unit_q_fwd_kernel(const float *__restrict__ q, int ld_q, BasicPtrs basics,
                                                                      float *__restrict__ logits, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const float *qrow = q + n * ld_q;
    float mine = 0.f, mine_hi = 0.f;                              // lane u keeps logit u (u < 32), lanes 0..7 also u+32
#pragma unroll
    for (int g = 0; g < 6; ++g) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(qrow + g * kC) + lane);
        const float c = __ldg(qrow + 6 * kC + g);
        const int nu = kGroupUnits[g], off = kGroupOffset[g];
        const float4 *row = reinterpret_cast<const float4 *>(basics.p[g] + n * nu * kC) + lane;
        for (int u = 0; u < nu; ++u) {
            const float4 v = __ldg(row + u * (kC / 4));
            float d = v.x * a.x + v.y * a.y + v.z * a.z + v.w * a.w;
            d = dc_warp_sum(d) + c;
            const int o = off + u;
            if (o < 32) { if (lane == o) mine = d; } else { if (lane == o - 32) mine_hi = d; }
        }
    }
    logits[n * kMaxUnits + lane] = mine;
    if (lane < kMaxUnits - 32) logits[n * kMaxUnits + 32 + lane] = mine_hi;
}

// s[n, g*128 + j] = sum_u dlogits[n, off_g + u] basic_g[n,u,j],  s[n, 768 + g] = sum_u dlogits[n, off_g + u]  (zeros elsewhere):
// d_att = s [W_0 | ... | W_5 | b_0..b_5]^T is then one GEMM over tokens.  Tokens that did not use the head write zeros, read nothing.
__global__ void __launch_bounds__(kThreadsE) target_unit_q_bwd_kernel(const float *__restrict__ dlogits, BasicPtrs basics,
                                                                      float *__restrict__ s, int ld_s, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const float g_lo = dlogits[n * kMaxUnits + lane];
    const float g_hi = lane < kMaxUnits - 32 ? dlogits[n * kMaxUnits + 32 + lane] : 0.f;
    const bool any = __any_sync(0xffffffffu, g_lo != 0.f || g_hi != 0.f);
    float4 *srow = reinterpret_cast<float4 *>(s + n * ld_s) + lane;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!any) {
#pragma unroll
        for (int g = 0; g < 7; ++g) srow[g * (kC / 4)] = zero;
        return;
    }
    float sig[6];
#pragma unroll
    for (int g = 0; g < 6; ++g) {
        const int nu = kGroupUnits[g], off = kGroupOffset[g];
        const float4 *row = reinterpret_cast<const float4 *>(basics.p[g] + n * nu * kC) + lane;
        float4 acc = zero;
        float sg = 0.f;
        for (int u = 0; u < nu; ++u) {
            const int o = off + u;
            const float gv = __shfl_sync(0xffffffffu, o < 32 ? g_lo : g_hi, o & 31);
            const float4 v = __ldg(row + u * (kC / 4));
            acc.x = fmaf(gv, v.x, acc.x); acc.y = fmaf(gv, v.y, acc.y); acc.z = fmaf(gv, v.z, acc.z); acc.w = fmaf(gv, v.w, acc.w);
            sg += gv;
        }
        srow[g * (kC / 4)] = acc;
        sig[g] = sg;
    }
    float4 tail = zero;                                            // columns 768..895: the six sums, then zeros
    if (lane == 0) tail = make_float4(sig[0], sig[1], sig[2], sig[3]);
    if (lane == 1) tail = make_float4(sig[4], sig[5], 0.f, 0.f);
    srow[6 * (kC / 4)] = tail;
}

}  // namespace

extern "C" int dc_unit_basic_fwd(const float *units, const float *w_b, const float *b_b, float *basic, int64_t R,
                                 dc_stream_t stream) {
    DC_REQUIRE(units && w_b && b_b && basic && R > 0, DC_EINVAL, "dc_unit_basic_fwd: bad arguments");
    DC_REQUIRE(((uintptr_t)basic & 15) == 0, DC_EINVAL, "dc_unit_basic_fwd: output must be 16-byte aligned");
    unit_basic_fwd_kernel<<<grid_rows(0), kThreadsE, 0, dc_cu_stream(stream)>>>(units, w_b, b_b, basic, R);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" size_t dc_unit_basic_bwd_workspace_bytes(void) { return (size_t)4 * 1024 * kC * (kIn + 1) * sizeof(float); }

extern "C" int dc_unit_basic_bwd(const float *d_basic, const float *basic, const float *units, float *dw_b, float *db_b,
                                 int64_t R, int accumulate, void *workspace, dc_stream_t stream) {
    DC_REQUIRE(d_basic && basic && units && dw_b && db_b && workspace && R > 0, DC_EINVAL, "dc_unit_basic_bwd: bad arguments");
    DC_REQUIRE((((uintptr_t)d_basic | (uintptr_t)basic) & 15) == 0, DC_EINVAL, "dc_unit_basic_bwd: inputs must be 16-byte aligned");
    cudaStream_t st = dc_cu_stream(stream);
    float *partial = reinterpret_cast<float *>(workspace);
    int blocks;
    if ((((uintptr_t)units) & 15) == 0) {                      // bulk copies need 16-byte aligned sources
        blocks = dc_sm_count();
        // per-device attribute: set on every call (a process-wide "done" flag breaks the second GPU of a process)
        DC_CUDA(cudaFuncSetAttribute(unit_basic_bwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
        unit_basic_bwd_tma_kernel<<<blocks, kThreadsE, kBwdSmem, st>>>(d_basic, basic, units, R, partial);
    } else {
        blocks = 2 * dc_sm_count();
        unit_basic_bwd_kernel<<<blocks, kThreadsE, 0, st>>>(d_basic, basic, units, R, partial);
    }
    DC_LAUNCH_OK();
    unit_basic_bwd_reduce_kernel<<<(kC * (kIn + 1) + 255) / 256, 256, 0, st>>>(partial, blocks, dw_b, db_b, accumulate);
    DC_LAUNCH_OK();
    return DC_OK;
}

int dc_unit_basic_reduce(const float *partial, int nblocks, float *dw_b, float *db_b, int accumulate, cudaStream_t st) {
    unit_basic_bwd_reduce_kernel<<<(kC * (kIn + 1) + 255) / 256, 256, 0, st>>>(partial, nblocks, dw_b, db_b, accumulate);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_env_fwd(const float *env, const float *w_e, const float *b_e, float *out, int ld_out, int64_t N,
                          dc_stream_t stream) {
    DC_REQUIRE(env && w_e && b_e && out && N > 0 && ld_out >= kC && ld_out % 4 == 0 && ((uintptr_t)out & 15) == 0, DC_EINVAL,
               "dc_env_fwd: bad arguments");
    env_fwd_kernel<<<grid_rows(0), kThreadsE, 0, dc_cu_stream(stream)>>>(env, w_e, b_e, out, ld_out, N);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" size_t dc_env_bwd_workspace_bytes(void) { return (size_t)1024 * kC * (kEnvIn + 1) * sizeof(float); }

extern "C" int dc_env_bwd(const float *d_out, const float *out, int ld, const float *env, float *dw_e, float *db_e, int64_t N,
                          void *workspace, dc_stream_t stream) {
    DC_REQUIRE(d_out && out && env && dw_e && db_e && workspace && N > 0 && ld >= kC && ld % 4 == 0, DC_EINVAL,
               "dc_env_bwd: bad arguments");
    DC_REQUIRE((((uintptr_t)d_out | (uintptr_t)out) & 15) == 0, DC_EINVAL, "dc_env_bwd: inputs must be 16-byte aligned");
    cudaStream_t st = dc_cu_stream(stream);
    float *partial = reinterpret_cast<float *>(workspace);
    const int blocks = 2 * dc_sm_count();                          // <= 1024 (workspace bound)
    env_bwd_kernel<<<blocks, kThreadsE, 0, st>>>(d_out, out, ld, env, N, partial);
    DC_LAUNCH_OK();
    env_bwd_reduce_kernel<<<(kC * (kEnvIn + 1) + 31) / 32, 256, 0, st>>>(partial, blocks, dw_e, db_e);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_unit_max_fwd(const float *emb, int64_t tok_stride, int units, float *xmax, float *xmax_copy, int ld_x,
                               uint8_t *argmax, int64_t N, dc_stream_t stream) {
    DC_REQUIRE(emb && xmax && argmax && N > 0 && units >= 1 && units <= 255 && tok_stride >= (int64_t)units * kC, DC_EINVAL,
               "dc_unit_max_fwd: bad arguments");
    DC_REQUIRE(((uintptr_t)emb & 15) == 0 && ((uintptr_t)xmax & 15) == 0 && ld_x % 4 == 0 && tok_stride % 4 == 0 &&
                   (!xmax_copy || ((uintptr_t)xmax_copy & 15) == 0), DC_EINVAL, "dc_unit_max_fwd: alignment");
    unit_max_fwd_kernel<<<(unsigned)((N + kWarps - 1) / kWarps), kThreadsE, 0, dc_cu_stream(stream)>>>(
        emb, tok_stride, units, xmax, ld_x, xmax_copy, argmax, N);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_unit_max_bwd(float *d_emb, int64_t tok_stride, const float *d_xmax, const float *d_xmax_copy, int ld_dx,
                               const uint8_t *argmax, int64_t N, dc_stream_t stream) {
    DC_REQUIRE(d_emb && d_xmax && argmax && N > 0, DC_EINVAL, "dc_unit_max_bwd: bad arguments");
    DC_REQUIRE(((uintptr_t)d_xmax & 15) == 0 && ld_dx % 4 == 0 && (!d_xmax_copy || ((uintptr_t)d_xmax_copy & 15) == 0), DC_EINVAL,
               "dc_unit_max_bwd: alignment");
    unit_max_bwd_kernel<<<(unsigned)((N + kWarps - 1) / kWarps), kThreadsE, 0, dc_cu_stream(stream)>>>(
        d_emb, tok_stride, d_xmax, d_xmax_copy, ld_dx, argmax, N);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_target_unit_fwd(const float *att, const float *ue, float *logits, int64_t N, dc_stream_t stream) {
    DC_REQUIRE(att && ue && logits && N > 0, DC_EINVAL, "dc_target_unit_fwd: bad arguments");
    DC_REQUIRE((((uintptr_t)att | (uintptr_t)ue) & 15) == 0, DC_EINVAL, "dc_target_unit_fwd: alignment");
    target_unit_fwd_kernel<<<(unsigned)((N + kWarps - 1) / kWarps), kThreadsE, 0, dc_cu_stream(stream)>>>(att, ue, logits, N);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_target_unit_q_fwd(const float *q, int ld_q, const float *const basics[6], float *logits, int64_t N,
                                    dc_stream_t stream) {
    DC_REQUIRE(q && basics && logits && N > 0 && ld_q >= 7 * kC && ld_q % 4 == 0, DC_EINVAL, "dc_target_unit_q_fwd: bad arguments");
    BasicPtrs bp;
    for (int g = 0; g < 6; ++g) {
        DC_REQUIRE(basics[g] && ((uintptr_t)basics[g] & 15) == 0, DC_EINVAL, "dc_target_unit_q_fwd: basic[%d] null / unaligned", g);
        bp.p[g] = basics[g];
    }
    DC_REQUIRE(((uintptr_t)q & 15) == 0, DC_EINVAL, "dc_target_unit_q_fwd: alignment");
    target_unit_q_fwd_kernel<<<(unsigned)((N + kWarps - 1) / kWarps), kThreadsE, 0, dc_cu_stream(stream)>>>(q, ld_q, bp, logits, N);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_target_unit_q_bwd(const float *dlogits, const float *const basics[6], float *s, int ld_s, int64_t N,
                                    dc_stream_t stream) {
    DC_REQUIRE(dlogits && basics && s && N > 0 && ld_s >= 7 * kC && ld_s % 4 == 0, DC_EINVAL, "dc_target_unit_q_bwd: bad arguments");
    BasicPtrs bp;
    for (int g = 0; g < 6; ++g) {
        DC_REQUIRE(basics[g] && ((uintptr_t)basics[g] & 15) == 0, DC_EINVAL, "dc_target_unit_q_bwd: basic[%d] null / unaligned", g);
        bp.p[g] = basics[g];
    }
    DC_REQUIRE(((uintptr_t)s & 15) == 0, DC_EINVAL, "dc_target_unit_q_bwd: alignment");
    target_unit_q_bwd_kernel<<<(unsigned)((N + kWarps - 1) / kWarps), kThreadsE, 0, dc_cu_stream(stream)>>>(dlogits, bp, s, ld_s, N);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_target_unit_bwd(const float *dlogits, const float *att, const float *ue, float *d_att, float *d_ue,
                                  int64_t N, dc_stream_t stream) {
    DC_REQUIRE(dlogits && att && ue && d_att && N > 0, DC_EINVAL, "dc_target_unit_bwd: bad arguments");
    DC_REQUIRE((((uintptr_t)att | (uintptr_t)ue | (uintptr_t)d_att | (uintptr_t)d_ue) & 15) == 0, DC_EINVAL,
               "dc_target_unit_bwd: alignment");
    target_unit_bwd_kernel<<<(unsigned)((N + kWarps - 1) / kWarps), kThreadsE, 0, dc_cu_stream(stream)>>>(dlogits, att, ue,
                                                                                                         d_att, d_ue, N);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_unit_grad_assemble(const float *dlogits, const float *att, const float *d_xmax, int ld_dx,
                                     const uint8_t *argmax, float *d_ue, int64_t N, dc_stream_t stream) {
    DC_REQUIRE(d_ue && N > 0 && (!dlogits || att) && (!d_xmax || argmax), DC_EINVAL, "dc_unit_grad_assemble: bad arguments");
    DC_REQUIRE((((uintptr_t)att | (uintptr_t)d_xmax | (uintptr_t)d_ue) & 15) == 0 && ld_dx % 4 == 0, DC_EINVAL,
               "dc_unit_grad_assemble: alignment");
    unit_grad_assemble_kernel<<<(unsigned)((N + kWarps - 1) / kWarps), kThreadsE, 0, dc_cu_stream(stream)>>>(
        dlogits, att, d_xmax, ld_dx, argmax, d_ue, N);
    DC_LAUNCH_OK();
    return DC_OK;
}
