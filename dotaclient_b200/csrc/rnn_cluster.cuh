// Cluster-resident tensor-core recurrence for H = 256 (GRU and LSTM), forward and backward.
//
// H = 256 is the reference's own width (policy.py:66 nn.GRU(256, 256)); W_hh is 768 KB (GRU) / 1 MB (LSTM) in fp32 and
// cannot live in one SM.  At this width the step IS a dense contraction -- [sequences, 256] x [256, G*256] per step --
// and fp32 FFMA cannot reach even a third of the HBM roofline (512 x 256 x 1024 MACs / step = 3.7 us on all 148 SMs), so
// the mat-vec runs on tcgen05 with the same 3xTF32 split as the other dense layers (fp32-level accuracy, ~1e-6).
//
//   * a CLUSTER of 8 CTAs owns kNB = 32 sequences for all S steps; CTA r owns hidden units [32r, 32r+32): its 4 x 32
//     rows of W_hh (forward) / 4 x 32 columns (backward) stay on chip for the whole launch -- the tf32 hi half in
//     TENSOR MEMORY (256 columns, the A operand of tcgen05.mma [d], [a_tmem], b-desc), the lo half in shared memory
//     (128 KB of K-major SWIZZLE_128B tiles);
//   * forward  D^T[(g,u)][b] = W_hh[(g,u)][:] . h[b][:]      M = 128 rows, N = 32 sequences, K = 256: every CTA needs the
//     whole h of its 32 sequences -- an all-gather.  h_t is an OUTPUT of the layer anyway (ybuf), so each CTA writes its
//     32-unit slice to ybuf, the cluster barrier (release/acquire) publishes it, and every CTA reads the [32 x 256] tile
//     back from L2, splits it hi/lo and stores the B-operand tiles;
//   * backward D^T[k][b] = sum_{j in own 128 gate columns} W_hh[j][k] dg[b][j]   2 x (M = 128), N = 32, K = 128: the B
//     operand (this CTA's own gate gradients) is local; the partial sums over the 8 CTAs are exchanged through a small
//     L2-resident scratch (reduce-scatter, fixed summation order => deterministic);
//   * gate math: thread = (unit, 4 sequences), 128-byte coalesced global accesses; the per-step global inputs are
//     prefetched one step ahead into registers while the MMAs run.
// Per step and CTA: 96 tcgen05.mma (M128 N32 K8), one cluster barrier, one L2 round trip.  ncu (profiles/r2_rnn_cluster_ncu.md):
// a chain of MMAs into ONE accumulator is latency-bound (~100 cycles per dependent MMA vs ~30 of tensor-pipe work), so every
// K-panel accumulates into its own TMEM accumulator (8 independent chains of 12) and the epilogue adds them; ONE thread issuing
// all 96 MMAs is itself a bottleneck (~100 cycles of scalar work per instruction), so eight threads issue one chain each from
// register-resident descriptors; the cluster barrier is split (arrive.release right after the exchanged slice is stored,
// wait.acquire after the remaining stores).
// Algorithmic HBM bytes per token: forward 4*(G+1)*H, backward 8*(G+1)*H.
#pragma once
#include "dc_common.cuh"

namespace dc_rnnc {

constexpr int kH = 256;
constexpr int kCL = 8;                  // CTAs per cluster
constexpr int kNB = 32;                 // sequences per cluster = MMA N
constexpr int kThreads = 512;
constexpr int kPanelA = 128 * 128;      // bytes of one A tile: 128 rows x 32 tf32 (one SWIZZLE_128B row each)
constexpr int kPanelB = kNB * 128;      // bytes of one B tile:  32 rows x 32 tf32

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (same encoding as csrc/gemm_tf32x3.cu): 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// D = f32, A = B = tf32, both K-major, N = 32 (>>3 at bit 17), M = 128 (>>4 at bit 24)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ float tf32_rna(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// All threads of all CTAs of the cluster.  release/acquire at cluster scope: global writes made before the barrier by any
// thread of the cluster are visible to every thread of the cluster after it.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_barrier() { cluster_arrive(); cluster_wait(); }
__device__ __forceinline__ unsigned char *align1024(unsigned char *p) {
    return reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}
// one 16-byte chunk (4 tf32) of row `row`, 16-byte chunk index `c` (0..7) of a K-major SWIZZLE_128B tile
__device__ __forceinline__ int swz(int row, int c) { return row * 128 + ((c ^ (row & 7)) << 4); }

__device__ __forceinline__ void tmem_alloc_512(uint32_t *slot, int warp) {
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_free_512(uint32_t tmem_base, int warp) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- forward ------------------------------------------------------------------------------------------------------------
// TMEM: [0,256) W_hh slice, hi half: lane rho = g*32 + u  <->  row g*H + 32*rank + u, column = k;  [256,512) eight
// accumulators of 32 columns, one per K-panel.
// shared: W lo half (8 K-panels of [128 rows x 32]), h hi / lo (8 K-panels of [32 sequences x 32]), transposition scratch.
struct FwdSmem {
    static constexpr size_t wlo = 0;
    static constexpr size_t hhi = wlo + 8 * kPanelA;
    static constexpr size_t hlo = hhi + 8 * kPanelB;
    static constexpr size_t scratch = hlo + 8 * kPanelB;          // [4 gates][32 sequences][32 units] fp32
    static constexpr size_t bars = scratch + 4 * kNB * 32 * 4;
    static constexpr size_t total = bars + 64 + 1024;             // + alignment slack
};
constexpr int kNP = 2;                                            // (sequence, unit) pairs per thread in the gate phase

template <int G>
__global__ void __launch_bounds__(kThreads, 1) fwd_cluster_kernel(float *gates, const float *__restrict__ w_hh,
                                                                   const float *__restrict__ b_hh, float *ybuf, float *cbuf,
                                                                   int B, int S) {
    constexpr int H = kH, GH = G * kH, NP = kNP;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = align1024(smem_raw);
    unsigned char *wlo = base + FwdSmem::wlo, *hhi = base + FwdSmem::hhi, *hlo = base + FwdSmem::hlo;
    float *scratch = reinterpret_cast<float *>(base + FwdSmem::scratch);
    uint64_t *mma_done = reinterpret_cast<uint64_t *>(base + FwdSmem::bars);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mma_done + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = blockIdx.x % kCL, b0 = (blockIdx.x / kCL) * kNB;

    if (tid == 0) {
        mbar_init(mma_done, 8);                                           // one commit per issuing thread
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tmem_alloc_512(tmem_slot, warp);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_w = tmem_base, tmem_acc = tmem_base + 256;
    // MMA issue: lane 0 of warp p issues the 12 MMAs of K-panel p into accumulator p.  ncu (v2): ONE thread issuing all 96 MMAs
    // of a step needs ~100 cycles of scalar work per instruction (descriptor arithmetic, R2UR moves) for ~30 cycles of
    // tensor-pipe work; eight threads issue concurrently and keep their (step-invariant) descriptors in registers.
    const bool issuer = lane == 0 && warp < 8;
    const uint64_t d_alo = make_desc(smem_u32(wlo) + (warp & 7) * kPanelA), d_bhi = make_desc(smem_u32(hhi) + (warp & 7) * kPanelB),
                   d_blo = make_desc(smem_u32(hlo) + (warp & 7) * kPanelB);
    const uint32_t t_acc = tmem_acc + 32 * (warp & 7), t_ahi = tmem_w + 32 * (warp & 7);

    {   // resident weights, once: warp (q, kq) fills lanes [32q, 32q+32), columns [64*kq, 64*kq+64)
        const int q = warp & 3, kq = warp >> 2, rho = q * 32 + lane;
        const bool valid = q < G;                                          // q == gate index; GRU has no 4th gate: zero rows
        const float *wrow = w_hh + (size_t)(q * H + rank * 32 + lane) * H;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
#pragma unroll 2
        for (int k0 = kq * 64; k0 < kq * 64 + 64; k0 += 8) {
            float w[8], hi[8], lo[8];
            if (valid) {
                const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wrow + k0)), w1 = __ldg(reinterpret_cast<const float4 *>(wrow + k0 + 4));
                w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { hi[e] = tf32_rna(w[e]); lo[e] = w[e] - hi[e]; }
            tmem_st8(tmem_w + lane_addr + k0, hi);
            unsigned char *panel = wlo + (size_t)(k0 >> 5) * kPanelA;
            const int c0 = (k0 & 31) >> 2;
            *reinterpret_cast<float4 *>(panel + swz(rho, c0)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4 *>(panel + swz(rho, c0 + 1)) = make_float4(lo[4], lo[5], lo[6], lo[7]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // gate phase: thread = (unit ul, sequences sb + 16 i)
    const int ul = lane, unit = rank * 32 + ul, sb = warp;
    float bias[G], c_reg[NP], h_reg[NP], cur[NP][G], nxt[NP][G];
    bool live[NP];
#pragma unroll
    for (int g = 0; g < G; ++g) bias[g] = __ldg(b_hh + g * H + unit);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int b = b0 + sb + 16 * i;
        live[i] = b < B;
        c_reg[i] = (G == 4 && live[i]) ? cbuf[(size_t)b * H + unit] : 0.f;
        h_reg[i] = (G == 3 && live[i]) ? ybuf[(size_t)b * H + unit] : 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            cur[i][g] = live[i] ? gates[(size_t)b * GH + g * H + unit] : 0.f;
            nxt[i][g] = 0.f;
        }
    }
    // h-tile loader: thread = (sequence hb, 16-byte chunk hc) of K-panels hp0, hp0+2, hp0+4, hp0+6
    const int hc = tid & 7, hb = (tid >> 3) & 31, hp0 = tid >> 8;
    const bool hlive = b0 + hb < B;
    cluster_barrier();

    for (int t = 0; t < S; ++t) {
        // ---- A: gather h_{t-1} [32 x 256] (ybuf slot t) from L2, split, store the B-operand tiles
        {
            const float *src = ybuf + ((size_t)t * B + b0 + hb) * H + 4 * hc;
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = hlive ? __ldcg(reinterpret_cast<const float4 *>(src + 32 * (hp0 + 2 * j))) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int off = swz(hb, hc);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int p = hp0 + 2 * j;
                float4 hi, lo;
                hi.x = tf32_rna(v[j].x); hi.y = tf32_rna(v[j].y); hi.z = tf32_rna(v[j].z); hi.w = tf32_rna(v[j].w);
                lo.x = v[j].x - hi.x; lo.y = v[j].y - hi.y; lo.z = v[j].z - hi.z; lo.w = v[j].w - hi.w;
                *reinterpret_cast<float4 *>(hhi + p * kPanelB + off) = hi;
                *reinterpret_cast<float4 *>(hlo + p * kPanelB + off) = lo;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- B: 96 MMAs, 12 per issuing thread (K-panel = accumulator = warp index)
        if (issuer) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                umma_ss(t_acc, d_alo + 2 * ks, d_bhi + 2 * ks, kIdesc, ks != 0);                 // small terms first
                umma_ts(t_acc, t_ahi + 8 * ks, d_blo + 2 * ks, kIdesc, 1u);
                umma_ts(t_acc, t_ahi + 8 * ks, d_bhi + 2 * ks, kIdesc, 1u);
            }
            umma_commit(mma_done);
        }
        __syncwarp();
        // prefetch the next step's i2h pre-activations while the tensor core works
        if (t + 1 < S) {
#pragma unroll
            for (int i = 0; i < NP; ++i)
#pragma unroll
                for (int g = 0; g < G; ++g)
                    nxt[i][g] = live[i] ? gates[((size_t)(t + 1) * B + b0 + sb + 16 * i) * GH + g * H + unit] : 0.f;
        }
        mbar_wait(mma_done, t & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- C: sum of the 8 accumulators -> scratch [gate][sequence][unit] -> gate math
        {
            const int q = warp & 3, cgp = warp >> 2;                               // TMEM lane quadrant (= gate), 8-column group
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + 8 * cgp;
            uint32_t r[8][8];                                                      // all eight loads in flight, one wait
#pragma unroll
            for (int p = 0; p < 8; ++p) tmem_ld8(taddr + 32 * p, r[p]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float sum[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = __uint_as_float(r[0][j]);
#pragma unroll
                for (int p = 1; p < 8; ++p) a += __uint_as_float(r[p][j]);         // fixed order: deterministic
                sum[j] = a;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) scratch[(q * kNB + 8 * cgp + j) * 32 + lane] = sum[j];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        float act[NP][G + 1];                                                      // activated gates (+ c | hn) of this step
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int bb = sb + 16 * i;
            float pre[G];
#pragma unroll
            for (int g = 0; g < G; ++g) pre[g] = bias[g] + scratch[(g * kNB + bb) * 32 + ul];
            float hnew;
            if (G == 3) {
                const float r = dc_sigmoid(cur[i][0] + pre[0]);
                const float z = dc_sigmoid(cur[i][1] + pre[1]);
                const float n = dc_tanh(cur[i][2] + r * pre[2]);
                hnew = (1.0f - z) * n + z * h_reg[i];
                act[i][0] = r; act[i][1] = z; act[i][2] = n; act[i][G] = pre[2];   // W_hn h + b_hn -> cbuf slot t+1
                h_reg[i] = hnew;
            } else {
                const float ig = dc_sigmoid(cur[i][0] + pre[0]);
                const float fg = dc_sigmoid(cur[i][1] + pre[1]);
                const float gg = dc_tanh(cur[i][2] + pre[2]);
                const float og = dc_sigmoid(cur[i][G - 1] + pre[G - 1]);
                c_reg[i] = fg * c_reg[i] + ig * gg;
                hnew = og * dc_tanh(c_reg[i]);
                act[i][0] = ig; act[i][1] = fg; act[i][2] = gg; act[i][G - 1] = og; act[i][G] = c_reg[i];
            }
            if (live[i]) ybuf[((size_t)(t + 1) * B + b0 + bb) * H + unit] = hnew;  // the slice the other CTAs wait for
        }
        // ---- D: publish this CTA's slice of h_t (release), then write the rest of the step's outputs behind the barrier
        cluster_arrive();
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (live[i]) {
                const size_t tok = (size_t)t * B + b0 + sb + 16 * i;
                float *gout = gates + tok * GH + unit;
#pragma unroll
                for (int g = 0; g < G; ++g) gout[g * H] = act[i][g];
                cbuf[(tok + B) * H + unit] = act[i][G];
            }
#pragma unroll
            for (int g = 0; g < G; ++g) cur[i][g] = nxt[i][g];
        }
        cluster_wait();
    }
    tmem_free_512(tmem_base, warp);
}

// ---- backward -----------------------------------------------------------------------------------------------------------
// TMEM: [0,256) W_hh^T slice, hi half, two M tiles: tile m, lane rho <-> k = 128 m + rho, column kappa = g*32 + u <->
// j = g*H + 32*rank + u;  [256,512) eight accumulators of 32 columns (M tile m, K-panel g at 256 + 32*(4m + g)).
// shared: lo half (2 x 4 K-panels), gate-gradient tile hi / lo.
struct BwdSmem {
    static constexpr size_t wlo = 0;
    static constexpr size_t ghi = wlo + 8 * kPanelA;
    static constexpr size_t glo = ghi + 4 * kPanelB;
    static constexpr size_t bars = glo + 4 * kPanelB;
    static constexpr size_t total = bars + 64 + 1024;
};

inline size_t bwd_workspace_bytes(int B) { return (size_t)2 * ((B + kNB - 1) / kNB) * kCL * kNB * kH * sizeof(float); }

template <int G>
__global__ void __launch_bounds__(kThreads, 1) bwd_cluster_kernel(float *gates, const float *__restrict__ w_hh, const float *ybuf,
                                                                   float *cbuf, const float *__restrict__ dy,
                                                                   const float *__restrict__ dhn, const float *__restrict__ dcn,
                                                                   float *__restrict__ dh0, float *__restrict__ dc0, float *part,
                                                                   int B, int S) {
    constexpr int H = kH, GH = G * kH, NP = kNP;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = align1024(smem_raw);
    unsigned char *wlo = base + BwdSmem::wlo, *ghi = base + BwdSmem::ghi, *glo = base + BwdSmem::glo;
    uint64_t *mma_done = reinterpret_cast<uint64_t *>(base + BwdSmem::bars);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mma_done + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = blockIdx.x % kCL, cl = blockIdx.x / kCL, ncl = gridDim.x / kCL, b0 = cl * kNB;

    if (tid == 0) {
        mbar_init(mma_done, 2 * G);                                       // one commit per issuing thread
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tmem_alloc_512(tmem_slot, warp);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_w = tmem_base, tmem_acc = tmem_base + 256;
    // MMA issue: lane 0 of warp c < 2G issues the 12 MMAs of chain c = (M tile c / G, K-panel c % G) (see the forward kernel)
    const bool issuer = lane == 0 && warp < 2 * G;
    const int im = (warp % (2 * G)) / G, ip = (warp % (2 * G)) % G;
    const uint64_t d_alo = make_desc(smem_u32(wlo) + (im * 4 + ip) * kPanelA), d_bhi = make_desc(smem_u32(ghi) + ip * kPanelB),
                   d_blo = make_desc(smem_u32(glo) + ip * kPanelB);
    const uint32_t t_acc = tmem_acc + 32 * (4 * im + ip), t_ahi = tmem_w + im * 128 + ip * 32;

    const int q = warp & 3, mt = (warp >> 2) & 1, wh = warp >> 3;          // TMEM lane quadrant, M tile, half (K range / columns)
    {   // resident W_hh^T slice, once: lane rho of tile mt <-> k; a warp reads 32 consecutive k of one row j (128 B)
        const int rho = q * 32 + lane, k = mt * 128 + rho;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
#pragma unroll 2
        for (int kap0 = wh * 64; kap0 < wh * 64 + 64; kap0 += 8) {
            float w[8], hi[8], lo[8];
            const int g = kap0 >> 5;
#pragma unroll
            for (int e = 0; e < 8; ++e)
                w[e] = g < G ? __ldg(w_hh + (size_t)(g * H + rank * 32 + (kap0 & 31) + e) * H + k) : 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { hi[e] = tf32_rna(w[e]); lo[e] = w[e] - hi[e]; }
            tmem_st8(tmem_w + lane_addr + mt * 128 + kap0, hi);
            unsigned char *panel = wlo + (size_t)(mt * 4 + g) * kPanelA;
            const int c0 = (kap0 & 31) >> 2;
            *reinterpret_cast<float4 *>(panel + swz(rho, c0)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4 *>(panel + swz(rho, c0 + 1)) = make_float4(lo[4], lo[5], lo[6], lo[7]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // gate phase: thread = (unit ul, sequences sb + 16 i)
    const int ul = lane, unit = rank * 32 + ul, sb = warp;
    bool live[NP];
    float dh_carry[NP], dc_carry[NP], c_cur[NP];
    // per-step inputs, prefetched one step ahead: saved gates, dy, aux0 (LSTM c_{t-1} | GRU hn), aux1 (GRU h_{t-1})
    float cg[NP][G], cdy[NP], ca0[NP], ca1[NP], ng[NP][G], ndy[NP], na0[NP], na1[NP];
    auto fetch = [&](int t, float (&fg)[NP][G], float (&fdy)[NP], float (&fa0)[NP], float (&fa1)[NP]) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const size_t tok = (size_t)t * B + b0 + sb + 16 * i;
            if (live[i]) {
#pragma unroll
                for (int g = 0; g < G; ++g) fg[i][g] = gates[tok * GH + g * H + unit];
                fdy[i] = __ldg(dy + tok * H + unit);
                if (G == 4) {
                    fa0[i] = cbuf[tok * H + unit];                                 // c_{t-1} (slot t)
                    fa1[i] = 0.f;
                } else {
                    fa0[i] = cbuf[(tok + B) * H + unit];                           // hn (slot t+1)
                    fa1[i] = ybuf[tok * H + unit];                                 // h_{t-1} (slot t)
                }
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) fg[i][g] = 0.f;
                fdy[i] = fa0[i] = fa1[i] = 0.f;
            }
        }
    };
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int b = b0 + sb + 16 * i;
        live[i] = b < B;
        dh_carry[i] = (live[i] && dhn) ? dhn[(size_t)b * H + unit] : 0.f;
        dc_carry[i] = (G == 4 && live[i] && dcn) ? dcn[(size_t)b * H + unit] : 0.f;
        c_cur[i] = (G == 4 && live[i]) ? cbuf[((size_t)S * B + b) * H + unit] : 0.f;
    }
    fetch(S - 1, cg, cdy, ca0, ca1);
    fetch(S - 1, ng, ndy, na0, na1);                                               // (initialises the second buffer)
    cluster_barrier();

    for (int it = 0; it < S; ++it) {
        const int t = S - 1 - it;
        // ---- recurrent gradient: fixed-order sum of the 8 CTAs' partials of the previous step, then the gate gradients
        const float *pprev = part + ((size_t)(((it + 1) & 1) * ncl + cl) * kCL) * kNB * H;
        float dgi[NP][G], daux[NP];                                                // global outputs of this step, stored behind the MMAs
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int bb = sb + 16 * i;
            float dh = cdy[i] + dh_carry[i];
            if (it > 0) {
                float pv[kCL];
#pragma unroll
                for (int r = 0; r < kCL; ++r) pv[r] = __ldcg(pprev + ((size_t)r * kNB + bb) * H + unit);
#pragma unroll
                for (int r = 0; r < kCL; ++r) dh += pv[r];
            }
            float d[4] = {0.f, 0.f, 0.f, 0.f};                                     // gradients wrt the h2h pre-activations
#pragma unroll
            for (int g = 0; g < G; ++g) dgi[i][g] = 0.f;
            daux[i] = 0.f;
            if (live[i]) {
                if (G == 3) {
                    const float r = cg[i][0], z = cg[i][1], n = cg[i][2], hn = ca0[i], hprev = ca1[i];
                    const float dpn = dh * (1.0f - z) * (1.0f - n * n);
                    const float dpz = dh * (hprev - n) * z * (1.0f - z);
                    const float dpr = dpn * hn * r * (1.0f - r);
                    const float dghn = dpn * r;
                    dgi[i][0] = dpr; dgi[i][1] = dpz; dgi[i][2] = dpn;             // dgi
                    daux[i] = dghn;                                                // n-gate part of dgh -> cbuf slot t+1
                    d[0] = dpr; d[1] = dpz; d[2] = dghn;
                    dh_carry[i] = dh * z;
                } else {
                    const float ig = cg[i][0], fg = cg[i][1], gg = cg[i][2], og = cg[i][G - 1], cprev = ca0[i];
                    const float tc = dc_tanh(c_cur[i]);
                    const float dc = dc_carry[i] + dh * og * (1.0f - tc * tc);
                    const float dpi = dc * gg * ig * (1.0f - ig);
                    const float dpf = dc * cprev * fg * (1.0f - fg);
                    const float dpg = dc * ig * (1.0f - gg * gg);
                    const float dpo = dh * tc * og * (1.0f - og);
                    dgi[i][0] = dpi; dgi[i][1] = dpf; dgi[i][2] = dpg; dgi[i][G - 1] = dpo;
                    d[0] = dpi; d[1] = dpf; d[2] = dpg; d[3] = dpo;
                    dc_carry[i] = dc * fg;
                    c_cur[i] = cprev;
                    dh_carry[i] = 0.f;
                }
            }
            // B operand: row = sequence bb, column kappa = g*32 + ul  ->  K-panel g, 16-byte chunk ul/4, word ul%4
            const int off = swz(bb, ul >> 2) + (ul & 3) * 4;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float hi = tf32_rna(d[g]);
                *reinterpret_cast<float *>(ghi + g * kPanelB + off) = hi;
                *reinterpret_cast<float *>(glo + g * kPanelB + off) = d[g] - hi;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (issuer) {                                                              // 2 x G chains of 12, one issuing thread each
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                umma_ss(t_acc, d_alo + 2 * ks, d_bhi + 2 * ks, kIdesc, ks != 0);
                umma_ts(t_acc, t_ahi + 8 * ks, d_blo + 2 * ks, kIdesc, 1u);
                umma_ts(t_acc, t_ahi + 8 * ks, d_bhi + 2 * ks, kIdesc, 1u);
            }
            umma_commit(mma_done);
        }
        __syncwarp();
        // behind the MMAs: this step's global outputs, then the next step's inputs
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (!live[i]) continue;
            const size_t tok = (size_t)t * B + b0 + sb + 16 * i;
            float *gout = gates + tok * GH + unit;
#pragma unroll
            for (int g = 0; g < G; ++g) gout[g * H] = dgi[i][g];
            if (G == 3) cbuf[(tok + B) * H + unit] = daux[i];
        }
        if (t > 0) fetch(t - 1, ng, ndy, na0, na1);
        mbar_wait(mma_done, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {   // partial dh_{t-1}[b][k] of this CTA's 128 gate columns -> scratch [buffer][cluster][rank][b][k]
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + 32 * (4 * mt) + 16 * wh;
            uint32_t r[G][16];                                                     // all G loads in flight, one wait
#pragma unroll
            for (int p = 0; p < G; ++p) tmem_ld16(taddr + 32 * p, r[p]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float sum[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float a = __uint_as_float(r[0][j]);
#pragma unroll
                for (int p = 1; p < G; ++p) a += __uint_as_float(r[p][j]);
                sum[j] = a;
            }
            float *dst = part + (((size_t)((it & 1) * ncl + cl) * kCL + rank) * kNB + 16 * wh) * H + mt * 128 + q * 32 + lane;
#pragma unroll
            for (int b = 0; b < 16; ++b) __stcg(dst + (size_t)b * H, sum[b]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        cluster_arrive();
#pragma unroll
        for (int i = 0; i < NP; ++i) {
#pragma unroll
            for (int g = 0; g < G; ++g) cg[i][g] = ng[i][g];
            cdy[i] = ndy[i]; ca0[i] = na0[i]; ca1[i] = na1[i];
        }
        cluster_wait();
    }
    // gradient of the initial state
    const float *plast = part + ((size_t)(((S - 1) & 1) * ncl + cl) * kCL) * kNB * H;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        if (!live[i]) continue;
        const int bb = sb + 16 * i;
        float dh = dh_carry[i];
#pragma unroll
        for (int r = 0; r < kCL; ++r) dh += __ldcg(plast + ((size_t)r * kNB + bb) * H + unit);
        if (dh0) dh0[(size_t)(b0 + bb) * H + unit] = dh;
        if (G == 4 && dc0) dc0[(size_t)(b0 + bb) * H + unit] = dc_carry[i];
    }
    tmem_free_512(tmem_base, warp);
}

inline bool cluster_supported(int H) { return H == kH; }

template <typename K, typename... Args>
inline int launch_cluster(K kern, int B, size_t smem, cudaStream_t st, Args... args) {
    DC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(((B + kNB - 1) / kNB) * kCL));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DC_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    return DC_OK;
}

inline int launch_fwd(int cell, float *gates, const float *w_hh, const float *b_hh, float *ybuf, float *cbuf, int B, int S,
                      cudaStream_t st) {
    if (cell == DC_CELL_GRU) return launch_cluster(fwd_cluster_kernel<3>, B, FwdSmem::total, st, gates, w_hh, b_hh, ybuf, cbuf, B, S);
    return launch_cluster(fwd_cluster_kernel<4>, B, FwdSmem::total, st, gates, w_hh, b_hh, ybuf, cbuf, B, S);
}
inline int launch_bwd(int cell, float *gates, const float *w_hh, const float *ybuf, float *cbuf, const float *dy, const float *dhn,
                      const float *dcn, float *dh0, float *dc0, float *part, int B, int S, cudaStream_t st) {
    if (cell == DC_CELL_GRU)
        return launch_cluster(bwd_cluster_kernel<3>, B, BwdSmem::total, st, gates, w_hh, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, part, B, S);
    return launch_cluster(bwd_cluster_kernel<4>, B, BwdSmem::total, st, gates, w_hh, ybuf, cbuf, dy, dhn, dcn, dh0, dc0, part, B, S);
}

}  // namespace dc_rnnc
