// fp32-accurate GEMM on the 5th-generation tensor cores:  C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (ReLU optional)
//
// Replaces the library SGEMM behind the dense layers of the hot path whose shapes are real contractions
// -- first of all the input-to-hidden GEMM of the recurrent layer (x W_ih^T + b_ih inside nn.GRU / nn.LSTM,
// policy.py:66,141; the one place north_star puts tensor cores) -- while keeping fp32-level accuracy, so
// that the parity tolerances against the fp32 reference hold:
//
//   3xTF32 split.  a = a_hi + a_lo with a_hi = rna_tf32(a), a_lo = a - a_hi (exact in fp32);
//   a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, three tcgen05.mma.kind::tf32 per k-step into one fp32
//   TMEM accumulator.  The dropped a_lo*b_lo term and the truncation of the lo parts are O(2^-22).
//
// Three kernels share the building blocks below (one persistent CTA per SM, warp-specialised, 17 warps):
//   gemm_tf32x3_kernel        any K: A and B both stream through a 3-stage shared-memory ring (4 tiles of 16 KB per stage:
//                             A_hi, A_lo, B_hi, B_lo); used for K > 128 (affine_pre_rnn, the i2h data gradient).
//   gemm_tf32x3_wtmem_kernel  K <= 128 (nearly every layer of this model): the TRANSPOSED product with the weight block
//                             resident in tensor memory as the A operand; activations through a 6-stage ring; an epilogue
//                             thread owns one output feature.  See the comment above the kernel.
//   gemm_wgrad_atmem_kernel   weight gradient dW = dY^T X (split-K), dY^T fed to the MMAs from tensor memory.
// (The round-1 variants with the weight block / both wgrad operands in shared memory were measured slower and are gone.)
// Roles in every kernel:
//   PRODUCER warps (groups of 4 taking alternate k-chunks, so the global-load latency of one chunk overlaps the split
//              arithmetic of another): coalesced global loads of a [128 x 32] fp32 chunk, hi/lo split on the CUDA cores,
//              stores into the canonical SWIZZLE_128B shared-memory layout (or tcgen05.st into TMEM), fence.proxy.async,
//              one elected mbarrier arrive per warp.  (TMA cannot do this step: the split is arithmetic, so the data passes
//              through registers anyway.)
//   MMA ISSUER (one warp): an elected lane issues 12 tcgen05.mma (4 k-steps x 3 products) per 32-wide k-chunk from
//              shared-memory descriptors / TMEM addresses; tcgen05.commit releases the stage / publishes the accumulator.
//   EPILOGUE   (4 warps): tcgen05.ld 32x32b, + bias, ReLU, global stores.  Two TMEM accumulator buffers (2 x 128 columns)
//              overlap the epilogue of tile i with the main loop of tile i+1.
// Tensor-pipe work: 2*M*N*K*3 flops; HBM: 4*(M*K + N*K + M*N) bytes.  For the K=128 GEMMs of this model the kernels are
// HBM-bound even with the 3x flops (measured: 5.2-5.5 TB/s on the 2M x 128 x 128 layers).
#include "dc_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;           // BK floats = 128 bytes = one swizzle-128B row
constexpr int kStages = 3;
constexpr int kTileBytes = BM * BK * 4;              // 16 KB
constexpr int kStageBytes = 4 * kTileBytes;          // A_hi, A_lo, B_hi, B_lo
constexpr int kProducerThreads = 128;
constexpr int kProducerGroups = 3;
constexpr int kMmaWarp = 4 * kProducerGroups;
constexpr int kThreads = (kMmaWarp + 1 + 4) * 32;
constexpr int kAccCols = 128, kTmemCols = 256;       // two accumulator buffers
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + 128 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void umma_commit(uint64_t *bar) {   // arrives on `bar` when all prior MMAs of this thread retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (8 rows * 128 B = 1024)
// | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), both K-major,
// N>>3 at bit 17, M>>4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ float tf32_rna(float v) {
    // round-to-nearest, ties away from zero, to 10 mantissa bits: add half an ulp to the magnitude, clear the low 13 bits.
    // Same result as cvt.rna.tf32.f32 for finite inputs (which ptxas expands to this plus an |v| < inf test); inf stays inf,
    // a NaN may become inf here but v - hi is then NaN, so it still poisons the product.
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}

// Row addressing of an operand: plain (row r at r*ld) or two-level (rows grouped in blocks of `rpb` rows that are
// `bs` floats apart: row r at (r / rpb) * bs + (r % rpb) * ld).  The two-level form lets a GEMM read or write one
// unit group of the [N, 40, 128] unit-embedding tensor in place (policy.py:130-131's torch.cat disappears).
struct RowMap {
    int ld;
    int rpb;          // 0 = plain
    long long bs;
    unsigned mul;     // r / rpb without a divide: t = umulhi(r, mul); q = (t + ((r - t) >> 1)) >> sh   (exact for all 32-bit r,
    int sh;           // rpb >= 2; mul = floor(2^32 (2^s - rpb) / rpb) + 1, s = ceil(log2 rpb), sh = s - 1) -- set by make_rowmap
    __device__ __forceinline__ size_t off(int r) const {
        if (rpb == 0) return (size_t)r * ld;
        const unsigned n = (unsigned)r, t = __umulhi(n, mul);
        const unsigned q = (t + ((n - t) >> 1)) >> sh;
        return (size_t)q * (size_t)bs + (size_t)(n - q * (unsigned)rpb) * ld;
    }
};

// K-major operand chunk: rows [row0, row0+128) x cols [k0, k0+32) of a row-major fp32 matrix.  tile_load_k issues the
// eight 16-byte loads of this thread (rows >= rows_total read as zero); tile_store_k splits hi/lo and stores both
// swizzle-128B tiles.  Keeping the two halves apart lets a producer issue the NEXT chunk's loads before it touches the
// current chunk's data (register double buffering), which hides the global-load latency ncu showed as 64 % long-scoreboard.
__device__ __forceinline__ void tile_load_k(const float *__restrict__ src, RowMap map, int row0, int rows_total, int k0, int t,
                                            float4 (&v)[8]) {
    const int c = t & 7, r0 = t >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + 16 * i;
        v[i] = (row0 + r < rows_total) ? __ldg(reinterpret_cast<const float4 *>(src + map.off(row0 + r) + k0) + c)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
__device__ __forceinline__ void tile_store_k(const float4 (&v)[8], unsigned char *dst_hi, unsigned char *dst_lo, int t) {
    const int c = t & 7, r0 = t >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + 16 * i;
        const int off = r * 128 + ((c ^ (r & 7)) << 4);
        float4 hi, lo;
        hi.x = tf32_rna(v[i].x); hi.y = tf32_rna(v[i].y); hi.z = tf32_rna(v[i].z); hi.w = tf32_rna(v[i].w);
        lo.x = v[i].x - hi.x; lo.y = v[i].y - hi.y; lo.z = v[i].z - hi.z; lo.w = v[i].w - hi.w;
        *reinterpret_cast<float4 *>(dst_hi + off) = hi;
        *reinterpret_cast<float4 *>(dst_lo + off) = lo;
    }
}
__device__ __forceinline__ void produce_tile(const float *__restrict__ src, RowMap map, int row0, int rows_total, int k0,
                                             unsigned char *dst_hi, unsigned char *dst_lo, int t) {
    float4 v[8];
    tile_load_k(src, map, row0, rows_total, k0, t, v);
    tile_store_k(v, dst_hi, dst_lo, t);
}

// Epilogue store of 32 consecutive outputs of one row held by one thread.  `wide` = 256-bit stores (STG.256, sm_100):
// every lane then writes whole 32-byte sectors; with 16-byte stores each sector is touched twice (ncu: 32 sectors per
// request, half filled).
__device__ __forceinline__ void store_row32(float *dst, const uint32_t (&r)[32], const float *bias, int relu, bool wide) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        float o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = __uint_as_float(r[j + q]);
        if (bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + j)), b1 = __ldg(reinterpret_cast<const float4 *>(bias + j + 4));
            o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w; o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
        }
        if (relu) {
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = fmaxf(o[q], 0.f);
        }
        if (wide) {
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + j), "f"(o[0]), "f"(o[1]), "f"(o[2]),
                         "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                         : "memory");
        } else {
            *reinterpret_cast<float4 *>(dst + j) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(dst + j + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) gemm_tf32x3_kernel(const float *__restrict__ A, RowMap amap,
                                                                  const float *__restrict__ B, int ldb,
                                                                  const float *__restrict__ bias, float *__restrict__ C,
                                                                  RowMap cmap, int M, int N, int K, int relu, bool wide,
                                                                  int ksplit, long long c_split_stride) {
    // ksplit > 1: split-K -- work item = (K range sp, output tile); range sp writes its partial product to
    // C + sp * c_split_stride (the caller sums the partials in a fixed order), bias goes with range 0.
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + (size_t)kStages * kStageBytes);
    uint64_t *full = bars, *empty = bars + kStages, *acc_full = bars + 2 * kStages, *acc_empty = bars + 2 * kStages + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (M + BM - 1) / BM, n_blocks = N / BN, k_chunks = K / BK / ksplit;      // chunks per K range
    const int tiles_mn = m_blocks * n_blocks, tiles_total = tiles_mn * ksplit;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], kProducerThreads / 32); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {   // TMEM allocation: one full warp; the address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kMmaWarp) {
        // ===== PRODUCERS =====  group g takes the k-chunks whose running index is == g (mod kProducerGroups)
        const int t = threadIdx.x & (kProducerThreads - 1), g = warp >> 2;
        uint32_t chunk = 0;
        for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
            const int tmn = tile % tiles_mn, kc0 = (tile / tiles_mn) * k_chunks;
            const int m0 = (tmn / n_blocks) * BM, n0 = (tmn % n_blocks) * BN;
            for (int kc = 0; kc < k_chunks; ++kc, ++chunk) {
                if ((int)(chunk % kProducerGroups) != g) continue;
                const int stage = chunk % kStages;
                const uint32_t phase = (chunk / kStages) & 1;
                mbar_wait(&empty[stage], phase ^ 1);
                unsigned char *st = tiles + (size_t)stage * kStageBytes;
                produce_tile(A, amap, m0, M, (kc0 + kc) * BK, st, st + kTileBytes, t);
                produce_tile(B, RowMap{ldb, 0, 0, 0, 0}, n0, N, (kc0 + kc) * BK, st + 2 * kTileBytes, st + 3 * kTileBytes, t);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);        // one arrival per warp (128 single arrivals serialise on the barrier)
            }
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA ISSUER =====
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++it) {
            const int a = it & 1;
            mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);                    // epilogue drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + a * kAccCols;
            for (int kc = 0; kc < k_chunks; ++kc) {
                mbar_wait(&full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t base = smem_u32(tiles + (size_t)stage * kStageBytes);
                    const uint64_t a_hi = make_desc(base), a_lo = make_desc(base + kTileBytes);
                    const uint64_t b_hi = make_desc(base + 2 * kTileBytes), b_lo = make_desc(base + 3 * kTileBytes);
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {                      // UMMA_K = 8 tf32 = 32 bytes = 2 x 16 B
                        const uint64_t adv = (uint64_t)(ks * 2);
                        const uint32_t first = (kc | ks) != 0;
                        umma_tf32(tmem_d, a_lo + adv, b_hi + adv, kIdesc, first);   // small terms first
                        umma_tf32(tmem_d, a_hi + adv, b_lo + adv, kIdesc, 1u);
                        umma_tf32(tmem_d, a_hi + adv, b_hi + adv, kIdesc, 1u);
                    }
                    umma_commit(&empty[stage]);                                 // stage reusable once these MMAs retire
                    if (kc == k_chunks - 1) umma_commit(&acc_full[a]);          // accumulator complete
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== EPILOGUE (4 warps; TMEM lane quadrant = warp % 4) =====
        const int q = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++it) {
            const int a = it & 1;
            const int tmn = tile % tiles_mn, sp = tile / tiles_mn;
            const int m0 = (tmn / n_blocks) * BM, n0 = (tmn % n_blocks) * BN;
            mbar_wait(&acc_full[a], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + q * 32 + lane;
            float *crow = C + (size_t)sp * (size_t)c_split_stride + cmap.off(row < M ? row : 0) + n0;
            const float *tile_bias = (bias && sp == 0) ? bias : nullptr;
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * kAccCols + cb);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < M) store_row32(crow + cb, r, tile_bias ? tile_bias + n0 + cb : nullptr, relu, wide);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}


constexpr int kMaxResChunks = 4;                     // K <= 128: the weight block fits the tensor-memory-resident kernel

// =================================================================================================
// K <= 128, WEIGHTS IN TENSOR MEMORY, transposed product.  D^T[feature][token] = W[feature][k] . X[token][k]:
//   * the 128 x K weight block (hi and lo halves) is written ONCE per CTA into TMEM (tcgen05.st, one lane = one output
//     feature) and is the A operand of every MMA (tcgen05.mma [d], [a_tmem], b-desc) -- it costs no shared memory and no
//     shared-memory bandwidth;
//   * the activations stream through a 6-stage shared-memory ring as the B operand (K-major swizzle-128B tiles, hi/lo),
//     loaded with fully coalesced 16-byte accesses;
//   * the accumulator is feature-major, so an epilogue thread owns ONE output feature: its bias is a register and a warp
//     store writes 32 consecutive floats of one output row -- one 128-byte wavefront, fully coalesced.
// History (profiles/r1_gemm_ncu.md): with the activations as the A operand one thread owns one ROW, and both its loads
// (tcgen05.st wants lane == row) and its stores touch 32 different lines per instruction; ncu showed the LSU data pipe at
// 80 % with DRAM at 50 % and the MMA warp waiting 40 % of the time for the epilogue to drain the accumulator.
// TMEM map (512 columns): [0,256) two accumulators (128 tokens each), [256,384) W_hi, [384,512) W_lo.
constexpr int kWStages = 6;
constexpr int kWStageBytes = 2 * kTileBytes;                       // X_hi, X_lo
constexpr size_t kSmemBytesWTmem = (size_t)kWStages * kWStageBytes + 1024 + 256;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
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
// 32 tokens x this thread's feature: C[row(tok0 + j)][col] = r[j] + bias (ReLU); a warp writes 128 contiguous bytes per token
template <bool kFull>
__device__ __forceinline__ void store_feature32_impl(float *__restrict__ C, RowMap cmap, int tok0, int n, int col, const uint32_t (&r)[32],
                                                     float bias, int relu, int lane) {
    if (kFull && cmap.ld == 128 && (cmap.rpb == 0 || cmap.rpb == 16)) {
        // the two layouts nearly all rows of this model go through (a plain [M,128] matrix; 16-unit groups of the
        // [N,40,128] unit embedding): 32 rows = one or two runs of rows 512 bytes apart -> immediate store offsets,
        // 2 instructions per row (ncu: the generic path's address arithmetic made the epilogue ~40 % of the kernel's instructions)
        float *p0, *p1;
        if (cmap.rpb == 0) {
            p0 = C + (size_t)tok0 * 128 + col;
            p1 = p0 + 16 * 128;
        } else {                                                   // tok0 % 16 == 0 (tiles start on multiples of 128 rows)
            p0 = C + (size_t)(tok0 >> 4) * (size_t)cmap.bs + col;
            p1 = p0 + cmap.bs;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float o = __uint_as_float(r[j]) + bias;
            if (relu) o = fmaxf(o, 0.f);
            (j < 16 ? p0 : p1)[(j & 15) * 128] = o;
        }
    } else if (cmap.rpb == 0) {
        float *p = C + (size_t)tok0 * cmap.ld + col;
        const size_t ld = cmap.ld;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (kFull || j < n) {
                float o = __uint_as_float(r[j]) + bias;
                if (relu) o = fmaxf(o, 0.f);
                p[j * ld] = o;
            }
        }
    } else {
        // two-level rows: lane j works out the offset of token tok0 + j once (one division), the loop broadcasts it --
        // independent shuffles instead of a serial wrap-around pointer walk
        const unsigned long long mine = (unsigned long long)cmap.off(tok0 + min(lane, n - 1));
        const uint32_t lo = (uint32_t)mine, hi = (uint32_t)(mine >> 32);
        float *base = C + col;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const unsigned long long off = ((unsigned long long)__shfl_sync(0xffffffffu, hi, j) << 32) | __shfl_sync(0xffffffffu, lo, j);
            if (kFull || j < n) {
                float o = __uint_as_float(r[j]) + bias;
                if (relu) o = fmaxf(o, 0.f);
                base[off] = o;
            }
        }
    }
}
__device__ __forceinline__ void store_feature32(float *__restrict__ C, RowMap cmap, int tok0, int M, int col, const uint32_t (&r)[32],
                                                float bias, int relu, int lane) {
    const int n = M - tok0;                                       // valid tokens in this group
    if (n >= 32) store_feature32_impl<true>(C, cmap, tok0, 32, col, r, bias, relu, lane);
    else if (n > 0) store_feature32_impl<false>(C, cmap, tok0, n, col, r, bias, relu, lane);
}

// NU > 0: "unit max" epilogue for the unit-embedding layers (policy.py:101-127): the rows are (token, unit) pairs, NU units
// per token; instead of storing the [M, 128] embedding the epilogue reduces every token's NU rows to their maximum (+ bias)
// and the index of the maximising unit: C[token * cmap.ld + feature] (and C_copy, policy.py:127), argmax[token * 128 + feature].
// Tiles advance by kTileRows = 128 - 128 % NU rows (125 for the 5-unit group) so that a token never straddles two tiles.
template <int NU>
__global__ void __launch_bounds__(kThreads, 1) gemm_tf32x3_wtmem_kernel(const float *__restrict__ A, RowMap amap,
                                                                        const float *__restrict__ B, int ldb,
                                                                        const float *__restrict__ bias, float *__restrict__ C,
                                                                        RowMap cmap, int M, int N, int K, int relu,
                                                                        float *__restrict__ C_copy, uint8_t *__restrict__ argmax) {
    constexpr int kTileRows = NU > 0 ? BM - BM % NU : BM;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + (size_t)kWStages * kWStageBytes);
    uint64_t *full = bars, *empty = bars + kWStages, *acc_full = bars + 2 * kWStages, *acc_empty = bars + 2 * kWStages + 2;
    uint64_t *w_ready = bars + 2 * kWStages + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kWStages + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (M + kTileRows - 1) / kTileRows, n_blocks = N / BN, k_chunks = K / BK;
    const int n_blk = blockIdx.x % n_blocks, m_first = blockIdx.x / n_blocks, m_step = gridDim.x / n_blocks;
    const int n0 = n_blk * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWStages; ++s) { mbar_init(&full[s], kProducerThreads / 32); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        mbar_init(w_ready, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_w_hi = tmem_base + 2 * kAccCols, tmem_w_lo = tmem_w_hi + 128;

    if (warp < kMmaWarp) {
        // ===== PRODUCERS =====
        const int t = threadIdx.x & (kProducerThreads - 1), g = warp >> 2;
        if (g == 0) {
            // the weight block -> TMEM, once: thread = output feature n0 + 32*(warp&3) + lane = its TMEM lane
            const float *wrow = B + (size_t)(n0 + (warp & 3) * 32 + lane) * ldb;
            const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
            for (int k0 = 0; k0 < K; k0 += 8) {
                const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wrow + k0)), w1 = __ldg(reinterpret_cast<const float4 *>(wrow + k0 + 4));
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                float hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { hi[q] = tf32_rna(w[q]); lo[q] = w[q] - hi[q]; }
                tmem_st8(tmem_w_hi + lane_addr + k0, hi);
                tmem_st8(tmem_w_lo + lane_addr + k0, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(w_ready);
        }
        // activation chunks: running index c = tile_iter * k_chunks + kc; this group takes c == g (mod groups).  Two register
        // buffers in ping-pong: a buffer is refilled (two rounds ahead) as soon as it has been stored to shared memory.
        const int n_my_tiles = m_first < m_blocks ? (m_blocks - m_first + m_step - 1) / m_step : 0;
        const uint32_t total_chunks = (uint32_t)n_my_tiles * k_chunks;
        float4 va[8], vb[8];
        auto fetch = [&](uint32_t cc, float4 (&buf)[8]) {
            if (cc < total_chunks) tile_load_k(A, amap, (m_first + (int)(cc / k_chunks) * m_step) * kTileRows, M, (int)(cc % k_chunks) * BK, t, buf);
        };
        auto emit = [&](uint32_t cc, const float4 (&buf)[8]) {
            const int stage = cc % kWStages;
            mbar_wait(&empty[stage], ((cc / kWStages) & 1) ^ 1);
            unsigned char *st = tiles + (size_t)stage * kWStageBytes;
            tile_store_k(buf, st, st + kTileBytes, t);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        };
        fetch(g, va);
        fetch(g + kProducerGroups, vb);
        for (uint32_t c = g; c < total_chunks; c += 2 * kProducerGroups) {
            emit(c, va);
            fetch(c + 2 * kProducerGroups, va);
            if (c + kProducerGroups < total_chunks) {
                emit(c + kProducerGroups, vb);
                fetch(c + 3 * kProducerGroups, vb);
            }
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA ISSUER =====
        mbar_wait(w_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t c = 0;
        int it = 0;
        for (int mb = m_first; mb < m_blocks; mb += m_step, ++it) {
            const int a = it & 1;
            mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + a * kAccCols;
            for (int kc = 0; kc < k_chunks; ++kc, ++c) {
                const int stage = c % kWStages;
                mbar_wait(&full[stage], (c / kWStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t xbase = smem_u32(tiles + (size_t)stage * kWStageBytes);
                    const uint64_t x_hi = make_desc(xbase), x_lo = make_desc(xbase + kTileBytes);
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);             // 32 bytes along K inside the 128-byte swizzle row
                        const uint32_t kcol = kc * BK + ks * 8;
                        const uint32_t first = (kc | ks) != 0;
                        umma_tf32_ts(tmem_d, tmem_w_lo + kcol, x_hi + adv, kIdesc, first);
                        umma_tf32_ts(tmem_d, tmem_w_hi + kcol, x_lo + adv, kIdesc, 1u);
                        umma_tf32_ts(tmem_d, tmem_w_hi + kcol, x_hi + adv, kIdesc, 1u);
                    }
                    umma_commit(&empty[stage]);
                    if (kc == k_chunks - 1) umma_commit(&acc_full[a]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== EPILOGUE =====  thread = output feature; accumulator columns = the tile's 128 tokens
        const int q = warp & 3;
        const int col = n0 + q * 32 + lane;
        const float bias_f = bias ? __ldg(bias + col) : 0.f;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        int it = 0;
        for (int mb = m_first; mb < m_blocks; mb += m_step, ++it) {
            const int a = it & 1;
            const int m0 = mb * kTileRows;
            mbar_wait(&acc_full[a], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(a * kAccCols);
            if constexpr (NU > 0) {
                // thread = output feature; the accumulator columns are the tile's rows = (token, unit) pairs in order
                const int tok0 = m0 / NU, n_tok = M / NU;
                float best = 0.f;
                int best_u = 0;
#pragma unroll
                for (int ld = 0; ld < 4; ++ld) {
                    uint32_t r[32];
                    tmem_ld32(taddr + 32 * ld, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (ld == 3) {                                                  // accumulator fully read: hand it back early
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[a]);
                    }
                    if constexpr (NU == 16) {
                        // two whole tokens per load: maximum by a 4-level tree, arg-max = first unit that equals it (torch.max) --
                        // independent instructions instead of a 16-long dependent compare/select chain
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float v[16];
#pragma unroll
                            for (int u = 0; u < 16; ++u) v[u] = __uint_as_float(r[16 * h + u]);
                            float m8[8], m4[4];
#pragma unroll
                            for (int u = 0; u < 8; ++u) m8[u] = fmaxf(v[2 * u], v[2 * u + 1]);
#pragma unroll
                            for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m8[2 * u], m8[2 * u + 1]);
                            const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                            int i8[8], i4[4];
#pragma unroll
                            for (int u = 0; u < 8; ++u) i8[u] = v[2 * u] == m ? 2 * u : (v[2 * u + 1] == m ? 2 * u + 1 : 16);
#pragma unroll
                            for (int u = 0; u < 4; ++u) i4[u] = min(i8[2 * u], i8[2 * u + 1]);
                            const int am = min(min(i4[0], i4[1]), min(i4[2], i4[3]));
                            const int tok = tok0 + 2 * ld + h;
                            if (tok < n_tok) {
                                const float o = m + bias_f;
                                C[(size_t)tok * cmap.ld + col] = o;
                                if (C_copy) C_copy[(size_t)tok * cmap.ld + col] = o;
                                argmax[(size_t)tok * BN + col] = (uint8_t)am;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int row = ld * 32 + j;                            // compile-time after unrolling
                            if (row < kTileRows) {
                                const int u = row % NU, tok = tok0 + row / NU;
                                const float v = __uint_as_float(r[j]);
                                if (u == 0 || v > best) { best = v; best_u = u; }   // first maximum wins (torch.max)
                                if (u == NU - 1 && tok < n_tok) {
                                    const float o = best + bias_f;
                                    C[(size_t)tok * cmap.ld + col] = o;
                                    if (C_copy) C_copy[(size_t)tok * cmap.ld + col] = o;
                                    argmax[(size_t)tok * BN + col] = (uint8_t)best_u;
                                }
                            }
                        }
                    }
                }
            } else {
                uint32_t ra[32], rb[32];
                tmem_ld32(taddr, ra);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_ld32(taddr + 32, rb);
                store_feature32(C, cmap, m0, M, col, ra, bias_f, relu, lane);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_ld32(taddr + 64, ra);
                store_feature32(C, cmap, m0 + 32, M, col, rb, bias_f, relu, lane);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_ld32(taddr + 96, rb);
                store_feature32(C, cmap, m0 + 64, M, col, ra, bias_f, relu, lane);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");     // accumulator fully read: hand it back early
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[a]);
                store_feature32(C, cmap, m0 + 96, M, col, rb, bias_f, relu, lane);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// =================================================================================================
// Unit-embedding layer, backward DATA path, fused end to end (policy.py:100-127,152-153 backwards).  In the reference's
// autograd the gradient of the [N, units, 128] embedding is materialised, multiplied by W_g, masked by the ReLU of the
// `basic` layer and reduced into dW_b / db_b -- four passes over [N*units, 128] tensors.  Here none of them exists in memory:
//   * PRODUCERS generate the embedding gradient straight into the swizzled shared-memory tiles:
//       d_emb[(n,u), c] = (argmax[n,c] == u ? d_xmax[n,c] : 0)  +  dlogits[n,u] * att[n,c]
//     (max-pool routing + the target-unit head's rank-1 part); the sources are one gradient row, one arg-max row and one
//     attention row per TOKEN, re-read from L1/L2 by the token's units and prefetched into L2 one tile ahead.
//   * D^T[j][(n,u)] = sum_c W_g^T[j][c] d_emb[(n,u)][c] on the tensor cores exactly as in gemm_tf32x3_wtmem_kernel (W_g^T
//     resident in tensor memory, 3xTF32), accumulator feature-major.
//   * EPILOGUE thread = one feature j of the basic layer: it recomputes the ReLU mask of every row from the row's 12 raw unit
//     features (the same FMA chain as unit_basic_fwd_kernel, bit-identical to the forward value) and accumulates dW_b[j][0..12)
//     and db_b[j] in registers across ALL of the CTA's tiles; one [128][13] partial per epilogue half per CTA at the end,
//     summed in a fixed order by the basic layer's reduce kernel.  The raw features of a tile (<= 6 KB, contiguous) arrive in
//     shared memory by one 1-D bulk copy (cp.async.bulk + mbarrier complete_tx), double-buffered one tile ahead -- broadcast
//     global loads would expose a DRAM latency per row.
// HBM traffic: d_xmax + arg-max + att + dlogits + units, about 1/30 of the four dense passes.  The kernel is bound by the
// epilogue's CUDA-core work (~25 instructions per row and feature), hence 8 epilogue warps (two per tensor-memory lane
// quadrant, each taking half of a tile's rows) and only 2 producer groups.
// Roles: warps 0-7 producers (2 groups), warp 8 MMA issuer, warps 9-16 epilogue.
constexpr int kDGroups = 2;
constexpr int kDMmaWarp = 4 * kDGroups;
constexpr int kDIn = 12;                                           // raw features per unit (policy.py:56)
constexpr int kDUnitsTile = BM * kDIn;                             // floats per staged units tile
constexpr size_t kSmemBytesDgrad = (size_t)kWStages * kWStageBytes + 1024 + 256 + 2 * kDUnitsTile * sizeof(float);

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {   // 16-byte aligned, multiple of 16
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {          // explicit ld.shared (volatile: stays behind the mbarrier wait)
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Producer side of one chunk [128 rows x 32 channels] of d_emb, in two steps so that the global loads of a chunk are ONE batch of
// independent instructions issued two chunks ahead of their use (a first version loaded row by row behind bounds tests: eight
// dependent round trips to L2 per chunk, 27,000 cycles per tile, profiles/r2_encoder_bwd.md):
//   demb_fetch  raw sources -> registers.  For the 16- and 5-unit groups a thread owns (token, 16-byte channel slice) pairs: the
//               token's gradient slice, its 4 arg-max bytes, its attention slice and the dlogits of the pair's units (8 of the
//               16 / all 5) -- 18 registers for 8 rows of the chunk, instead of 8 finished rows.  Tokens past the end are clamped
//               for the loads and zeroed in the second step.
//   demb_store  d_emb rows from the raw registers, hi/lo split, swizzle-128B stores (row r, slice c at r*128 + ((c ^ r%8) << 4)).
// The 1-unit groups (row == token) keep finished rows (float4 v[8], the mapping of tile_load_k) with clamped, unconditional loads.
template <int NU>
struct DembRaw {
    static constexpr int kP = NU == 16 ? 1 : 2;                   // (token, slice) pairs per thread
    static constexpr int kU = NU == 16 ? 8 : NU;                  // units of a pair this thread generates
    float4 d[kP], d2[kP], a[kP];                                  // d2: second gradient source (16-unit form only; else added at fetch)
    uint32_t am[kP];
    float g[kP][kU];
    bool ok[kP];
};
template <>
struct DembRaw<1> {
    float4 v[8];
};

template <int NU>
__device__ __forceinline__ void demb_fetch(DembRaw<NU> &raw, const float *__restrict__ dx, const float *__restrict__ dx2, int ld_dx,
                                           const uint8_t *__restrict__ argmax, const float *__restrict__ dl, int ld_dl,
                                           const float *__restrict__ att, int tok0, int n_tok, int k0, int t) {
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (NU == 1) {
        const int c = t & 7, r0 = t >> 3;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = tok0 + r0 + 16 * i;
            const bool ok = n < n_tok;
            const size_t nn = (size_t)min(n, n_tok - 1);
            float4 d = dx != nullptr ? __ldg(reinterpret_cast<const float4 *>(dx + nn * ld_dx + k0) + c) : zero4;
            if (dx2 != nullptr) {
                const float4 d2 = __ldg(reinterpret_cast<const float4 *>(dx2 + nn * ld_dx + k0) + c);
                d.x += d2.x; d.y += d2.y; d.z += d2.z; d.w += d2.w;
            }
            const float4 a4 = dl != nullptr ? __ldg(reinterpret_cast<const float4 *>(att + nn * 128 + k0) + c) : zero4;
            const float gl = dl != nullptr ? __ldg(dl + nn * ld_dl) : 0.f;
            raw.v[i] = ok ? make_float4(fmaf(gl, a4.x, d.x), fmaf(gl, a4.y, d.y), fmaf(gl, a4.z, d.z), fmaf(gl, a4.w, d.w)) : zero4;
        }
    } else {
        constexpr int kTileToks = (BM - BM % NU) / NU;
#pragma unroll
        for (int k = 0; k < DembRaw<NU>::kP; ++k) {
            const int p = t + kProducerThreads * k;                // NU == 16: one pair per thread, two threads (unit halves) per pair
            const int c = p & 7, tokl = NU == 16 ? p >> 4 : p >> 3;
            const int u0 = NU == 16 ? ((p >> 3) & 1) * 8 : 0;
            const bool active = tokl < kTileToks;
            const int n = tok0 + tokl;
            raw.ok[k] = active && n < n_tok;
            const size_t nn = (size_t)min(n, n_tok - 1);
            float4 d = (active && dx != nullptr) ? __ldg(reinterpret_cast<const float4 *>(dx + nn * ld_dx + k0) + c) : zero4;
            const float4 d2 = (active && dx2 != nullptr) ? __ldg(reinterpret_cast<const float4 *>(dx2 + nn * ld_dx + k0) + c) : zero4;
            if constexpr (NU == 16) {
                raw.d2[k] = d2;                                    // added in demb_store: no arithmetic on loaded values here, so that
            } else {                                               // the fetch never waits for its own loads
                d.x += d2.x; d.y += d2.y; d.z += d2.z; d.w += d2.w;
            }
            raw.d[k] = d;
            raw.am[k] = (active && dx != nullptr) ? __ldg(reinterpret_cast<const uint32_t *>(argmax + nn * 128 + k0) + c) : 0u;   // 4 arg-max bytes
            raw.a[k] = (active && dl != nullptr) ? __ldg(reinterpret_cast<const float4 *>(att + nn * 128 + k0) + c) : zero4;
#pragma unroll
            for (int j = 0; j < DembRaw<NU>::kU; ++j) raw.g[k][j] = (active && dl != nullptr) ? __ldg(dl + nn * ld_dl + u0 + j) : 0.f;
        }
    }
}

__device__ __forceinline__ void store_hi_lo(unsigned char *dst_hi, unsigned char *dst_lo, int off, float4 o) {
    float4 hi, lo;
    hi.x = tf32_rna(o.x); hi.y = tf32_rna(o.y); hi.z = tf32_rna(o.z); hi.w = tf32_rna(o.w);
    lo.x = o.x - hi.x; lo.y = o.y - hi.y; lo.z = o.z - hi.z; lo.w = o.w - hi.w;
    *reinterpret_cast<float4 *>(dst_hi + off) = hi;
    *reinterpret_cast<float4 *>(dst_lo + off) = lo;
}

template <int NU>
__device__ __forceinline__ void demb_store(const DembRaw<NU> &raw, unsigned char *dst_hi, unsigned char *dst_lo, int t) {
    if constexpr (NU == 1) {
        tile_store_k(raw.v, dst_hi, dst_lo, t);
    } else {
        constexpr int kTileRows = BM - BM % NU, kTileToks = kTileRows / NU;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < DembRaw<NU>::kP; ++k) {
            const int p = t + kProducerThreads * k;
            const int c = p & 7, tokl = NU == 16 ? p >> 4 : p >> 3;
            const int u0 = NU == 16 ? ((p >> 3) & 1) * 8 : 0;
            if (tokl < kTileToks) {
                float4 d = raw.d[k];
                if constexpr (NU == 16) { d.x += raw.d2[k].x; d.y += raw.d2[k].y; d.z += raw.d2[k].z; d.w += raw.d2[k].w; }
                const float4 a = raw.a[k];
                const uint32_t am = raw.am[k];
#pragma unroll
                for (int j = 0; j < DembRaw<NU>::kU; ++j) {
                    const int u = u0 + j, r = tokl * NU + u;
                    const float gl = raw.g[k][j];
                    float4 o;
                    o.x = fmaf(gl, a.x, (int)(am & 0xffu) == u ? d.x : 0.f);
                    o.y = fmaf(gl, a.y, (int)((am >> 8) & 0xffu) == u ? d.y : 0.f);
                    o.z = fmaf(gl, a.z, (int)((am >> 16) & 0xffu) == u ? d.z : 0.f);
                    o.w = fmaf(gl, a.w, (int)(am >> 24) == u ? d.w : 0.f);
                    store_hi_lo(dst_hi, dst_lo, r * 128 + ((c ^ (r & 7)) << 4), raw.ok[k] ? o : zero4);
                }
            }
        }
        if constexpr (kTileRows < BM) {                            // the rows of the MMA tile past the last whole token stay zero
            if (t < (BM - kTileRows) * 8) {
                const int r = kTileRows + (t >> 3), c = t & 7;
                store_hi_lo(dst_hi, dst_lo, r * 128 + ((c ^ (r & 7)) << 4), zero4);
            }
        }
    }
}

// one 32-row group LD of a tile: accumulator columns [32 LD, 32 LD + 32) -> registers, then per row the recomputed ReLU mask and
// the rank-1 updates of this thread's dW_b row.  `us` = shared-window address of the tile's raw unit features (all lanes read the same
// address: broadcast); kFull = every row of the tile exists (the hot path: no per-row bounds test).
template <int NU, int LD, bool kFull>
__device__ __forceinline__ void dgrad_rows32(const uint32_t (&r)[32], uint32_t us, int rows_valid, const float (&wb)[kDIn], float bb,
                                             float2 (&acc2)[kDIn / 2], float &accb) {
    constexpr int kTileRows = BM - BM % NU;
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
        constexpr int row0 = 32 * LD;
        const int row = row0 + jj;
        if (row < kTileRows && (kFull || row < rows_valid)) {
            const uint32_t up = us + row * (kDIn * 4);
            const float4 u0 = lds128(up), u1 = lds128(up + 16), u2 = lds128(up + 32);
            float pre = bb;                                        // the forward value of this (row, feature), before the ReLU
            pre = fmaf(u0.x, wb[0], pre); pre = fmaf(u0.y, wb[1], pre); pre = fmaf(u0.z, wb[2], pre); pre = fmaf(u0.w, wb[3], pre);
            pre = fmaf(u1.x, wb[4], pre); pre = fmaf(u1.y, wb[5], pre); pre = fmaf(u1.z, wb[6], pre); pre = fmaf(u1.w, wb[7], pre);
            pre = fmaf(u2.x, wb[8], pre); pre = fmaf(u2.y, wb[9], pre); pre = fmaf(u2.z, wb[10], pre); pre = fmaf(u2.w, wb[11], pre);
            const float gval = pre > 0.f ? __uint_as_float(r[jj]) : 0.f;
            const float2 g2 = make_float2(gval, gval);
            acc2[0] = __ffma2_rn(g2, make_float2(u0.x, u0.y), acc2[0]);
            acc2[1] = __ffma2_rn(g2, make_float2(u0.z, u0.w), acc2[1]);
            acc2[2] = __ffma2_rn(g2, make_float2(u1.x, u1.y), acc2[2]);
            acc2[3] = __ffma2_rn(g2, make_float2(u1.z, u1.w), acc2[3]);
            acc2[4] = __ffma2_rn(g2, make_float2(u2.x, u2.y), acc2[4]);
            acc2[5] = __ffma2_rn(g2, make_float2(u2.z, u2.w), acc2[5]);
            accb += gval;
        }
    }
}
template <int NU, int LD>
__device__ __forceinline__ void dgrad_group(uint32_t taddr, uint64_t *acc_empty_bar, bool last, int lane, uint32_t us, int rows_valid,
                                            const float (&wb)[kDIn], float bb, float2 (&acc2)[kDIn / 2], float &accb) {
    constexpr int kTileRows = BM - BM % NU;
    uint32_t r[32];
    tmem_ld32(taddr + 32 * LD, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (last) {                                                    // this warp's share of the accumulator is read: hand it back early
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty_bar);
    }
    if (rows_valid >= kTileRows) dgrad_rows32<NU, LD, true>(r, us, rows_valid, wb, bb, acc2, accb);
    else dgrad_rows32<NU, LD, false>(r, us, rows_valid, wb, bb, acc2, accb);
}

template <int NU>
__global__ void __launch_bounds__(kThreads, 1) unit_dgrad_fused_kernel(const float *__restrict__ dx, const float *__restrict__ dx2, int ld_dx,
                                                                       const uint8_t *__restrict__ argmax,
                                                                       const float *__restrict__ dl, int ld_dl,
                                                                       const float *__restrict__ att,
                                                                       const float *__restrict__ Wt, const float *__restrict__ units,
                                                                       const float *__restrict__ w_b, const float *__restrict__ b_b,
                                                                       int n_tok, float *__restrict__ partial) {
    constexpr int kTileRows = BM - BM % NU, kTileToks = kTileRows / NU;
    constexpr int K = BN, k_chunks = K / BK;                      // the embedding layers are 128 x 128
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + (size_t)kWStages * kWStageBytes);
    uint64_t *full = bars, *empty = bars + kWStages, *acc_full = bars + 2 * kWStages, *acc_empty = bars + 2 * kWStages + 2;
    uint64_t *w_ready = bars + 2 * kWStages + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kWStages + 5);
    uint64_t *u_full = bars + 2 * kWStages + 6, *u_empty = bars + 2 * kWStages + 8;            // units staging (2 buffers)
    float *units_s = reinterpret_cast<float *>(tiles + (size_t)kWStages * kWStageBytes + 256);  // [2][128 * 12]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = n_tok * NU;
    const int m_blocks = (M + kTileRows - 1) / kTileRows, m_first = blockIdx.x, m_step = gridDim.x;
    const int n_my_tiles = m_first < m_blocks ? (m_blocks - m_first + m_step - 1) / m_step : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWStages; ++s) { mbar_init(&full[s], kProducerThreads / 32); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8);
            mbar_init(&u_full[a], 1); mbar_init(&u_empty[a], 8);
        }
        mbar_init(w_ready, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kDMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_w_hi = tmem_base + 2 * kAccCols, tmem_w_lo = tmem_w_hi + 128;

    if (warp < kDMmaWarp) {
        // ===== PRODUCERS =====
        const int t = threadIdx.x & (kProducerThreads - 1), g = warp >> 2;
        if (g == 0) {
            // W_g^T -> TMEM, once: thread = feature j of the basic layer = its TMEM lane
            const float *wrow = Wt + (size_t)((warp & 3) * 32 + lane) * K;
            const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
            for (int k0 = 0; k0 < K; k0 += 8) {
                const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wrow + k0)), w1 = __ldg(reinterpret_cast<const float4 *>(wrow + k0 + 4));
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                float hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { hi[e] = tf32_rna(w[e]); lo[e] = w[e] - hi[e]; }
                tmem_st8(tmem_w_hi + lane_addr + k0, hi);
                tmem_st8(tmem_w_lo + lane_addr + k0, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(w_ready);
        }
        // d_emb chunks: running index c = tile_iter * k_chunks + kc; group g takes c == g (mod kDGroups); the RAW sources of two
        // chunks wait in registers (ping-pong) while earlier chunks are converted and stored
        const uint32_t total_chunks = (uint32_t)n_my_tiles * k_chunks;
        DembRaw<NU> ra, rb;
        auto fetch = [&](uint32_t cc, DembRaw<NU> &raw) {
            if (cc >= total_chunks) return;
            const int ti = (int)(cc / k_chunks), kc = (int)(cc % k_chunks);
            const int tok0 = (m_first + ti * m_step) * kTileToks;
            if (kc == 0 && ti + 1 < n_my_tiles) {
                // the sources of this CTA's NEXT tile -> L2 (they are first touched here; the chunk loads then hit L2, not DRAM)
                const int ntok0 = tok0 + m_step * kTileToks;
                for (int idx = t; idx < kTileToks * 4; idx += kProducerThreads) {           // 512-byte rows: 4 lines per token
                    const int n = min(ntok0 + (idx >> 2), n_tok - 1), line = (idx & 3) * 32;
                    if (dx != nullptr) prefetch_l2(dx + (size_t)n * ld_dx + line);
                    if (dx2 != nullptr) prefetch_l2(dx2 + (size_t)n * ld_dx + line);
                    if (dl != nullptr) prefetch_l2(att + (size_t)n * 128 + line);
                }
                for (int idx = t; idx < kTileToks; idx += kProducerThreads) {
                    const int n = min(ntok0 + idx, n_tok - 1);
                    if (NU > 1 && dx != nullptr) prefetch_l2(argmax + (size_t)n * 128);
                    if (dl != nullptr) prefetch_l2(dl + (size_t)n * ld_dl);
                }
            }
            demb_fetch<NU>(raw, dx, dx2, ld_dx, argmax, dl, ld_dl, att, tok0, n_tok, kc * BK, t);
        };
        auto emit = [&](uint32_t cc, const DembRaw<NU> &raw) {
            const int stage = cc % kWStages;
            mbar_wait(&empty[stage], ((cc / kWStages) & 1) ^ 1);
            unsigned char *st = tiles + (size_t)stage * kWStageBytes;
            demb_store<NU>(raw, st, st + kTileBytes, t);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        };
        fetch(g, ra);
        fetch(g + kDGroups, rb);
        for (uint32_t c = g; c < total_chunks; c += 2 * kDGroups) {
            emit(c, ra);
            fetch(c + 2 * kDGroups, ra);
            if (c + kDGroups < total_chunks) {
                emit(c + kDGroups, rb);
                fetch(c + 3 * kDGroups, rb);
            }
        }
    } else if (warp == kDMmaWarp) {
        // ===== MMA ISSUER =====
        mbar_wait(w_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t c = 0;
        int it = 0;
        for (int mb = m_first; mb < m_blocks; mb += m_step, ++it) {
            const int a = it & 1;
            mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + a * kAccCols;
            for (int kc = 0; kc < k_chunks; ++kc, ++c) {
                const int stage = c % kWStages;
                mbar_wait(&full[stage], (c / kWStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t xbase = smem_u32(tiles + (size_t)stage * kWStageBytes);
                    const uint64_t x_hi = make_desc(xbase), x_lo = make_desc(xbase + kTileBytes);
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);
                        const uint32_t kcol = kc * BK + ks * 8;
                        const uint32_t first = (kc | ks) != 0;
                        umma_tf32_ts(tmem_d, tmem_w_lo + kcol, x_hi + adv, kIdesc, first);
                        umma_tf32_ts(tmem_d, tmem_w_hi + kcol, x_lo + adv, kIdesc, 1u);
                        umma_tf32_ts(tmem_d, tmem_w_hi + kcol, x_hi + adv, kIdesc, 1u);
                    }
                    umma_commit(&empty[stage]);
                    if (kc == k_chunks - 1) umma_commit(&acc_full[a]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== EPILOGUE =====  thread = feature j of the basic layer; half h of the warps takes rows [64h, 64h + 64) of a tile
        const int e = warp - (kDMmaWarp + 1), q4 = warp & 3, half = e >> 2;
        const int col = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const bool stager = e == 0 && lane == 0;                  // the thread that issues the bulk copies of the raw unit features
        auto stage_units = [&](int j) {                            // tile number j of this CTA -> buffer j & 1
            const int m0 = (m_first + j * m_step) * kTileRows;
            const uint32_t bytes = (uint32_t)min(kTileRows, M - m0) * kDIn * 4;   // rows are 48 bytes: always a multiple of 16
            mbar_expect_tx(&u_full[j & 1], bytes);
            bulk_g2s(units_s + (j & 1) * kDUnitsTile, units + (size_t)m0 * kDIn, bytes, &u_full[j & 1]);
        };
        if (stager) {
            if (n_my_tiles > 0) stage_units(0);
            if (n_my_tiles > 1) stage_units(1);
        }
        float wb[kDIn];
#pragma unroll
        for (int k = 0; k < kDIn; ++k) wb[k] = __ldg(w_b + col * kDIn + k);
        const float bb = __ldg(b_b + col);
        float2 acc2[kDIn / 2];                                    // (dW_b[j][2kk], dW_b[j][2kk+1])
#pragma unroll
        for (int kk = 0; kk < kDIn / 2; ++kk) acc2[kk] = make_float2(0.f, 0.f);
        float accb = 0.f;
        for (int it = 0; it < n_my_tiles; ++it) {
            const int a = it & 1;
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            const int rows_valid = M - (m_first + it * m_step) * kTileRows;
            const uint32_t us = smem_u32(units_s + a * kDUnitsTile);   // shared-window address of this tile's raw unit features
            mbar_wait(&u_full[a], ph);                             // this tile's raw unit features have landed
            mbar_wait(&acc_full[a], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(a * kAccCols);
            if (half == 0) {
                dgrad_group<NU, 0>(taddr, &acc_empty[a], false, lane, us, rows_valid, wb, bb, acc2, accb);
                dgrad_group<NU, 1>(taddr, &acc_empty[a], true, lane, us, rows_valid, wb, bb, acc2, accb);
            } else {
                dgrad_group<NU, 2>(taddr, &acc_empty[a], false, lane, us, rows_valid, wb, bb, acc2, accb);
                dgrad_group<NU, 3>(taddr, &acc_empty[a], true, lane, us, rows_valid, wb, bb, acc2, accb);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&u_empty[a]);               // this warp is done with the staged features
            if (stager && it + 2 < n_my_tiles) {                   // refill the buffer once all 8 warps have released it
                mbar_wait(&u_empty[a], ph);
                stage_units(it + 2);
            }
            __syncwarp();
        }
        float *prow = partial + ((size_t)(blockIdx.x * 2 + half) * BN + col) * (kDIn + 1);
#pragma unroll
        for (int kk = 0; kk < kDIn / 2; ++kk) { prow[2 * kk] = acc2[kk].x; prow[2 * kk + 1] = acc2[kk].y; }
        prow[kDIn] = accb;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kDMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// =================================================================================================
// Weight-gradient GEMM:  dW[No, Ni] = dY[T, No]^T * X[T, Ni]   and   db[No] = column sums of dY.
//
// The contraction runs over the TOKEN dimension, so both operands are MN-major in memory (features contiguous).
// tcgen05 takes them as they are (instruction-descriptor a_major = b_major = MN).  For 32-bit operands the only
// MN-major shared-memory layout is SWIZZLE_128B_BASE32B (cute: Layout_MN_SW128_32B_Atom): atoms of 4 tokens x 32
// features (4 rows of 128 B) whose 32-byte chunks are XOR-swizzled with the row index.  A [32 tokens x 128 features]
// chunk is 8 x 4 such atoms -- the natural row-major order, no transposition.  Split-K: the SMs are divided over the
// (No/128 x Ni/128) output tiles and then over interleaved 32-token chunks; every CTA accumulates a 128x128 fp32
// partial in TMEM, writes it to the workspace, and a small second kernel sums the partials in a fixed order
// (deterministic, unlike atomics).  The producers that stream dY also accumulate its column sums, so the bias
// gradient costs no extra pass over dY.
constexpr uint32_t kIdescMN = kIdesc | (1u << 15) | (1u << 16);
constexpr int kAtomBytes = 512;                       // 4 tokens x 32 features x 4 B
constexpr int kKGroupBytes = 4 * kAtomBytes;          // the four 32-feature atoms of one 4-token group
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
    // LBO = stride between 32-feature atoms (512 B); SBO = stride between 4-token groups (2 KB); layout 1 = SW128_BASE32B
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(kAtomBytes >> 4) << 16) | ((uint64_t)(kKGroupBytes >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}

// MN-major operand chunk: tokens [t0, t0+32) x features [f0, f0+128) of a row-major [T, ld] matrix.  Load and
// split/store halves are separate for register double buffering (see tile_load_k).  tile_store_mn optionally adds the
// chunk to this thread's running column sums (4 features).
__device__ __forceinline__ void tile_load_mn(const float *__restrict__ src, RowMap map, int t0, int T, int f0, int t,
                                             float4 (&v)[8]) {
    const int l = t & 31, w = t >> 5;                  // lane -> 4 features, warp -> token (mod 4)
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (map.rpb == 0 || map.rpb >= 4) {
        // this thread's rows are t0+w, t0+w+4, ...: one offset computation, then a walk (+4 rows, at most one block wrap per step)
        const float *p = src + map.off(t0 + w) + f0 + 4 * l;
        const long long step = 4ll * map.ld, wrap = map.rpb > 0 ? map.bs - (long long)map.rpb * map.ld : 0;
        int rem = 0;
        if (map.rpb > 0) { const unsigned n = (unsigned)(t0 + w), tq = __umulhi(n, map.mul); rem = (int)(n - ((tq + ((n - tq) >> 1)) >> map.sh) * (unsigned)map.rpb); }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = (t0 + w + 4 * i < T) ? __ldg(reinterpret_cast<const float4 *>(p)) : zero4;
            p += step;
            if (map.rpb > 0) { rem += 4; if (rem >= map.rpb) { rem -= map.rpb; p += wrap; } }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int tok = w + 4 * i;
        v[i] = (t0 + tok < T) ? __ldg(reinterpret_cast<const float4 *>(src + map.off(t0 + tok) + f0) + l) : zero4;
    }
}
__device__ __forceinline__ void tile_store_mn(const float4 (&v)[8], unsigned char *dst_hi, unsigned char *dst_lo, int t,
                                              float4 *colsum) {
    const int l = t & 31, w = t >> 5;
    const int mi = l >> 3, c32 = (l & 7) >> 1, half = l & 1;     // 32-feature atom, 32-byte chunk, 16-byte half
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int tok = w + 4 * i;
        const int row = tok & 3, kj = tok >> 2;
        const int off = kj * kKGroupBytes + mi * kAtomBytes + row * 128 + ((c32 ^ row) << 5) + (half << 4);
        float4 hi, lo;
        hi.x = tf32_rna(v[i].x); hi.y = tf32_rna(v[i].y); hi.z = tf32_rna(v[i].z); hi.w = tf32_rna(v[i].w);
        lo.x = v[i].x - hi.x; lo.y = v[i].y - hi.y; lo.z = v[i].z - hi.z; lo.w = v[i].w - hi.w;
        *reinterpret_cast<float4 *>(dst_hi + off) = hi;
        *reinterpret_cast<float4 *>(dst_lo + off) = lo;
        if (colsum) { colsum->x += v[i].x; colsum->y += v[i].y; colsum->z += v[i].z; colsum->w += v[i].w; }
    }
}

// ---- weight gradient, dY^T as the A operand in TENSOR MEMORY ------------------------------------------------------------
// ncu on gemm_wgrad_kernel: splitting BOTH operands through shared memory costs 64 KB of stores + 96 KB of tensor-core reads
// per 32-token chunk (LSU shared wavefronts 27 % + tensor-core shared wavefronts 35 % + the global loads on the same L1 data
// path) for 32 KB of HBM data.  Here the dY side never touches shared memory: a producer thread owns ONE output feature o
// (= its TMEM lane), loads dY[t][o] for the chunk's 32 tokens -- a warp load is 32 consecutive features of one token, 128
// coalesced bytes -- splits hi/lo in registers and writes 32 + 32 TMEM columns with tcgen05.st; the MMAs take A from TMEM
// (tcgen05.mma [d], [a], b-desc).  Only X goes through the (now 6-stage) shared-memory ring as the MN-major B operand.
// Roles: warps 0-7 two A-producer groups (quadrant = warp & 3; also the epilogue), warps 8-15 two B-producer groups,
// warp 16 MMA issuer.  TMEM: [0,128) accumulator, [128,384) four A stages of 64 columns.
constexpr int kGAStages = 4, kGBStages = 6;
constexpr int kGMmaWarp = 16;
constexpr int kThreadsG = (kGMmaWarp + 1) * 32;                     // 544
constexpr size_t kSmemBytesWgradA = (size_t)kGBStages * 2 * kTileBytes + 1024 + 256 + 2 * 128 * sizeof(float);
constexpr uint32_t kIdescBMN = kIdesc | (1u << 16);                  // A: TMEM (K-major by construction), B: MN-major shared memory

//
// NU > 0 ("routed"): dY is not read, it is the max-pool routing of a unit-embedding layer GENERATED in the A producer --
// row (token n, unit u), feature o:  (argmax[n,o] == u) ? dY[n*ld + o] (+ dY2[n*ld + o]) : 0  with dY/dY2 the [n_tokens, ld]
// gradients of the group maximum (policy.py:102-127) -- so the dense [n_tokens*NU, 128] gradient never exists in memory.
// A K-chunk then holds whole tokens: kRPC = (32 / NU) * NU rows (30 for the 5-unit group, the last two rows are zero).
template <int NU>
__global__ void __launch_bounds__(kThreadsG, 1) gemm_wgrad_atmem_kernel(const float *__restrict__ dY, RowMap ymap,
                                                                        const float *__restrict__ dY2,
                                                                        const uint8_t *__restrict__ argmax,
                                                                        const float *__restrict__ X, RowMap xmap, int T, int No,
                                                                        int Ni, int nsplit, float *__restrict__ part_w,
                                                                        float *__restrict__ part_b) {
    constexpr int kTPC = NU > 0 ? BK / NU : 0;                  // tokens per K-chunk (routed form)
    constexpr int kRPC = NU > 0 ? kTPC * NU : BK;               // rows per K-chunk
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + (size_t)kGBStages * 2 * kTileBytes);
    uint64_t *a_full = bars, *a_empty = bars + kGAStages, *b_full = bars + 2 * kGAStages, *b_empty = b_full + kGBStages;
    uint64_t *acc_full = b_empty + kGBStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);
    float *colsum_s = reinterpret_cast<float *>(acc_full + 2);              // [2 A groups][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blocks = Ni / BN, tiles_mn = (No / BM) * n_blocks;
    const int tile = blockIdx.x % tiles_mn, split = blockIdx.x / tiles_mn;
    const int m0 = (tile / n_blocks) * BM, n0 = (tile % n_blocks) * BN;
    const int chunks_total = (T + kRPC - 1) / kRPC;
    const int my_chunks = split < chunks_total ? (chunks_total - split + nsplit - 1) / nsplit : 0;   // chunk = split + j*nsplit

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGAStages; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kGBStages; ++s) { mbar_init(&b_full[s], 4); mbar_init(&b_empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kGMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_a0 = tmem_base + kAccCols;

    if (warp < 8) {
        // ===== A PRODUCERS (dY^T -> TMEM) =====  group ga takes chunks j == ga (mod 2); thread = output feature
        const int ga = warp >> 2, q = warp & 3;
        const int o = m0 + q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const bool want_b = part_b != nullptr && n0 == 0;
        float cs = 0.f;
        float va[32], vb[32];
        auto fetch = [&](int j, float (&buf)[32]) {
            if (j >= my_chunks) return;
            if constexpr (NU > 0) {
                // whole tokens tok0 .. tok0+kTPC-1: one gradient value and one arg-max byte per token, expanded to NU rows
                const int tok0 = (split + j * nsplit) * kTPC, n_tok = T / NU;
                float dv[kTPC];
                int am[kTPC];
#pragma unroll
                for (int k = 0; k < kTPC; ++k) {
                    const bool ok = tok0 + k < n_tok;
                    const size_t at = (size_t)(tok0 + k) * ymap.ld + o;
                    dv[k] = ok ? __ldg(dY + at) : 0.f;
                    if (ok && dY2 != nullptr) dv[k] += __ldg(dY2 + at);
                    am[k] = ok ? (int)__ldg(argmax + (size_t)(tok0 + k) * 128 + o) : 255;      // argmax rows are 128 bytes (No == 128)
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) buf[i] = (i < kRPC && am[i / NU] == i % NU) ? dv[i / NU] : 0.f;
            } else {
                const int t0 = (split + j * nsplit) * BK;
                // rows t0 .. t0+31 of dY: offset of the first, then a walk (+1 row, block wrap for two-level maps)
                const float *p = dY + ymap.off(t0) + o;
                int rem = 0;
                if (ymap.rpb > 0) { const unsigned n = (unsigned)t0, tq = __umulhi(n, ymap.mul); rem = (int)(n - ((tq + ((n - tq) >> 1)) >> ymap.sh) * (unsigned)ymap.rpb); }
                const long long wrap = ymap.rpb > 0 ? ymap.bs - (long long)ymap.rpb * ymap.ld : 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    buf[i] = (t0 + i < T) ? __ldg(p) : 0.f;
                    p += ymap.ld;
                    if (ymap.rpb > 0 && ++rem == ymap.rpb) { rem = 0; p += wrap; }
                }
            }
        };
        auto emit = [&](int j, const float (&buf)[32]) {
            const int stage = j % kGAStages;
            mbar_wait(&a_empty[stage], ((j / kGAStages) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tcol = tmem_a0 + stage * 64 + lane_addr;
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
                float hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { hi[e] = tf32_rna(buf[8 * k8 + e]); lo[e] = buf[8 * k8 + e] - hi[e]; cs += buf[8 * k8 + e]; }
                tmem_st8(tcol + 8 * k8, hi);
                tmem_st8(tcol + 32 + 8 * k8, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[stage]);
        };
        fetch(ga, va);
        fetch(ga + 2, vb);
        for (int j = ga; j < my_chunks; j += 4) {
            emit(j, va);
            fetch(j + 4, va);
            if (j + 2 < my_chunks) {
                emit(j + 2, vb);
                fetch(j + 6, vb);
            }
        }
        if (want_b) colsum_s[ga * 128 + q * 32 + lane] = cs;
        if (ga == 0) {
            // ===== EPILOGUE: TMEM partial -> workspace [split][tile][128][128] =====
            float *prow = part_w + (((size_t)split * tiles_mn + tile) * BM + q * 32 + lane) * BN;
            if (my_chunks > 0) {
                mbar_wait(acc_full, 0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += 32) {
                uint32_t r[32];
                if (my_chunks > 0) {
                    tmem_ld32(tmem_base + lane_addr + (uint32_t)cb, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) r[jj] = 0u;
                }
                store_row32(prow + cb, r, nullptr, 0, true);                  // workspace rows are 512-byte aligned
            }
        }
    } else if (warp < kGMmaWarp) {
        // ===== B PRODUCERS (X -> MN-major shared-memory tiles) =====  group gb takes chunks j == gb (mod 2)
        const int gb = (warp - 8) >> 2, t = threadIdx.x & (kProducerThreads - 1);
        float4 v[8], vn[8];
        int j = gb;
        auto load_x = [&](int jj, float4 (&buf)[8]) {              // rows of chunk jj; past its kRPC rows (or T) read as zero
            const int t0 = (split + jj * nsplit) * kRPC;
            tile_load_mn(X, xmap, t0, min(T, t0 + kRPC), n0, t, buf);
        };
        if (j < my_chunks) load_x(j, v);
        for (; j < my_chunks; j += 2) {
            if (j + 2 < my_chunks) load_x(j + 2, vn);
            const int stage = j % kGBStages;
            mbar_wait(&b_empty[stage], ((j / kGBStages) & 1) ^ 1);
            unsigned char *st = tiles + (size_t)stage * 2 * kTileBytes;
            tile_store_mn(v, st, st + kTileBytes, t, nullptr);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[stage]);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = vn[e];
        }
    } else {
        // ===== MMA ISSUER =====
        for (int j = 0; j < my_chunks; ++j) {
            const int sa = j % kGAStages, sb = j % kGBStages;
            mbar_wait(&a_full[sa], (j / kGAStages) & 1);
            mbar_wait(&b_full[sb], (j / kGBStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t base = smem_u32(tiles + (size_t)sb * 2 * kTileBytes);
                const uint32_t a_hi0 = tmem_a0 + sa * 64, a_lo0 = a_hi0 + 32;
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {                          // 8 tokens = two 4-token groups per MMA
                    const uint32_t adv = ks * 2 * kKGroupBytes;
                    const uint64_t b_hi = make_desc_mn(base + adv), b_lo = make_desc_mn(base + kTileBytes + adv);
                    const uint32_t first = (j | ks) != 0;
                    umma_tf32_ts(tmem_base, a_lo0 + 8 * ks, b_hi, kIdescBMN, first);
                    umma_tf32_ts(tmem_base, a_hi0 + 8 * ks, b_lo, kIdescBMN, 1u);
                    umma_tf32_ts(tmem_base, a_hi0 + 8 * ks, b_hi, kIdescBMN, 1u);
                }
                umma_commit(&a_empty[sa]);
                umma_commit(&b_empty[sb]);
                if (j == my_chunks - 1) umma_commit(acc_full);
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (part_b != nullptr && n0 == 0 && threadIdx.x < 128)           // column sums of this CTA's share of dY (fixed order)
        part_b[(size_t)split * No + m0 + threadIdx.x] = colsum_s[threadIdx.x] + colsum_s[128 + threadIdx.x];
    if (warp == kGMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// dW[o][i] (+)= sum_split part_w[split][tile][o%128][i%128];  db[o] (+)= sum_split part_b[split][o].
// One float4 of [dW | db] per thread column, the splits dealt to 8 thread rows (s = y, y+8, ...) whose partial sums are
// then added in row order: a fixed summation tree (deterministic) with 8x the loads in flight of a serial loop.
constexpr int kRedRows = 8;
__global__ void __launch_bounds__(32 * kRedRows) wgrad_reduce_kernel(const float *__restrict__ part_w, const float *__restrict__ part_b,
                                                                     int nsplit, int No, int Ni, float *__restrict__ dW, int ldw,
                                                                     float *__restrict__ db, int accumulate) {
    __shared__ float4 sh[kRedRows][32];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int n_blocks = Ni / BN, tiles_mn = (No / BM) * n_blocks;
    const int total4 = No * Ni / 4, bias4 = db != nullptr ? No / 4 : 0;
    const int idx = blockIdx.x * 32 + x;
    const float *src = nullptr;
    float *dst = nullptr;
    size_t stride = 0;
    if (idx < total4) {
        const int o = (idx * 4) / Ni, i = (idx * 4) % Ni;
        const int tile = (o / BM) * n_blocks + i / BN;
        src = part_w + ((size_t)tile * BM + o % BM) * BN + i % BN;
        stride = (size_t)tiles_mn * BM * BN;
        dst = dW + (size_t)o * ldw + i;
    } else if (idx < total4 + bias4) {
        src = part_b + (size_t)(idx - total4) * 4;
        stride = (size_t)No;
        dst = db + (size_t)(idx - total4) * 4;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src != nullptr) {
#pragma unroll 4
        for (int s = y; s < nsplit; s += kRedRows) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src + (size_t)s * stride));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    sh[y][x] = acc;
    __syncthreads();
    if (y == 0 && dst != nullptr) {
        float4 t = accumulate ? *reinterpret_cast<const float4 *>(dst) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kRedRows; ++r) { t.x += sh[r][x].x; t.y += sh[r][x].y; t.z += sh[r][x].z; t.w += sh[r][x].w; }
        *reinterpret_cast<float4 *>(dst) = t;
    }
}

}  // namespace

extern "C" int dc_gemm_tf32x3_supported(int64_t M, int N, int K) { return M > 0 && N > 0 && K > 0 && N % BN == 0 && K % BK == 0; }

// Row map for (ld, rows_per_block, block_stride); one-row blocks are plain rows `block_stride` apart.
static RowMap make_rowmap(int ld, int64_t rpb, int64_t bs) {
    RowMap m{ld, 0, 0, 0, 0};
    if (rpb <= 0) return m;
    if (rpb == 1) { m.ld = (int)bs; return m; }                    // rowmap_ok: bs < 2^31 in this case
    m.rpb = (int)rpb;
    m.bs = (long long)bs;
    int s = 0;
    while ((1ull << s) < (unsigned long long)rpb) ++s;            // ceil(log2 rpb) >= 1
    m.mul = (unsigned)((((unsigned long long)1 << 32) * (((unsigned long long)1 << s) - (unsigned long long)rpb)) / (unsigned long long)rpb + 1);
    m.sh = s - 1;
    return m;
}

static int gemm_impl(const float *A, RowMap amap, const float *B, int ldb, const float *bias, float *C, RowMap cmap,
                     int64_t M, int N, int K, int relu, dc_stream_t stream) {
    const int lda = amap.ld, ldc = cmap.ld;
    DC_REQUIRE(A && B && C, DC_EINVAL, "dc_gemm_tf32x3: null pointer");
    DC_REQUIRE(dc_gemm_tf32x3_supported(M, N, K) && M < (1ll << 31) - BM, DC_EUNSUPPORTED,
               "dc_gemm_tf32x3: need N %% 128 == 0 and K %% 32 == 0 (M=%lld N=%d K=%d)", (long long)M, N, K);
    DC_REQUIRE(lda >= K && ldb >= K && ldc >= N && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, DC_EINVAL,
               "dc_gemm_tf32x3: leading dimensions must be >= extent and multiples of 4 floats");
    DC_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0),
               DC_EINVAL, "dc_gemm_tf32x3: pointers must be 16-byte aligned");
    // 256-bit epilogue stores need 32-byte aligned rows
    const bool wide = ((uintptr_t)C & 31) == 0 && ldc % 8 == 0 && (cmap.rpb == 0 || cmap.bs % 8 == 0) &&
                      (!bias || ((uintptr_t)bias & 31) == 0);
    const int tiles = (int)((M + BM - 1) / BM) * (N / BN);
    const int n_blocks = N / BN, m_blocks = (int)((M + BM - 1) / BM);
    // the shared-memory opt-in is a per-device attribute: set it on every call (cheap) rather than caching it per process
    if (K / BK <= kMaxResChunks && n_blocks <= dc_sm_count()) {          // K <= 128: weights resident in tensor memory
        int per_col = dc_sm_count() / n_blocks;                          // CTAs per column block
        if (per_col > m_blocks) per_col = m_blocks;
        DC_CUDA(cudaFuncSetAttribute(gemm_tf32x3_wtmem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesWTmem));
        gemm_tf32x3_wtmem_kernel<0><<<per_col * n_blocks, kThreads, kSmemBytesWTmem, dc_cu_stream(stream)>>>(A, amap, B, ldb, bias, C, cmap,
                                                                                                         (int)M, N, K, relu, nullptr, nullptr);
        DC_LAUNCH_OK();
        return DC_OK;
    }
    DC_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    const int grid = tiles < dc_sm_count() ? tiles : dc_sm_count();
    gemm_tf32x3_kernel<<<grid, kThreads, kSmemBytes, dc_cu_stream(stream)>>>(A, amap, B, ldb, bias, C, cmap, (int)M, N, K, relu, wide, 1, 0);
    DC_LAUNCH_OK();
    return DC_OK;
}

// Split-K form for the skinny products of the step-wise recurrence (csrc/rnn_stepwise.cuh): part[sp] = A[:, K range sp] *
// B[:, K range sp]^T for sp < ksplit, each [M, N] with leading dimension N, `ksplit` chosen by the caller so that
// tiles x ksplit fills the SMs.  Library-internal (C++ linkage).
int dc_gemm_tf32x3_splitk(const float *A, int lda, const float *B, int ldb, float *part, int64_t M, int N, int K, int ksplit,
                          bool first_call, cudaStream_t st) {
    DC_REQUIRE(A && B && part && ksplit >= 1 && N % BN == 0 && K % (BK * ksplit) == 0 && M > 0, DC_EINVAL,
               "dc_gemm_tf32x3_splitk: bad arguments (M=%lld N=%d K=%d ksplit=%d)", (long long)M, N, K, ksplit);
    if (first_call)                                   // per-device attribute; the step loop calls this thousands of times
        DC_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    const int tiles = (int)((M + BM - 1) / BM) * (N / BN) * ksplit;
    const int grid = tiles < dc_sm_count() ? tiles : dc_sm_count();
    const bool wide = ((uintptr_t)part & 31) == 0 && N % 8 == 0 && ((size_t)M * N) % 8 == 0;
    gemm_tf32x3_kernel<<<grid, kThreads, kSmemBytes, st>>>(A, make_rowmap(lda, 0, 0), B, ldb, nullptr, part, make_rowmap(N, 0, 0), (int)M,
                                                            N, K, 0, wide, ksplit, (long long)M * N);
    DC_LAUNCH_OK();
    return DC_OK;
}

static int wgrad_splits(int No, int Ni) {
    const int tiles_mn = (No / BM) * (Ni / BN);
    const int n = dc_sm_count() / tiles_mn;
    return n < 1 ? 1 : n;
}

extern "C" size_t dc_gemm_wgrad_workspace_bytes(int No, int Ni) {
    if (No <= 0 || Ni <= 0 || No % BM || Ni % BN) return 0;
    const int nsplit = wgrad_splits(No, Ni);
    return ((size_t)nsplit * No * Ni + (size_t)nsplit * No) * sizeof(float);
}

static int wgrad_impl(const float *dY, RowMap ymap, const float *X, RowMap xmap, int64_t T, int No, int Ni, float *dW,
                      int ldw, float *db, int accumulate, void *workspace, dc_stream_t stream) {
    const int ldy = ymap.ld, ldx = xmap.ld;
    DC_REQUIRE(dY && X && dW && workspace, DC_EINVAL, "dc_gemm_wgrad_tf32x3: null pointer");
    DC_REQUIRE(T > 0 && T < (1ll << 31) - BK && No > 0 && Ni > 0 && No % BM == 0 && Ni % BN == 0, DC_EUNSUPPORTED,
               "dc_gemm_wgrad_tf32x3: need No %% 128 == 0 and Ni %% 128 == 0 (T=%lld No=%d Ni=%d)", (long long)T, No, Ni);
    DC_REQUIRE(ldy >= No && ldx >= Ni && ldw >= Ni && ldy % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0, DC_EINVAL,
               "dc_gemm_wgrad_tf32x3: bad leading dimension");
    DC_REQUIRE(((uintptr_t)dY & 15) == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)dW & 15) == 0 && ((uintptr_t)workspace & 15) == 0 &&
                   ((uintptr_t)db & 15) == 0,
               DC_EINVAL, "dc_gemm_wgrad_tf32x3: pointers must be 16-byte aligned");
    const int tiles_mn = (No / BM) * (Ni / BN), nsplit = wgrad_splits(No, Ni);
    float *part_w = reinterpret_cast<float *>(workspace);
    float *part_b = db ? part_w + (size_t)nsplit * No * Ni : nullptr;
    cudaStream_t st = dc_cu_stream(stream);
    DC_CUDA(cudaFuncSetAttribute(gemm_wgrad_atmem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesWgradA));
    gemm_wgrad_atmem_kernel<0><<<tiles_mn * nsplit, kThreadsG, kSmemBytesWgradA, st>>>(dY, ymap, nullptr, nullptr, X, xmap, (int)T, No, Ni,
                                                                                      nsplit, part_w, part_b);
    DC_LAUNCH_OK();
    const int total4 = No * Ni / 4 + (db ? No / 4 : 0);
    wgrad_reduce_kernel<<<(total4 + 31) / 32, 32 * kRedRows, 0, st>>>(part_w, part_b, nsplit, No, Ni, dW, ldw, db, accumulate);
    DC_LAUNCH_OK();
    return DC_OK;
}

static bool rowmap_ok(int64_t rpb, int64_t bs, int ld) {
    return rpb == 0 || (rpb > 0 && rpb < (1 << 30) && bs >= rpb * (int64_t)ld && bs % 4 == 0 && (rpb > 1 || bs < (1ll << 31)));
}

extern "C" int dc_gemm_tf32x3(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc,
                              int64_t M, int N, int K, int relu, dc_stream_t stream) {
    return gemm_impl(A, make_rowmap(lda, 0, 0), B, ldb, bias, C, make_rowmap(ldc, 0, 0), M, N, K, relu, stream);
}

// Same GEMM with two-level row addressing of A and/or C (rows_per_block = 0 selects the plain form).
extern "C" int dc_gemm_tf32x3_blocked(const float *A, int lda, int64_t a_rows_per_block, int64_t a_block_stride,
                                      const float *B, int ldb, const float *bias, float *C, int ldc,
                                      int64_t c_rows_per_block, int64_t c_block_stride, int64_t M, int N, int K, int relu,
                                      dc_stream_t stream) {
    DC_REQUIRE(rowmap_ok(a_rows_per_block, a_block_stride, lda) && rowmap_ok(c_rows_per_block, c_block_stride, ldc), DC_EINVAL,
               "dc_gemm_tf32x3_blocked: bad block addressing");
    return gemm_impl(A, make_rowmap(lda, a_rows_per_block, a_block_stride), B, ldb, bias, C,
                     make_rowmap(ldc, c_rows_per_block, c_block_stride), M, N, K, relu, stream);
}

// Unit-embedding layer with the max-pool fused into the epilogue (policy.py:101-127): basic [N*n_units, 128] x W[128,128]^T
// -> xmax[n*ld_x + c] = max_u(emb[n,u,c]) + b[c] (and xmax_copy), argmax[n*128 + c]; the [N*n_units, 128] embedding is never stored.
extern "C" int dc_gemm_unit_max(const float *basic, const float *w, const float *bias, float *xmax, float *xmax_copy, int ld_x,
                                uint8_t *argmax, int64_t n_tokens, int n_units, dc_stream_t stream) {
    DC_REQUIRE(basic && w && bias && xmax && argmax && n_tokens > 0, DC_EINVAL, "dc_gemm_unit_max: null pointer / empty input");
    DC_REQUIRE(n_units == 5 || n_units == 16, DC_EUNSUPPORTED, "dc_gemm_unit_max: n_units=%d (5 or 16; 1-unit groups are plain GEMMs)", n_units);
    DC_REQUIRE(ld_x >= BN && n_tokens * n_units < (1ll << 31) - BM, DC_EINVAL, "dc_gemm_unit_max: bad ld_x / size");
    DC_REQUIRE(((uintptr_t)basic & 15) == 0 && ((uintptr_t)w & 15) == 0, DC_EINVAL, "dc_gemm_unit_max: pointers must be 16-byte aligned");
    const int M = (int)(n_tokens * n_units);
    const int tile_rows = BM - BM % n_units;
    int grid = (M + tile_rows - 1) / tile_rows;
    if (grid > dc_sm_count()) grid = dc_sm_count();
    const RowMap amap = make_rowmap(BN, 0, 0), cmap = make_rowmap(ld_x, 0, 0);
    cudaStream_t st = dc_cu_stream(stream);
    if (n_units == 5) {
        DC_CUDA(cudaFuncSetAttribute(gemm_tf32x3_wtmem_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesWTmem));
        gemm_tf32x3_wtmem_kernel<5><<<grid, kThreads, kSmemBytesWTmem, st>>>(basic, amap, w, BN, bias, xmax, cmap, M, BN, BN, 0, xmax_copy, argmax);
    } else {
        DC_CUDA(cudaFuncSetAttribute(gemm_tf32x3_wtmem_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesWTmem));
        gemm_tf32x3_wtmem_kernel<16><<<grid, kThreads, kSmemBytesWTmem, st>>>(basic, amap, w, BN, bias, xmax, cmap, M, BN, BN, 0, xmax_copy, argmax);
    }
    DC_LAUNCH_OK();
    return DC_OK;
}

// Backward data path of one unit-embedding layer, fused (unit_dgrad_fused_kernel): dW_b[128,12] (+)= G^T units, db_b (+)= colsum G,
//   G[(n,u), j] = relu'(basic[(n,u), j]) * sum_c ( R[(n,u), c] + dlogits[n*ld_dl + u] * att[n*128 + c] ) W_g[c, j],
// R = the max-pool routing of (d_xmax (+ d_xmax2), argmax) (d_xmax NULL: no routing, e.g. the enemy-tower layer, policy.py:127);
// dlogits / att NULL: the target-unit head was not used.  w_t = W_g^T [128,128].  Workspace: dc_unit_basic_bwd_workspace_bytes().
extern "C" int dc_unit_dgrad_fused(const float *d_xmax, const float *d_xmax2, int ld_dx, const uint8_t *argmax, const float *dlogits,
                                   int ld_dl, const float *att, const float *w_t, const float *units, const float *w_b,
                                   const float *b_b, int64_t n_tokens, int n_units, float *dw_b, float *db_b, int accumulate,
                                   void *workspace, dc_stream_t stream) {
    DC_REQUIRE(w_t && units && w_b && b_b && dw_b && db_b && workspace && n_tokens > 0, DC_EINVAL, "dc_unit_dgrad_fused: null pointer / empty input");
    DC_REQUIRE(n_units == 1 || n_units == 5 || n_units == 16, DC_EUNSUPPORTED, "dc_unit_dgrad_fused: n_units=%d (1, 5 or 16)", n_units);
    DC_REQUIRE((d_xmax == nullptr && d_xmax2 == nullptr) || (d_xmax != nullptr && ld_dx >= BN && ld_dx % 4 == 0), DC_EINVAL,
               "dc_unit_dgrad_fused: bad d_xmax / ld_dx");
    DC_REQUIRE(d_xmax == nullptr || n_units == 1 || argmax != nullptr, DC_EINVAL, "dc_unit_dgrad_fused: routing needs the arg-max");
    DC_REQUIRE((dlogits == nullptr) == (att == nullptr) && (dlogits == nullptr || ld_dl >= n_units), DC_EINVAL,
               "dc_unit_dgrad_fused: dlogits and att come together");
    DC_REQUIRE(n_tokens * n_units < (1ll << 31) - BM, DC_EINVAL, "dc_unit_dgrad_fused: too many rows");
    DC_REQUIRE(((uintptr_t)d_xmax & 15) == 0 && ((uintptr_t)d_xmax2 & 15) == 0 && ((uintptr_t)argmax & 3) == 0 && ((uintptr_t)att & 15) == 0 &&
                   ((uintptr_t)w_t & 15) == 0 && ((uintptr_t)units & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
               DC_EINVAL, "dc_unit_dgrad_fused: d_xmax, att, w_t, units must be 16-byte aligned (arg-max 4-byte)");
    const int n_tok = (int)n_tokens;
    const int tile_rows = BM - BM % n_units;
    int grid = (int)((n_tokens * n_units + tile_rows - 1) / tile_rows);
    if (grid > dc_sm_count()) grid = dc_sm_count();
    float *partial = reinterpret_cast<float *>(workspace);
    cudaStream_t st = dc_cu_stream(stream);
#define DC_LAUNCH_DGRAD(NU)                                                                                                          \
    do {                                                                                                                              \
        DC_CUDA(cudaFuncSetAttribute(unit_dgrad_fused_kernel<NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesDgrad)); \
        unit_dgrad_fused_kernel<NU><<<grid, kThreads, kSmemBytesDgrad, st>>>(d_xmax, d_xmax2, ld_dx, argmax, dlogits, ld_dl, att, w_t,    \
                                                                            units, w_b, b_b, n_tok, partial);                        \
    } while (0)
    if (n_units == 1) DC_LAUNCH_DGRAD(1);
    else if (n_units == 5) DC_LAUNCH_DGRAD(5);
    else DC_LAUNCH_DGRAD(16);
#undef DC_LAUNCH_DGRAD
    DC_LAUNCH_OK();
    return dc_unit_basic_reduce(partial, 2 * grid, dw_b, db_b, accumulate, st);     // two epilogue halves per CTA
}

extern "C" int dc_gemm_wgrad_tf32x3(const float *dY, int ldy, const float *X, int ldx, int64_t T, int No, int Ni, float *dW,
                                    int ldw, float *db, int accumulate, void *workspace, dc_stream_t stream) {
    return wgrad_impl(dY, make_rowmap(ldy, 0, 0), X, make_rowmap(ldx, 0, 0), T, No, Ni, dW, ldw, db, accumulate, workspace, stream);
}

// Weight gradient of one unit-embedding layer from the gradient of its max-pool (policy.py:101-127), without the dense
// [n_tokens*n_units, 128] gradient of the embedding: dW[128,128] = R^T basic, db[128] = column sums of R, where
// R[(n,u), c] = (argmax[n,c] == u) ? d_xmax[n*ld_dx + c] (+ d_xmax2[n*ld_dx + c]) : 0 is generated inside the kernel.
extern "C" int dc_unit_wgrad_routed(const float *d_xmax, const float *d_xmax2, int ld_dx, const uint8_t *argmax, const float *basic,
                                    int64_t n_tokens, int n_units, float *dW, float *db, void *workspace, dc_stream_t stream) {
    DC_REQUIRE(d_xmax && argmax && basic && dW && db && workspace && n_tokens > 0, DC_EINVAL, "dc_unit_wgrad_routed: null pointer / empty input");
    DC_REQUIRE(n_units == 5 || n_units == 16, DC_EUNSUPPORTED, "dc_unit_wgrad_routed: n_units=%d (5 or 16; a 1-unit group is dc_gemm_wgrad_tf32x3)", n_units);
    DC_REQUIRE(ld_dx >= BM && n_tokens * n_units < (1ll << 31) - BK, DC_EINVAL, "dc_unit_wgrad_routed: bad ld_dx / size");
    DC_REQUIRE(((uintptr_t)basic & 15) == 0 && ((uintptr_t)dW & 15) == 0 && ((uintptr_t)db & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
               DC_EINVAL, "dc_unit_wgrad_routed: pointers must be 16-byte aligned");
    const int No = BM, Ni = BN, nsplit = wgrad_splits(No, Ni);
    const int T = (int)(n_tokens * n_units);
    float *part_w = reinterpret_cast<float *>(workspace);
    float *part_b = part_w + (size_t)nsplit * No * Ni;
    const RowMap ymap = make_rowmap(ld_dx, 0, 0), xmap = make_rowmap(BN, 0, 0);
    cudaStream_t st = dc_cu_stream(stream);
    if (n_units == 5) {
        DC_CUDA(cudaFuncSetAttribute(gemm_wgrad_atmem_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesWgradA));
        gemm_wgrad_atmem_kernel<5><<<nsplit, kThreadsG, kSmemBytesWgradA, st>>>(d_xmax, ymap, d_xmax2, argmax, basic, xmap, T, No, Ni, nsplit,
                                                                                part_w, part_b);
    } else {
        DC_CUDA(cudaFuncSetAttribute(gemm_wgrad_atmem_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesWgradA));
        gemm_wgrad_atmem_kernel<16><<<nsplit, kThreadsG, kSmemBytesWgradA, st>>>(d_xmax, ymap, d_xmax2, argmax, basic, xmap, T, No, Ni, nsplit,
                                                                                 part_w, part_b);
    }
    DC_LAUNCH_OK();
    wgrad_reduce_kernel<<<(No * Ni / 4 + No / 4 + 31) / 32, 32 * kRedRows, 0, st>>>(part_w, part_b, nsplit, No, Ni, dW, Ni, db, 0);
    DC_LAUNCH_OK();
    return DC_OK;
}

extern "C" int dc_gemm_wgrad_tf32x3_blocked(const float *dY, int ldy, int64_t y_rows_per_block, int64_t y_block_stride,
                                            const float *X, int ldx, int64_t T, int No, int Ni, float *dW, int ldw, float *db,
                                            int accumulate, void *workspace, dc_stream_t stream) {
    DC_REQUIRE(rowmap_ok(y_rows_per_block, y_block_stride, ldy), DC_EINVAL, "dc_gemm_wgrad_tf32x3_blocked: bad block addressing");
    return wgrad_impl(dY, make_rowmap(ldy, y_rows_per_block, y_block_stride), X, make_rowmap(ldx, 0, 0), T, No, Ni, dW, ldw,
                      db, accumulate, workspace, stream);
}
