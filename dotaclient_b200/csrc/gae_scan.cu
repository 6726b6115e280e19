// GAE-lambda advantages and rewards-to-go as a warp-shuffle segmented reverse scan.
//
// Replaces optimizer.py:53-64 (discount / advantage_returns), the per-step reward reduction
// np.sum(s_rewards, axis=1) (optimizer.py:397) and the zero bootstrap (optimizer.py:417-420).
//
// One warp per rollout (segment).  The warp walks the rollout backwards in 32-row tiles aligned
// to the END of the segment: loads are coalesced, the in-tile recurrence y_t = x_t + c*y_{t+1}
// is a 5-step Kogge-Stone scan over shuffles (multiplier c^d squared each step), and the carry
// of the later tile enters as c^(32-lane) * carry.  Accumulation is float64 and rounded once to
// fp32, like scipy.signal.lfilter (float64) + astype(float32); the TD residuals are formed in
// fp32 with the three roundings numpy applies (no FMA contraction).
//
// HBM traffic: (4*n_sub + 4) read + 8 written bytes per row -- 16 B/row with pre-summed rewards.
#include "dc_common.cuh"

namespace {

// numpy's pairwise float32 add-reduce over a contiguous axis for n < 128 (8 accumulators,
// combined as ((0+1)+(2+3))+((4+5)+(6+7)), remainder added sequentially).
__device__ __forceinline__ float np_sum_row(const float *__restrict__ p, int n) {
    if (n == 1) return p[0];
    if (n < 8) {
        float s = p[0];
        for (int i = 1; i < n; ++i) s = __fadd_rn(s, p[i]);
        return s;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = p[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], p[i + j]);
    }
    float s = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) s = __fadd_rn(s, p[i]);
    return s;
}

__global__ void __launch_bounds__(128) gae_scan_kernel(const float *__restrict__ rewards, int n_sub,
                                                        const float *__restrict__ values,
                                                        const int64_t *__restrict__ seg_off, int n_seg,
                                                        const float *__restrict__ boot_value,
                                                        const float *__restrict__ boot_reward, double gamma,
                                                        double lam, float *__restrict__ adv,
                                                        float *__restrict__ ret) {
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (seg >= n_seg) return;
    const int64_t lo = seg_off[seg], hi = seg_off[seg + 1];
    if (hi <= lo) return;
    const float gf = (float)gamma;
    const double ca = gamma * lam, cr = gamma;
    // c^(32-lane): weight of the incoming carry for this lane.
    double pa = 1.0, pr = 1.0;
    for (int i = 0; i < 32 - lane; ++i) { pa *= ca; pr *= cr; }
    const float boot = boot_value ? boot_value[seg] : 0.0f;
    // discount(rewards)[:-1] starts from the trailing reward element (optimizer.py:63, :419-420: 0)
    double carry_a = 0.0, carry_r = boot_reward ? (double)boot_reward[seg] : 0.0;
    float v_after = boot;  // value of the row following the current tile
    for (int64_t end = hi; end > lo; end -= 32) {
        const int64_t row = end - 32 + lane;
        const bool ok = row >= lo;
        float v = 0.f, r = 0.f;
        if (ok) {
            v = values[row];
            r = np_sum_row(rewards + row * (int64_t)n_sub, n_sub);
        }
        float v_next = __shfl_down_sync(0xffffffffu, v, 1);
        if (lane == 31) v_next = v_after;
        // deltas = rewards[:-1] + gamma*values[1:] - values[:-1], three fp32 roundings (optimizer.py:60)
        const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(gf, v_next)), v);
        double a = (double)delta, q = (double)r;
        double ma = ca, mr = cr;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double ua = __shfl_down_sync(0xffffffffu, a, d);
            const double uq = __shfl_down_sync(0xffffffffu, q, d);
            if (lane + d < 32) { a += ma * ua; q += mr * uq; }
            ma *= ma; mr *= mr;
        }
        a += pa * carry_a;
        q += pr * carry_r;
        if (ok) { adv[row] = (float)a; ret[row] = (float)q; }
        carry_a = __shfl_sync(0xffffffffu, a, 0);
        carry_r = __shfl_sync(0xffffffffu, q, 0);
        v_after = __shfl_sync(0xffffffffu, v, 0);
    }
}

}  // namespace

extern "C" int dc_gae_scan(const float *rewards, int n_sub, const float *values, const int64_t *seg_off,
                           int n_seg, const float *boot_value, const float *boot_reward, double gamma, double lam,
                           float *adv, float *ret, dc_stream_t stream) {
    DC_REQUIRE(n_seg >= 0 && n_sub >= 1 && n_sub < 128, DC_EINVAL, "dc_gae_scan: n_seg=%d n_sub=%d", n_seg, n_sub);
    if (n_seg == 0) return DC_OK;
    DC_REQUIRE(rewards && values && seg_off && adv && ret, DC_EINVAL, "dc_gae_scan: null pointer");
    const int warps = 4;
    gae_scan_kernel<<<(n_seg + warps - 1) / warps, warps * 32, 0, dc_cu_stream(stream)>>>(
        rewards, n_sub, values, seg_off, n_seg, boot_value, boot_reward, gamma, lam, adv, ret);
    DC_LAUNCH_OK();
    return DC_OK;
}
