/* Plain-C restatement of the reference GAE / returns computation.
 *
 * TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Never linked into the product.
 *
 * Follows optimizer.py:53-64 (TimZaman/dotaclient @ 8615b90):
 *   discount(x, g)          = lfilter([1], [1, -g], x[::-1])[::-1].astype(float32)      (:53-54)
 *   deltas                  = rewards[:-1] + gamma * values[1:] - values[:-1]   (fp32)     (:60)
 *   advantages              = discount(deltas, gamma*lam)                                 (:61)
 *   returns                 = discount(rewards, gamma)[:-1]                               (:63)
 * scipy's lfilter promotes the fp32 input to float64 (its coefficients are Python floats),
 * runs the direct-form-II-transposed recurrence y[n] = x[n] + g*y[n-1] in float64 and the
 * reference casts the result back to fp32.  rewards/values carry n+1 elements (the trailing
 * bootstrap entry, optimizer.py:417-420).
 *
 * numpy (NEP 50) evaluates the deltas in fp32 with gamma rounded to fp32:
 *   t = (float)gamma * v[i+1];  t = r[i] + t;  t = t - v[i]   -- each step rounded to fp32.
 */
#include <stddef.h>

void gae_ref(const float *rewards, const float *values, size_t n, double gamma, double lam,
             float *adv_out, float *ret_out)
{
    const float gf = (float)gamma;
    const double gl = gamma * lam;
    double a = 0.0;
    /* returns scan starts from the bootstrap element rewards[n] (discount(rewards)[:-1]) */
    double q = (double)rewards[n];
    for (size_t k = n; k-- > 0;) {
        volatile float t = gf * values[k + 1];
        t = rewards[k] + t;
        t = t - values[k];
        a = (double)t + gl * a;
        q = (double)rewards[k] + gamma * q;
        adv_out[k] = (float)a;
        ret_out[k] = (float)q;
    }
}
