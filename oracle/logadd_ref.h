/*
 * oracle/logadd_ref.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Deterministic fp32 log-add-exp: the log-domain counterpart of the reference's fp32 `+` when
 * equal paths are merged (atomicAdd at CTCBeamSearch.cu:488).  The reference has no log mode
 * (it underflows after ~30 frames); this function is the contract that replaces
 * it, written ONLY with IEEE-754 correctly rounded operations (add, mul, fma, div, rint) so a
 * CPU and a GPU evaluate it to the same bits.  Spec (DESIGN.md "log-add-exp"):
 *
 *   mx = max(a,b), mn = min(a,b);  mn == -inf  -> mx;   d = mn - mx;   d < -17.5 -> mx
 *   n = rint(d * log2e);  r = fma(n, -C1, d);  r = fma(n, -C2, r)          (Cephes ln2 split)
 *   e = (1 + r + r^2 * P5(r)) * 2^n                                        (Cephes expf poly)
 *   t = e / (2 + e);  w = t*t;  l = 2t * (1 + w/3 + w^2/5 + ... + w^6/13)  (= log1p(e))
 *   result = mx + l
 *
 * Build with -ffp-contract=off so the compiler never fuses the plain products below.
 */
#ifndef ORACLE_LOGADD_REF_H
#define ORACLE_LOGADD_REF_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float oracle_logaddexp(float a, float b) {
    float mx = a > b ? a : b;
    float mn = a > b ? b : a;
    if (mn == -INFINITY) return mx;
    float d = mn - mx;
    if (d < -17.5f) return mx;
    float n = rintf(d * 1.44269504088896341f);
    float r = fmaf(n, -0.693359375f, d);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float ex = fmaf(p, r2, r) + 1.0f;
    uint32_t sb = (uint32_t)((int32_t)n + 127) << 23;
    float scale;
    memcpy(&scale, &sb, 4);
    float e = ex * scale;
    float t = e / (2.0f + e);
    float w = t * t;
    float q = 1.0f / 13.0f;
    q = fmaf(q, w, 1.0f / 11.0f);
    q = fmaf(q, w, 1.0f / 9.0f);
    q = fmaf(q, w, 1.0f / 7.0f);
    q = fmaf(q, w, 1.0f / 5.0f);
    q = fmaf(q, w, 1.0f / 3.0f);
    q = fmaf(q, w, 1.0f);
    float l = (2.0f * t) * q;
    return mx + l;
}
#endif
