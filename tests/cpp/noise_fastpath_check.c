/* noise_fastpath_check.c -- error bound of the seeded Noise kernel's fast path (toucan_b200/csrc/tc_pointwise.cu:
 * fast_noise_scale), swept on the CPU.
 *
 * The kernel evaluates M = sqrt(-2 ln(r2) / r2) with Newton-refined reciprocal / rsqrt seeds and fused multiply-adds and
 * lets the exact IEEE chain (the oracle's sequence) decide whenever the fast result lies within 1024 double ulps of a
 * float rounding boundary.  That is bit-exact iff the fast result is always within 1024 ulps of the exact one.  This
 * program restates the fast path operation for operation (C99 fma == the device's fma.rn.f64; the SFU seeds are modelled
 * as the correctly rounded float value moved by -2 .. +2 float ulps: Newton's iteration forgets the seed), sweeps float
 * r2 values in (0, 1) and prints the largest distance in double ulps.
 *
 *   noise_fastpath_check <stride>     every stride-th float (1 = all 2^30-ish values, ~2 min on 8 cores)
 *
 * Test infrastructure: links oracle/_build/liboracle.so (o_det_log). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

double o_det_log(double x);

static float nudge(float v, int ulps)
{
    uint32_t b;
    memcpy(&b, &v, 4);
    b += (uint32_t)ulps;
    memcpy(&v, &b, 4);
    return v;
}
static double fast_rcp_d(double b, int n)
{
    double r = (double)nudge(1.0f / (float)b, n);
    r = fma(r, fma(-b, r, 1.0), r);
    r = fma(r, fma(-b, r, 1.0), r);
    return r;
}
static double fast_noise_scale(double x0, int n)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10, Lg1 = 6.666666666666735130e-01,
                 Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    uint64_t bits;
    memcpy(&bits, &x0, 8);
    int hx = (int)(bits >> 32);
    const uint32_t lx = (uint32_t)bits;
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    bits = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | lx;
    double x;
    memcpy(&x, &bits, 8);
    k += (i >> 20);
    const double f = x - 1.0, dk = (double)k;
    const double s = f * fast_rcp_d(2.0 + f, n);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    const double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double L = fma(dk, ln2_hi, -((hfsq - fma(s, hfsq + R, dk * ln2_lo)) - f));
    const double q = (-2.0 * L) * fast_rcp_d(x0, -n);
    double y = (double)nudge(1.0f / sqrtf((float)q), n);
    y = y * fma(-0.5 * q * y, y, 1.5);
    y = y * fma(-0.5 * q * y, y, 1.5);
    const double m = q * y;
    return fma(fma(-m, m, q), 0.5 * y, m);
}

int main(int argc, char** argv)
{
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 4099u;
    /* r2 = xr^2 + yr^2 with xr, yr multiples of 2^-19: the smallest non-zero value is 2^-38 */
    const float lo = ldexpf(1.0f, -38);
    uint32_t b0, b1;
    const float one = 1.0f;
    memcpy(&b0, &lo, 4);
    memcpy(&b1, &one, 4);
    double worst = 0.0;
    long long count = 0;
#pragma omp parallel for reduction(max : worst) reduction(+ : count) schedule(static)
    for (uint32_t b = b0; b < b1; b += stride) {
        float r2;
        memcpy(&r2, &b, 4);
        const double x = (double)r2;
        const double exact = sqrt(-2.0 * o_det_log(x) / x);
        int64_t eb;
        memcpy(&eb, &exact, 8);
        for (int n = -2; n <= 2; n += 2) {
            const double fast = fast_noise_scale(x, n);
            int64_t fb;
            memcpy(&fb, &fast, 8);
            const double d = fabs((double)(fb - eb));
            if (d > worst) worst = d;
        }
        ++count;
    }
    printf("values %lld max_ulps %.0f\n", count, worst);
    return worst < 256.0 ? 0 : 1;
}
