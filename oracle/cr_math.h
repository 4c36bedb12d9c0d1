/* oracle/cr_math.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Deterministic single-precision elementary functions.
 *
 * Why this exists: the reference's device code (libEyeRenderer3/shaders.cu) calls
 * cos/sin (:648-651, :425-426), acos (:431), atan2/asin (:743, :752), powf (:100-106,
 * :183-187) and cuRAND's Box-Muller (logf/sqrtf/sincosf, curand_normal.h:70-87) and is
 * built with --use_fast_math (CMakeLists.txt:142), i.e. with approximate intrinsics whose
 * bit patterns cannot be reproduced off-GPU.  To make "same ray, same hit, same pixel map"
 * checkable BIT-EXACTLY between this CPU oracle and the CUDA product, both sides evaluate
 * the SAME specified algorithms below using only IEEE-754 binary32 +,-,*,/,sqrt, fma,
 * round-to-nearest-even conversions and integer bit operations.  Every operation is
 * written out; compilers must not contract or reassociate
 * (gcc: -ffp-contract=off, no -ffast-math;  nvcc: -fmad=false, no --use_fast_math).
 *
 * Accuracy (checked against libm in tests/test_oracle_math.py): a few ulp, i.e. at least
 * as accurate as the fast-math intrinsics the reference build actually executes.
 *
 * Polynomial coefficient sets are the classic Cephes single-precision minimax sets
 * (public domain, S. Moshier), a published algorithm restated here.
 */
#ifndef CR_ORACLE_MATH_H
#define CR_ORACLE_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t crm_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    crm_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#define CRM_PI      3.14159265358979323846f   /* sutil M_PIf */
#define CRM_PIO2    1.5707963705062866f
#define CRM_PIO4    0.7853981852531433f
#define CRM_2OPI    0.6366197466850281f
#define CRM_PIO2_HI 1.5707963705062866f
#define CRM_PIO2_MID (-4.371138828673793e-08f)
#define CRM_PIO2_LO (-1.7763568394002505e-15f)

/* sin and cos of x (|x| < ~1e5 for full accuracy).  Cody-Waite 3-term reduction with fma,
 * then Cephes sinf/cosf kernels on [-pi/4, pi/4]. */
static inline void crm_sincosf(float x, float* sn, float* cs)
{
    float kf = rintf(x * CRM_2OPI);             /* round-to-nearest-even */
    int   k  = (int)kf;
    float r  = fmaf(-kf, CRM_PIO2_HI, x);
    r = fmaf(-kf, CRM_PIO2_MID, r);
    r = fmaf(-kf, CRM_PIO2_LO, r);
    float z  = r * r;
    /* sin kernel */
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    float s = fmaf(r * z, ps, r);
    /* cos kernel */
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    float c = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
    switch (k & 3) {
        case 0: *sn =  s; *cs =  c; break;
        case 1: *sn =  c; *cs = -s; break;
        case 2: *sn = -s; *cs = -c; break;
        default:*sn = -c; *cs =  s; break;
    }
}
static inline float crm_sinf(float x) { float s, c; crm_sincosf(x, &s, &c); return s; }
static inline float crm_cosf(float x) { float s, c; crm_sincosf(x, &s, &c); return c; }

/* natural log, x > 0 finite (denormals handled by pre-scaling). x<=0 -> -inf / NaN. */
static inline float crm_logf(float x)
{
    if (!(x > 0.0f)) return (x == 0.0f) ? -INFINITY : NAN;
    int e = 0;
    uint32_t u = crm_f2u(x);
    if (u < 0x00800000u) { x = x * 8388608.0f; u = crm_f2u(x); e = -23; }   /* denormal */
    if (u >= 0x7f800000u) return x;                                           /* inf */
    e += (int)(u >> 23) - 126;                       /* x = m * 2^e, m in [0.5,1) */
    float m = crm_u2f((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.70710678118654752440f) { e -= 1; m = (m + m) - 1.0f; }
    else                             { m = m - 1.0f; }
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = fmaf(p, m, -1.1514610310e-1f);
    p = fmaf(p, m,  1.1676998740e-1f);
    p = fmaf(p, m, -1.2420140846e-1f);
    p = fmaf(p, m,  1.4249322787e-1f);
    p = fmaf(p, m, -1.6668057665e-1f);
    p = fmaf(p, m,  2.0000714765e-1f);
    p = fmaf(p, m, -2.4999993993e-1f);
    p = fmaf(p, m,  3.3333331174e-1f);
    float y  = (m * z) * p;
    float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

/* e^x for x in about [-87, 88]; out of range saturates to 0 / inf. */
static inline float crm_expf(float x)
{
    if (x != x) return x;
    if (x > 88.72283905206835f) return INFINITY;
    if (x < -103.0f) return 0.0f;
    float n = floorf(fmaf(x, 1.44269504088896341f, 0.5f));
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float z = r * r;
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float y = fmaf(p, z, r) + 1.0f;
    /* scale by 2^n in two steps so that denormal results round once from a normal */
    int ni = (int)n;
    int n1 = ni / 2, n2 = ni - n1;
    y = y * crm_u2f((uint32_t)(n1 + 127) << 23);
    y = y * crm_u2f((uint32_t)(n2 + 127) << 23);
    return y;
}

/* x^y for x >= 0 (the only use: gamma 2.2 and 1/2.2 on [0,1], shaders.cu:100-106,183-187).
 * x == 0 -> 0 (y > 0); x == 1 -> 1 exactly. */
static inline float crm_powf(float x, float y)
{
    if (x != x || y != y) return NAN;
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : ((y == 0.0f) ? 1.0f : INFINITY);
    if (x < 0.0f) return NAN;
    return crm_expf(y * crm_logf(x));
}

/* asin / acos (Cephes asinf). |x| > 1 -> NaN (matters: the reference's projection arg-min
 * silently skips NaN angles, shaders.cu:431-441). */
static inline float crm_asinf(float xx)
{
    float a = fabsf(xx);
    if (!(a <= 1.0f)) return NAN;
    float x, z; int flag;
    if (a > 0.5f) { z = 0.5f * (1.0f - a); x = sqrtf(z); flag = 1; }
    else          { x = a; z = x * x; flag = 0; }
    float p = 4.2163199048e-2f;
    p = fmaf(p, z, 2.4181311049e-2f);
    p = fmaf(p, z, 4.5470025998e-2f);
    p = fmaf(p, z, 7.4953002686e-2f);
    p = fmaf(p, z, 1.6666752422e-1f);
    float r = fmaf(p * z, x, x);
    if (flag) { r = r + r; r = CRM_PIO2 - r; }
    return (xx < 0.0f) ? -r : r;
}
static inline float crm_acosf(float x)
{
    if (!(fabsf(x) <= 1.0f)) return NAN;
    if (x < -0.5f) return CRM_PI - 2.0f * crm_asinf(sqrtf(0.5f * (1.0f + x)));
    if (x >  0.5f) return 2.0f * crm_asinf(sqrtf(0.5f * (1.0f - x)));
    return CRM_PIO2 - crm_asinf(x);
}

/* atan for any finite x (Cephes atanf), atan2 with the usual quadrant rules. */
static inline float crm_atanf(float xx)
{
    float x = fabsf(xx), y;
    if (x > 2.414213562373095f)       { y = CRM_PIO2; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = CRM_PIO4; x = (x - 1.0f) / (x + 1.0f); }
    else                              { y = 0.0f; }
    float z = x * x;
    float p = 8.05374449538e-2f;
    p = fmaf(p, z, -1.38776856032e-1f);
    p = fmaf(p, z,  1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    y = y + fmaf(p * z, x, x);
    return (xx < 0.0f) ? -y : y;
}
static inline float crm_atan2f(float y, float x)
{
    if (x != x || y != y) return NAN;
    if (x == 0.0f) {
        if (y == 0.0f) return 0.0f;
        return (y > 0.0f) ? CRM_PIO2 : -CRM_PIO2;
    }
    float a = crm_atanf(y / x);
    if (x > 0.0f) return a;
    return (y >= 0.0f) ? a + CRM_PI : a - CRM_PI;
}

#endif
