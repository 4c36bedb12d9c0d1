// cr_math.h -- deterministic binary32 elementary functions for host and device.
//
// The reference's device code (libEyeRenderer3/shaders.cu) is compiled with --use_fast_math
// (CMakeLists.txt:142), so its cos/sin/acos/asin/atan2/powf/logf are hardware approximations
// whose bits cannot be reproduced on a CPU.  This renderer instead evaluates fixed, fully
// specified algorithms built only from IEEE +,-,*,/,sqrt, fma, rint/floor and integer bit
// operations, so a CPU checker can reproduce every ray, hit and pixel bit for bit.
// Build rules that make this hold: nvcc -fmad=false (no implicit contraction), no
// --use_fast_math, default -prec-div/-prec-sqrt/-ftz=false; every fused multiply-add below is
// an explicit fmaf().
//
// Kernels: Cody-Waite + Cephes single-precision minimax polynomials (public domain).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define CR_HD __host__ __device__ __forceinline__
#else
#define CR_HD inline
#endif

namespace crm {

CR_HD uint32_t f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
CR_HD float u2f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

constexpr float kPi      = 3.14159265358979323846f;   // sutil M_PIf
constexpr float kPiO2    = 1.5707963705062866f;
constexpr float kPiO4    = 0.7853981852531433f;
constexpr float k2OPi    = 0.6366197466850281f;
constexpr float kPiO2Hi  = 1.5707963705062866f;
constexpr float kPiO2Mid = -4.371138828673793e-08f;
constexpr float kPiO2Lo  = -1.7763568394002505e-15f;

CR_HD void sincos(float x, float& sn, float& cs)
{
    const float kf = rintf(x * k2OPi);
    const int k = static_cast<int>(kf);
    float r = fmaf(-kf, kPiO2Hi, x);
    r = fmaf(-kf, kPiO2Mid, r);
    r = fmaf(-kf, kPiO2Lo, r);
    const float z = r * r;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    const float s = fmaf(r * z, ps, r);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    const float c = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
    // quadrant k & 3: (sn, cs) = (s, c), (c, -s), (-s, -c), (-c, s) -- written as selects and sign flips (no branch: the lanes
    // of a warp fall into different quadrants); a sign flip is exact, so the bits are those of the four-way switch
    const bool swap = (k & 1) != 0;
    const float a = swap ? c : s, b = swap ? s : c;
    sn = (k & 2) ? -a : a;
    cs = ((k + 1) & 2) ? -b : b;
}

CR_HD float log(float x)
{
    if (!(x > 0.0f)) return (x == 0.0f) ? -INFINITY : NAN;
    int e = 0;
    uint32_t u = f2u(x);
    if (u < 0x00800000u) { x = x * 8388608.0f; u = f2u(x); e = -23; }
    if (u >= 0x7f800000u) return x;
    e += static_cast<int>(u >> 23) - 126;
    float m = u2f((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.70710678118654752440f) { e -= 1; m = (m + m) - 1.0f; }
    else { m = m - 1.0f; }
    const float z = m * m;
    float p = 7.0376836292e-2f;
    p = fmaf(p, m, -1.1514610310e-1f);
    p = fmaf(p, m, 1.1676998740e-1f);
    p = fmaf(p, m, -1.2420140846e-1f);
    p = fmaf(p, m, 1.4249322787e-1f);
    p = fmaf(p, m, -1.6668057665e-1f);
    p = fmaf(p, m, 2.0000714765e-1f);
    p = fmaf(p, m, -2.4999993993e-1f);
    p = fmaf(p, m, 3.3333331174e-1f);
    float y = (m * z) * p;
    const float fe = static_cast<float>(e);
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

CR_HD float exp(float x)
{
    if (x != x) return x;
    if (x > 88.72283905206835f) return INFINITY;
    if (x < -103.0f) return 0.0f;
    const float n = floorf(fmaf(x, 1.44269504088896341f, 0.5f));
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    const float z = r * r;
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float y = fmaf(p, z, r) + 1.0f;
    const int ni = static_cast<int>(n);
    const int n1 = ni / 2, n2 = ni - n1;
    y = y * u2f(static_cast<uint32_t>(n1 + 127) << 23);
    y = y * u2f(static_cast<uint32_t>(n2 + 127) << 23);
    return y;
}

// x^y for x >= 0 (gamma 2.2 and 1/2.2 only: shaders.cu:100-106,183-187)
CR_HD float powGeneric(float x, float y)      // the definition: exp(y * log(x)) with the special cases
{
    if (x != x || y != y) return NAN;
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : ((y == 0.0f) ? 1.0f : INFINITY);
    if (x < 0.0f) return NAN;
    return crm::exp(y * crm::log(x));
}
CR_HD float pow(float x, float y)
{
    // Fast path for the values this renderer actually raises (colours, gamma 2.2 and 1/2.2): normal x in
    // [2^-40, 2^40] and |y| <= 2.5, so |y*log x| < 70.  It performs EXACTLY the operations of
    // exp(y * log(x)) below with the branches that cannot be taken in that range removed (no NaN/zero/
    // denormal/overflow cases; the 2^n scaling is one exponent add instead of two exact multiplies), hence
    // the same bits -- checked exhaustively over the whole range by tests/test_host.py.
    const uint32_t ux = f2u(x);
    if (ux - 0x2B800000u <= 0x28000000u && fabsf(y) <= 2.5f) {
        int e = static_cast<int>(ux >> 23) - 126;
        float m = u2f((ux & 0x007fffffu) | 0x3f000000u);
        if (m < 0.70710678118654752440f) { e -= 1; m = (m + m) - 1.0f; }
        else { m = m - 1.0f; }
        const float z = m * m;
        float p = 7.0376836292e-2f;
        p = fmaf(p, m, -1.1514610310e-1f);
        p = fmaf(p, m, 1.1676998740e-1f);
        p = fmaf(p, m, -1.2420140846e-1f);
        p = fmaf(p, m, 1.4249322787e-1f);
        p = fmaf(p, m, -1.6668057665e-1f);
        p = fmaf(p, m, 2.0000714765e-1f);
        p = fmaf(p, m, -2.4999993993e-1f);
        p = fmaf(p, m, 3.3333331174e-1f);
        float l = (m * z) * p;
        const float fe = static_cast<float>(e);
        l = fmaf(fe, -2.12194440e-4f, l);
        l = fmaf(-0.5f, z, l);
        float lg = m + l;
        lg = fmaf(fe, 0.693359375f, lg);
        const float t = y * lg;
        const float n = floorf(fmaf(t, 1.44269504088896341f, 0.5f));
        float r = fmaf(n, -0.693359375f, t);
        r = fmaf(n, 2.12194440e-4f, r);
        const float zz = r * r;
        float q = 1.9875691500e-4f;
        q = fmaf(q, r, 1.3981999507e-3f);
        q = fmaf(q, r, 8.3334519073e-3f);
        q = fmaf(q, r, 4.1665795894e-2f);
        q = fmaf(q, r, 1.6666665459e-1f);
        q = fmaf(q, r, 5.0000001201e-1f);
        const float v = fmaf(q, zz, r) + 1.0f;
        return u2f(f2u(v) + (static_cast<uint32_t>(static_cast<int>(n)) << 23));
    }
    if (x != x || y != y) return NAN;
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : ((y == 0.0f) ? 1.0f : INFINITY);
    if (x < 0.0f) return NAN;
    return crm::exp(y * crm::log(x));
}

CR_HD float asin(float xx)
{
    const float a = fabsf(xx);
    if (!(a <= 1.0f)) return NAN;
    float x, z;
    bool flag;
    if (a > 0.5f) { z = 0.5f * (1.0f - a); x = sqrtf(z); flag = true; }
    else { x = a; z = x * x; flag = false; }
    float p = 4.2163199048e-2f;
    p = fmaf(p, z, 2.4181311049e-2f);
    p = fmaf(p, z, 4.5470025998e-2f);
    p = fmaf(p, z, 7.4953002686e-2f);
    p = fmaf(p, z, 1.6666752422e-1f);
    float r = fmaf(p * z, x, x);
    if (flag) { r = r + r; r = kPiO2 - r; }
    return (xx < 0.0f) ? -r : r;
}

// |x| > 1 -> NaN: the reference's nearest-ommatidium search silently skips NaN angles
// (shaders.cu:431-441), which this reproduces.
CR_HD float acos(float x)
{
    if (!(fabsf(x) <= 1.0f)) return NAN;
    if (x < -0.5f) return kPi - 2.0f * crm::asin(sqrtf(0.5f * (1.0f + x)));
    if (x > 0.5f) return 2.0f * crm::asin(sqrtf(0.5f * (1.0f - x)));
    return kPiO2 - crm::asin(x);
}

CR_HD float atan(float xx)
{
    float x = fabsf(xx), y;
    if (x > 2.414213562373095f) { y = kPiO2; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = kPiO4; x = (x - 1.0f) / (x + 1.0f); }
    else { y = 0.0f; }
    const float z = x * x;
    float p = 8.05374449538e-2f;
    p = fmaf(p, z, -1.38776856032e-1f);
    p = fmaf(p, z, 1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    y = y + fmaf(p * z, x, x);
    return (xx < 0.0f) ? -y : y;
}

CR_HD float atan2(float y, float x)
{
    if (x != x || y != y) return NAN;
    if (x == 0.0f) {
        if (y == 0.0f) return 0.0f;
        return (y > 0.0f) ? kPiO2 : -kPiO2;
    }
    const float a = crm::atan(y / x);
    if (x > 0.0f) return a;
    return (y >= 0.0f) ? a + kPi : a - kPi;
}

}  // namespace crm
