// cr_kernels.cu -- hand-written sm_100a kernels of the compound-eye render path.
//
// Replaces every live OptiX program of libEyeRenderer3/shaders.cu (reference lines cited at each
// function).  Compile with -fmad=false and without --use_fast_math: the arithmetic below is
// written one IEEE rounding per operation with explicit fmaf() where fusion is wanted, so the CPU
// checker reproduces rays, hits and pixel maps bit for bit (see cr_math.h).
//
// K0  k_rngInit            curand_init(42, id, 0) per sample stream   (shaders.cu:680-685), own jump tables
//     k_prepOmmatidia      per-ommatidium invariants of the sample-ray construction
// K0b k_buildEntries       per (frame, ommatidium): BVH subtree roots its sample cone can reach (entry frontier)
// K1  k_traceCompound      raygen + BVH traversal + shading, one sample ray per lane
//                          (shaders.cu:664-731 + 110-137 + 740-811)
// K1b k_sumSamples[Tma]    exact-order per-ommatidium sum (shaders.cu:341-347) + 8-bit row of single_dimension_fast;
//                          TMA bulk copies + mbarrier when S % 4 == 0, cp.async otherwise
// K2  k_projectVector/Raw  single_dimension[_fast], raw_ommatidial_samples (shaders.cu:354-406)
// K3  k_buildProjectionMap nearest-ommatidium map of the spherical modes, cached per eye/size
//     k_projectMap         map lookup + make_color, or ids                (shaders.cu:412-640)
// K4  k_camera             pinhole / panoramic / orthographic primary rays (shaders.cu:198-333)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "cr_device.h"
#include "cr_math.h"

namespace cr {

namespace {

// ------------------------------------------------------------------------------------------
// small vector helpers; plain versions follow sutil/vec_math.h (products then sums, left to right)
// ------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 vmuls(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 vcross(V3 a, V3 b)
{ return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float vlen(V3 a) { return sqrtf(vdot(a, a)); }
__device__ __forceinline__ V3 vnormalize(V3 v) { const float inv = 1.0f / sqrtf(vdot(v, v)); return vmuls(v, inv); }
// fused versions used by the intersection test (this renderer's own arithmetic; OptiX's is closed)
__device__ __forceinline__ float fdot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ V3 fcross(V3 a, V3 b)
{ return mk(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))); }

// ------------------------------------------------------------------------------------------
// Elementary-function policy.  FAST = false: the specified binary32 algorithms of cr_math.h (every bit
// reproducible by the CPU checker).  FAST = true (opt-in, crSetRenderMode): the hardware approximations the
// reference itself runs -- its device code is built with --use_fast_math (CMakeLists.txt:142) and cuRAND's
// Box-Muller calls __sincosf (curand_normal.h:70-87) -- MUFU sin/cos/lg2/ex2 instead of ~30-instruction
// polynomials.  Rays then differ from the exact mode in the last bits, so this mode is held to the
// north_star tolerance (max |dRGB| <= 1/255, mean <= 1e-4), not to bit-exactness.
// ------------------------------------------------------------------------------------------
template <bool FAST>
struct Fn {
    static __device__ __forceinline__ void sincos(float x, float& s, float& c)
    {
        if (FAST) __sincosf(x, &s, &c);
        else crm::sincos(x, s, c);
    }
    static __device__ __forceinline__ float log(float x) { return FAST ? __logf(x) : crm::log(x); }
    static __device__ __forceinline__ float pow(float x, float y) { return FAST ? __powf(x, y) : crm::pow(x, y); }
    static __device__ __forceinline__ float asin(float x) { return FAST ? asinf(x) : crm::asin(x); }
    static __device__ __forceinline__ float atan2(float y, float x) { return FAST ? atan2f(y, x) : crm::atan2(y, x); }
};

// ------------------------------------------------------------------------------------------
// cuRAND XORWOW on a 32-byte compact state (curand_kernel.h:150-156, :863-874;
// curand_normal.h:70-87,313-326; curand_uniform.h:69-72).  Streams/seeds as in shaders.cu:680-695.
// ------------------------------------------------------------------------------------------
struct Rng {
    uint32_t d, v0, v1, v2, v3, v4;
    int flag;
    float extra;
};
// The state array is streamed once per frame (64 B per ray in and out): evict-first accesses keep it
// from displacing BVH nodes in L1/L2.
__device__ __forceinline__ Rng rngLoad(const uint4* __restrict__ p)
{
    const uint4 a = __ldcs(p), b = __ldcs(p + 1);       // (one 256-bit streaming load measured slower here: 15.3 vs 15.5 Grays/s e2e)
    Rng r;
    r.d = a.x; r.v0 = a.y; r.v1 = a.z; r.v2 = a.w; r.v3 = b.x; r.v4 = b.y;
    r.flag = (int)b.z; r.extra = __uint_as_float(b.w);
    return r;
}
__device__ __forceinline__ void rngStore(uint4* __restrict__ p, const Rng& r)
{
    __stcs(p, make_uint4(r.d, r.v0, r.v1, r.v2));
    __stcs(p + 1, make_uint4(r.v3, r.v4, (uint32_t)r.flag, __float_as_uint(r.extra)));
}
__device__ __forceinline__ uint32_t rngNext(Rng& r)
{
    const uint32_t t = r.v0 ^ (r.v0 >> 2);
    r.v0 = r.v1; r.v1 = r.v2; r.v2 = r.v3; r.v3 = r.v4;
    r.v4 = (r.v4 ^ (r.v4 << 4)) ^ (t ^ (t << 1));
    r.d += 362437u;
    return r.v4 + r.d;
}
__device__ __forceinline__ float rngUniform(Rng& r)
{
    const uint32_t x = rngNext(r);
    return (float)x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}
template <bool FAST = false>
__device__ __forceinline__ float rngNormal(Rng& r)
{
    if (r.flag != 1) {
        const uint32_t x = rngNext(r);
        const uint32_t y = rngNext(r);
        const float u = (float)x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
        const float v = (float)y * (2.3283064e-10f * 6.2831855f) + ((2.3283064e-10f * 6.2831855f) / 2.0f);
        const float s = sqrtf(-2.0f * Fn<FAST>::log(u));
        float sn, cs;
        Fn<FAST>::sincos(v, sn, cs);
        r.extra = cs * s;
        r.flag = 1;
        return sn * s;
    }
    r.flag = 0;
    return r.extra;
}

// K0: one thread per stream.  Stream id = N*s + o (shaders.cu:668-669); stored at [o*S + s].
// v <- v * T^(n) (or T^(2^67 n) for firstLevel 0): one vector-matrix product per non-zero hex digit of n, with
// the host-built tables of cr_xorwow_jump.h (row = 2 x 16 B: 5 words + padding).  The digits are applied from
// the MOST significant down (the factors commute): the lanes of a warp hold consecutive ids, so they start
// from the same seed vector and share every digit but the last -- identical bit loops and broadcast row
// loads until the final digit, where only the matrix differs between lanes.
__device__ __forceinline__ void xorwowJump(uint32_t& v0, uint32_t& v1, uint32_t& v2, uint32_t& v3, uint32_t& v4,
                                           const uint4* __restrict__ table, int firstLevel, unsigned long long n)
{
    if (n == 0ull) return;
    for (int k = (63 - __clzll((long long)n)) >> 2; k >= 0; k--) {
        const int d = (int)((n >> (4 * k)) & 15ull);
        if (!d) continue;
        const uint4* m = table + ((size_t)(firstLevel + k) * 15 + (size_t)(d - 1)) * (160 * 2);
        uint32_t r0 = 0u, r1 = 0u, r2 = 0u, r3 = 0u, r4 = 0u;
        const uint32_t v[5] = {v0, v1, v2, v3, v4};
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const uint4* rows = m + i * 64;
            for (uint32_t w = v[i]; w; w &= w - 1u) {
                const int j = __ffs((int)w) - 1;
                const uint4 a = __ldg(rows + 2 * j);
                const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(rows + 2 * j + 1));
                r0 ^= a.x; r1 ^= a.y; r2 ^= a.z; r3 ^= a.w; r4 ^= b;
            }
        }
        v0 = r0; v1 = r1; v2 = r2; v3 = r3; v4 = r4;
    }
}

// K0: curand_init(42, id, 0) per sample stream (shaders.cu:680-685; curand_kernel.h:800-825) as seed
// scrambling + one jump of id subsequences.  Threads run over ommatidia fastest, so the lanes of a warp hold
// consecutive ids and share all but the lowest hex digit's matrix.
// firstFrame > 0 positions the stream as if `firstFrame` frames had already been rendered (pose sharding /
// restart): 2*(firstFrame & ~1) raw draws are skipped (3 + 1 draws per frame pair), plus one replayed frame when
// firstFrame is odd so that the Box-Muller cache is populated exactly as in the sequential run.
// An ommatidium-range shard (crSetOmmatidialShard) holds rows [oFirst, oFirst+N) of an eye of nGlobal
// ommatidia: the stream id keeps the GLOBAL indices, so a sharded frame equals the unsharded one.
__global__ void k_rngInit(uint4* __restrict__ rng, int N, int S, unsigned long long firstFrame, unsigned long long nGlobal,
                          unsigned long long oFirst, const uint4* __restrict__ jumpTable)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * S) return;
    const int s = (int)(i / N), o = (int)(i - (long long)s * N);
    const unsigned long long id = nGlobal * (unsigned long long)s + oFirst + (unsigned long long)o;
    // seed 42 (curand_kernel.h:800-812)
    const uint32_t s0 = 42u ^ 0xaad26b49u, s1 = 0u ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    Rng r;
    r.d = 6615241u + t1 + t0;
    r.v0 = 123456789u + t0; r.v1 = 362436069u ^ t0; r.v2 = 521288629u + t1; r.v3 = 88675123u ^ t1; r.v4 = 5783321u + t0;
    const unsigned long long skip = 2ull * (firstFrame & ~1ull);
    if (skip) {                                                                     // same for every stream
        xorwowJump(r.v0, r.v1, r.v2, r.v3, r.v4, jumpTable, 8, skip);
        r.d += 362437u * (uint32_t)skip;
    }
    xorwowJump(r.v0, r.v1, r.v2, r.v3, r.v4, jumpTable, 0, id);                    // whole subsequences: d unchanged
    r.flag = 0; r.extra = 0.0f;
    if (firstFrame & 1ull) { (void)rngNormal(r); (void)rngUniform(r); }
    rngStore(rng + 2 * ((size_t)o * (size_t)S + (size_t)s), r);
}

// ------------------------------------------------------------------------------------------
// Ommatidial sample ray (shaders.cu:648-662 generateOffsetRay/rotatePoint, :676-709)
// ------------------------------------------------------------------------------------------
#define CR_FWHM_SD_RATIO 2.35482004503094938202313865291f   /* shaders.cu:53 */
#ifndef CR_CONE_SIGMAS
#define CR_CONE_SIGMAS 4.0f
#endif
constexpr float kConeSigmas = CR_CONE_SIGMAS;   // sample rays with |splay| <= kConeSigmas*sd start from the entry frontier

template <bool FAST = false>
__device__ __forceinline__ V3 rotatePoint(V3 p, float angle, V3 axis)   // axis NOT re-normalised
{
    float sn, cs;
    Fn<FAST>::sincos(angle, sn, cs);
    const V3 a = vmuls(p, cs);
    const V3 b = vmuls(vcross(axis, p), sn);
    const V3 c = vmuls(axis, (1.0f - cs) * vdot(axis, p));
    return vadd(vadd(a, b), c);
}
struct Ray { V3 o, d; float tmin; };

// Per-ommatidium invariants of the sample-ray construction, hoisted out of the per-sample path
// (identical operations in identical order, so the rays are bit-identical to evaluating
// shaders.cu:652-709 per sample):
//   pre[0] = (relPos - normalize(axis)*focal, sd = acceptance / FWHM_SD_RATIO)
//   pre[1] = (axis, focal)          pre[2] = (perp, kConeSigmas*|sd| = splay bound of the entry cone)
//   pre[3] = (cross(perp, axis), dot(perp, axis))   -- rotatePoint(axis, splay, perp) without its cross and dot
__global__ void k_prepOmmatidia(const float4* __restrict__ omm, int N, float4* __restrict__ pre)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const float4 q0 = omm[2 * o], q1 = omm[2 * o + 1];
    const V3 relPos = mk(q0.x, q0.y, q0.z);
    const V3 axis = mk(q0.w, q1.x, q1.y);
    const float acceptance = q1.z, focal = q1.w;
    const float sd = acceptance / CR_FWHM_SD_RATIO;
    V3 perp = vcross(mk(0.0f, 1.0f, 0.0f), axis);
    if (perp.x + perp.y + perp.z == 0.0f) perp = mk(0.0f, 0.0f, 1.0f);   // exact-zero test on the SUM (:656)
    else perp = vnormalize(perp);
    const V3 rp = vsub(relPos, vmuls(vnormalize(axis), focal));
    const V3 pxa = vcross(perp, axis);                  // the two sample-independent factors of the first rotation
    const float pda = vdot(perp, axis);
    pre[kPreStride * o + 0] = make_float4(rp.x, rp.y, rp.z, sd);
    pre[kPreStride * o + 1] = make_float4(axis.x, axis.y, axis.z, focal);
    pre[kPreStride * o + 2] = make_float4(perp.x, perp.y, perp.z, kConeSigmas * fabsf(sd));
    pre[kPreStride * o + 3] = make_float4(pxa.x, pxa.y, pxa.z, pda);
}

template <bool FAST = false>
__device__ __forceinline__ Ray ommatidialRay(const float4 p0, const float4 p1, const float4 p2, const float4 p3, const DevicePose& P,
                                             Rng& rng, float& splayOut)
{
    const V3 rp = mk(p0.x, p0.y, p0.z);
    const V3 axis = mk(p1.x, p1.y, p1.z);
    const V3 perp = mk(p2.x, p2.y, p2.z);
    const float splay = rngNormal<FAST>(rng) * p0.w;
    splayOut = splay;
    const float axisAngle = rngUniform(rng) * crm::kPi;
    // = rotatePoint(axis, splay, perp) with cross(perp, axis) and dot(perp, axis) taken from the table
    float sSn, sCs;
    Fn<FAST>::sincos(splay, sSn, sCs);
    const V3 splayed = vadd(vadd(vmuls(axis, sCs), vmuls(mk(p3.x, p3.y, p3.z), sSn)), vmuls(perp, (1.0f - sCs) * p3.w));
    const V3 rd = rotatePoint<FAST>(splayed, axisAngle, axis);
    const V3 X = mk(P.xx, P.xy, P.xz), Y = mk(P.yx, P.yy, P.yz), Z = mk(P.zx, P.zy, P.zz);
    Ray r;
    r.o = vadd(vadd(vadd(mk(P.px, P.py, P.pz), vmuls(X, rp.x)), vmuls(Y, rp.y)), vmuls(Z, rp.z));
    r.d = vadd(vadd(vmuls(X, rd.x), vmuls(Y, rd.y)), vmuls(Z, rd.z));
    r.tmin = p1.w;                                                    // shaders.cu:721
    return r;
}

// ------------------------------------------------------------------------------------------
// Closest hit (stands in for optixTrace, shaders.cu:110-137): two-sided, closest t in
// (tmin, tmax], ties on t resolved to the LOWEST flattened primitive index so the result does not
// depend on traversal order.
//
// Box test: t = b*inv - o*inv as ONE fma per plane.  Conservative by construction:
//   inv  = rcp.approx(d) (<= 1 ulp error; only the box test uses it, never the triangle test)
//   invN = inv*(1-2^-21), invF = inv*(1+2^-21)  absorb the error of inv and the rounding of the fma
//          (both relative to |t|) for planes in front of the origin;
//   addN = -(o*invN) - E, addF = -(o*invF) + E with E = 2^-21*|o*inv| absorb the rounding of the
//          o*inv product (which is NOT relative to t).
// Leaf boxes are additionally padded at build time (cr_bvh.cu) for the inexactness of the
// triangle test itself, so BVH traversal returns exactly what a brute-force loop returns.
// ------------------------------------------------------------------------------------------
struct Hit { float t; int prim; float u, v; };

struct RayBox {
    float nix, niy, niz;      // invN
    float fix, fiy, fiz;      // invF
    float nax, nay, naz;      // addN
    float fax, fay, faz;      // addF
};

__device__ __forceinline__ float rcpApprox(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c)      // FMNMX3 on sm_100a
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__device__ __forceinline__ void setupAxis(float o, float d, float& invN, float& invF, float& addN, float& addF, int& neg)
{
    const float dc = (fabsf(d) < 1e-18f) ? copysignf(1e-18f, d) : d;
    const float inv = rcpApprox(dc);
    neg = (int)(__float_as_uint(dc) >> 31);         // direction negative -> near plane is max
    invN = inv * (1.0f - 4.76837158203125e-07f);
    invF = inv * (1.0f + 4.76837158203125e-07f);
    const float oN = o * invN, oF = o * invF;
    const float E = fmaxf(fabsf(oN), fabsf(oF)) * 4.76837158203125e-07f;
    addN = -oN - E;
    addF = -oF + E;
}

__device__ __forceinline__ bool triTest(const float4* __restrict__ tri, const V3 o, const V3 d, const float tmin,
                                        const float tlimit, float& t, float& u, float& v, int& prim)
{
    const float4 a = __ldg(tri), b = __ldg(tri + 1), c = __ldg(tri + 2);
    const V3 v0 = mk(a.x, a.y, a.z), e1 = mk(b.x, b.y, b.z), e2 = mk(c.x, c.y, c.z);
    const V3 p = fcross(d, e2);
    const float det = fdot(e1, p);
    if (!(det != 0.0f)) return false;
    const float inv = 1.0f / det;                                     // IEEE division: part of the exact result
    const V3 s = vsub(o, v0);
    const float uu = fdot(s, p) * inv;
    if (!(uu >= 0.0f && uu <= 1.0f)) return false;
    const V3 q = fcross(s, e1);
    const float vv = fdot(d, q) * inv;
    if (!(vv >= 0.0f && uu + vv <= 1.0f)) return false;
    const float tt = fdot(e2, q) * inv;
    if (!(tt > tmin && tt <= tlimit)) return false;
    t = tt; u = uu; v = vv; prim = __float_as_int(a.w);
    return true;
}

// Traversal stack: the first kSmemStack levels live in shared memory laid out [level][thread]
// (bank = thread, conflict-free, 32-bit addressing); deeper levels spill to a per-thread local
// array (L1-cached).  An LBVH over 63-bit codes plus index tie-breaks is at most 63+28 levels deep, and a
// depth-first walk that pushes one sibling per level never holds more entries than that (+3 frontier entries):
// kSmemStack + kLocalStack = 96 covers it, so push() needs no bound check.
#ifndef CR_TRACE_MIN_BLOCKS
#define CR_TRACE_MIN_BLOCKS 8
#endif
#ifndef CR_SMEM_STACK
#define CR_SMEM_STACK 10   // measured: 32 -> 16 -> 10 levels = 19.6 -> 20.1 Grays/s batched, 15.07 -> 15.2 -> 15.4 per-frame API:
#endif                     // with the entry frontier the stack stays shallow, and 5 KB/CTA leaves the SM a 192 KB L1

#ifndef CR_EXP_STATE
#define CR_EXP_STATE 0
#endif
#ifndef CR_INLINE_PHASED
#define CR_INLINE_PHASED 0   // the per-lane walk INSIDE the trace kernel: 0 = while-while (traceClosest), 1 = phase-switched
#endif
constexpr int kSmemStack = CR_SMEM_STACK;
constexpr int kLocalStack = 96 - CR_SMEM_STACK;   // shared + local levels >= 96 > the deepest possible LBVH (63 + 28 levels) + 3 entries
constexpr int kSentinel = (int)0x80000000;
constexpr unsigned kFullMask = 0xffffffffu;

// The stack pointer is a plain member and the overflow levels live in an array OUTSIDE the struct (round 2): with the array
// inside, the whole struct -- sp included -- was kept in local memory, and push/pop were a quarter of the walk's
// instructions (profiles/r02i_traceQueue_phased_ncu_summary.txt).
struct Stack {
    int* smem;          // &sStack[0][tid]; rows are kTraceThreads ints apart
    int* local;         // kLocalStack ints of thread-local memory for the levels beyond kSmemStack
    int sp;
    __device__ __forceinline__ void push(int v)
    {
        if (sp < kSmemStack) smem[sp * kTraceThreads] = v;
        else local[sp - kSmemStack] = v;
        sp++;
    }
    __device__ __forceinline__ int pop()
    {
        if (sp == 0) return kSentinel;
        sp--;
        return (sp < kSmemStack) ? smem[sp * kTraceThreads] : local[sp - kSmemStack];
    }
};

// One 64-byte node = two 256-bit read-only loads (LDG.E.256, new with sm_100): half the load instructions and
// half the L1 sector requests of four 128-bit loads when every lane fetches a different node.
#ifndef CR_NODE_LDG256
#define CR_NODE_LDG256 1
#endif
__device__ __forceinline__ void ldgNode(const float4* __restrict__ p, float4& n0, float4& n1, float4& n2, float4& n3)
{
#if CR_NODE_LDG256
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w) : "l"(p));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=f"(n3.x), "=f"(n3.y), "=f"(n3.z), "=f"(n3.w) : "l"(p + 2));
#else
    n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); n3 = __ldg(p + 3);
#endif
}

template <bool COUNT>
__device__ __forceinline__ Hit traceClosest(const float4* __restrict__ nodesAll, size_t variantStride,
                                            const float4* __restrict__ tris, const Ray& ray, const float tmax, int* sStackLane,
                                            int stackStride, int* nodeCount, int* triCount,
                                            const int4 entry = make_int4(0, kSentinel, kSentinel, kSentinel))
{
    RayBox rb;
    int sx, sy, sz;
    setupAxis(ray.o.x, ray.d.x, rb.nix, rb.fix, rb.nax, rb.fax, sx);
    setupAxis(ray.o.y, ray.d.y, rb.niy, rb.fiy, rb.nay, rb.fay, sy);
    setupAxis(ray.o.z, ray.d.z, rb.niz, rb.fiz, rb.naz, rb.faz, sz);
    // node variant of this ray's direction octant: (near, far) planes are pre-selected
    const float4* __restrict__ nodes = nodesAll + (size_t)(sx | (sy << 1) | (sz << 2)) * variantStride;
    Hit best;
    best.t = tmax; best.prim = -1; best.u = 0.0f; best.v = 0.0f;
    int deep[kLocalStack];
    Stack st;
    st.smem = sStackLane; st.local = deep; st.sp = 0;
    // Entry refs (k_buildEntries) are packed from .x and sorted near to far; the default is the root.
    int cur = entry.x;
    if (entry.y != kSentinel) {
        if (entry.z != kSentinel) {
            if (entry.w != kSentinel) st.push(entry.w);
            st.push(entry.z);
        }
        st.push(entry.y);
    }
    int nc = 0, tc = 0;
    for (;;) {
        while (cur >= 0) {
            const float4* np = nodes + 4 * (size_t)cur;
            float4 n0, n1, n2, n3;
            ldgNode(np, n0, n1, n2, n3);
            if (COUNT) nc++;
            const float tn0 = fmax3(fmaf(n0.x, rb.nix, rb.nax), fmaf(n0.z, rb.niy, rb.nay), fmaxf(fmaf(n2.x, rb.niz, rb.naz), ray.tmin));
            const float tf0 = fmin3(fmaf(n0.y, rb.fix, rb.fax), fmaf(n0.w, rb.fiy, rb.fay), fminf(fmaf(n2.y, rb.fiz, rb.faz), best.t));
            const float tn1 = fmax3(fmaf(n1.x, rb.nix, rb.nax), fmaf(n1.z, rb.niy, rb.nay), fmaxf(fmaf(n2.z, rb.niz, rb.naz), ray.tmin));
            const float tf1 = fmin3(fmaf(n1.y, rb.fix, rb.fax), fmaf(n1.w, rb.fiy, rb.fay), fminf(fmaf(n2.w, rb.fiz, rb.faz), best.t));
            const bool h0 = tn0 <= tf0, h1 = tn1 <= tf1;
            const int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool firstIs0 = tn0 <= tn1;          // near child first; ties -> child 0
                cur = firstIs0 ? r0 : r1;
                const int far = firstIs0 ? r1 : r0;
                st.push(far);
            } else if (h0) cur = r0;
            else if (h1) cur = r1;
            else cur = st.pop();
        }
        if (cur == kSentinel) break;
        const int x = ~cur;
        const int first = x >> 3, cnt = (x & 7) + 1;
        for (int k = 0; k < cnt; k++) {
            float t, u, v;
            int prim;
            if (COUNT) tc++;
            if (triTest(tris + 3 * (size_t)(first + k), ray.o, ray.d, ray.tmin, best.t, t, u, v, prim)) {
                if (t < best.t || best.prim < 0 || prim < best.prim) { best.t = t; best.prim = prim; best.u = u; best.v = v; }
            }
        }
        cur = st.pop();
        if (cur == kSentinel) break;
    }
    if (COUNT) { *nodeCount = nc; *triCount = tc; }
    return best;
}

// ------------------------------------------------------------------------------------------
// Phase-switched traversal (round 2).  In the while-while walk above a lane that has reached a leaf waits until EVERY lane
// of its warp has reached one: with the unequal walks of rays that graze the scene the node loop runs at mean/max of the
// per-lane step counts -- 10.5 of 32 lanes (profiles/r02g_traceQueue_whilewhile_ncu_summary.txt).  Here the warp votes
// before every node step and leaves the node loop as soon as fewer than `nodeLanes` lanes still want a node while
// somebody holds a leaf; the leaf holders test their triangles and pop, and everybody re-enters the node loop together.
// Each lane performs exactly the operations of the while-while walk in the same order (a lane holding a leaf does not
// look ahead), only the interleaving between lanes changes: same hits, same node/triangle counts.
// The node step is written branch-light: one predicated push, one predicated pop, selects for the rest -- the four-way
// branch of the walk above spent a quarter of its instructions on stack code at 3 to 5 lanes.
// Must be called by all 32 lanes of a warp.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int nodeStep(const float4* __restrict__ nodes, const int cur, const RayBox& rb, const float tmin,
                                        const float tbest, Stack& st)
{
    const float4* np = nodes + 4 * (size_t)cur;
    float4 n0, n1, n2, n3;
    ldgNode(np, n0, n1, n2, n3);
    const float tn0 = fmax3(fmaf(n0.x, rb.nix, rb.nax), fmaf(n0.z, rb.niy, rb.nay), fmaxf(fmaf(n2.x, rb.niz, rb.naz), tmin));
    const float tf0 = fmin3(fmaf(n0.y, rb.fix, rb.fax), fmaf(n0.w, rb.fiy, rb.fay), fminf(fmaf(n2.y, rb.fiz, rb.faz), tbest));
    const float tn1 = fmax3(fmaf(n1.x, rb.nix, rb.nax), fmaf(n1.z, rb.niy, rb.nay), fmaxf(fmaf(n2.z, rb.niz, rb.naz), tmin));
    const float tf1 = fmin3(fmaf(n1.y, rb.fix, rb.fax), fmaf(n1.w, rb.fiy, rb.fay), fminf(fmaf(n2.w, rb.fiz, rb.faz), tbest));
    const bool h0 = tn0 <= tf0, h1 = tn1 <= tf1;
    const int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
    const bool firstIs0 = h0 && (!h1 || tn0 <= tn1);      // near child first; ties -> child 0 (as traceClosest)
    int next = r0, other = r1;
    if (!firstIs0) { next = r1; other = r0; }
    if (h0 && h1) st.push(other);
    if (!(h0 || h1)) next = st.pop();
    return next;
}

__device__ __forceinline__ void leafStep(const float4* __restrict__ tris, const int leafRef, const V3 o, const V3 d, const float tmin,
                                         Hit& best, int& triCount)
{
    const int x = ~leafRef;
    const int first = x >> 3, cnt = (x & 7) + 1;
    for (int k = 0; k < cnt; k++) {
        float t, u, v;
        int prim;
        triCount++;
        if (triTest(tris + 3 * (size_t)(first + k), o, d, tmin, best.t, t, u, v, prim)) {
            if (t < best.t || best.prim < 0 || prim < best.prim) { best.t = t; best.prim = prim; best.u = u; best.v = v; }
        }
    }
}

template <bool COUNT>
__device__ __forceinline__ Hit traceClosestPhased(const float4* __restrict__ nodesAll, size_t variantStride,
                                                  const float4* __restrict__ tris, const Ray& ray, const float tmax, int* sStackLane,
                                                  int stackStride, int* nodeCount, int* triCount, const int4 entry, const int nodeLanes)
{
    RayBox rb;
    int sx, sy, sz;
    setupAxis(ray.o.x, ray.d.x, rb.nix, rb.fix, rb.nax, rb.fax, sx);
    setupAxis(ray.o.y, ray.d.y, rb.niy, rb.fiy, rb.nay, rb.fay, sy);
    setupAxis(ray.o.z, ray.d.z, rb.niz, rb.fiz, rb.naz, rb.faz, sz);
    const float4* __restrict__ nodes = nodesAll + (size_t)(sx | (sy << 1) | (sz << 2)) * variantStride;
    Hit best;
    best.t = tmax; best.prim = -1; best.u = 0.0f; best.v = 0.0f;
    int deep[kLocalStack];
    Stack st;
    st.smem = sStackLane; st.local = deep; st.sp = 0;
    int cur = entry.x;
    if (entry.y != kSentinel) {
        if (entry.z != kSentinel) {
            if (entry.w != kSentinel) st.push(entry.w);
            st.push(entry.z);
        }
        st.push(entry.y);
    }
    int nc = 0, tc = 0;
    for (;;) {
        unsigned leaves;
        for (;;) {                                                  // one vote per node step while enough lanes want one
            const int nWant = __popc(__ballot_sync(kFullMask, cur >= 0));
            if (nWant < nodeLanes) {
                leaves = __ballot_sync(kFullMask, cur < 0 && cur != kSentinel);
                if (nWant == 0 || leaves != 0u) break;
            }
            if (cur >= 0) {
                cur = nodeStep(nodes, cur, rb, ray.tmin, best.t, st);
                if (COUNT) nc++;
            }
        }
        if (leaves == 0u) break;                                    // nobody wants a node, nobody holds a leaf
        if (cur < 0 && cur != kSentinel) {
            leafStep(tris, cur, ray.o, ray.d, ray.tmin, best, tc);
            cur = st.pop();
        }
    }
    if (COUNT) { *nodeCount = nc; *triCount = tc; }
    return best;
}

// ------------------------------------------------------------------------------------------
// Closest hit through an ommatidium's CANDIDATE LIST (k_buildEntries, second stage).  The sample cone of most
// ommatidia reaches only a handful of BVH leaves; for those the frontier pass flattens the part of the tree the
// cone can reach into a short list of "pre-leaf" elements (node index, which of its two children are reachable
// leaves).  The 32 lanes of a warp -- 32 samples of that ommatidium -- then run
//   phase 1, in lockstep over the list (no stack, no divergence, one node address per octant variant for the whole
//            warp): each lane tests its own ray against the listed child boxes and notes the leaves it hits as bits;
//   phase 2, per lane: each lane walks ITS bits and tests the triangles of ITS leaves -- different triangles in
//            different lanes under the same instructions.
// The per-lane stack walk below the frontier, whose lanes drift out of phase (14.7 of 32 lanes live in its node loop,
// 8.5 on its stack: profiles/r01e_final_k1_ncu_summary.txt), is gone for these warps.  A lane tests every listed leaf
// whose box its ray passes, i.e. a superset of the leaves the stack walk would test (which also prunes by the
// closest hit so far) -- and with the lowest-primitive tie rule the closest hit, hence every output bit, is the same.
// Must be called by all 32 lanes of the warp; `list` is warp-uniform.
// ------------------------------------------------------------------------------------------
constexpr int kListMax = 15;                 // elements per (frame, ommatidium); the record is 1 + kListMax ints = 64 B
constexpr int kListStride = kListMax + 1;
constexpr int kListFallback = -1;            // header value: no list -- walk the entry frontier per lane
static_assert(2 * kListMax <= 32, "two bits per element in the per-lane hit mask");
static_assert(kSmemStack >= 1, "the warp's leaf refs live in row 0 of the shared stack");

template <bool COUNT>
__device__ __forceinline__ Hit traceList(const float4* __restrict__ nodesAll, size_t variantStride, const float4* __restrict__ tris,
                                         const Ray& ray, const float tmax, int* sWarp, const int lane, int* nodeCount, int* triCount,
                                         const int* __restrict__ list, const int n)
{
    Hit best;
    best.t = tmax; best.prim = -1; best.u = 0.0f; best.v = 0.0f;
    if (n == 0) {                                                  // the cone reaches no leaf: every sample misses
        if (COUNT) { *nodeCount = 0; *triCount = 0; }
        return best;
    }
    RayBox rb;
    int sx, sy, sz;
    setupAxis(ray.o.x, ray.d.x, rb.nix, rb.fix, rb.nax, rb.fax, sx);
    setupAxis(ray.o.y, ray.d.y, rb.niy, rb.fiy, rb.nay, rb.fay, sy);
    setupAxis(ray.o.z, ray.d.z, rb.niz, rb.fiz, rb.naz, rb.faz, sz);
    const float4* __restrict__ nodes = nodesAll + (size_t)(sx | (sy << 1) | (sz << 2)) * variantStride;
    unsigned bits = 0u;
    __syncwarp();                                                   // the previous frame's phase 2 is done with the refs
    for (int e = 0; e < n; e++) {
        const int w = __ldg(list + 1 + e);                          // node << 2 | reachable-leaf mask of its two children
        const float4* np = nodes + 4 * (size_t)(w >> 2);
        float4 n0, n1, n2, n3;
        ldgNode(np, n0, n1, n2, n3);
        const float tn0 = fmax3(fmaf(n0.x, rb.nix, rb.nax), fmaf(n0.z, rb.niy, rb.nay), fmaxf(fmaf(n2.x, rb.niz, rb.naz), ray.tmin));
        const float tf0 = fmin3(fmaf(n0.y, rb.fix, rb.fax), fmaf(n0.w, rb.fiy, rb.fay), fminf(fmaf(n2.y, rb.fiz, rb.faz), tmax));
        const float tn1 = fmax3(fmaf(n1.x, rb.nix, rb.nax), fmaf(n1.z, rb.niy, rb.nay), fmaxf(fmaf(n2.z, rb.niz, rb.naz), ray.tmin));
        const float tf1 = fmin3(fmaf(n1.y, rb.fix, rb.fax), fmaf(n1.w, rb.fiy, rb.fay), fminf(fmaf(n2.w, rb.fiz, rb.faz), tmax));
        const unsigned h0 = ((w & 1) != 0 && tn0 <= tf0) ? 1u : 0u, h1 = ((w & 2) != 0 && tn1 <= tf1) ? 2u : 0u;
        bits |= (h0 | h1) << (2 * e);
        if (lane == 0) *reinterpret_cast<int2*>(sWarp + 2 * e) = make_int2(__float_as_int(n3.x), __float_as_int(n3.y));
    }
    __syncwarp();
    int tc = 0;
    while (bits != 0u) {
        const int b = __ffs((int)bits) - 1;
        bits &= bits - 1u;
        const int x = ~sWarp[b];                                    // leaf ref of (element b >> 1, child b & 1)
        const int first = x >> 3, cnt = (x & 7) + 1;
        for (int k = 0; k < cnt; k++) {
            float t, u, v;
            int prim;
            if (COUNT) tc++;
            if (triTest(tris + 3 * (size_t)(first + k), ray.o, ray.d, ray.tmin, best.t, t, u, v, prim)) {
                if (t < best.t || best.prim < 0 || prim < best.prim) { best.t = t; best.prim = prim; best.u = u; best.v = v; }
            }
        }
    }
    __syncwarp();                                                   // every lane is done with the refs: the next frame may be a per-lane
    if (COUNT) { *nodeCount = n; *triCount = tc; }                  // walk, whose stack level 0 is this very row (racecheck, r03b)
    return best;
}

// ------------------------------------------------------------------------------------------
// Shading: closest hit (shaders.cu:779-811 live part; cuda/LocalGeometry.h:55-156) and the two
// miss programs (shaders.cu:740-756).  params.lighting is hard-wired false in the reference
// (libEyeRenderer.cpp:98), so the result is the base colour.
// ------------------------------------------------------------------------------------------
template <bool FAST = false>
__device__ __forceinline__ V3 linearize(V3 c) { return mk(Fn<FAST>::pow(c.x, 2.2f), Fn<FAST>::pow(c.y, 2.2f), Fn<FAST>::pow(c.z, 2.2f)); }

template <bool FAST = false>
__device__ __forceinline__ V3 shadeHit(const DeviceScene& sc, const Hit& h)
{
    const uint4 pr = __ldg(sc.prims + h.prim);
    const MeshRec* m = sc.meshes + pr.w;
    const float w0 = 1.0f - h.u - h.v;
    if (m->colorType != -1) {
        const float4 c0 = __ldg(sc.colors + pr.x), c1 = __ldg(sc.colors + pr.y), c2 = __ldg(sc.colors + pr.z);
        const V3 col = mk(w0 * c0.x + h.u * c1.x + h.v * c2.x, w0 * c0.y + h.u * c1.y + h.v * c2.y,
                          w0 * c0.z + h.u * c1.z + h.v * c2.z);
        return linearize<FAST>(col);
    }
    if (m->hasTex) {
        float uu = h.u, vv = h.v;                                     // LocalGeometry.h:97-103
        if (m->hasUV) {
            const float2 t0 = __ldg(sc.uvs + pr.x), t1 = __ldg(sc.uvs + pr.y), t2 = __ldg(sc.uvs + pr.z);
            uu = w0 * t0.x + h.u * t1.x + h.v * t2.x;
            vv = w0 * t0.y + h.u * t1.y + h.v * t2.y;
        }
        const float4 tx = tex2D<float4>((cudaTextureObject_t)m->tex, uu, vv);
        return linearize<FAST>(mk(tx.x, tx.y, tx.z));
    }
    return mk(m->baseColor[0], m->baseColor[1], m->baseColor[2]);
}

template <bool FAST = false>
__device__ __forceinline__ V3 shadeMiss(int shader, V3 rayDir)
{
    const V3 dir = vnormalize(rayDir);
    if (shader == 1) {                                                // __miss__simple_sky
        const float mix = fminf(fmaxf(0.0f, (Fn<FAST>::asin(dir.y) * 2.0f) / crm::kPi), 1.0f);
        const float i255 = 1.0f / 255.0f;                             // sutil float3/float = * reciprocal
        const V3 upper = mk(1.0f * i255, 31.0f * i255, 117.0f * i255);
        const V3 lower = mk((143.0f * i255) * 0.8f, (179.0f * i255) * 0.8f, (203.0f * i255) * 0.8f);
        return vadd(vmuls(lower, 1.0f - mix), vmuls(upper, mix));
    }
    const float border = 0.01f;                                       // __miss__default_background
    if (fabsf(dir.x) < border || fabsf(dir.y) < border || fabsf(dir.z) < border) return mk(0.0f, 0.0f, 0.0f);
    return mk((Fn<FAST>::atan2(dir.z, dir.x) + crm::kPi) / (crm::kPi * 2.0f), (Fn<FAST>::asin(dir.y) + crm::kPi / 2.0f) / crm::kPi, 0.0f);
}

template <bool FAST = false>
__device__ __forceinline__ uchar4 makeColor(float r, float g, float b)   // shaders.cu:180-189
{
    constexpr float ex = static_cast<float>(1.0 / static_cast<double>(2.2f));
    return make_uchar4(static_cast<unsigned char>(Fn<FAST>::pow(fminf(fmaxf(r, 0.0f), 1.0f), ex) * 255.0f),
                       static_cast<unsigned char>(Fn<FAST>::pow(fminf(fmaxf(g, 0.0f), 1.0f), ex) * 255.0f),
                       static_cast<unsigned char>(Fn<FAST>::pow(fminf(fmaxf(b, 0.0f), 1.0f), ex) * 255.0f), 255u);
}
__device__ __forceinline__ uchar4 makeColorMode(bool fast, float r, float g, float b)
{ return fast ? makeColor<true>(r, g, b) : makeColor<false>(r, g, b); }

// ------------------------------------------------------------------------------------------
// K0b.  Entry frontier.  All sample rays of one ommatidium leave ONE origin inside a narrow cone
// around its axis (angle to the axis <= |splay|: a rotation by splay moves a vector by at most
// splay, the second rotation is about the axis itself).  Instead of walking the top of the BVH once
// per sample, one thread per (frame, ommatidium) walks it once per frame with a conservative
// cone/box test and leaves up to kEntryK subtree roots; the samples start there.  The closest
// hit is unchanged: a subtree is dropped only when its box lies wholly outside one face of a
// square pyramid circumscribing the cone widened by 1 % + 1 mrad (orders of magnitude above the
// rounding of either test), so no ray of the cone could have passed the slab test of that box.
// Rays drawn outside kConeSigmas, eyes whose axes are not unit length, poses that are not
// orthonormal (then the bound above does not hold), negative focal offsets (tmin < 0 admits hits behind
// the origin) and cones wider than 1 rad start at the root.
// ------------------------------------------------------------------------------------------
#ifndef CR_ENTRY_BUDGET
#define CR_ENTRY_BUDGET 4
#endif
#ifndef CR_ENTRY_PREFETCH
#define CR_ENTRY_PREFETCH 0   // measured: 46 us instead of 36 us per headline frame with the prefetch (the extra L1 fills cost more than they hide)
#endif
#ifndef CR_ENTRY_V2
#define CR_ENTRY_V2 1         // descent with per-lane slots, ballots and the centre / half-extent cone test (0: the first form)
#endif
constexpr int kEntryK = 4;                      // slots of the int4 record
constexpr int kEntryBudget = CR_ENTRY_BUDGET;   // entries actually handed out (<= kEntryK)

struct ConePyramid { V3 apex, axis, n0, n1, n2, n3; };

__device__ __forceinline__ bool coneMayTouchBox(const ConePyramid& P, const V3 bmin, const V3 bmax)
{
    const V3 lo = vsub(bmin, P.apex), hi = vsub(bmax, P.apex);
    const float scale = fmaxf(fmaxf(fmaxf(fabsf(P.apex.x), fabsf(P.apex.y)), fabsf(P.apex.z)),
                              fmaxf(fmaxf(fmaxf(fabsf(bmin.x), fabsf(bmin.y)), fmaxf(fabsf(bmin.z), fabsf(bmax.x))),
                                    fmaxf(fabsf(bmax.y), fabsf(bmax.z))));
    const float tol = scale * 1.52587890625e-05f;      // 2^-16 of the coordinate magnitude >> rounding of lo/hi
    const V3 n[4] = {P.n0, P.n1, P.n2, P.n3};
#pragma unroll
    for (int i = 0; i < 4; i++) {                       // least signed distance of the box to face i
        const float m = fminf(n[i].x * lo.x, n[i].x * hi.x) + fminf(n[i].y * lo.y, n[i].y * hi.y) + fminf(n[i].z * lo.z, n[i].z * hi.z);
        if (m > tol) return false;                      // (NaN compares false: never culls)
    }
    const float M = fmaxf(P.axis.x * lo.x, P.axis.x * hi.x) + fmaxf(P.axis.y * lo.y, P.axis.y * hi.y) + fmaxf(P.axis.z * lo.z, P.axis.z * hi.z);
    if (M < -tol) return false;                         // wholly behind the apex
    return true;
}

// The same test in centre / half-extent form (round 2, last session): least signed distance of the box to face i =
// n_i.c - |n_i|.h, greatest extent along the axis = axis.c + |axis|.h, with c = box centre - apex, h = half extents.  52
// instead of 91 instructions per child box, and axis.c is the sort key the pass needs anyway.  The tolerance is one constant
// per ommatidium -- 2^-16 of the largest coordinate magnitude of the apex or of ANY box of the scene -- hence never smaller
// than the per-box tolerance above: the test culls a subset of what that one culls, and that one culls no box a ray of the
// cone can touch.  (Its own rounding: a few ulp of the coordinate magnitude, 2^-23, against the 2^-16 of the tolerance.)
struct ConePlanes {
    V3 apex, axis, absAxis;
    V3 n0, n1, n2, n3, a0, a1, a2, a3;      // face normals and their absolute values
    float tol;
};
// (bmin, bmax) may arrive as (near, far) per axis -- the octant copies of the nodes store them that way: the centre is
// symmetric in the two and the half extent is taken as |far - centre|, so no swap back is needed.)
__device__ __forceinline__ bool coneMayTouchBoxCH(const ConePlanes& P, const V3 bmin, const V3 bmax, float& key)
{
    const V3 mid = vmuls(vadd(bmin, bmax), 0.5f);
    const V3 c = vsub(mid, P.apex), d = vsub(bmax, mid), h = mk(fabsf(d.x), fabsf(d.y), fabsf(d.z));
    key = vdot(P.axis, c);                              // (the same expression as the sort key of the first form)
    const float m0 = fdot(P.n0, c) - fdot(P.a0, h), m1 = fdot(P.n1, c) - fdot(P.a1, h);
    const float m2 = fdot(P.n2, c) - fdot(P.a2, h), m3 = fdot(P.n3, c) - fdot(P.a3, h);
    const float M = key + fdot(P.absAxis, h);
    // (NaN compares false: never culls)
    return !(m0 > P.tol) && !(m1 > P.tol) && !(m2 > P.tol) && !(m3 > P.tol) && !(M < -P.tol);
}

// Child boxes of a node read from the copy of direction octant (sx, sy, sz): that copy stores (near, far) per axis, i.e.
// (max, min) on the axes whose sign bit is set (cr_bvh.cu k_emitNodes) -- swap them back.
__device__ __forceinline__ void nodeBoxes(const float4 n0, const float4 n1, const float4 n2, const bool sx, const bool sy, const bool sz,
                                          V3& min0, V3& max0, V3& min1, V3& max1)
{
    min0 = mk(sx ? n0.y : n0.x, sy ? n0.w : n0.z, sz ? n2.y : n2.x);
    max0 = mk(sx ? n0.x : n0.y, sy ? n0.z : n0.w, sz ? n2.x : n2.y);
    min1 = mk(sx ? n1.y : n1.x, sy ? n1.w : n1.z, sz ? n2.w : n2.z);
    max1 = mk(sx ? n1.x : n1.y, sy ? n1.z : n1.w, sz ? n2.z : n2.w);
}

// kEntryK lanes work on one (frame, ommatidium): lane j owns slot j of the frontier, fetches that
// node and tests its two child boxes; the four results are exchanged by shuffles and every lane
// applies the same replacement rules to its replica of the list.  The descent is a chain of
// dependent node fetches, so the slots advancing together cut its length from (entries x depth)
// to about the depth.
struct EntryList {
    int ref[kEntryK];
    float key[kEntryK];
    unsigned fin;      // bit j: slot j is final (a leaf child, or no room to split)
    int n;
};
__device__ __forceinline__ void entryAppend(EntryList& L, int ref, float key, bool fin)
{
#pragma unroll
    for (int k = 0; k < kEntryK; k++)
        if (k == L.n) { L.ref[k] = ref; L.key[k] = key; }
    if (fin) L.fin |= 1u << L.n;
    L.n++;
}

__global__ void __launch_bounds__(128) k_buildEntries(const DeviceScene sc, const EyeParams ep, int4* __restrict__ entries,
                                                      int* __restrict__ lists)
{
    asm volatile("griddepcontrol.launch_dependents;");      // the trace kernel may move in; it waits before it reads the entries
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long idx = tid / kEntryK;
    const int lane = (int)(tid % kEntryK);
    const int F = ep.poses ? ep.nFrames : 1;
    const bool live = idx < (long long)ep.N * F;
    const int f = live ? (int)(idx / ep.N) : 0, o = live ? (int)(idx - (long long)f * ep.N) : 0;
    DevicePose P = ep.pose;
    if (ep.poses) P = ep.poses[f];
    const float4 p0 = __ldg(ep.pre + kPreStride * o), p1 = __ldg(ep.pre + kPreStride * o + 1), p2 = __ldg(ep.pre + kPreStride * o + 2);
    const V3 rp = mk(p0.x, p0.y, p0.z), a = mk(p1.x, p1.y, p1.z);
    const V3 X = mk(P.xx, P.xy, P.xz), Y = mk(P.yx, P.yy, P.yz), Z = mk(P.zx, P.zy, P.zz);
    const float half = p2.w * 1.01f + 1.0e-3f;
    const float tolU = 1.0e-4f;
    bool ok = live && half <= 1.0f;
    ok = ok && p1.w >= 0.0f;                    // tmin = focal offset < 0 lets a ray hit geometry BEHIND its origin: no forward cone
    ok = ok && fabsf(vdot(a, a) - 1.0f) <= tolU;
    ok = ok && fabsf(vdot(X, X) - 1.0f) <= tolU && fabsf(vdot(Y, Y) - 1.0f) <= tolU && fabsf(vdot(Z, Z) - 1.0f) <= tolU;
    ok = ok && fabsf(vdot(X, Y)) <= tolU && fabsf(vdot(X, Z)) <= tolU && fabsf(vdot(Y, Z)) <= tolU;
    ConePyramid C;
    C.apex = vadd(vadd(vadd(mk(P.px, P.py, P.pz), vmuls(X, rp.x)), vmuls(Y, rp.y)), vmuls(Z, rp.z));   // = Ray::o of every sample
    C.axis = vnormalize(vadd(vadd(vmuls(X, a.x), vmuls(Y, a.y)), vmuls(Z, a.z)));
    const V3 t = fabsf(C.axis.x) < 0.57f ? mk(1.0f, 0.0f, 0.0f) : (fabsf(C.axis.y) < 0.57f ? mk(0.0f, 1.0f, 0.0f) : mk(0.0f, 0.0f, 1.0f));
    const V3 u = vnormalize(vcross(C.axis, t));
    const V3 v = vcross(C.axis, u);
    float sh, ch;
    crm::sincos(half, sh, ch);
    const V3 back = vmuls(C.axis, -sh);
    C.n0 = vadd(vmuls(u, ch), back);
    C.n1 = vadd(vmuls(u, -ch), back);
    C.n2 = vadd(vmuls(v, ch), back);
    C.n3 = vadd(vmuls(v, -ch), back);

    // The pass reads the node copy of the CONE AXIS's direction octant -- the copy most of this ommatidium's sample rays will
    // read in the trace kernel (the first form swaps its (near, far) planes back to (min, max), the second form's test is
    // symmetric in them): what it fetches is then warm in L2 for
    // the rays (and, while the camera moves slowly, already warm from the previous frame's rays), instead of living in copy 0
    // that no ray of this cone touches.
    const bool csx = C.axis.x < 0.0f, csy = C.axis.y < 0.0f, csz = C.axis.z < 0.0f;
    const float4* __restrict__ coneNodes = sc.nodes + (size_t)((csx ? 1 : 0) | (csy ? 2 : 0) | (csz ? 4 : 0)) * sc.nodeVariantStride;

    EntryList L;
#if CR_ENTRY_V2
    // Second form of the descent (same replacement rules, about 250 instead of 390 instructions per level -- the pass is a
    // chain of dependent instructions of ~1.7 warps per scheduler, so its time is its instruction count): each lane keeps only
    // ITS slot (ref, key, final); the group's flags travel as ballots, every lane replays the four replacement decisions on
    // those bit fields to learn which (old slot, child) lands in its slot, and pulls that entry with shuffles.
    ConePlanes CP;
    CP.apex = C.apex; CP.axis = C.axis; CP.absAxis = mk(fabsf(C.axis.x), fabsf(C.axis.y), fabsf(C.axis.z));
    CP.n0 = C.n0; CP.n1 = C.n1; CP.n2 = C.n2; CP.n3 = C.n3;
    CP.a0 = mk(fabsf(C.n0.x), fabsf(C.n0.y), fabsf(C.n0.z)); CP.a1 = mk(fabsf(C.n1.x), fabsf(C.n1.y), fabsf(C.n1.z));
    CP.a2 = mk(fabsf(C.n2.x), fabsf(C.n2.y), fabsf(C.n2.z)); CP.a3 = mk(fabsf(C.n3.x), fabsf(C.n3.y), fabsf(C.n3.z));
    CP.tol = fmaxf(fmaxf(fmaxf(fabsf(C.apex.x), fabsf(C.apex.y)), fabsf(C.apex.z)), sc.boundsAbsMax) * 1.52587890625e-05f;
    const int gbase = (int)(threadIdx.x & 31u) & ~(kEntryK - 1);       // first lane of my group within the warp
    int myRef = lane == 0 ? 0 : kSentinel, nList = 1;
    float myKey = 0.0f;
    bool myFin = !ok;                                                  // not ok: the root stays the only entry
    for (int iter = 0; iter < ep.entryMaxLevels; iter++) {             // (stopping early leaves a coarser but equally valid frontier)
        const bool active = lane < nList && !myFin;
        const unsigned bAct = __ballot_sync(0xffffffffu, active);
        if (bAct == 0u) break;                                         // warp-uniform exit
        int r0 = kSentinel, r1 = kSentinel;
        float k0 = 0.0f, k1 = 0.0f;
        bool h0 = false, h1 = false, stop = false;
        if (active) {
            const float4* np = coneNodes + 4 * (size_t)myRef;
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
            r0 = __float_as_int(n3.x); r1 = __float_as_int(n3.y);
            // (near, far) planes as stored in the cone axis's octant copy: see coneMayTouchBoxCH
            h0 = coneMayTouchBoxCH(CP, mk(n0.x, n0.z, n2.x), mk(n0.y, n0.w, n2.y), k0);
            h1 = coneMayTouchBoxCH(CP, mk(n1.x, n1.z, n2.z), mk(n1.y, n1.w, n2.w), k1);
            stop = (h0 && r0 < 0) || (h1 && r1 < 0);                   // never hand out leaves: their triangles would be tested without the per-ray box test
        }
        const unsigned gAct = (bAct >> gbase) & 15u;
        const unsigned gH0 = (__ballot_sync(0xffffffffu, h0) >> gbase) & 15u, gH1 = (__ballot_sync(0xffffffffu, h1) >> gbase) & 15u;
        const unsigned gStop = (__ballot_sync(0xffffffffu, stop) >> gbase) & 15u, gFin = (__ballot_sync(0xffffffffu, myFin) >> gbase) & 15u;
        // (branch-free: the groups of a warp take different paths, so branches here would run every path for everybody)
        int cnt = 0, src = 0, kind = -1;                               // my new entry: old slot src; kind 0 = that entry itself, 1 / 2 = its child 0 / 1
        bool newFin = false;
#pragma unroll
        for (int i = 0; i < kEntryK; i++) {
            const int inList = i < nList ? 1 : 0;
            const int a = (int)((gAct >> i) & 1u), c0 = (int)((gH0 >> i) & 1u), c1 = (int)((gH1 >> i) & 1u);
            const int keep = (a ^ 1) | (int)((gStop >> i) & 1u) | (cnt + c0 + c1 + (nList - 1 - i) > kEntryBudget ? 1 : 0);
            const int nOut = inList * (keep ? 1 : c0 + c1);
            const int pos = lane - cnt;
            const bool mine = pos >= 0 && pos < nOut;
            const int kindHere = keep ? 0 : ((pos == 0 && c0) ? 1 : 2);
            src = mine ? i : src;
            kind = mine ? kindHere : kind;
            newFin = mine ? (keep && (a || ((gFin >> i) & 1u))) : newFin;
            cnt += nOut;
        }
        const int sl = gbase + src;
        const int pRef = __shfl_sync(0xffffffffu, myRef, sl), pR0 = __shfl_sync(0xffffffffu, r0, sl), pR1 = __shfl_sync(0xffffffffu, r1, sl);
        const float pKey = __shfl_sync(0xffffffffu, myKey, sl), pK0 = __shfl_sync(0xffffffffu, k0, sl), pK1 = __shfl_sync(0xffffffffu, k1, sl);
        myRef = kind < 0 ? kSentinel : (kind == 0 ? pRef : (kind == 1 ? pR0 : pR1));
        myKey = kind < 0 ? 0.0f : (kind == 0 ? pKey : (kind == 1 ? pK0 : pK1));
        myFin = newFin;
        nList = cnt;
    }
#pragma unroll
    for (int k = 0; k < kEntryK; k++) {
        L.ref[k] = __shfl_sync(0xffffffffu, myRef, gbase + k);
        L.key[k] = __shfl_sync(0xffffffffu, myKey, gbase + k);
    }
    L.n = nList; L.fin = 0u;
#else
#pragma unroll
    for (int k = 0; k < kEntryK; k++) { L.ref[k] = kSentinel; L.key[k] = 0.0f; }
    L.ref[0] = 0;
    L.n = 1;
    L.fin = ok ? 0u : 1u;                       // not ok: the root stays the only entry
    for (int iter = 0; iter < ep.entryMaxLevels; iter++) {      // (stopping early leaves a coarser but equally valid frontier)
        int myRef = kSentinel;
#pragma unroll
        for (int k = 0; k < kEntryK; k++) if (k == lane) myRef = L.ref[k];
        const bool active = lane < L.n && !((L.fin >> lane) & 1u);
        if (__ballot_sync(0xffffffffu, active) == 0u) break;       // warp-uniform exit
        int r0 = kSentinel, r1 = kSentinel, flags = 0;
        float k0 = 0.0f, k1 = 0.0f;
        if (active) {
            const float4* np = coneNodes + 4 * (size_t)myRef;
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
            V3 min0, max0, min1, max1;
            nodeBoxes(n0, n1, n2, csx, csy, csz, min0, max0, min1, max1);
            r0 = __float_as_int(n3.x); r1 = __float_as_int(n3.y);
#if CR_ENTRY_PREFETCH
            // The descent is a chain of dependent node fetches: start pulling both children into L1 now, so that the cone
            // tests below overlap the next level's memory latency instead of preceding it.
            if (r0 >= 0) { const float4* c = sc.nodes + 4 * (size_t)r0; asm volatile("prefetch.global.L1 [%0];" ::"l"(c)); asm volatile("prefetch.global.L1 [%0];" ::"l"(c + 2)); }
            if (r1 >= 0) { const float4* c = sc.nodes + 4 * (size_t)r1; asm volatile("prefetch.global.L1 [%0];" ::"l"(c)); asm volatile("prefetch.global.L1 [%0];" ::"l"(c + 2)); }
#endif
            const bool h0 = coneMayTouchBox(C, min0, max0), h1 = coneMayTouchBox(C, min1, max1);
            k0 = vdot(C.axis, vsub(vmuls(vadd(min0, max0), 0.5f), C.apex));
            k1 = vdot(C.axis, vsub(vmuls(vadd(min1, max1), 0.5f), C.apex));
            // never hand out leaves: their triangles would be tested without the per-ray box test
            const bool stop = (h0 && r0 < 0) || (h1 && r1 < 0);
            flags = 1 | (h0 ? 2 : 0) | (h1 ? 4 : 0) | (stop ? 8 : 0);
        }
        EntryList Nw;
#pragma unroll
        for (int k = 0; k < kEntryK; k++) { Nw.ref[k] = kSentinel; Nw.key[k] = 0.0f; }
        Nw.n = 0; Nw.fin = 0u;
        const int nOld = L.n;
#pragma unroll
        for (int i = 0; i < kEntryK; i++) {
            const int fl = __shfl_sync(0xffffffffu, flags, i, kEntryK);
            const int c0 = __shfl_sync(0xffffffffu, r0, i, kEntryK), c1 = __shfl_sync(0xffffffffu, r1, i, kEntryK);
            const float q0 = __shfl_sync(0xffffffffu, k0, i, kEntryK), q1 = __shfl_sync(0xffffffffu, k1, i, kEntryK);
            if (i < nOld) {
                const int cnt = ((fl >> 1) & 1) + ((fl >> 2) & 1);
                const bool wasFin = (L.fin >> i) & 1u;
                if (!(fl & 1)) entryAppend(Nw, L.ref[i], L.key[i], wasFin);
                else if ((fl & 8) || Nw.n + cnt + (nOld - 1 - i) > kEntryBudget) entryAppend(Nw, L.ref[i], L.key[i], true);
                else {
                    if (fl & 2) entryAppend(Nw, c0, q0, false);
                    if (fl & 4) entryAppend(Nw, c1, q1, false);
                }
            }
        }
        L = Nw;
    }
#endif
    // near to far along the axis (kEntryK = 4: fixed compare-exchange network, empty slots last); every lane of the
    // group holds the same list and sorts it the same way
#pragma unroll
    for (int k = 0; k < kEntryK; k++) if (k >= L.n) L.key[k] = 3.0e38f;
    auto cswap = [&](int x, int y) {
        if (L.key[x] > L.key[y]) { const float tk = L.key[x]; L.key[x] = L.key[y]; L.key[y] = tk; const int tr = L.ref[x]; L.ref[x] = L.ref[y]; L.ref[y] = tr; }
    };
    cswap(0, 1); cswap(2, 3); cswap(0, 2); cswap(1, 3); cswap(1, 2);
    if (live && lane == 0) entries[idx] = make_int4(L.ref[0], L.ref[1], L.ref[2], L.ref[3]);
    if (lists == nullptr) return;                                  // (uniform)

    // Stage 2: candidate list.  Lane j walks the subtree of entry j depth-first with the same cone test, flattening
    // what the cone can reach into "pre-leaf" elements (node << 2 | mask of its children that are reachable LEAVES);
    // reachable internal children are descended into, nearer child first, so the list runs roughly near to far.  The
    // four partial lists are concatenated in entry order.  The walk gives up -- header kListFallback, K1 then walks the
    // frontier per lane -- as soon as the cone turns out to reach more than kListMax elements, after kListVisits nodes
    // in one subtree, or when the frontier itself was a fallback (ok == false).  Header 0 = the cone reaches no leaf at
    // all (sky): K1 skips traversal.
    constexpr int kListVisits = 28, kListStack = 16;
    int stack[kListStack];
    int el[kListMax];
    int sp = 0, n = 0, visits = 0;
    bool good = ok;
    {
        int myRef = kSentinel;
#pragma unroll
        for (int k = 0; k < kEntryK; k++) if (k == lane && k < L.n) myRef = L.ref[k];
        if (myRef != kSentinel && good) stack[sp++] = myRef;
    }
    while (sp > 0 && good) {
        const int node = stack[--sp];
        if (++visits > kListVisits) { good = false; break; }
        const float4* np = coneNodes + 4 * (size_t)node;
        const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
        V3 min0, max0, min1, max1;
        nodeBoxes(n0, n1, n2, csx, csy, csz, min0, max0, min1, max1);
        const bool h0 = coneMayTouchBox(C, min0, max0), h1 = coneMayTouchBox(C, min1, max1);
        const int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
        const int mask = ((h0 && r0 < 0) ? 1 : 0) | ((h1 && r1 < 0) ? 2 : 0);
        if (mask) {
            if (n == kListMax) { good = false; break; }
            el[n++] = (node << 2) | mask;
        }
        const bool d0 = h0 && r0 >= 0, d1 = h1 && r1 >= 0;
        if (d0 && d1) {
            if (sp + 2 > kListStack) { good = false; break; }
            const float k0 = vdot(C.axis, vsub(vmuls(vadd(min0, max0), 0.5f), C.apex));
            const float k1 = vdot(C.axis, vsub(vmuls(vadd(min1, max1), 0.5f), C.apex));
            const bool near0 = k0 <= k1;
            stack[sp++] = near0 ? r1 : r0;
            stack[sp++] = near0 ? r0 : r1;
        } else if (d0 || d1) {
            if (sp + 1 > kListStack) { good = false; break; }
            stack[sp++] = d0 ? r0 : r1;
        }
    }
    // concatenate: offsets = exclusive prefix of the group's counts; any failure or overflow voids the whole list
    int before = 0, total = 0;
    bool allGood = true;
#pragma unroll
    for (int j = 0; j < kEntryK; j++) {
        const int cj = __shfl_sync(0xffffffffu, n, j, kEntryK);
        const int gj = __shfl_sync(0xffffffffu, good ? 1 : 0, j, kEntryK);
        if (j < lane) before += cj;
        total += cj;
        allGood = allGood && gj != 0;
    }
    allGood = allGood && total <= kListMax;
    if (!live) return;
    int* out = lists + (size_t)idx * kListStride;
    if (allGood)
        for (int i = 0; i < n; i++) out[1 + before + i] = el[i];
    if (lane == 0) out[0] = allGood ? total : kListFallback;
}

// ------------------------------------------------------------------------------------------
// K1.  Persistent warps; work unit = 32 consecutive sample rays r = o*S + s (lanes hold
// consecutive samples of one ommatidium -> coherent rays, one coalesced 1 KB RNG-state read and
// one coalesced 384 B colour write per warp).  No block-level synchronisation.
//   FUSED = false: per-sample colour/S (shaders.cu:730) goes to the [o][s] sample buffer and K1b sums it IN SAMPLE
//                  ORDER, which is exactly the reference's sequential fp32 sum (getSummedOmmatidiumData, :341-347).
//   FUSED = true : (S % 32 == 0) the 32 samples of the warp are summed here by a shuffle butterfly and lane 0 writes
//                  ONE float4 partial per warp and frame to partials[f][o][s/32]; k_sumPartials adds the S/32
//                  partials of a row in a fixed order.  12 B/ray of sample traffic (written here, read back by K1b:
//                  ~85 % of this path's DRAM bytes) become 0.5 B/ray.  The order of the additions differs from the
//                  reference's, so the float RGB agrees to rounding (~1e-7 relative), not bit for bit; it is a
//                  FIXED order, restated by the checker (oracle.fused_sum), so the mode is still reproducible.
// Traversal: warps whose 32 lanes are samples of one ommatidium with a candidate list (k_buildEntries, stage 2) test
// that list (traceList); everything else walks the BVH per lane from the entry frontier or the root (traceClosest).
// ------------------------------------------------------------------------------------------
template <bool DUMP, bool MULTI, bool FUSED, bool FAST, bool GROUPED = false>
__global__ void __launch_bounds__(kTraceThreads, CR_TRACE_MIN_BLOCKS) k_traceCompound(const DeviceScene sc, const EyeParams ep)
{
    __shared__ int sStack[kSmemStack][kTraceThreads];
    __shared__ uint4 sRng[MULTI ? 2 : 1][MULTI ? kTraceThreads : 1];   // RNG state parked here while a ray is traced (batches)
    const unsigned total = (unsigned)ep.N * (unsigned)ep.S;
    const float invS = 1.0f / (float)(uint32_t)ep.S;
    const unsigned stride = gridDim.x * kTraceThreads;
    const int F = MULTI ? ep.nFrames : 1;
    const int lane = (int)(threadIdx.x & 31u);
    int* sWarp = &sStack[0][threadIdx.x & ~31u];
    // S % 32 == 0: the 32 rays of a warp are samples of ONE ommatidium (and `r < total` is warp-uniform)
    const bool warpRows = (ep.S & 31) == 0;
    // Work distribution (round 2): a warp's unit is 32 consecutive rays (x F frames in a batch); units are handed out in
    // CHUNKS of ep.chunkUnits (1 in a batch) -- the first chunk of every warp by its position in the grid, the following
    // ones from a global counter (one atomic per chunk, issued a whole chunk before its answer is needed).  With the static grid-stride split the warps finished up to 15 % apart (sky, listed, walking and queued units
    // cost between 1x and 6x; sm__warps_active 41.8 of 50 %, profiles/r02j_k1_queue_ncu_summary.txt).
    const unsigned gridWarps = gridDim.x * (kTraceThreads / 32u);
    // Small frames (batches): the launch's F frames are cut into G groups of groupFrames (an EVEN number of) frames and a unit
    // is (32 rays, one group) -- G times as many units for the counter to balance.  The streams of a group start where the
    // previous group leaves them: its lanes load the launch's starting state and step it over the draws of the frames before
    // their group (2 draws per frame on average: 3 + 1 per frame pair, the Box-Muller cache empty at every even frame --
    // the renderer only splits launches that start at an even frame).  Only the last group writes states, to rngOut.
    const unsigned rayUnits = (total + 31u) >> 5;
    const unsigned groups = (MULTI && GROUPED) ? (unsigned)max(1, ep.frameGroups) : 1u;
    const unsigned nUnits = rayUnits * groups;
    const unsigned chunkUnits = (MULTI || ep.smSeq != nullptr) ? 1u : (unsigned)max(1, ep.chunkUnits);
    const unsigned nChunks = (nUnits + chunkUnits - 1u) / chunkUnits;
    // Two chunks are known ahead (the first two of every warp by its position in the grid): while chunk i is traced, the
    // RNG state of chunk i+1 is already on its way to L2 and the counter is asked for chunk i+2.
    //
    // SM-affine hand-out (ep.smSeq != nullptr; chunkUnits = 1): the global counter hands out BLOCKS of 32 consecutive units and
    // a block belongs to one slot (an SM).  A warp draws a ticket k from its slot's counter: block
    // j = k / 32 of the slot, unit k % 32 of that block.  Whoever draws the first ticket of a block fetches the block's number
    // from the global counter and publishes it in smTab[slot][j] tagged with the launch's epoch (so the table is never cleared);
    // the other 31 ticket holders read it there -- one or two units later, because tickets, like chunks, are drawn two ahead.
    // Every unit is still traced exactly once and a ray's result does not depend on who traces it: no bit changes.
    // Termination: block numbers need not grow with j (two warps may reach the global counter out of order), so a warp leaves
    // only when BOTH units it holds are beyond the end, and draws a new ticket only while its current unit is real -- the warp
    // that holds the last real ticket of a slot keeps drawing until the slot's sequence runs into a block beyond the end.
    // Tickets per slot <= units + 2 per warp, hence j < units/32 + gridWarps/16 + 2 <= smCap (sized by the host).
    const bool smMode = ep.smSeq != nullptr;                               // (uniform)
    asm volatile("griddepcontrol.launch_dependents;");                    // (the reduction kernel waits for this grid's completion itself)
    // (the slot is read once: re-reading %smid per ticket measured slower -- 23.4 vs 24.3 Grays/s batched -- and a first version
    //  that derived it with two integer modulos per unit spent 5 % of the per-frame kernel's instructions on them)
    unsigned mySlot = 0u;
    if (smMode) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        mySlot = min(smid, ep.smSlots - 1u);                               // (smSlots = %nsmid: the clamp never acts)
    }
    auto takeTicket = [&]() -> unsigned {                                  // (lane 0)
        const unsigned slot = mySlot;
        const unsigned k = atomicAdd(ep.smSeq + slot, 1u);
        if ((k & 31u) == 0u) {
            const unsigned nb = atomicAdd(ep.workCounter, 1u);
            if ((k >> 5) >= ep.smCap) __trap();                            // (cannot happen: see the bound above)
            __stcg(ep.smTab + (size_t)slot * ep.smCap + (k >> 5), ((unsigned long long)ep.smEpoch << 32) | nb);
        }
        return k;
    };
    auto unitOfTicket = [&](unsigned k) -> unsigned {                      // (lane 0)
        if (k == 0xffffffffu) return 0xffffffffu;
        const unsigned nBlocks32 = (nUnits + 31u) >> 5;
        const volatile unsigned long long* p = ep.smTab + (size_t)mySlot * ep.smCap + (k >> 5);
        unsigned long long v;
        do { v = *p; } while ((unsigned)(v >> 32) != ep.smEpoch);
        const unsigned nb = (unsigned)v;
        return nb < nBlocks32 ? (nb << 5) + (k & 31u) : 0xffffffffu;
    };
    unsigned chunk = blockIdx.x * (kTraceThreads / 32u) + (threadIdx.x >> 5);
    unsigned nextChunk = chunk + gridWarps;
    if (smMode) {                                                          // nextChunk holds a TICKET until the top of the loop
        unsigned t0 = 0u, t1 = 0u;
        if (lane == 0) { t0 = takeTicket(); t1 = takeTicket(); t0 = unitOfTicket(t0); }
        chunk = __shfl_sync(kFullMask, t0, 0);
        nextChunk = t1;
    }
    // Programmatic dependent launch: everything above ran beside the frontier pass; its entries are read from here on.
    // (waiting only before the first read of the entries -- a unit's state load and ray construction ahead of it -- measured the
    //  same: 0.550 vs 0.551 ms per frame)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (;;) {                                                             // (every branch on chunk / nextChunk is warp-uniform)
      if (smMode) {
          unsigned u = 0u;
          if (lane == 0) u = unitOfTicket(nextChunk);
          nextChunk = __shfl_sync(kFullMask, u, 0);
          if (chunk >= nChunks && nextChunk >= nChunks) break;
      } else if (chunk >= nChunks) break;
      unsigned afterNext = nextChunk + gridWarps;                           // static split when there is no counter
      if (smMode) { afterNext = 0xffffffffu; if (lane == 0 && chunk < nChunks) afterNext = takeTicket(); }
      else if (ep.workCounter != nullptr && lane == 0) afterNext = atomicAdd(ep.workCounter, 1u) + 2u * gridWarps;
      if (groups == 1u) {
          const size_t rn = ((size_t)nextChunk * chunkUnits << 5) + (unsigned)lane;
          if (rn < total) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.rng + 2 * rn));
      }
      if (chunk < nChunks)
      for (unsigned k = 0; k < chunkUnits; k++) {
        const unsigned unit = chunk * chunkUnits + k;
        if (unit >= nUnits) break;                                          // (warp-uniform)
        const unsigned group = (GROUPED && groups > 1u) ? unit / rayUnits : 0u;   // group-major: every ray unit of group 0 first
        const unsigned r0 = ((unit - group * rayUnits) << 5) + (unsigned)lane;
#if CR_INLINE_PHASED
        // a warp stays whole (the phased traversal votes across its lanes): the lanes of the last warp beyond the last
        // ray (N*S % 32 != 0) redo ray total-1 and store nothing
        const bool valid = r0 < total;
        const unsigned r = valid ? r0 : total - 1u;
#else
        constexpr bool valid = true;
        const unsigned r = r0;
        if (r0 < total) {
#endif
        if (k + 1u < chunkUnits && r0 + 32u < total) {                      // warm L2 with the next unit's RNG state
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.rng + 2 * (size_t)(r0 + 32u)));
        }
        const unsigned o = r / (unsigned)ep.S;
        uint4* statePtr = ep.rng + 2 * (size_t)r;
#if CR_EXP_STATE >= 2      // timing experiment only (wrong streams): no state load
        Rng rng;
        rng.d = r; rng.v0 = r * 2654435761u; rng.v1 = r ^ 0x9e3779b9u; rng.v2 = r + 77u; rng.v3 = ~r; rng.v4 = r * 40503u + 1u; rng.flag = 0; rng.extra = 0.0f;
#else
        Rng rng = rngLoad(statePtr);
#endif
        const int fBegin = (MULTI && GROUPED) ? (int)group * ep.groupFrames : 0;
        const int fEnd = (MULTI && GROUPED) ? min(F, fBegin + (groups > 1u ? ep.groupFrames : F)) : F;
        if (MULTI && GROUPED && group > 0u) {                      // the draws of the frames before this group
            for (int i = 2 * fBegin; i > 0; i--) (void)rngNext(rng);
        }
        // Consecutive frames of one sample stream are processed back to back by the same lane: the
        // state is loaded and stored once per batch, and one launch covers F frames (poses).
        for (int f = fBegin; f < fEnd; f++) {
            DevicePose pose = ep.pose;
            if (MULTI) {
                const float4* pp = reinterpret_cast<const float4*>(ep.poses + f);
                const float4 a = __ldg(pp), b = __ldg(pp + 1), c = __ldg(pp + 2);
                pose.px = a.x; pose.py = a.y; pose.pz = a.z; pose.xx = a.w;
                pose.xy = b.x; pose.xz = b.y; pose.yx = b.z; pose.yy = b.w;
                pose.yz = c.x; pose.zx = c.y; pose.zy = c.z; pose.zz = c.w;
            }
            const float4* pp4 = ep.pre + kPreStride * (size_t)o;        // (re-read per frame: L1 hits, no registers held across the walk)
            const float4 p0 = __ldg(pp4), p1 = __ldg(pp4 + 1), p2 = __ldg(pp4 + 2), p3 = __ldg(pp4 + 3);
            float splay;
            const Ray ray = ommatidialRay<FAST>(p0, p1, p2, p3, pose, rng, splay);
            const bool inCone = fabsf(splay) <= p2.w;
            if (MULTI) {   // park the state in shared memory: its 8 registers are dead while the ray is traced
                sRng[0][threadIdx.x] = make_uint4(rng.d, rng.v0, rng.v1, rng.v2);
                sRng[1][threadIdx.x] = make_uint4(rng.v3, rng.v4, (uint32_t)rng.flag, __float_as_uint(rng.extra));
            } else if (valid) {
#if CR_EXP_STATE < 1       // timing experiment only: no state store
                rngStore(statePtr, rng);
#endif
            }
            int nNode = 0, nTri = 0;
            Hit h;
            // Candidate list of this (frame, ommatidium): used when the warp is 32 samples of that ommatidium and every
            // one of them fell inside the kConeSigmas cone the list was built for (else: the per-lane walk below).
            const int* list = nullptr;
            int listN = kListFallback;
            if (warpRows && ep.lists != nullptr) {
                list = ep.lists + ((size_t)f * ep.entryFrameStride + o) * kListStride;
                listN = __ldg(list);
                if (listN != kListFallback && !__all_sync(kFullMask, inCone)) listN = kListFallback;
            }
            bool queued = false;
            if (listN != kListFallback) {
                h = traceList<DUMP>(sc.nodes, sc.nodeVariantStride, sc.tris, ray, kTMax, sWarp, lane, &nNode, &nTri, list, listN);
            } else {
                // No list: this warp-frame's 32 rays walk the BVH one by one -- the walk whose lanes drift apart.  In batches
                // the whole warp-frame goes to the wavefront queue instead (k_traceQueue refills finished lanes with new
                // rays; k_shadeQueue shades and reduces in this warp's lane order, so the result bits do not change).
                if (!DUMP && ep.queueRays != nullptr && list != nullptr) {             // warp-uniform
                    unsigned base = 0u;
                    if (lane == 0) base = atomicAdd(ep.queueCounters, 32u);
                    base = __shfl_sync(kFullMask, base, 0);
                    if (base + 32u <= ep.queueCap) {
                        // record: origin, tmin | direction, (frame*N + ommatidium) with the in-cone flag on top; the warp's
                        // position in its row (block of 32 samples) goes to the per-warp header for k_shadeQueue
                        const unsigned slot = base + (unsigned)lane;
                        const unsigned id = ((unsigned)f * (unsigned)ep.N + o) | (inCone ? 0x80000000u : 0u);
                        __stcs(ep.queueRays + 2 * (size_t)slot, make_float4(ray.o.x, ray.o.y, ray.o.z, ray.tmin));
                        __stcs(ep.queueRays + 2 * (size_t)slot + 1, make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(id)));
                        if (lane == 0) ep.queueWarps[base >> 5] = (int)((r - o * (unsigned)ep.S) >> 5);
                        queued = true;
                    }
                }
                if (!queued) {
                    int4 entry = make_int4(0, kSentinel, kSentinel, kSentinel);
                    if (ep.entries != nullptr && inCone) entry = __ldg(ep.entries + (size_t)f * ep.entryFrameStride + o);
#if CR_INLINE_PHASED
                    h = traceClosestPhased<DUMP>(sc.nodes, sc.nodeVariantStride, sc.tris, ray, kTMax, &sStack[0][threadIdx.x], kTraceThreads,
                                                 &nNode, &nTri, entry, ep.nodeLanes);
#else
                    if (entry.x == kSentinel) {             // empty frontier (the cone reaches nothing: sky): no box-test set-up either
                        h.t = kTMax; h.prim = -1; h.u = 0.0f; h.v = 0.0f;
                    } else {
                        h = traceClosest<DUMP>(sc.nodes, sc.nodeVariantStride, sc.tris, ray, kTMax, &sStack[0][threadIdx.x], kTraceThreads,
                                               &nNode, &nTri, entry);
                    }
#endif
                }
            }
            if (queued) {                                   // (warp-uniform) shaded and reduced by k_shadeQueue
                if (MULTI) {
                    const uint4 a = sRng[0][threadIdx.x], b = sRng[1][threadIdx.x];
                    rng.d = a.x; rng.v0 = a.y; rng.v1 = a.z; rng.v2 = a.w; rng.v3 = b.x; rng.v4 = b.y;
                    rng.flag = (int)b.z; rng.extra = __uint_as_float(b.w);
                }
                continue;
            }
            const V3 col = (h.prim >= 0) ? shadeHit<FAST>(sc, h) : shadeMiss<FAST>(sc.missShader, ray.d);
            float cx = col.x * invS, cy = col.y * invS, cz = col.z * invS;                       // shaders.cu:730
            if (FUSED) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {        // butterfly: level d adds lane i and lane i^d; all lanes end with the same sum
                    cx += __shfl_xor_sync(kFullMask, cx, d);
                    cy += __shfl_xor_sync(kFullMask, cy, d);
                    cz += __shfl_xor_sync(kFullMask, cz, d);
                }
                if (lane == 0) {
                    const unsigned blocksPerRow = (unsigned)ep.S >> 5;
                    const unsigned blk = (r - o * (unsigned)ep.S) >> 5;
                    __stcs(ep.partials + ((size_t)f * (unsigned)ep.N + o) * blocksPerRow + blk, make_float4(cx, cy, cz, 0.0f));
                }
            } else if (valid) {
                float* dst = ep.samples + 3 * ((size_t)f * total + r);
                __stcs(dst, cx); __stcs(dst + 1, cy); __stcs(dst + 2, cz);
            }
            if (MULTI) {
                const uint4 a = sRng[0][threadIdx.x], b = sRng[1][threadIdx.x];
                rng.d = a.x; rng.v0 = a.y; rng.v1 = a.z; rng.v2 = a.w; rng.v3 = b.x; rng.v4 = b.y;
                rng.flag = (int)b.z; rng.extra = __uint_as_float(b.w);
            }
            if (DUMP && valid) {
                const unsigned s = r - o * (unsigned)ep.S;
                const size_t id = (size_t)ep.N * s + o;                                 // reference stream-id order
                ep.dumpOrigins[3 * id] = ray.o.x; ep.dumpOrigins[3 * id + 1] = ray.o.y; ep.dumpOrigins[3 * id + 2] = ray.o.z;
                ep.dumpDirs[3 * id] = ray.d.x; ep.dumpDirs[3 * id + 1] = ray.d.y; ep.dumpDirs[3 * id + 2] = ray.d.z;
                ep.dumpHits[id] = make_int4(h.prim, __float_as_int(h.t), __float_as_int(h.u), __float_as_int(h.v));
                if (ep.dumpCounts) ep.dumpCounts[id] = make_int2(nNode, nTri);   // nodes fetched / triangles tested by THIS kernel
            }
        }
        if (MULTI && valid && (!GROUPED || group + 1u == groups)) rngStore(ep.rngOut + 2 * (size_t)r, rng);
#if !CR_INLINE_PHASED
        }
#endif
      }
      chunk = nextChunk;
      nextChunk = __shfl_sync(kFullMask, afterNext, 0);
    }
    // The last warp of the grid to leave rearms the counters for the next launch (no memset between frames).
    if (ep.workCounter != nullptr && lane == 0) {
        if (atomicAdd(ep.workCounter + 1, 1u) == gridWarps - 1u) {
            atomicExch(ep.workCounter, 0u); atomicExch(ep.workCounter + 1, 0u);
            if (smMode) for (unsigned i = 0; i < ep.smSlots; i++) atomicExch(ep.smSeq + i, 0u);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Wavefront queue (batches).  K1 pushes the warp-frames it would have walked per lane; k_traceQueue traces them with
// DYNAMIC RAY FETCH (persistent warps whose lanes pull the next ray of the queue when fewer than ep.queueRefillBelow lanes
// are still busy) and PHASE SWITCHING (the node loop is left as soon as fewer than ep.nodeLanes lanes want a node while
// somebody holds a leaf), so the node loop stays populated however unequal the walks are -- the grazing ommatidia this
// path exists for visit 5 to 60 nodes per ray.  The closest hit does
// not depend on which lane or kernel finds it (same box test, same triangle test, lowest-primitive tie rule), and
// k_shadeQueue turns hits into colours and sums them warp-frame by warp-frame in K1's lane order: same output bits.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTraceThreads, CR_TRACE_MIN_BLOCKS) k_traceQueue(const DeviceScene sc, const EyeParams ep)
{
    __shared__ int sStack[kSmemStack][kTraceThreads];
    const unsigned n = min(ep.queueCounters[0], ep.queueCap & ~31u);
    const int lane = (int)(threadIdx.x & 31u);
    const int nodeLanes = ep.nodeLanes, refillBelow = ep.queueRefillBelow;
    int deep[kLocalStack];
    Stack st;
    st.smem = &sStack[0][threadIdx.x]; st.local = deep; st.sp = 0;
    RayBox rb;
    rb.nix = rb.niy = rb.niz = rb.fix = rb.fiy = rb.fiz = rb.nax = rb.nay = rb.naz = rb.fax = rb.fay = rb.faz = 0.0f;
    V3 ro = mk(0.0f, 0.0f, 0.0f), rd = mk(0.0f, 0.0f, 1.0f);
    float tmin = 0.0f;
    const float4* __restrict__ nodes = sc.nodes;
    Hit best;
    best.t = kTMax; best.prim = -1; best.u = 0.0f; best.v = 0.0f;
    int cur = kSentinel;
    unsigned slot = 0u;
    bool active = false, exhausted = false;
    int tc = 0;
    for (;;) {
        // ---- retire the finished rays, hand new ones to the idle lanes
        if (active && cur == kSentinel) {
            __stcs(ep.queueHits + slot, make_int4(best.prim, __float_as_int(best.t), __float_as_int(best.u), __float_as_int(best.v)));
            active = false;
        }
        unsigned live = __ballot_sync(kFullMask, active);
        if (!exhausted && __popc(live) < refillBelow) {
            const unsigned need = ~live;
            const int leader = __ffs((int)need) - 1;
            unsigned base = 0u;
            if (lane == leader) base = atomicAdd(ep.queueCounters + 1, (unsigned)__popc(need));
            base = __shfl_sync(kFullMask, base, leader);
            if (base + (unsigned)__popc(need) >= n) exhausted = true;
            const unsigned my = base + (unsigned)__popc(need & ((1u << lane) - 1u));
            if (!active && my < n) {
                const float4 q0 = __ldcs(ep.queueRays + 2 * (size_t)my), q1 = __ldcs(ep.queueRays + 2 * (size_t)my + 1);
                ro = mk(q0.x, q0.y, q0.z); tmin = q0.w; rd = mk(q1.x, q1.y, q1.z);
                const unsigned idw = __float_as_uint(q1.w);
                int sx, sy, sz;
                setupAxis(ro.x, rd.x, rb.nix, rb.fix, rb.nax, rb.fax, sx);
                setupAxis(ro.y, rd.y, rb.niy, rb.fiy, rb.nay, rb.fay, sy);
                setupAxis(ro.z, rd.z, rb.niz, rb.fiz, rb.naz, rb.faz, sz);
                nodes = sc.nodes + (size_t)(sx | (sy << 1) | (sz << 2)) * sc.nodeVariantStride;
                best.t = kTMax; best.prim = -1; best.u = 0.0f; best.v = 0.0f;
                st.sp = 0;
                int4 entry = make_int4(0, kSentinel, kSentinel, kSentinel);
                if (ep.entries != nullptr && (idw & 0x80000000u)) entry = __ldg(ep.entries + (idw & 0x7fffffffu));
                cur = entry.x;
                if (entry.y != kSentinel) {
                    if (entry.z != kSentinel) {
                        if (entry.w != kSentinel) st.push(entry.w);
                        st.push(entry.z);
                    }
                    st.push(entry.y);
                }
                slot = my;
                active = true;
            }
            live = __ballot_sync(kFullMask, active);
        }
        if (live == 0u) break;
        // ---- leaf phase: the lanes that hold a leaf test its triangles and pop
        if (cur < 0 && cur != kSentinel) {
            leafStep(sc.tris, cur, ro, rd, tmin, best, tc);
            cur = st.pop();
        }
        // ---- node phase: one vote per step; below the threshold the warp leaves as soon as somebody holds a leaf or has
        //      finished and can be replaced (a second vote, taken only then)
        for (;;) {
            const int nWant = __popc(__ballot_sync(kFullMask, cur >= 0));
            if (nWant < nodeLanes) {
                if (nWant == 0) break;
                if (__any_sync(kFullMask, cur < 0 && (cur != kSentinel || (active && !exhausted)))) break;
            }
            if (cur >= 0) cur = nodeStep(nodes, cur, rb, tmin, best.t, st);
        }
    }
}

// Shading + reduction of the queued warp-frames: 32 consecutive queue slots = the 32 lanes of the K1 warp that pushed them.
template <bool FUSED, bool FAST>
__global__ void __launch_bounds__(128) k_shadeQueue(const DeviceScene sc, const EyeParams ep)
{
    const unsigned n = min(ep.queueCounters[0], ep.queueCap & ~31u);
    const float invS = 1.0f / (float)(uint32_t)ep.S;
    const int lane = (int)(threadIdx.x & 31u);
    for (unsigned q = blockIdx.x * 128u + threadIdx.x; q < n; q += gridDim.x * 128u) {        // n % 32 == 0: warp-uniform
        const int4 hw = __ldcs(ep.queueHits + q);
        const float4 q1 = __ldcs(ep.queueRays + 2 * (size_t)q + 1);
        Hit h;
        h.prim = hw.x; h.t = __int_as_float(hw.y); h.u = __int_as_float(hw.z); h.v = __int_as_float(hw.w);
        const V3 col = (h.prim >= 0) ? shadeHit<FAST>(sc, h) : shadeMiss<FAST>(sc.missShader, mk(q1.x, q1.y, q1.z));
        float cx = col.x * invS, cy = col.y * invS, cz = col.z * invS;
        const size_t row = __float_as_uint(q1.w) & 0x7fffffffu;                                // frame*N + ommatidium
        const unsigned blk = (unsigned)__ldg(ep.queueWarps + (q >> 5));                         // block of 32 samples within the row
        if (FUSED) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                cx += __shfl_xor_sync(kFullMask, cx, d);
                cy += __shfl_xor_sync(kFullMask, cy, d);
                cz += __shfl_xor_sync(kFullMask, cz, d);
            }
            if (lane == 0) __stcs(ep.partials + row * ((unsigned)ep.S >> 5) + blk, make_float4(cx, cy, cz, 0.0f));
        } else {
            float* dst = ep.samples + 3 * (row * (unsigned)ep.S + 32u * blk + (unsigned)lane);
            __stcs(dst, cx); __stcs(dst + 1, cy); __stcs(dst + 2, cz);
        }
    }
}

// K1b of the fused mode: summed[f][o] = fixed-order sum of the S/32 warp partials of the row, one warp per row.
// Lane l first adds partials l, l+32, l+64, ... in ascending order, then the same butterfly as in K1 combines the 32
// lanes (S <= 1024: one partial per lane).  Also writes the 8-bit row of single_dimension_fast / pose batches.
// Rows with few partials (S <= 512) share a warp: a row takes W = the power of two >= S/32 lanes, 32/W rows per warp.
// The value is that of the one-warp-per-row form bit for bit: there the lanes beyond S/32 hold +0.0f, so the butterfly
// levels d >= W only add zeros (x + 0.0f = x: the partial sums of colours are never -0.0f) and the levels d < W pair the
// same lanes in the same order as the sub-warp butterfly here.  (S = 64: 16 rows per warp; one warp per row spent
// 0.73 ms per 256-frame launch of the 6 374-ommatidia eye on two partials per row, profiles/r02v_cfg5_metrics.csv.)
constexpr int kSumPartialWarps = 32;         // warps per CTA of k_sumPartials
template <bool FAST>
__global__ void __launch_bounds__(32 * kSumPartialWarps) k_sumPartials(const float4* __restrict__ partials, int NF, int blocksPerRow,
                                                                       int lanesPerRow, float4* __restrict__ summed,
                                                                       uchar4* __restrict__ fastRow, int fastRowCount,
                                                                       uchar4* __restrict__ fastRowHost)
{
    __shared__ uchar4 sPx[32 * kSumPartialWarps];
    asm volatile("griddepcontrol.wait;" ::: "memory");      // (programmatic dependent launch: the partials of the trace kernel)
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
    const int rowsPerWarp = 32 / lanesPerRow, rowsPerCta = rowsPerWarp * kSumPartialWarps;
    const int row0 = (int)blockIdx.x * rowsPerCta;
    const int sub = lane / lanesPerRow, l = lane - sub * lanesPerRow;      // my row within the warp, my lane within the row
    const int rowInCta = warp * rowsPerWarp + sub, row = row0 + rowInCta;
    float cx = 0.0f, cy = 0.0f, cz = 0.0f;
    if (row < NF) {
        const float4* p = partials + (size_t)row * blocksPerRow;
        for (int k = l; k < blocksPerRow; k += lanesPerRow) {               // (lanesPerRow == 32 whenever blocksPerRow > 32)
            const float4 q = __ldcs(p + k);
            cx += q.x; cy += q.y; cz += q.z;
        }
    }
    for (int d = lanesPerRow >> 1; d > 0; d >>= 1) {                        // (all 32 lanes take part: uniform trip count)
        cx += __shfl_xor_sync(kFullMask, cx, d);
        cy += __shfl_xor_sync(kFullMask, cy, d);
        cz += __shfl_xor_sync(kFullMask, cz, d);
    }
    if (row < NF && l == 0) summed[row] = make_float4(cx, cy, cz, 0.0f);
    if (fastRow != nullptr) {                                // (uniform)
        if (lanesPerRow >= 4) {
            // every lane of the row holds the three sums: lanes 0..2 gamma-encode one channel each (the pow is ~60 instructions)
            constexpr float ex = static_cast<float>(1.0 / static_cast<double>(2.2f));
            const float c = l == 0 ? cx : (l == 1 ? cy : cz);
            const unsigned v = l < 3 ? (unsigned)static_cast<unsigned char>(Fn<FAST>::pow(fminf(fmaxf(c, 0.0f), 1.0f), ex) * 255.0f) : 0u;
            const int base = lane - l;
            const unsigned g = __shfl_sync(kFullMask, v, base + 1), b = __shfl_sync(kFullMask, v, base + 2);
            if (row < NF && l == 0) sPx[rowInCta] = make_uchar4((unsigned char)v, (unsigned char)g, (unsigned char)b, 255u);
        } else if (row < NF && l == 0) {
            sPx[rowInCta] = makeColor<FAST>(cx, cy, cz);
        }
    }
    if (fastRow == nullptr) return;                          // (uniform)
    __syncthreads();
    // the CTA's pixels as coalesced stores -- to the device frame and, when the caller reads every frame, straight into the
    // pinned (mapped) host frame: 128-byte PCIe writes instead of a copy queued behind the kernel (or 10 000 writes of
    // 4 bytes: 22 us)
    const int t = (int)threadIdx.x;
    if (t < rowsPerCta && row0 + t < NF && row0 + t < fastRowCount) {
        const uchar4 px = sPx[t];
        fastRow[row0 + t] = px;
        if (fastRowHost != nullptr) fastRowHost[row0 + t] = px;
    }
}

// K1b: summed[f][o].ch = ((0 + c[f][o][0]) + c[f][o][1]) + ... + c[f][o][S-1]: the reference's sequential
// fp32 order (shaders.cu:341-347), so each (row, channel) is one dependent chain of S additions.
// Two forms: the TMA kernel below (k_sumSamplesTma, bulk copies + mbarrier) whenever S % 4 == 0, and this
// cp.async form for every other S (row segments that are not 16-byte aligned).
// A CTA owns kSumRows consecutive (frame, ommatidium) rows.  Its 96 threads stream tiles of
// kSumChunk samples per row into shared memory with cp.async -- coalesced 384-byte pieces of the
// 12*S contiguous bytes of a row ([row][s][3] layout), two tiles in flight, no registers held --
// and the first 3*kSumRows threads (row, channel) walk their chains through the landed tile.
// Row stride = 3 mod 32 floats: those 24 lanes hit 24 different banks.
constexpr int kSumRows = 8, kSumChunk = 128, kSumThreads = 96, kSumStride = 3 * kSumChunk + 3, kSumStages = 3;

__device__ __forceinline__ void cpAsync4(float* smemDst, const float* gmemSrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smemDst)), "l"(gmemSrc));
}

template <bool FAST>
__global__ void __launch_bounds__(kSumThreads) k_sumSamples(const float* __restrict__ samples, int NF, int S, float4* __restrict__ summed,
                                                            uchar4* __restrict__ fastRow, int fastRowCount, uchar4* __restrict__ fastRowHost)
{
    __shared__ float tile[kSumStages][kSumRows * kSumStride];
    const int row0 = blockIdx.x * kSumRows;
    const int rows = min(kSumRows, NF - row0);
    const int t = threadIdx.x;
    const size_t rowFloats = 3 * (size_t)S;
    const float* base = samples + rowFloats * (size_t)row0;
    const int nChunks = (S + kSumChunk - 1) / kSumChunk;
    auto issue = [&](int c) {
        if (c < nChunks) {
            const int nf = 3 * min(kSumChunk, S - c * kSumChunk);
            const float* src = base + 3 * (size_t)c * kSumChunk;
            float* dst = tile[c % kSumStages];
            for (int r = 0; r < rows; r++)
#pragma unroll
                for (int k = t; k < 3 * kSumChunk; k += kSumThreads)
                    if (k < nf) cpAsync4(dst + r * kSumStride + k, src + rowFloats * r + k);
        }
        asm volatile("cp.async.commit_group;");
    };
    issue(0);
    issue(1);
    const int myRow = t / 3, ch = t - 3 * myRow;
    const bool summing = myRow < rows;               // threads 0 .. 3*rows-1, all in warp 0
    float sum = 0.0f;
    for (int c = 0; c < nChunks; c++) {
        asm volatile("cp.async.wait_group 1;");        // tile c has landed (tile c+1 may still be in flight)
        __syncthreads();                                // ... for every thread; and everyone is done with tile c-1
        issue(c + 2);                                   // reuses the stage of tile c-1
        if (summing) {
            const float* q = &tile[c % kSumStages][myRow * kSumStride + ch];
            const int ns = min(kSumChunk, S - c * kSumChunk);
            if (ns == kSumChunk) {
#pragma unroll 16
                for (int k = 0; k < kSumChunk; k++) sum += q[3 * k];
            } else {
                for (int k = 0; k < ns; k++) sum += q[3 * k];
            }
        }
    }
    if (summing) reinterpret_cast<float*>(summed + row0 + myRow)[ch] = sum;
    // single_dimension_fast / pose-batch rows: pixel x IS ommatidium x (shaders.cu:389-406), so the 8-bit row is
    // written here instead of by a separate projection launch.  Warp 0 holds (row, channel) at lane 3*row+ch.
    if (fastRow != nullptr && t < 32) {
        const float g = __shfl_down_sync(0xffffffffu, sum, 1), b = __shfl_down_sync(0xffffffffu, sum, 2);
        if (summing && ch == 0 && row0 + myRow < fastRowCount) {
            const uchar4 px = makeColor<FAST>(sum, g, b);
            fastRow[row0 + myRow] = px;
            if (fastRowHost != nullptr) fastRowHost[row0 + myRow] = px;
        }
    }
}

// K1b, TMA form (used when S % 4 == 0, i.e. when every row segment is 16-byte aligned): one warp per kSumRows rows.
// Lane 0 is the producer: per tile it arms the stage's mbarrier with the expected byte count and issues one
// bulk asynchronous copy (cp.async.bulk, the 1-D TMA path: no per-thread addresses, no registers) per row segment
// of kTmaChunk samples = 768 contiguous bytes.  The copies complete on the mbarrier; the warp waits on its phase
// bit and lanes 0 .. 3*rows-1 walk their (row, channel) chains through the landed tile while the next two tiles
// are in flight.  Row stride = 4 mod 32 floats: 16-byte aligned for the bulk copy and conflict-free for the walk.
#ifndef CR_TMA_CHUNK
#define CR_TMA_CHUNK 64
#endif
constexpr int kTmaChunk = CR_TMA_CHUNK;          // samples per row segment: 64 -> 18.6 KB per CTA, the whole 10k-ommatidia frame is one wave
constexpr int kTmaStride = 3 * kTmaChunk + 4;

__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbarArriveExpectTx(uint64_t* bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smemAddr(smemDst)), "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nCR_MBAR_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra CR_MBAR_DONE_%=;\nbra CR_MBAR_WAIT_%=;\nCR_MBAR_DONE_%=:\n}"
                 ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}

template <bool FAST>
__global__ void __launch_bounds__(32) k_sumSamplesTma(const float* __restrict__ samples, int NF, int S, float4* __restrict__ summed,
                                                      uchar4* __restrict__ fastRow, int fastRowCount, uchar4* __restrict__ fastRowHost)
{
    __shared__ __align__(128) float tile[kSumStages][kSumRows * kTmaStride];
    __shared__ __align__(8) uint64_t full[kSumStages];
    const int row0 = blockIdx.x * kSumRows;
    const int rows = min(kSumRows, NF - row0);
    const int lane = threadIdx.x;
    const size_t rowFloats = 3 * (size_t)S;
    const float* base = samples + rowFloats * (size_t)row0;
    const int nChunks = (S + kTmaChunk - 1) / kTmaChunk;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kSumStages; st++) mbarInit(&full[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int c) {
        if (c < nChunks && lane == 0) {
            const unsigned bytes = 12u * (unsigned)min(kTmaChunk, S - c * kTmaChunk);     // S % 4 == 0: a multiple of 16
            uint64_t* bar = &full[c % kSumStages];
            float* dst = tile[c % kSumStages];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                  // earlier generic reads of this stage are done
            mbarArriveExpectTx(bar, bytes * (unsigned)rows);
            for (int r = 0; r < rows; r++) bulkLoad(dst + r * kTmaStride, base + rowFloats * r + 3 * (size_t)c * kTmaChunk, bytes, bar);
        }
    };
    issue(0);
    issue(1);
    const int myRow = lane / 3, ch = lane - 3 * myRow;
    const bool summing = myRow < rows;
    float sum = 0.0f;
    for (int c = 0; c < nChunks; c++) {
        mbarWait(&full[c % kSumStages], (unsigned)((c / kSumStages) & 1));
        issue(c + 2);                                   // reuses the stage of tile c-1 (all lanes left it at the __syncwarp below)
        if (summing) {
            const float* q = &tile[c % kSumStages][myRow * kTmaStride + ch];
            const int ns = min(kTmaChunk, S - c * kTmaChunk);
            if (ns == kTmaChunk) {
#pragma unroll 16
                for (int k = 0; k < kTmaChunk; k++) sum += q[3 * k];
            } else {
                for (int k = 0; k < ns; k++) sum += q[3 * k];
            }
        }
        __syncwarp();
    }
    if (summing) reinterpret_cast<float*>(summed + row0 + myRow)[ch] = sum;
    if (fastRow != nullptr) {
        const float g = __shfl_down_sync(0xffffffffu, sum, 1), b = __shfl_down_sync(0xffffffffu, sum, 2);
        if (summing && ch == 0 && row0 + myRow < fastRowCount) {
            const uchar4 px = makeColor<FAST>(sum, g, b);
            fastRow[row0 + myRow] = px;
            if (fastRowHost != nullptr) fastRowHost[row0 + myRow] = px;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2: vector projections (shaders.cu:375-406) and raw samples (shaders.cu:354-369)
// ------------------------------------------------------------------------------------------
__global__ void k_projectVector(int mode, bool fast, const float4* __restrict__ summed, int N, uchar4* __restrict__ frame, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W || y >= H) return;
    uint32_t idx;
    if (mode == PROJ_SINGLE_DIM_FAST) {
        if (y > 0 || x >= N) return;                                  // other pixels keep their contents
        idx = (uint32_t)x;
    } else {
        idx = (uint32_t)(((unsigned long long)(uint32_t)x * (unsigned long long)N) / (unsigned long long)(uint32_t)W);
    }
    const float4 c = __ldg(summed + idx);
    frame[(size_t)y * W + x] = makeColorMode(fast, c.x, c.y, c.z);
}

__global__ void k_projectRaw(bool fast, const float* __restrict__ samples, int N, int S, uchar4* __restrict__ frame, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W || y >= H || y >= S || x >= N) return;
    const float* p = samples + 3 * ((size_t)x * S + y);             // sample buffer is laid out [o][s]
    frame[(size_t)y * W + x] = makeColorMode(fast, p[0], p[1], p[2]);
}

// ------------------------------------------------------------------------------------------
// K3: nearest-ommatidium map of the spherical projections (shaders.cu:412-448, 454-490, 496-541,
// 548-593, 600-640): first index minimising acos(dot(a,v)/(|a||v|)), strict '<', NaN never wins,
// index 0 is the unconditional initial candidate.  The map depends only on (eye, W, H, mode), so
// it is built once and cached instead of being recomputed per frame (O(W*H*N) acos).
// ------------------------------------------------------------------------------------------
constexpr int kMapThreads = 128;

__global__ void __launch_bounds__(kMapThreads)
k_buildProjectionMap(int mode, const float4* __restrict__ omm, int N, uint32_t* __restrict__ map, int W, int H)
{
    __shared__ float4 sA[kMapThreads];     // (a.xyz, |a|)
    __shared__ float sPx[kMapThreads];     // relativePosition.x (split eligibility)
    const long long p = (long long)blockIdx.x * kMapThreads + threadIdx.x;
    const bool live = p < (long long)W * H;
    const int x = live ? (int)(p % W) : 0, y = live ? (int)(p / W) : 0;
    const float uvx = (float)x / (float)W;
    float dx, dy;
    if (mode == PROJ_SPH_SPLIT_ORIENTATIONWISE) {                     // shaders.cu:505-513
        const float sx = uvx * 2.0f, sy = ((float)y / (float)H) * 1.0f;
        const float sub = sx > 1.0f ? 1.0f : 0.0f;
        dx = (sx - sub) * 2.0f - 1.0f;
        dy = sy * 2.0f - 1.0f;
    } else {
        dx = 2.0f * uvx - 1.0f;
        dy = 2.0f * ((float)y / (float)H) - 1.0f;
    }
    const float ax = dx * (-crm::kPi) + crm::kPi / 2.0f;
    const float ay = dy * (crm::kPi / 2.0f) + 0.0f;
    float sax, cax, say, cay;
    crm::sincos(ax, sax, cax);
    crm::sincos(ay, say, cay);
    const V3 usp = mk(cax * cay, say, sax * cay);
    const float lu = vlen(usp);
    const bool byPos = (mode == PROJ_SPH_POSITIONWISE || mode == PROJ_SPH_POSITIONWISE_IDS);
    const bool split = (mode == PROJ_SPH_SPLIT_ORIENTATIONWISE);
    uint32_t closest = 0;
    float smallest = 0.0f;
    for (int base = 0; base < N; base += kMapThreads) {
        const int i = base + threadIdx.x;
        __syncthreads();
        if (i < N) {
            const float4 q0 = __ldg(omm + 2 * i), q1 = __ldg(omm + 2 * i + 1);
            const V3 a = byPos ? mk(q0.x, q0.y, q0.z) : mk(q0.w, q1.x, q1.y);
            sA[threadIdx.x] = make_float4(a.x, a.y, a.z, vlen(a));
            sPx[threadIdx.x] = q0.x;
        }
        __syncthreads();
        const int cnt = min(kMapThreads, N - base);
        for (int k = 0; k < cnt; k++) {
            const float4 a = sA[k];
            const float ang = crm::acos(vdot(mk(a.x, a.y, a.z), usp) / (a.w * lu));
            const int idx = base + k;
            if (idx == 0) { smallest = ang; continue; }
            bool eligible = true;
            if (split) { const float px = sPx[k]; eligible = (px > 0.0f && uvx > 0.5f) || (px < 0.0f && uvx < 0.5f); }
            if (eligible && ang < smallest) { smallest = ang; closest = (uint32_t)idx; }
        }
    }
    if (live) map[p] = closest;
}

__global__ void k_projectMap(bool ids, bool fast, const uint32_t* __restrict__ map, const float4* __restrict__ summed,
                             uchar4* __restrict__ frame, long long nPix)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPix) return;
    const uint32_t idx = __ldg(map + p);
    if (ids) {                                                        // shaders.cu:583-592
        frame[p] = make_uchar4((unsigned char)(idx >> 24), (unsigned char)((idx >> 16) & 0xff),
                               (unsigned char)((idx >> 8) & 0xff), (unsigned char)(idx & 0xff));
    } else {
        const float4 c = __ldg(summed + idx);
        frame[p] = makeColorMode(fast, c.x, c.y, c.z);
    }
}

// ------------------------------------------------------------------------------------------
// K4: ordinary cameras (shaders.cu:198-333), tmin 0.01, one primary ray per pixel.
// ------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(128)
k_camera(const DeviceScene sc, int kind, const DevicePose P, float s0, float s1, float s2, uchar4* __restrict__ frame, int W, int H)
{
    __shared__ int sStack[kSmemStack][128];
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)W * H) return;
    const int x = (int)(p % W), y = (int)(p / W);
    const float dx = 2.0f * (((float)x + 0.0f) / (float)W) - 1.0f;
    const float dy = 2.0f * (((float)y + 0.0f) / (float)H) - 1.0f;
    const V3 X = mk(P.xx, P.xy, P.xz), Y = mk(P.yx, P.yy, P.yz), Z = mk(P.zx, P.zy, P.zz), C = mk(P.px, P.py, P.pz);
    Ray ray;
    if (kind == 0) {
        ray.d = vadd(vadd(vmuls(Z, s2), vmuls(vmuls(X, dx), s0)), vmuls(vmuls(Y, dy), s1));
        ray.o = C;
    } else if (kind == 1) {
        const float ax = dx * (-crm::kPi) + crm::kPi / 2.0f, ay = dy * (crm::kPi / 2.0f) + 0.0f;
        float sax, cax, say, cay;
        Fn<FAST>::sincos(ax, sax, cax);
        Fn<FAST>::sincos(ay, say, cay);
        const V3 od = mk(cax * cay, say, sax * cay);
        ray.d = vnormalize(vadd(vadd(vmuls(X, od.x), vmuls(Y, od.y)), vmuls(Z, od.z)));
        ray.o = vadd(C, vmuls(ray.d, s0));
    } else {
        ray.d = Z;
        ray.o = vadd(vadd(C, vmuls(vmuls(X, dx), s0)), vmuls(vmuls(Y, dy), s1));
    }
    ray.tmin = 0.01f;
    const Hit h = traceClosest<false>(sc.nodes, sc.nodeVariantStride, sc.tris, ray, kTMax, &sStack[0][threadIdx.x], 128, nullptr, nullptr);
    const V3 col = (h.prim >= 0) ? shadeHit<FAST>(sc, h) : shadeMiss<FAST>(sc.missShader, ray.d);
    frame[p] = makeColor<FAST>(col.x, col.y, col.z);
}

// ------------------------------------------------------------------------------------------
// Debug / parity kernels (reached only through the crDebug* calls)
// ------------------------------------------------------------------------------------------
__global__ void k_traceRays(const DeviceScene sc, const float* __restrict__ origins, const float* __restrict__ dirs,
                            const float* __restrict__ tmins, int n, int4* __restrict__ hits)
{
    __shared__ int sStack[kSmemStack][128];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray ray;
    ray.o = mk(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
    ray.d = mk(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    ray.tmin = tmins[i];
    int nc = 0, tc = 0;
    const Hit h = traceClosest<true>(sc.nodes, sc.nodeVariantStride, sc.tris, ray, kTMax, &sStack[0][threadIdx.x], 128, &nc, &tc);
    hits[2 * i] = make_int4(h.prim, __float_as_int(h.t), __float_as_int(h.u), __float_as_int(h.v));
    hits[2 * i + 1] = make_int4(nc, tc, 0, 0);
}

__global__ void k_sampleTexture(unsigned long long tex, const float* __restrict__ uv, int n, float4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = tex2D<float4>((cudaTextureObject_t)tex, uv[2 * i], uv[2 * i + 1]);
}

__global__ void k_evalMath(int fn, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = a[i], y = b ? b[i] : 0.0f;
    float r = 0.0f, s, c;
    switch (fn) {
        case 0: crm::sincos(x, s, c); r = s; break;
        case 1: crm::sincos(x, s, c); r = c; break;
        case 2: r = crm::log(x); break;
        case 3: r = crm::exp(x); break;
        case 4: r = crm::pow(x, y); break;
        case 5: r = crm::asin(x); break;
        case 6: r = crm::acos(x); break;
        case 7: r = crm::atan2(x, y); break;
        default: break;
    }
    out[i] = r;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void launchRngInit(uint4* rng, int N, int S, unsigned long long firstFrame, unsigned long long nGlobal, unsigned long long oFirst,
                   const uint4* jumpTable, cudaStream_t stream)
{
    const long long n = (long long)N * S;
    if (n <= 0) return;
    const int tpb = 128;
    k_rngInit<<<(unsigned)((n + tpb - 1) / tpb), tpb, 0, stream>>>(rng, N, S, firstFrame, nGlobal, oFirst, jumpTable);
}

void launchPrepOmmatidia(const float4* omm, int N, float4* pre, cudaStream_t stream)
{
    if (N <= 0) return;
    k_prepOmmatidia<<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(omm, N, pre);
}

// Programmatic dependent launch (sm_90+): with the attribute, a kernel may be scheduled as soon as every CTA of the kernel
// before it in the stream has executed griddepcontrol.launch_dependents (or exited); what it reads from that kernel it
// reads behind its own griddepcontrol.wait, which returns when the earlier grid has completed and its writes are visible.
// The frontier pass releases the trace kernel at once -- whose warps draw their first tickets and prefetch their first
// states while the pass's latency chain runs -- and the trace kernel releases the reduction kernel, whose CTAs then move in as
// the trace kernel's leave: two launch latencies off the frame.
template <typename... KArgs, typename... Args>
static void launchMaybePdl(void (*kern)(KArgs...), unsigned grid, unsigned block, cudaStream_t stream, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kern, args...);
}

template <bool FUSED, bool FAST>
static void launchTraceT(const DeviceScene& sc, const EyeParams& eye, int grid, cudaStream_t stream)
{
    if (eye.pdl) {
        if (eye.poses && eye.frameGroups > 1) launchMaybePdl(k_traceCompound<false, true, FUSED, FAST, true>, (unsigned)grid, kTraceThreads, stream, true, sc, eye);
        else if (eye.poses) launchMaybePdl(k_traceCompound<false, true, FUSED, FAST, false>, (unsigned)grid, kTraceThreads, stream, true, sc, eye);
        else launchMaybePdl(k_traceCompound<false, false, FUSED, FAST, false>, (unsigned)grid, kTraceThreads, stream, true, sc, eye);
        return;
    }
    if (eye.poses && eye.frameGroups > 1) k_traceCompound<false, true, FUSED, FAST, true><<<grid, kTraceThreads, 0, stream>>>(sc, eye);
    else if (eye.poses) k_traceCompound<false, true, FUSED, FAST><<<grid, kTraceThreads, 0, stream>>>(sc, eye);
    else k_traceCompound<false, false, FUSED, FAST><<<grid, kTraceThreads, 0, stream>>>(sc, eye);
}

template <bool FAST>
static void launchSumT(const EyeParams& eye, cudaStream_t stream)
{
    const long long nf = (long long)eye.N * eye.nFrames;
    if (eye.fused) {
        const int bpr = eye.S >> 5;
        int lanesPerRow = 1;
        while (lanesPerRow < bpr && lanesPerRow < 32) lanesPerRow <<= 1;
        const long long rowsPerCta = (32 / lanesPerRow) * (long long)kSumPartialWarps;
        launchMaybePdl(k_sumPartials<FAST>, (unsigned)((nf + rowsPerCta - 1) / rowsPerCta), 32u * kSumPartialWarps, stream, eye.pdl && eye.queueRays == nullptr,
                       (const float4*)eye.partials, (int)nf, bpr, lanesPerRow, eye.summed, eye.fastRow, eye.fastRowCount, eye.fastRowHost);
        return;
    }
    static const bool useTma = [] { const char* e = getenv("CR_SUM_TMA"); return e ? atoi(e) != 0 : true; }();
    const unsigned sumGrid = (unsigned)((nf + kSumRows - 1) / kSumRows);
    if (useTma && eye.S % 4 == 0)
        k_sumSamplesTma<FAST><<<sumGrid, 32, 0, stream>>>(eye.samples, (int)nf, eye.S, eye.summed, eye.fastRow, eye.fastRowCount, eye.fastRowHost);
    else
        k_sumSamples<FAST><<<sumGrid, kSumThreads, 0, stream>>>(eye.samples, (int)nf, eye.S, eye.summed, eye.fastRow, eye.fastRowCount, eye.fastRowHost);
}

// K1 + K1b.  eye.fused (needs S % 32 == 0 and eye.partials) selects the in-kernel reduction, eye.fast the hardware
// elementary functions; the per-ray dump exists for the ordered single-frame kernel only.
void launchTraceCompound(const DeviceScene& sc, const EyeParams& eye, int gridBlocks, cudaStream_t stream, cudaEvent_t afterTrace)
{
    const long long total = (long long)eye.N * eye.S;
    if (total <= 0) return;
    const long long need = (total + kTraceThreads - 1) / kTraceThreads;
    const int grid = (int)(need < gridBlocks ? need : gridBlocks);
    if (eye.dumpHits) {
        if (eye.fast) k_traceCompound<true, false, false, true><<<grid, kTraceThreads, 0, stream>>>(sc, eye);
        else k_traceCompound<true, false, false, false><<<grid, kTraceThreads, 0, stream>>>(sc, eye);
    } else if (eye.fused) {
        if (eye.fast) launchTraceT<true, true>(sc, eye, grid, stream);
        else launchTraceT<true, false>(sc, eye, grid, stream);
    } else {
        if (eye.fast) launchTraceT<false, true>(sc, eye, grid, stream);
        else launchTraceT<false, false>(sc, eye, grid, stream);
    }
    if (eye.queueRays != nullptr && !eye.dumpHits) {                 // the queued warp-frames: trace with dynamic fetch, then shade
        k_traceQueue<<<gridBlocks, kTraceThreads, 0, stream>>>(sc, eye);
        if (eye.fused) {
            if (eye.fast) k_shadeQueue<true, true><<<gridBlocks, 128, 0, stream>>>(sc, eye);
            else k_shadeQueue<true, false><<<gridBlocks, 128, 0, stream>>>(sc, eye);
        } else {
            if (eye.fast) k_shadeQueue<false, true><<<gridBlocks, 128, 0, stream>>>(sc, eye);
            else k_shadeQueue<false, false><<<gridBlocks, 128, 0, stream>>>(sc, eye);
        }
    }
    if (afterTrace != nullptr) cudaEventRecord(afterTrace, stream);
    if (eye.fast) launchSumT<true>(eye, stream);
    else launchSumT<false>(eye, stream);
}

void launchBuildEntries(const DeviceScene& sc, const EyeParams& eye, int4* entries, int* lists, cudaStream_t stream)
{
    const long long total = (long long)eye.N * (eye.poses ? eye.nFrames : 1) * kEntryK;
    if (total <= 0) return;
    k_buildEntries<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(sc, eye, entries, lists);
}
int candidateListStride() { return kListStride; }

int traceKernelOccupancy()
{
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_traceCompound<false, true, false, false>, kTraceThreads, 0);
    return n > 0 ? n : 1;
}

// Upper bound of %smid on this device (%nsmid).  SM identifiers need not be contiguous -- a B200 has 148 SMs but may number
// them beyond 148 -- so the per-SM ticket slots of the trace kernel are sized by this, not by the SM count.
__global__ void k_smIdLimit(unsigned* out)
{
    unsigned n;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n));
    *out = n;
}
unsigned deviceSmIdLimit(cudaStream_t stream)
{
    unsigned* d = nullptr;
    unsigned h = 0;
    if (cudaMalloc(&d, sizeof(unsigned)) != cudaSuccess) return 0;
    k_smIdLimit<<<1, 1, 0, stream>>>(d);
    if (cudaMemcpyAsync(&h, d, sizeof(unsigned), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) h = 0;
    cudaFree(d);
    return h;
}

void launchProjectVector(int mode, bool fast, const float4* summed, int N, uchar4* frame, int W, int H, cudaStream_t stream)
{
    if (W <= 0 || H <= 0) return;
    const int rows = (mode == PROJ_SINGLE_DIM_FAST) ? 1 : H;
    dim3 grid((unsigned)((W + 127) / 128), (unsigned)rows);
    k_projectVector<<<grid, 128, 0, stream>>>(mode, fast, summed, N, frame, W, H);
}

void launchProjectRaw(bool fast, const float* samples, int N, int S, uchar4* frame, int W, int H, cudaStream_t stream)
{
    if (W <= 0 || H <= 0) return;
    const int rows = H < S ? H : S;
    dim3 grid((unsigned)((W + 127) / 128), (unsigned)rows);
    k_projectRaw<<<grid, 128, 0, stream>>>(fast, samples, N, S, frame, W, H);
}

void launchBuildProjectionMap(int mode, const float4* omm, int N, uint32_t* map, int W, int H, cudaStream_t stream)
{
    const long long nPix = (long long)W * H;
    if (nPix <= 0) return;
    k_buildProjectionMap<<<(unsigned)((nPix + kMapThreads - 1) / kMapThreads), kMapThreads, 0, stream>>>(mode, omm, N, map, W, H);
}

void launchProjectMap(bool ids, bool fast, const uint32_t* map, const float4* summed, uchar4* frame, int W, int H, cudaStream_t stream)
{
    const long long nPix = (long long)W * H;
    if (nPix <= 0) return;
    k_projectMap<<<(unsigned)((nPix + 255) / 256), 256, 0, stream>>>(ids, fast, map, summed, frame, nPix);
}

void launchCamera(const DeviceScene& sc, int kind, bool fast, const DevicePose& pose, float s0, float s1, float s2, uchar4* frame, int W,
                  int H, cudaStream_t stream)
{
    const long long nPix = (long long)W * H;
    if (nPix <= 0) return;
    if (fast) k_camera<true><<<(unsigned)((nPix + 127) / 128), 128, 0, stream>>>(sc, kind, pose, s0, s1, s2, frame, W, H);
    else k_camera<false><<<(unsigned)((nPix + 127) / 128), 128, 0, stream>>>(sc, kind, pose, s0, s1, s2, frame, W, H);
}

void launchTraceRays(const DeviceScene& sc, const float* origins, const float* dirs, const float* tmins, int n, int4* hits,
                     cudaStream_t stream)
{
    if (n <= 0) return;
    k_traceRays<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(sc, origins, dirs, tmins, n, hits);
}

void launchSampleTexture(unsigned long long tex, const float* uv, int n, float4* out, cudaStream_t stream)
{
    if (n <= 0) return;
    k_sampleTexture<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(tex, uv, n, out);
}

void launchEvalMath(int fn, const float* a, const float* b, float* out, int n, cudaStream_t stream)
{
    if (n <= 0) return;
    k_evalMath<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(fn, a, b, out, n);
}

}  // namespace cr
