// cr_bvh.cu -- GPU LBVH builder (sm_100a).
//
// Replaces the closed OptiX acceleration-structure build the reference calls per mesh and per
// scene (libEyeRenderer3/MulticamScene.cpp:1052-1341 buildMeshAccels, :1350-1428
// buildInstanceAccel).  Static instances were flattened to world space by the loader, so ONE
// tree covers the scene:
//   1. per-triangle bounds + 63-bit Morton code of the box centre (21 bits / axis)
//   2. stable LSD radix sort of (code, triangle) pairs, 8 bits per pass, written here
//      (block histograms -> exclusive scan -> stable scatter with warp match ranking)
//   3. Karras 2012 hierarchy from the sorted codes (ties broken by position)
//   4. bottom-up AABB refit with one atomic arrival counter per internal node
//   5. emit: 64-byte nodes holding both child boxes (x8 direction-octant variants, near/far
//      pre-selected), subtrees of <= leafSize triangles folded
//      into leaf references, and 48-byte (v0, e1, e2, prim) triangle records in tree order.
// HBM-bound integer/byte work; every kernel is a flat grid-stride or one-thread-per-item pass.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>
#include <stdexcept>
#include <string>

#include "cr_device.h"

namespace cr {

#define CR_CUDA(x)                                                                                          \
    do {                                                                                                    \
        cudaError_t e_ = (x);                                                                               \
        if (e_ != cudaSuccess)                                                                              \
            throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(e_) + " at " #x);     \
    } while (0)

namespace {

// ---------------------------------------------------------------------------------- step 1
__device__ __forceinline__ unsigned long long expandBits21(unsigned long long v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

struct BuildConst {
    float smin[3];
    float invExt[3];
    float padAbs;
};

__global__ void k_triBoundsMorton(const float* __restrict__ pos, const uint32_t* __restrict__ idx, int nTris, BuildConst bc,
                                  float4* __restrict__ leafMin, float4* __restrict__ leafMax,
                                  unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTris) return;
    const uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
    float mn[3], mx[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float p0 = pos[3 * (size_t)i0 + a], p1 = pos[3 * (size_t)i1 + a], p2 = pos[3 * (size_t)i2 + a];
        mn[a] = fminf(p0, fminf(p1, p2));
        mx[a] = fmaxf(p0, fmaxf(p1, p2));
    }
    unsigned long long code = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float c = 0.5f * (mn[a] + mx[a]);
        float n = (c - bc.smin[a]) * bc.invExt[a];
        n = fminf(fmaxf(n, 0.0f), 1.0f);
        unsigned long long q = (unsigned long long)(n * 2097151.0f);
        code |= expandBits21(q) << (2 - a);
        // conservative padding: the Moller-Trumbore test may accept points a few ulps outside the
        // exact triangle; the box must still contain them (see DESIGN.md "Traversal exactness")
        const float m = fmaxf(fabsf(mn[a]), fabsf(mx[a]));
        const float pad = m * 9.5367431640625e-07f + bc.padAbs;
        mn[a] -= pad;
        mx[a] += pad;
    }
    leafMin[t] = make_float4(mn[0], mn[1], mn[2], 0.0f);
    leafMax[t] = make_float4(mx[0], mx[1], mx[2], 0.0f);
    keys[t] = code;
    vals[t] = (uint32_t)t;
}

// ---------------------------------------------------------------------------------- step 2
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 8;                       // elements per thread
constexpr int kSortTile = kSortThreads * kSortRounds;

__global__ void __launch_bounds__(kSortThreads)
k_sortHist(const unsigned long long* __restrict__ keys, int n, int shift, uint32_t* __restrict__ blockHist, int numBlocks)
{
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortRounds; r++) {
        const int i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&hist[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    blockHist[(size_t)threadIdx.x * numBlocks + blockIdx.x] = hist[threadIdx.x];
}

// Exclusive scan of the [digit][block] histogram in digit-major order, in two levels (round 2: the single-CTA scan of all
// 256 * numBlocks counters was 107 us of every pass): CTA d scans row d in place and leaves its total, one more CTA scans
// the 256 totals into digitBase; k_sortScatter adds the two.
__global__ void __launch_bounds__(256) k_scanRows(uint32_t* __restrict__ data, int numBlocks, uint32_t* __restrict__ rowTotal)
{
    __shared__ uint32_t warpSums[8];
    __shared__ uint32_t carry;
    uint32_t* row = data + (size_t)blockIdx.x * numBlocks;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < numBlocks; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t v = (i < numBlocks) ? row[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warpSums[warp] = x;
        __syncthreads();
        uint32_t warpOff = 0;
        for (int w = 0; w < warp; w++) warpOff += warpSums[w];
        const uint32_t c = carry;
        if (i < numBlocks) row[i] = c + warpOff + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = c + warpOff + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) rowTotal[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(256) k_scanTotals(const uint32_t* __restrict__ rowTotal, uint32_t* __restrict__ digitBase)
{
    __shared__ uint32_t warpSums[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t v = rowTotal[threadIdx.x];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warpSums[warp] = x;
    __syncthreads();
    uint32_t warpOff = 0;
    for (int w = 0; w < warp; w++) warpOff += warpSums[w];
    digitBase[threadIdx.x] = warpOff + x - v;
}

__global__ void __launch_bounds__(kSortThreads)
k_sortScatter(const unsigned long long* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
              unsigned long long* __restrict__ keysOut, uint32_t* __restrict__ valsOut, int n, int shift,
              const uint32_t* __restrict__ blockOffsets, const uint32_t* __restrict__ digitBase, int numBlocks)
{
    __shared__ uint32_t base[256];
    __shared__ uint32_t warpCnt[kSortThreads / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    base[threadIdx.x] = blockOffsets[(size_t)threadIdx.x * numBlocks + blockIdx.x] + digitBase[threadIdx.x];
    const int tileBase = blockIdx.x * kSortTile;
    for (int r = 0; r < kSortRounds; r++) {
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; w++) warpCnt[w][threadIdx.x] = 0;
        __syncthreads();
        const int i = tileBase + r * kSortThreads + threadIdx.x;
        const bool valid = i < n;
        unsigned long long key = 0;
        uint32_t val = 0, digit = 0;
        if (valid) { key = keysIn[i]; val = valsIn[i]; digit = (uint32_t)(key >> shift) & 255u; }
        // rank among same-digit elements of this warp (stable: lower lanes first)
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? digit : 0xffffffffu);
        const uint32_t rankInWarp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rankInWarp == 0) warpCnt[warp][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t off = base[digit] + rankInWarp;
            for (int w = 0; w < warp; w++) off += warpCnt[w][digit];
            keysOut[off] = key;
            valsOut[off] = val;
        }
        __syncthreads();
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; w++) add += warpCnt[w][threadIdx.x];
        base[threadIdx.x] += add;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- step 3
__device__ __forceinline__ int deltaKey(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ b));
}

// node numbering: internal nodes 0 .. n-2, leaves n-1 .. 2n-2 (leaf k = sorted position k)
__global__ void k_buildHierarchy(const unsigned long long* __restrict__ keys, int n, int* __restrict__ parent,
                                 int2* __restrict__ children, int2* __restrict__ ranges)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int dl = deltaKey(keys, n, i, i - 1), dr = deltaKey(keys, n, i, i + 1);
    const int d = (dr > dl) ? 1 : -1;
    const int dmin = (d > 0) ? dl : dr;
    int lmax = 2;
    while (deltaKey(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (deltaKey(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = deltaKey(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (deltaKey(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    children[i] = make_int2(left, right);
    ranges[i] = make_int2(lo, hi);
    parent[left] = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}

// ---------------------------------------------------------------------------------- step 4
__global__ void k_refit(int n, const uint32_t* __restrict__ sortedVals, const float4* __restrict__ leafMin,
                        const float4* __restrict__ leafMax, const int* __restrict__ parent, const int2* __restrict__ children,
                        float4* __restrict__ boxMin, float4* __restrict__ boxMax, int* __restrict__ arrivals)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t t = sortedVals[k];
    const int leaf = n - 1 + k;
    boxMin[leaf] = leafMin[t];
    boxMax[leaf] = leafMax[t];
    __threadfence();
    int node = parent[leaf];
    while (node >= 0) {
        if (atomicAdd(&arrivals[node], 1) == 0) return;   // first child to arrive stops
        __threadfence();
        const int2 c = children[node];
        const float4 a0 = __ldcg(&boxMin[c.x]), a1 = __ldcg(&boxMax[c.x]);
        const float4 b0 = __ldcg(&boxMin[c.y]), b1 = __ldcg(&boxMax[c.y]);
        boxMin[node] = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), 0.0f);
        boxMax[node] = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), 0.0f);
        __threadfence();
        node = parent[node];
    }
}

// ---------------------------------------------------------------------------------- step 5
__device__ __forceinline__ int childRef(int c, int n, const int2* __restrict__ ranges, int leafSize)
{
    if (c >= n - 1) return ~(((c - (n - 1)) << 3) | 0);           // single-triangle leaf
    const int2 r = ranges[c];
    const int cnt = r.y - r.x + 1;
    if (cnt <= leafSize) return ~((r.x << 3) | (cnt - 1));          // fold the subtree into one leaf
    return c;
}

// Every node is emitted 8 times, once per ray-direction sign octant v = sx | sy<<1 | sz<<2, with
// each child box stored as (near, far) for that octant: near = max, far = min on axes whose
// direction component is negative.  The traversal loop then needs no per-node selects.  All rays
// of a frame leave one eye, so an octant's rays visit their own region of the tree and the cached
// working set does not grow with the 8x footprint (HBM is plentiful: 64 B x 8 per node).
__global__ void k_emitNodes(int n, const int2* __restrict__ children, const int2* __restrict__ ranges,
                            const float4* __restrict__ boxMin, const float4* __restrict__ boxMax, int leafSize,
                            float4* __restrict__ nodes, size_t variantStride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 c = children[i];
    const float4 a0 = boxMin[c.x], a1 = boxMax[c.x], b0 = boxMin[c.y], b1 = boxMax[c.y];
    const int r0 = childRef(c.x, n, ranges, leafSize), r1 = childRef(c.y, n, ranges, leafSize);
    const float4 refs = make_float4(__int_as_float(r0), __int_as_float(r1), 0.0f, 0.0f);
#pragma unroll
    for (int v = 0; v < 8; v++) {
        const bool sx = v & 1, sy = v & 2, sz = v & 4;
        float4* out = nodes + (size_t)v * variantStride + 4 * (size_t)i;
        out[0] = make_float4(sx ? a1.x : a0.x, sx ? a0.x : a1.x, sy ? a1.y : a0.y, sy ? a0.y : a1.y);
        out[1] = make_float4(sx ? b1.x : b0.x, sx ? b0.x : b1.x, sy ? b1.y : b0.y, sy ? b0.y : b1.y);
        out[2] = make_float4(sz ? a1.z : a0.z, sz ? a0.z : a1.z, sz ? b1.z : b0.z, sz ? b0.z : b1.z);
        out[3] = refs;
    }
}

// n == 1 (or the whole scene fits one leaf): a root whose child 0 is the leaf and child 1 is empty
__global__ void k_emitSingleRoot(int n, const float4* __restrict__ rootMin, const float4* __restrict__ rootMax,
                                 float4* __restrict__ nodes)
{
    const float4 a0 = rootMin[0], a1 = rootMax[0];
    const float inf = __int_as_float(0x7f800000);
    const int r0 = ~((0 << 3) | (n - 1));
    for (int v = 0; v < 8; v++) {
        const bool sx = v & 1, sy = v & 2, sz = v & 4;
        float4* out = nodes + 4 * (size_t)v;
        // child 1 is empty: near = +inf, far = -inf in every octant
        out[0] = make_float4(sx ? a1.x : a0.x, sx ? a0.x : a1.x, sy ? a1.y : a0.y, sy ? a0.y : a1.y);
        out[1] = make_float4(inf, -inf, inf, -inf);
        out[2] = make_float4(sz ? a1.z : a0.z, sz ? a0.z : a1.z, inf, -inf);
        out[3] = make_float4(__int_as_float(r0), __int_as_float(r0), 0.0f, 0.0f);
    }
}

__global__ void k_emitTris(const float* __restrict__ pos, const uint32_t* __restrict__ idx, int n,
                           const uint32_t* __restrict__ sortedVals, float4* __restrict__ tris)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t t = sortedVals[k];
    const uint32_t i0 = idx[3 * (size_t)t], i1 = idx[3 * (size_t)t + 1], i2 = idx[3 * (size_t)t + 2];
    const float v0x = pos[3 * (size_t)i0], v0y = pos[3 * (size_t)i0 + 1], v0z = pos[3 * (size_t)i0 + 2];
    const float v1x = pos[3 * (size_t)i1], v1y = pos[3 * (size_t)i1 + 1], v1z = pos[3 * (size_t)i1 + 2];
    const float v2x = pos[3 * (size_t)i2], v2y = pos[3 * (size_t)i2 + 1], v2z = pos[3 * (size_t)i2 + 2];
    tris[3 * (size_t)k + 0] = make_float4(v0x, v0y, v0z, __int_as_float((int)t));
    tris[3 * (size_t)k + 1] = make_float4(v1x - v0x, v1y - v0y, v1z - v0z, 0.0f);
    tris[3 * (size_t)k + 2] = make_float4(v2x - v0x, v2y - v0y, v2z - v0z, 0.0f);
}

template <typename T>
T* dalloc(size_t n)
{
    T* p = nullptr;
    CR_CUDA(cudaMalloc(&p, sizeof(T) * (n ? n : 1)));
    return p;
}
// The build's temporaries come from the device's stream-ordered pool (kept between builds: release threshold = everything):
// cudaMalloc / cudaFree cost milliseconds each and made the reported build time mostly allocator time.
template <typename T>
T* talloc(size_t n, cudaStream_t stream)
{
    static bool poolReady = false;
    if (!poolReady) {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        poolReady = true;
    }
    T* p = nullptr;
    CR_CUDA(cudaMallocAsync(&p, sizeof(T) * (n ? n : 1), stream));
    return p;
}

}  // namespace

BvhBuildResult buildLbvh(const float* dPositions, const uint32_t* dIndices, int nTris, const float sceneMin[3],
                         const float sceneMax[3], int leafSize, cudaStream_t stream)
{
    BvhBuildResult out;
    out.nTris = nTris;
    if (leafSize < 1) leafSize = 1;
    if (leafSize > 8) leafSize = 8;
    if (nTris >= (1 << 28)) throw std::runtime_error("scene too large: the leaf encoding holds 2^28 triangles");
    // The first launch of a kernel in a process loads it (lazy module loading, ~10 ms per kernel): do that before the clock
    // starts, so that buildMs is the build and not the loader (the first scene of a process used to report 125 ms).
    static bool loaded = false;
    if (!loaded) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, k_triBoundsMorton); cudaFuncGetAttributes(&fa, k_sortHist); cudaFuncGetAttributes(&fa, k_scanRows);
        cudaFuncGetAttributes(&fa, k_scanTotals); cudaFuncGetAttributes(&fa, k_sortScatter); cudaFuncGetAttributes(&fa, k_buildHierarchy);
        cudaFuncGetAttributes(&fa, k_refit); cudaFuncGetAttributes(&fa, k_emitNodes); cudaFuncGetAttributes(&fa, k_emitSingleRoot);
        cudaFuncGetAttributes(&fa, k_emitTris);
        cudaGetLastError();
        loaded = true;
    }
    const auto t0 = std::chrono::steady_clock::now();
    if (nTris == 0) {
        // a root with two empty children: every ray misses
        out.nNodes = 1;
        out.nodes = dalloc<float4>(4 * 8);
        out.tris = dalloc<float4>(3);
        const float inf = INFINITY;
        const int emptyRef = ~0;   // first 0, count 1 -- never visited because the boxes are empty
        float4 h[32];
        for (int v = 0; v < 8; v++) {
            h[4 * v + 0] = make_float4(inf, -inf, inf, -inf);
            h[4 * v + 1] = make_float4(inf, -inf, inf, -inf);
            h[4 * v + 2] = make_float4(inf, -inf, inf, -inf);
            h[4 * v + 3] = make_float4(0, 0, 0, 0);
            memcpy(&h[4 * v + 3].x, &emptyRef, 4);
            memcpy(&h[4 * v + 3].y, &emptyRef, 4);
        }
        CR_CUDA(cudaMemcpyAsync(out.nodes, h, sizeof h, cudaMemcpyHostToDevice, stream));
        CR_CUDA(cudaMemsetAsync(out.tris, 0, sizeof(float4) * 3, stream));
        CR_CUDA(cudaStreamSynchronize(stream));
        return out;
    }
    const int n = nTris;
    BuildConst bc;
    float extMax = 0.0f;
    for (int a = 0; a < 3; a++) {
        bc.smin[a] = sceneMin[a];
        const float ext = sceneMax[a] - sceneMin[a];
        bc.invExt[a] = ext > 0.0f ? 1.0f / ext : 0.0f;
        extMax = fmaxf(extMax, ext);
    }
    bc.padAbs = extMax * 1.1920929e-07f;
    // Morton quantisation grid: one uniform cube of the largest extent (default) keeps cells
    // cubical, so flat scenes (terrain) are not split along their thin axis at the top levels;
    // CR_MORTON_PER_AXIS=1 restores per-axis normalisation.
    const char* perAxis = getenv("CR_MORTON_PER_AXIS");
    if (!(perAxis && atoi(perAxis) != 0))
        for (int a = 0; a < 3; a++) bc.invExt[a] = extMax > 0.0f ? 1.0f / extMax : 0.0f;

    float4* leafMin = talloc<float4>(n, stream);
    float4* leafMax = talloc<float4>(n, stream);
    unsigned long long* keysA = talloc<unsigned long long>(n, stream);
    unsigned long long* keysB = talloc<unsigned long long>(n, stream);
    uint32_t* valsA = talloc<uint32_t>(n, stream);
    uint32_t* valsB = talloc<uint32_t>(n, stream);
    const int tpb = 256;
    const int gridN = (n + tpb - 1) / tpb;
    k_triBoundsMorton<<<gridN, tpb, 0, stream>>>(dPositions, dIndices, n, bc, leafMin, leafMax, keysA, valsA);

    const int numBlocks = (n + kSortTile - 1) / kSortTile;
    uint32_t* blockHist = talloc<uint32_t>((size_t)256 * numBlocks + 512, stream);   // + 256 row totals + 256 digit bases
    uint32_t* rowTotal = blockHist + (size_t)256 * numBlocks;
    uint32_t* digitBase = rowTotal + 256;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = pass * 8;
        k_sortHist<<<numBlocks, kSortThreads, 0, stream>>>(keysA, n, shift, blockHist, numBlocks);
        k_scanRows<<<256, 256, 0, stream>>>(blockHist, numBlocks, rowTotal);
        k_scanTotals<<<1, 256, 0, stream>>>(rowTotal, digitBase);
        k_sortScatter<<<numBlocks, kSortThreads, 0, stream>>>(keysA, valsA, keysB, valsB, n, shift, blockHist, digitBase, numBlocks);
        std::swap(keysA, keysB);
        std::swap(valsA, valsB);
    }
    // keysA / valsA now hold the sorted pairs (8 passes = even number of swaps back to A)

    out.tris = dalloc<float4>((size_t)3 * n);
    k_emitTris<<<gridN, tpb, 0, stream>>>(dPositions, dIndices, n, valsA, out.tris);

    float4* boxMin = talloc<float4>((size_t)2 * n, stream);
    float4* boxMax = talloc<float4>((size_t)2 * n, stream);
    if (n == 1 || n <= leafSize) {
        // whole scene in one leaf: reduce the leaf boxes on the host (tiny)
        std::vector<float4> hMin(n), hMax(n);
        CR_CUDA(cudaMemcpyAsync(hMin.data(), leafMin, sizeof(float4) * n, cudaMemcpyDeviceToHost, stream));
        CR_CUDA(cudaMemcpyAsync(hMax.data(), leafMax, sizeof(float4) * n, cudaMemcpyDeviceToHost, stream));
        CR_CUDA(cudaStreamSynchronize(stream));
        float4 mn = hMin[0], mx = hMax[0];
        for (int i = 1; i < n; i++) {
            mn.x = fminf(mn.x, hMin[i].x); mn.y = fminf(mn.y, hMin[i].y); mn.z = fminf(mn.z, hMin[i].z);
            mx.x = fmaxf(mx.x, hMax[i].x); mx.y = fmaxf(mx.y, hMax[i].y); mx.z = fmaxf(mx.z, hMax[i].z);
        }
        CR_CUDA(cudaMemcpyAsync(boxMin, &mn, sizeof mn, cudaMemcpyHostToDevice, stream));
        CR_CUDA(cudaMemcpyAsync(boxMax, &mx, sizeof mx, cudaMemcpyHostToDevice, stream));
        out.nNodes = 1;
        out.nodes = dalloc<float4>(4 * 8);
        k_emitSingleRoot<<<1, 1, 0, stream>>>(n, boxMin, boxMax, out.nodes);
    } else {
        int* parent = talloc<int>((size_t)2 * n, stream);
        int2* children = talloc<int2>(n, stream);
        int2* ranges = talloc<int2>(n, stream);
        int* arrivals = talloc<int>(n, stream);
        CR_CUDA(cudaMemsetAsync(arrivals, 0, sizeof(int) * n, stream));
        const int gridI = (n - 1 + tpb - 1) / tpb;
        k_buildHierarchy<<<gridI, tpb, 0, stream>>>(keysA, n, parent, children, ranges);
        k_refit<<<gridN, tpb, 0, stream>>>(n, valsA, leafMin, leafMax, parent, children, boxMin, boxMax, arrivals);
        out.nNodes = n - 1;
        out.nodes = dalloc<float4>((size_t)4 * (n - 1) * 8);
        k_emitNodes<<<gridI, tpb, 0, stream>>>(n, children, ranges, boxMin, boxMax, leafSize, out.nodes, (size_t)4 * (n - 1));
        CR_CUDA(cudaStreamSynchronize(stream));
        cudaFreeAsync(parent, stream); cudaFreeAsync(children, stream); cudaFreeAsync(ranges, stream); cudaFreeAsync(arrivals, stream);
    }
    CR_CUDA(cudaStreamSynchronize(stream));
    CR_CUDA(cudaGetLastError());
    cudaFreeAsync(leafMin, stream); cudaFreeAsync(leafMax, stream); cudaFreeAsync(keysA, stream); cudaFreeAsync(keysB, stream);
    cudaFreeAsync(valsA, stream); cudaFreeAsync(valsB, stream);
    cudaFreeAsync(blockHist, stream); cudaFreeAsync(boxMin, stream); cudaFreeAsync(boxMax, stream);
    CR_CUDA(cudaStreamSynchronize(stream));
    out.buildMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return out;
}

}  // namespace cr
