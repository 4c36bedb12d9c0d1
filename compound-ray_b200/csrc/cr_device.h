// cr_device.h -- device-side data layout shared by the builder, the kernels and the renderer.
//
// HBM layout (all buffers 256-byte aligned by cudaMalloc):
//   nodes   float4[8][4*nNodes] 64 B per BVH2 node, both child boxes inline, stored once per
//           ray-direction sign octant v = sx | sy<<1 | sz<<2 with (near, far) planes pre-selected
//           (variant 0 = (min, max) on every axis):
//             n[0] = (c0.near.x, c0.far.x, c0.near.y, c0.far.y)
//             n[1] = (c1.near.x, c1.far.x, c1.near.y, c1.far.y)
//             n[2] = (c0.near.z, c0.far.z, c1.near.z, c1.far.z)
//             n[3] = (ref0, ref1, unused, unused) as int bits
//           child ref >= 0 : internal node index;  ref < 0 : leaf,  x = ~ref,
//           first triangle = x >> 3 (position in the sorted array), count = (x & 7) + 1
//   tris    float4[3*nTris]    48 B per triangle in BVH (Morton) order:
//             t[0] = (v0.xyz, flattened primitive index as int bits)
//             t[1] = (e1.xyz = v1 - v0, 0)     t[2] = (e2.xyz = v2 - v0, 0)
//   prims   uint4[nTris]       per flattened primitive: global vertex ids i0,i1,i2 and mesh group
//   uvs     float2[nVerts]     colors float4[nVerts]   (only when some mesh has them)
//   meshes  MeshRec[nMeshes]
// Compound eye (per camera):
//   omm     float4[2*N]        the 32-byte ommatidium rows as loaded
//   pre     float4[4*N]        per-ommatidium ray invariants (origin offset, sd, axis, focal, perp, cone bound, perp x axis, perp . axis)
//   entries int4[F*N]          per (frame, ommatidium): up to 4 BVH subtree roots its sample cone can reach
//                              (near to far, packed from .x, 0x80000000 = none)
//   lists   int[F*N][16]       per (frame, ommatidium): candidate list -- [0] = element count n (0..15; -1 = no list, walk the
//                              frontier), [1..n] = node << 2 | mask of that node's children that are leaves the cone can reach
//   rng     uint4[2*N*S]       32 B compact XORWOW state per sample stream, laid out [o][s]
//             r[0] = (d, v0, v1, v2)   r[1] = (v3, v4, boxmuller_flag, boxmuller_extra bits)
//   samples float[3*N*S]       per-sample colour/S, laid out [o][s] (12 B per ray, written by K1; ordered mode)
//   partials float4[N*S/32]    fused mode instead: one butterfly sum of 32 samples per warp (0.5 B per ray)
//   queue   float4[2*cap] + int4[cap]   wavefront queue (batches): rays of the ommatidia without a candidate list, traced by
//                              k_traceQueue with dynamic ray fetch and shaded by k_shadeQueue (48 B per queued ray)
//   summed  float4[N]          per-ommatidium RGB: sequential sum of the samples (K1b)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cr {

struct MeshRec {
    int colorType;      // -1 none
    int hasUV;
    int hasTex;
    int pad;
    unsigned long long tex;   // cudaTextureObject_t
    float baseColor[4];
};

struct DeviceScene {
    const float4* nodes = nullptr;
    const float4* tris = nullptr;
    const uint4* prims = nullptr;
    const float2* uvs = nullptr;
    const float4* colors = nullptr;
    const MeshRec* meshes = nullptr;
    size_t nodeVariantStride = 0;   // float4 units between direction-octant variants of the node array
    int nNodes = 0;
    int nTris = 0;
    int missShader = 0;
    float boundsAbsMax = 0.0f;      // largest |coordinate| of any node box (vertex bounds + the leaf padding): scale of the frontier pass's tolerance
};

struct alignas(16) DevicePose {
    float px, py, pz;
    float xx, xy, xz;
    float yx, yy, yz;
    float zx, zy, zz;
};

struct EyeParams {
    const float4* pre = nullptr;  // kPreStride float4 per ommatidium (k_prepOmmatidia)
    uint4* rng = nullptr;
    uint4* rngOut = nullptr;      // batches: where the states go after the launch (= rng unless the launch is cut into frame groups)
    int frameGroups = 1;          // batches of small frames: the launch's frames in this many groups of groupFrames (even) frames,
    int groupFrames = 0;          // a work unit = (32 rays, one group); needs the launch to start at an even frame
    float4* summed = nullptr;
    float* samples = nullptr;     // [o][s][3] per-sample colour/S
    // optional per-ray dump (debug / parity): reference stream-id order [N*s+o]
    float* dumpOrigins = nullptr;
    float* dumpDirs = nullptr;
    int4* dumpHits = nullptr;     // (prim, t bits, u bits, v bits)
    int2* dumpCounts = nullptr;   // (BVH nodes fetched, triangles tested) per sample ray
    int N = 0;
    int S = 0;
    int nFrames = 1;              // frames (poses) covered by one launch
    const DevicePose* poses = nullptr;   // device array [nFrames]; nullptr: use `pose`
    uchar4* fastRow = nullptr;           // when set: K1b also writes make_color(summed[i]) for i < fastRowCount
    int fastRowCount = 0;
    uchar4* fastRowHost = nullptr;       // when set: the same pixels also go straight to the mapped pinned host frame
    const int4* entries = nullptr;       // [nFrames][N] entry frontier (k_buildEntries); nullptr: start at the root
    unsigned entryFrameStride = 0;       // rows between the frames of `entries` / `lists`: N, or 0 when every frame of the launch has the same pose
    // wavefront queue of the warp-frames whose ommatidium has no candidate list (k_traceCompound -> k_traceQueue -> k_shadeQueue)
    float4* queueRays = nullptr;         // 2 float4 per ray: (origin, tmin), (direction, frame*N + ommatidium | inCone << 31)
    int* queueWarps = nullptr;           // per 32 queue slots (one pushed warp-frame): its block of 32 samples within the row
    int4* queueHits = nullptr;           // (prim, t bits, u bits, v bits) per queued ray
    unsigned* queueCounters = nullptr;   // [0] rays pushed by K1 (multiples of 32), [1] rays handed out by k_traceQueue
    unsigned queueCap = 0;               // rays the queue holds; a warp that finds it full walks inline
    int entryMaxLevels = 256;            // k_buildEntries: levels the frontier may descend (each is one dependent node fetch)
    int chunkUnits = 1;                  // units of 32 rays per work-counter fetch of the single-frame trace kernel
    unsigned* workCounter = nullptr;     // [0] chunks of ray units handed out beyond every warp's first, [1] warps that have left the
                                         // kernel (the last one zeroes both for the next launch); nullptr: static split
    // SM-affine hand-out (round 2): blocks of 32 consecutive units go to ONE SM, whose warps share them through a per-SM ticket
    // counter -- at S = 1024 the 32 resident warps of an SM then trace the 32 x 32 samples of one ommatidium together, so the nodes
    // its cone reaches are fetched into that SM's L1 once instead of into 32 different ones.
    unsigned* smSeq = nullptr;           // [smSlots] tickets handed out per slot (slot = SM); zeroed by the last warp to leave
    unsigned long long* smTab = nullptr; // [smSlots][smCap]: (launch epoch << 32 | block of 32 units) of the slot's j-th block
    unsigned smSlots = 0, smCap = 0, smEpoch = 0;
    int queueRefillBelow = 24;           // k_traceQueue fetches new rays when fewer lanes than this are still walking
    int nodeLanes = 16;                  // phase switch of the per-lane BVH walk: leave the node loop for the pending leaves when
                                         // fewer lanes than this still want a node (1 = classic while-while)
    float4* partials = nullptr;          // fused reduction: [nFrames][N][S/32] per-warp sums of 32 samples (K1 -> k_sumPartials)
    const int* lists = nullptr;          // [nFrames][N][16] candidate lists (k_buildEntries stage 2): header = element count, -1 = none
    bool pdl = false;                    // launch the trace and reduction kernels as programmatic dependents of the kernel before them
    bool fused = false;                  // in-kernel reduction (needs S % 32 == 0) instead of the ordered per-sample buffer
    bool fast = false;                   // hardware elementary functions instead of cr_math.h
    DevicePose pose;
};

constexpr int kTraceThreads = 128;
constexpr int kPreStride = 4;     // float4 per ommatidium in the `pre` table
constexpr float kTMax = 1e16f;

struct BvhBuildResult {
    float4* nodes = nullptr;
    float4* tris = nullptr;
    int nNodes = 0;
    int nTris = 0;
    double buildMs = 0.0;
};

// Builds the LBVH on `stream` from world-space positions + indices already on the device.
// sceneMin/sceneMax bound all vertices.  leafSize in [1,8].
BvhBuildResult buildLbvh(const float* dPositions, const uint32_t* dIndices, int nTris, const float sceneMin[3],
                         const float sceneMax[3], int leafSize, cudaStream_t stream);

// projection modes (suffix after "__raygen__compound_projection_", libEyeRenderer3/shaders.cu:354-640)
enum Projection : int {
    PROJ_RAW_SAMPLES = 0, PROJ_SINGLE_DIM = 1, PROJ_SINGLE_DIM_FAST = 2, PROJ_SPH_POSITIONWISE = 3,
    PROJ_SPH_ORIENTATIONWISE = 4, PROJ_SPH_SPLIT_ORIENTATIONWISE = 5, PROJ_SPH_ORIENTATIONWISE_IDS = 6,
    PROJ_SPH_POSITIONWISE_IDS = 7, PROJ_UNKNOWN = -1
};

// kernel launchers (cr_kernels.cu)
void launchRngInit(uint4* rng, int N, int S, unsigned long long firstFrame, unsigned long long nGlobal, unsigned long long oFirst,
                   const uint4* jumpTable, cudaStream_t stream);
void launchPrepOmmatidia(const float4* omm, int N, float4* pre, cudaStream_t stream);
void launchBuildEntries(const DeviceScene& sc, const EyeParams& eye, int4* entries, int* lists, cudaStream_t stream);
int candidateListStride();    // ints per (frame, ommatidium) record of the candidate lists
void launchTraceCompound(const DeviceScene& sc, const EyeParams& eye, int gridBlocks, cudaStream_t stream,
                         cudaEvent_t afterTrace = nullptr);   // afterTrace: recorded between the trace and the reduction kernels
void launchProjectVector(int mode, bool fast, const float4* summed, int N, uchar4* frame, int W, int H, cudaStream_t stream);
void launchProjectRaw(bool fast, const float* samples, int N, int S, uchar4* frame, int W, int H, cudaStream_t stream);
void launchBuildProjectionMap(int mode, const float4* omm, int N, uint32_t* map, int W, int H, cudaStream_t stream);
void launchProjectMap(bool ids, bool fast, const uint32_t* map, const float4* summed, uchar4* frame, int W, int H, cudaStream_t stream);
void launchCamera(const DeviceScene& sc, int kind, bool fast, const DevicePose& pose, float s0, float s1, float s2, uchar4* frame,
                  int W, int H, cudaStream_t stream);
void launchTraceRays(const DeviceScene& sc, const float* origins, const float* dirs, const float* tmins, int n, int4* hits,
                     cudaStream_t stream);
void launchSampleTexture(unsigned long long tex, const float* uv, int n, float4* out, cudaStream_t stream);
void launchEvalMath(int fn, const float* a, const float* b, float* out, int n, cudaStream_t stream);
int traceKernelOccupancy();   // resident CTAs per SM for the compound trace kernel
unsigned deviceSmIdLimit(cudaStream_t stream);   // %nsmid: SM identifiers are below this (they need not be contiguous)

}  // namespace cr
