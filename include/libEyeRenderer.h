/* include/libEyeRenderer.h -- C ABI of libEyeRenderer3.so (B200-native build).
 *
 * Drop-in boundary: the 35 extern "C" entry points of the reference's
 * libEyeRenderer3/libEyeRenderer.h:16-67 with identical names, argument order/meaning and
 * return conventions, so python-examples/eyeRendererHelperFunctions.py (ctypes signatures at
 * :40-71) and the python-examples run unchanged.  Each declaration cites the reference
 * definition it replaces (libEyeRenderer3/libEyeRenderer.cpp).  One implicit global renderer per
 * process, not thread-safe, every call synchronous -- as in the reference.
 *
 * Differences that are deliberate:
 *   - headless: loading the library creates no window/GL context (reference: libEyeRenderer.cpp:88-90);
 *     displayFrame() is a no-op;
 *   - C++ exceptions never cross this boundary: failures are printed to stderr as
 *     "[PyEye] ERROR: ..." and the call returns its neutral value (0 / false / "" / NULL);
 *   - there is no CPU fallback: rendering calls fail loudly when no CUDA device is present.
 *
 * Symbols prefixed cr* are ADDITIONS (never substitutes) for batch rendering, multi-GPU use and
 * parity testing.
 */
#ifndef LIB_EYE_RENDERER_3_B200_H
#define LIB_EYE_RENDERER_3_B200_H

#include <stddef.h>
#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

/* libEyeRenderer.h:8-14 -- 8 floats, identical to one .eye line and to the device row */
struct OmmatidiumPacket {
    float posX, posY, posZ;
    float dirX, dirY, dirZ;
    float acceptanceAngle;
    float focalpointOffset;
};

/* layout-identical to CUDA's float3 (the reference returns float3 by value, libEyeRenderer.h:65-66) */
#if defined(__VECTOR_TYPES_H__)
typedef float3 crFloat3;
#else
typedef struct crFloat3 { float x, y, z; } crFloat3;
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- configuration / frame loop ---- */
void setVerbosity(bool v);                      /* libEyeRenderer.cpp:211-214 */
void loadGlTFscene(const char* filepath);       /* :215-220  ASCII .gltf, builds the BVH */
void stop(void);                                /* :289-295  frees device state */
void setRenderSize(int w, int h);               /* :221-228 */
double renderFrame(void);                       /* :229-242  returns host-timed milliseconds */
void displayFrame(void);                        /* :243-258  no-op (headless) */
void saveFrameAs(char* ppmFilename);            /* :259-269  binary P6, top-down */
unsigned char* getFramePointer(void);           /* :270-275  uchar4[H][W], row 0 = bottom, library-owned */

/* ---- camera control ---- */
size_t getCameraCount(void);                    /* :306-309 */
void nextCamera(void);                          /* :310-313 */
void previousCamera(void);                      /* :322-325 */
size_t getCurrentCameraIndex(void);             /* :314-317 */
const char* getCurrentCameraName(void);         /* :318-321  library-owned */
void gotoCamera(int index);                     /* :326-329  wraps modulo the camera count */
bool gotoCameraByName(char* name);              /* :330-340  on a miss: false, camera 0 selected */
void setCameraPosition(float x, float y, float z);                      /* :341-344 */
void getCameraPosition(float* x, float* y, float* z);                   /* :345-351 (float& == float*) */
void setCameraLocalSpace(float lxx, float lxy, float lxz,               /* :352-359  x, y, z axes */
                         float lyx, float lyy, float lyz,
                         float lzx, float lzy, float lzz);
void rotateCameraAround(float angle, float axisX, float axisY, float axisZ);         /* :360-363 radians */
void rotateCameraLocallyAround(float angle, float axisX, float axisY, float axisZ);  /* :364-367 */
void translateCamera(float x, float y, float z);                        /* :368-371 */
void translateCameraLocally(float x, float y, float z);                 /* :372-375 */
void resetCameraPose(void);                                             /* :376-379 */
void setCameraPose(float posX, float posY, float posZ,                  /* :380-388 reset; rot X,Y,Z; move */
                   float rotX, float rotY, float rotZ);

/* ---- compound eye ---- */
bool isCompoundEyeActive(void);                                         /* :393-396 */
void setCurrentEyeSamplesPerOmmatidium(int s);                          /* :397-403 max(1,s); resets RNG streams */
int getCurrentEyeSamplesPerOmmatidium(void);                            /* :404-411 -1 if not compound */
void changeCurrentEyeSamplesPerOmmatidiumBy(int s);                     /* :412-418 */
size_t getCurrentEyeOmmatidialCount(void);                              /* :419-426 0 if not compound */
void setOmmatidia(struct OmmatidiumPacket* omms, size_t count);         /* :427-452 copied; caller keeps ownership */
const char* getCurrentEyeDataPath(void);                                /* :453-460 "" if not compound */
void setCurrentEyeShaderName(char* name);                               /* :461-468 projection suffix */

/* ---- scene queries ---- */
bool isInsideHitGeometry(float x, float y, float z, char* name);        /* :470-473 */
crFloat3 getGeometryMaxBounds(char* name);                              /* :474-477 */
crFloat3 getGeometryMinBounds(char* name);                              /* :478-481 */

/* =====================  additions (not in the reference)  ===================== */
/* Select the CUDA device (default: env CR_DEVICE, else 0). Must precede the first GPU use. */
void crSetDevice(int device);
/* Float RGB per ommatidium of the last compound frame (sum over samples of colour/S): 3*N floats. */
void crGetOmmatidialData(float* outRgb);
/* Render `count` consecutive frames of the current compound eye, one per pose.  poses: 12 floats
 * each (position, x axis, y axis, z axis).  Output row p = the single_dimension_fast row of pose p
 * (4 bytes per ommatidium).  outDevice (a CUDA device pointer) takes precedence over outHost.
 * Returns host-timed milliseconds. */
double crRenderPoseBatch(const float* poses, size_t count, unsigned char* outHost, void* outDevice);
/* outDevice is written from the library's own (non-blocking) CUDA stream and the call returns after that stream is idle:
 * the buffer must not have work pending on any other stream when the call is made (e.g. a torch.zeros fill still
 * queued on torch's stream -- synchronise that stream first), and may be read from any stream once the call returned.
 * crGetOmmatidialData after this call returns the float RGB of the LAST pose of the batch. */
/* Position the RNG streams as if `frame` frames had already been rendered (pose sharding/restart);
 * takes effect at the next stream initialisation, which this call forces, and is consumed by it: any LATER
 * re-initialisation (setCurrentEyeSamplesPerOmmatidium, setOmmatidia with a new count, crSetOmmatidialShard) restarts
 * at frame 0 as the reference does. */
void crSetFirstFrame(uint64_t frame);
/* Ommatidium-range sharding (one pose, large N*S; SURVEY 8e secondary partition): declare that the rows given to
 * setOmmatidia are rows [firstIndex, firstIndex+count) of an eye of globalCount ommatidia.  Sample stream ids then
 * use the global indices (id = globalCount*s + firstIndex + o, shaders.cu:680-685), so the gathered per-ommatidium
 * results equal the unsharded frame bit for bit.  globalCount = 0 switches back.  Forces a stream initialisation.
 * The next render fails (logged, nothing drawn) unless firstIndex + count <= globalCount and globalCount * S < 2^31. */
void crSetOmmatidialShard(uint64_t globalCount, uint64_t firstIndex);
/* ---- multi-GPU data plane: one process per GPU, NCCL all-gather of the per-pose rows, owned by the library.
 * NCCL is bound at run time (libnccl.so.2 -- a copy already loaded into the process, e.g. torch's, is reused; CR_NCCL_LIB
 * overrides), so nothing here is needed, or loaded, for single-GPU use.  Functions returning int give 0 on success, -1 on
 * failure (logged).  The reference has no counterpart: its scripts render on one GPU. */
/* Rank 0: fills the 128-byte NCCL unique id, which the launcher passes to every rank by its own means (MPI,
 * torch.distributed broadcast, a file). */
int crCommGetUniqueId(void* out128);
/* Every rank, collectively: joins the communicator on the library's device (crSetDevice first).  nRanks == 1 creates no
 * communicator and does not touch NCCL (id128 is ignored; do not call crCommGetUniqueId for it either: a one-rank job then
 * runs exactly as without the data plane). */
int crCommInit(const void* id128, int nRanks, int rank);
void crCommDestroy(void);
int crCommRank(void);                 /* 0 without a communicator */
int crCommSize(void);                 /* 1 without a communicator */
int crCommNcclVersion(void);          /* e.g. 22809; 0 when NCCL cannot be loaded */
/* recvDevice[q * bytesPerRank ...] := rank q's sendDevice on every rank (ncclAllGather on the library's communication
 * stream; sendDevice may be the rank's own slot of recvDevice).  Returns when recvDevice is complete.  Use after
 * crRenderPoseBatch(..., outDevice = send slot) for ommatidium-range or hand-made pose shards. */
int crAllGatherRows(const void* sendDevice, void* recvDevice, size_t bytesPerRank);
/* One logical `count`-pose run over all ranks of the communicator (or this GPU alone without one).  Every rank passes the
 * SAME poses (12 floats each); rank r renders the contiguous block [r*count/R ...) with its sample streams positioned at
 * firstFrame + its first pose, so pose k is frame firstFrame + k of every stream for ANY number of ranks and the result
 * does not depend on R.  The block is rendered in chunks of chunkPoses poses (0 = one chunk); each chunk's rows are
 * gathered into their final places on the communication stream while the next chunk is traced.  Every rank receives all
 * count*N*4 bytes, in pose order, in outDevice (device memory) and/or outHost.  Returns host milliseconds (-1 on failure);
 * crGetLastTraceMs then gives the device time of this rank's trace launches. */
double crRenderPoseBatchSharded(const float* poses, size_t count, unsigned char* outHost, void* outDevice, size_t chunkPoses,
                                uint64_t firstFrame);
/* Render modes of the compound path; a negative argument leaves that switch as it is.  Both default to 0 (also preset by
 * the environment variables CR_REDUCE=fused and CR_FAST_MATH=1, for scripts that cannot call this).
 *   fusedReduction 0: every sample's colour/S is stored and summed in sample order -- the reference's sequential fp32
 *                     sum (shaders.cu:341-347, :730); every output bit equals the CPU checker.
 *                  1: the trace kernel sums the 32 samples of a warp with a shuffle butterfly and a second, tiny kernel
 *                     adds the S/32 partial sums in a fixed order (needs S % 32 == 0, else the ordered path runs).  Rays,
 *                     hits and per-sample colours are unchanged; the float RGB per ommatidium agrees with mode 0 to
 *                     fp32 rounding of a different addition order (~1e-7 relative; 8-bit frames differ in < 1e-4 of
 *                     the bytes, by one step).  No per-sample buffer: 12 B/ray of HBM traffic become 0.5 B/ray.
 *   fastMath       0: sin/cos/log/pow/asin/atan2 are the specified binary32 algorithms of csrc/cr_math.h.
 *                  1: the hardware approximations (__sincosf, __logf, __powf) the reference itself runs -- its device
 *                     code is built with --use_fast_math (CMakeLists.txt:142).  Rays differ in the last bits; the mode
 *                     is held to max |dRGB| <= 1/255, mean <= 1e-4 against the checker, not to bit-exactness.
 * crGetRenderMode returns fusedReduction | fastMath << 1. */
void crSetRenderMode(int fusedReduction, int fastMath);
int crGetRenderMode(void);
/* CUDA-event time of the last compound trace launch(es), milliseconds (crRenderPoseBatch: always; renderFrame: from the
 * first call of this function on -- until then the host-side frame time is returned, so that untimed loops pay no events). */
double crGetLastTraceMs(void);
/* Kernels launched by this library so far. */
unsigned long long crGetLaunchCount(void);
/* Build time of the BVH of the loaded scene, milliseconds. */
double crGetBvhBuildMs(void);
/* Frames covered by each trace launch of the last crRenderPoseBatch call. */
int crGetLastBatchFrames(void);

/* ---- parity / debug access (used by tests and the roofline counters only) ---- */
size_t crDebugGetTriangleCount(void);
size_t crDebugGetVertexCount(void);
size_t crDebugGetMeshCount(void);
void crDebugCopyTriangles(float* out9);          /* host loader output: v0, e1, e2 per flattened prim */
void crDebugCopyTriangleMesh(int32_t* out);      /* mesh group of each flattened primitive */
void crDebugCopyMeshInfo(int32_t* out4, float* baseColor4);  /* per mesh: colorType, hasUV, texture, material */
void crDebugCopyCornerAttributes(float* uv6, float* col12);  /* per prim: 3 x uv, 3 x rgba (either may be NULL) */
void crDebugCopyCameraPose(float* out12);        /* current camera: position, x, y, z axes */
void crDebugCopyCameraScale(float* out3);
int crDebugGetCameraKind(void);                  /* 0 perspective, 1 panoramic, 2 orthographic, 3 compound */
void crDebugCopyOmmatidia(float* out8);
bool crDebugDecodeImageFile(const char* path, int* w, int* h);  /* PNG / baseline JPEG through the loader's decoder */
void crDebugCopyDecodedImage(unsigned char* outRgba);
int crDebugGetMissShader(void);
size_t crDebugGetTextureCount(void);
void crDebugGetTextureSize(int index, int* w, int* h);
void crDebugCopyTexture(int index, unsigned char* outRgba);   /* decoded RGBA8, row 0 first */
size_t crDebugGetBvhNodeCount(void);
void crDebugCopyBvh(float* nodes16, float* tris12);
/* Host evaluation of curand_init(seed, subsequence, offset) with the library's own jump-ahead tables (the ones
 * k_rngInit uses on the device): out6 = d, v[0..4].  No GPU needed; subsequence < 2^32. */
void crDebugXorwowInit(uint64_t seed, uint64_t subsequence, uint64_t offset, uint32_t* out6);
void crDebugSetRayDump(bool on);
/* A/B switch of the per-ommatidium entry frontier (default: on for S >= 8 and N*S >= 786432 rays per frame);
 * negative thresholds keep the current value. */
void crDebugSetEntryFrontier(int on, int minSamples, long long minRaysPerFrame);
/* Switch of the per-ommatidium candidate lists (they need the entry frontier and S % 32 == 0): the frontier pass also
 * flattens what an ommatidium's sample cone can reach into a list of <= 15 pre-leaf BVH nodes, which the 32 samples a warp
 * holds of that ommatidium test in lockstep instead of walking the tree per lane.  0 = never (default since round 2: with the
 * trace kernel's dynamic work distribution they no longer pay), 1 = in launches that cover at least 4 frames (the extra
 * latency of building them hides behind a batch, not behind one synchronous frame), 2 = always.  The closest hit is the
 * same either way. */
void crDebugSetCandidateLists(int on);
/* The candidate lists of the last trace launch: 16 ints per (frame, ommatidium) -- [0] = element count (-1: none, the
 * frontier is walked per lane; 0: the cone reaches no leaf), [1..] = node << 2 | reachable-leaf mask.  Returns the
 * number of records copied (<= records; 0 when the launch built none). */
size_t crDebugCopyCandidateLists(int32_t* out, size_t records);
/* Wavefront queue (launches that build candidate lists): the warp-frames whose ommatidium has NO list -- cones that graze
 * the scene, whose 32 per-lane walks drift far apart -- are not walked inside the trace kernel; their rays go to a queue,
 * k_traceQueue traces it with dynamic ray fetch (a lane that finishes pulls the next ray) and k_shadeQueue shades and
 * reduces them in the pushing warp's lane order.  Same hits, same output bits.  on: 0 never (default: with the trace
 * kernel's dynamic work distribution the queue measured slower), 1 in launches that build candidate lists; refillBelow
 * (1..32, <= 0 keeps it): lanes still walking below which a warp fetches new rays; queueFraction (< 0 keeps it): share
 * of a launch's rays the queue is sized for -- warps that find it full walk inline. */
void crDebugSetWavefront(int on, int refillBelow, double queueFraction);
/* Phase switch of the per-lane BVH walk (trace kernel and k_traceQueue): the warp leaves its node loop for the pending
 * leaves as soon as fewer than `lanes` (1..32) lanes still want a node; 1 = classic while-while (a lane at a leaf waits
 * for every other lane).  Each lane's own sequence of node and triangle tests does not change: same hits. */
void crDebugSetNodeLanes(int lanes);
/* Batched launches of small frames (default on): when a launch has too few units of 32 rays for the trace kernel's work
 * counter to balance, its frames are cut into groups and a unit becomes (32 rays, one group of frames); the lanes of a later
 * group step the launch's starting stream states over the draws of the frames before it.  Same frames bit for bit. */
void crDebugSetFrameGroups(int on);
/* Read-ahead for a standing camera (default on): once three consecutive renderFrame calls found the same pose, eye and
 * sample count, the following single_dimension_fast frames of up to 4M rays are rendered several at a time in one
 * batched launch of at most `budgetMs` (default 1.5; <= 0 keeps it) and handed out one per call; anything that changes
 * (pose, eye, samples, mode, size, a batch call, crSetFirstFrame ...) drops what is left and rewinds the sample streams to
 * the frame the caller has reached.  The frames are those of one-at-a-time rendering bit for bit; renderFrame's return
 * value is the time THAT call took (the launching call carries the batch). */
void crDebugSetReadAhead(int on, double budgetMs);
/* Device-side breakdown of renderFrame for a compound eye: with the profile on, event marks are recorded inside the frame
 * and crDebugFrameBreakdown returns the milliseconds of [frontier pass (+ counter reset), trace kernel(s), reduction kernel]
 * of the last frame (-1 when there is none). */
void crDebugSetFrameProfile(int on);
void crDebugFrameBreakdown(float* out3);
/* Trace kernel work distribution: 1 (default) = units of 32 rays handed out through a global counter, 0 = static
 * grid-stride split.  Which warp traces a ray does not enter its result. */
void crDebugSetDynamicChunks(int on);
/* ... and, on top of the counter, SM-affine hand-out: blocks of 32 consecutive units stay on ONE SM, whose warps share them
 * through a per-SM ticket counter (at S = 1024 an SM's 32 resident warps trace one ommatidium together: its nodes are fetched
 * into that SM's L1 once).  on = 1 (default), in launches of at least minBlocksPerSm (default 48) blocks per SM.  Result-neutral. */
void crDebugSetSmAffine(int on, int minBlocksPerSm);
/* single_dimension_fast rows written by the reduction kernel straight into the pinned (mapped) host frame when the caller
 * reads every frame: 1 (default) on, 0 = always a device-to-host copy behind the frame's kernels.  Same bytes. */
void crDebugSetZeroCopy(int on);
unsigned long long crDebugLastQueuedRays(void);  /* rays of the last trace launch that went through the queue */
size_t crDebugCopyLastRays(float* origins3, float* dirs3, int32_t* hits4);  /* stream-id order N*s+o */
size_t crDebugCopyLastRayCounts(int32_t* counts2); /* (BVH nodes fetched, triangles tested) per dumped ray, same order */
void crDebugCopyRngStates(uint32_t* out8);       /* d, v0..v4, flag, extra bits; stream-id order */
void crDebugTraceRays(const float* origins3, const float* dirs3, const float* tmins, int n, int32_t* hits8);
void crDebugCopyProjectionMap(uint32_t* out);    /* W*H ommatidium indices of the cached map */
void crDebugSampleTexture(int index, const float* uv2, int n, float* outRgba4);  /* hardware tex2D at n (u,v) */
void crDebugEvalMath(int fn, const float* a, const float* b, float* out, int n);

#ifdef __cplusplus
}
#endif
#endif
