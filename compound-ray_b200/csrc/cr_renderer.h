// cr_renderer.h -- the single implicit renderer behind the C ABI.
//
// Replaces the global state + frame loop of libEyeRenderer3/libEyeRenderer.cpp:81-92,152-195,
// 229-242 and the per-camera device buffers of cameras/CompoundEye.cpp:30-62,98-183.  Headless:
// no GLFW/GL (the reference creates a window at load time, libEyeRenderer.cpp:88-90).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "cr_device.h"
#include "cr_scene.h"
#include "cr_xorwow_jump.h"

namespace cr {

struct CompoundState {                 // device side of one CompoundEye camera
    float4* dOmm = nullptr;
    float4* dPre = nullptr;            // per-ommatidium ray invariants
    uint4* dRng = nullptr;
    float4* dSummed = nullptr;
    const float4* dLastSummed = nullptr;   // per-ommatidium RGB of the most recent frame: dSummed, or the last row of a pose batch
    float* dSamples = nullptr;         // per-sample colour/S, [o][s]
    uint32_t* dMap = nullptr;          // cached pixel -> ommatidium map
    int mapMode = -2, mapW = 0, mapH = 0;
    uint64_t mapEyeVersion = 0, eyeVersion = 1;
    int N = 0, S = 1;
    int rngN = 0, rngS = 0;            // allocation the RNG/summed buffers were sized for
    uint4* dRngAlt = nullptr; size_t rngAltCap = 0;   // second state array: launches cut into frame groups write the final states there (then swapped)
    bool ommDirty = true;
    bool randomsConfigured = false;    // cameras/CompoundEyeDataTypes.h:12
    uint64_t frameIndex = 0;           // frames rendered since the streams were (re)initialised
    uint64_t firstFrame = 0;           // frame offset applied at the next stream initialisation
    uint64_t shardGlobalN = 0, shardFirst = 0;   // ommatidium-range shard: rows [shardFirst, shardFirst+N) of shardGlobalN (0 = whole eye)
    // multi-frame batch buffers (crRenderPoseBatch)
    float* dBatchSamples = nullptr; float4* dBatchSummed = nullptr; DevicePose* dBatchPoses = nullptr;
    size_t batchSampleCap = 0, batchSummedCap = 0, batchPoseCap = 0;
    int4* dEntries = nullptr; size_t entryCap = 0;   // entry frontier [frames][N] (k_buildEntries)
    int* dLists = nullptr; size_t listCap = 0;       // candidate lists [frames][N][16] (k_buildEntries stage 2)
    size_t listsLast = 0;                            // records written by the last launch (0: lists not built)
    bool entriesValid = false; bool entriesLists = false; uint64_t entriesEyeVersion = 0;   // what dEntries/dLists row 0 was built for:
    bool lastPoseValid = false; DevicePose lastPose{}; uint64_t lastPoseEyeVersion = 0; int standingFrames = 0;   // consecutive single frames from one pose
    DevicePose entriesPose{};                                                               // a single frame of this eye at this pose
    float4* dPartials = nullptr; size_t partialCap = 0;   // fused reduction: [frames][N][S/32] warp partials
    // Read-ahead for a standing camera (renderFrame): frames [next, count) of one batched launch wait to be handed out; the sample
    // streams on the device are (count - next) frames ahead of frameIndex.  Dropping them rewinds the streams (needRewind).
    struct ReadAhead {
        int count = 0, next = 0, rowPixels = 0, S = 0, N = 0;
        DevicePose pose{};
        uint64_t eyeVersion = 0;
        bool fused = false, fast = false;
        double traceMsPerFrame = 0.0;
    } ahead;
    bool needRewind = false; uint64_t rewindTo = 0;
    int aheadFrames = 4;               // frames of the next read-ahead launch: doubles while batches are consumed to the end, back to 4 after a drop
    int aheadStreakNeeded = 2;         // standing frames required before a read-ahead: doubles whenever rendered frames had to be dropped
    uchar4* dAheadRows = nullptr; unsigned char* hAheadRows = nullptr; size_t aheadRowCap = 0;   // [frames][N] 8-bit rows, device + pinned host
    double lastSingleFrameMs = 0.0;                      // host time of the last frame rendered on its own
    // wavefront queue (k_traceCompound -> k_traceQueue -> k_shadeQueue): rays of the warp-frames without a candidate list
    float4* dQueueRays = nullptr; int4* dQueueHits = nullptr; int* dQueueWarps = nullptr; unsigned* dQueueCounters = nullptr; size_t queueCap = 0;
    // SM-affine hand-out of the trace kernel's work units (EyeParams::smSeq / smTab)
    unsigned* dSmSeq = nullptr; unsigned long long* dSmTab = nullptr; unsigned smSlots = 0, smCap = 0, smEpoch = 0;
    // debug dump buffers
    float* dDumpO = nullptr; float* dDumpD = nullptr; int4* dDumpH = nullptr; int2* dDumpC = nullptr; size_t dumpCap = 0;
};

class Renderer {
public:
    Renderer();
    ~Renderer();

    void setDevice(int dev);
    void loadScene(const std::string& path);
    void stop();
    void setRenderSize(int w, int h);
    double renderFrame();
    unsigned char* framePointer();
    void saveFrame(const std::string& path);

    HostScene& scene() { return scene_; }
    bool hasScene() const { return loaded_; }
    HostCamera& camera();
    size_t cameraCount();
    size_t cameraIndex() const { return current_; }
    void setCurrentCamera(int index);
    bool compoundActive();

    void setSamples(int s);
    int samples();
    size_t ommatidialCount();
    void setOmmatidia(const Ommatidium* omm, size_t count);

    bool verbose = true;
    bool dumpRays = false;
    int entryFrontier = 1;             // 0: every sample ray starts at the BVH root (A/B switch, crDebugSetEntryFrontier)
    int entryMinSamples = 8;           // below this S (or this many rays per launch) the frontier pass --
    long long entryMinRays = 3ll << 18; // a chain of ~20 dependent node fetches, ~25 us -- costs more than it saves
    // Render modes (crSetRenderMode; environment CR_REDUCE=fused / CR_FAST_MATH=1 preset them for unmodified scripts).
    // Default: ordered sum + cr_math.h = every output bit equal to the CPU checker.
    bool fusedReduce = false;          // K1 sums 32 samples per warp in-kernel (fixed order, RGB equal to rounding) -- no sample buffer
    bool fastMath = false;             // hardware sin/cos/log/pow, as the reference's --use_fast_math build
    int candidateLists = 0;            // per-ommatidium candidate lists: 0 never (default), 1 in batches of >= 4 frames, 2 always.  They paid
                                       // +1.7 % while the trace kernel split its work statically; with units from the work counter they
                                       // are +-0 on the headline workload and -7 % on the 6 374 x 64 pose batches (their pass costs more
                                       // than the lockstep test saves): profiles/r02x_small_batch_ab.json
    // Wavefront queue for the warp-frames that have no candidate list (they need the lists: same launches): 0 never
    // (default since the trace kernel hands its units out dynamically: 23.2 vs 22.3 Grays/s with the queue), 1 in launches
    // that build candidate lists, queueFraction = share of a launch's rays the queue is sized for
    // (a warp that finds it full walks inline), wavefrontRefill = lanes below which k_traceQueue fetches new rays.
    bool zeroCopyFrames = true;        // single_dimension_fast rows written straight into the pinned host frame (no D2H copy queued)
    // Standing camera: once three consecutive renderFrame calls found the same pose, eye and sample count, the following frames
    // (single_dimension_fast, small enough) are rendered several at a time in ONE batched launch -- the streams simply run
    // ahead -- and handed out one per call; any change drops what is left and rewinds the streams (k_rngInit at the frame
    // index actually consumed).  readAheadBudgetMs bounds the GPU time of one such launch, readAheadMaxRays the frame size.
    bool frameGroups = true;           // batched launches of small frames: units of (32 rays, a group of frames) instead of (32 rays, all frames)
    bool readAhead = true;
    double readAheadBudgetMs = 1.5;
    long long readAheadMaxRays = 4ll << 20;   // (a rewind re-initialises every stream: 0.65 ms per million)
    bool standingFrontier = true;      // small frames: build the frontier once the camera has stood still for three frames, then reuse it
    bool spinSync = false;             // cudaDeviceScheduleSpin (CR_SPIN_SYNC=1, before the first GPU use)
    int entryMaxLevels = 256;          // frontier pass: levels it may descend (a latency chain: one dependent node fetch per level)
    int chunkUnits = 1;                // single-frame trace kernel: units of 32 rays per counter fetch (larger chunks of one ommatidium's
                                       // units measured slower: 2 -> 623, 4 -> 723, 8 -> 993 us per headline frame)
    bool pdl = true;                   // programmatic dependent launch of the trace kernel behind the frontier pass and of the reduction behind the trace
    int smAffine = 1;                  // trace kernel: blocks of 32 units stay on one SM (per-SM tickets, EyeParams::smSeq); needs dynamicChunks
    int smAffineMinBlocks = 48;        // ... in launches of at least this many blocks per SM: an SM that holds a heavy block runs 32 latency-bound warps at once,
                                       // so the tail of a launch is longer than with mixed units (21 blocks per SM: 0.187 vs 0.148 ms per frame of 1000 x 3200)
    bool dynamicChunks = true;         // trace kernel: ray units handed out through a global counter instead of a static grid-stride split
    int wavefront = 0;
    int nodeLanes = 16;                // phase switch of the per-lane BVH walk (EyeParams::nodeLanes); 1 = classic while-while
    int wavefrontRefill = 24;
    double queueFraction = 0.35;
    int width() const { return W_; }
    int height() const { return H_; }

    // additive API
    void copyOmmatidialData(float* outRgb);                       // float RGB per ommatidium of the last frame
    double renderPoseBatch(const float* poses12, size_t count, unsigned char* outRgba, void* outDevice);
    static const std::vector<uint32_t>& xorwowTable();
    // multi-GPU data plane (cr_comm.cpp): NCCL communicator of one process per GPU
    void commUniqueId(void* out128);
    void commInit(const void* id128, int nRanks, int rank);
    void commDestroy();
    int commRank() const { return commRank_; }
    int commSize() const { return commSize_; }
    int commNcclVersion();
    void allGatherRows(const void* sendDevice, void* recvDevice, size_t bytesPerRank);
    double renderPoseBatchSharded(const float* poses12, size_t count, unsigned char* outHost, void* outDevice, size_t chunkPoses,
                                  uint64_t firstFrame);
    void setFirstFrame(uint64_t k);
    void setOmmatidialShard(uint64_t globalCount, uint64_t first);
    double lastTraceMs() { wantTraceEvents_ = true; return lastTraceMs_; }   // per-frame CUDA events only once somebody asks
    unsigned long long launchCount() const { return launches_; }
    double bvhBuildMs() const { return bvh_.buildMs; }
    int lastBatchFrames() const { return lastBatchFrames_; }

    // debug / parity access
    void debugCopyBvh(float* nodes, float* tris);
    int debugNodeCount() const { return bvh_.nNodes; }
    void debugCopyRngStates(uint32_t* out8);                      // [N*S][8] in reference stream-id order
    size_t debugCopyLastRayCounts(int32_t* counts2);
    size_t debugCopyCandidateLists(int32_t* out, size_t records);   // 16 ints per (frame, ommatidium) of the last launch
    bool profileFrame = false;                                      // extra event marks inside renderFrame (debugFrameBreakdown)
    void debugFrameBreakdown(float* out3);
    unsigned long long debugLastQueuedRays();                       // rays the last trace launch sent through the wavefront queue
    size_t debugCopyLastRays(float* origins, float* dirs, int32_t* hits4);
    void debugTraceRays(const float* origins, const float* dirs, const float* tmins, int n, int32_t* hits8);
    void debugCopyProjectionMap(uint32_t* out);
    void debugSampleTexture(int index, const float* uv, int n, float* out4);
    void debugEvalMath(int fn, const float* a, const float* b, float* out, int n);
    void ensureDevice();

private:
    void uploadScene();
    void freeScene();
    CompoundState& compoundState(size_t camIdx);
    void prepareCompound(CompoundState& cs, HostCamera& cam);
    void launchCompound(CompoundState& cs, const HostCamera& cam, const Pose& pose, uchar4* fastRow = nullptr, int fastRowCount = 0,
                        uchar4* fastRowHost = nullptr);
    void launchCompoundBatch(CompoundState& cs, const DevicePose* dPoses, int nFrames, float* dSamples, float4* dSummed,
                             uchar4* fastRow = nullptr, const Pose* samePose = nullptr);
    bool fusedActive(const CompoundState& cs, const HostCamera& cam) const;
    void ensurePartials(CompoundState& cs, size_t frames);
    void ensureBatchBuffers(CompoundState& cs, size_t F, bool fused);
    size_t batchFramesPerLaunch(const CompoundState& cs, size_t count, bool fused) const;
    void noteSinglePose(CompoundState& cs, const DevicePose& pose);
    bool consumeReadAhead(CompoundState& cs, const HostCamera& cam, bool eligible);
    bool launchReadAhead(CompoundState& cs, const HostCamera& cam);
    void dropReadAhead(CompoundState& cs);
    bool readAheadEligible(const CompoundState& cs, const HostCamera& cam) const;
    void ensureQueue(CompoundState& cs, size_t frames);
    size_t queueRaysFor(const CompoundState& cs, size_t frames) const;
    void attachQueue(CompoundState& cs, EyeParams& ep);
    void attachWorkCounter(CompoundState& cs, EyeParams& ep);
    bool entryFrontierActive(const CompoundState& cs, int frames) const;   // frames = poses covered by the launch
    void buildEntries(CompoundState& cs, EyeParams& ep);
    void project(CompoundState& cs, const HostCamera& cam);
    void ensureFrame();
    void freeCompound(CompoundState& cs);
    static DevicePose toDevicePose(const Pose& p);

    HostScene scene_;
    bool loaded_ = false;
    size_t current_ = 0;
    int W_ = 400, H_ = 400;                                       // libEyeRenderer.cpp:85-86
    int device_ = -1;
    bool deviceReady_ = false;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t evA_ = nullptr, evB_ = nullptr;
    cudaEvent_t evMark_[4] = {nullptr, nullptr, nullptr, nullptr};
    int numSMs_ = 148;
    unsigned smIdLimit_ = 0;            // %nsmid of the device (deviceSmIdLimit), queried at the first SM-affine launch
    int traceOcc_ = 8;

    // device scene
    DeviceScene dscene_;
    BvhBuildResult bvh_;
    float* dPositions_ = nullptr;
    uint32_t* dIndices_ = nullptr;
    uint4* dPrims_ = nullptr;
    float2* dUvs_ = nullptr;
    float4* dColors_ = nullptr;
    MeshRec* dMeshes_ = nullptr;
    uint4* dJumpTable_ = nullptr;                                 // cr_xorwow_jump.h tables (k_rngInit)
    std::vector<cudaArray_t> texArrays_;
    std::vector<cudaTextureObject_t> texObjects_;

    std::map<size_t, CompoundState> compound_;

    uchar4* dFrame_ = nullptr;
    unsigned char* hFrame_ = nullptr;                             // pinned, mapped into the device address space
    unsigned char* hFrameDev_ = nullptr;                          // device-side address of hFrame_
    bool hostMirrorsDevice_ = false;                              // hFrame_ == dFrame_ byte for byte (after a full copy / both zero)
    size_t frameCap_ = 0;
    bool hostFrameFresh_ = false;                                 // hFrame_ already holds the last rendered frame
    bool frameWasFetched_ = true;                                 // the caller read the previous frame (getFramePointer/saveFrameAs)
    static constexpr size_t kEagerFrameBytes = size_t(1) << 20;
    int frameW_ = 0, frameH_ = 0;

    void* comm_ = nullptr;                                        // ncclComm_t (cr_comm.cpp)
    cudaStream_t commStream_ = nullptr;
    int commRank_ = 0, commSize_ = 1;

    double lastTraceMs_ = 0.0;
    bool wantTraceEvents_ = false;                                // renderFrame records its event pair only after crGetLastTraceMs was used
    int lastBatchFrames_ = 1;
    unsigned lastQueueCap_ = 0;
    unsigned long long launches_ = 0;
};

Renderer& renderer();

}  // namespace cr
