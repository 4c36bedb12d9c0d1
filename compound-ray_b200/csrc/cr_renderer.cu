// cr_renderer.cu -- host orchestration: device buffers, frame loop, pose batches, debug access.
// See cr_renderer.h for the reference code this replaces.
#include "cr_renderer.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>

namespace cr {

#define CR_CUDA(x)                                                                                          \
    do {                                                                                                    \
        cudaError_t e_ = (x);                                                                               \
        if (e_ != cudaSuccess)                                                                              \
            throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(e_) + " at " #x);     \
    } while (0)

namespace {
template <typename T>
T* dallocT(size_t n)
{
    T* p = nullptr;
    CR_CUDA(cudaMalloc(&p, sizeof(T) * (n ? n : 1)));
    return p;
}
template <typename T>
void dfree(T*& p)
{
    if (p) cudaFree(p);
    p = nullptr;
}
int projectionFromName(const std::string& n)
{
    if (n == "raw_ommatidial_samples") return PROJ_RAW_SAMPLES;
    if (n == "single_dimension") return PROJ_SINGLE_DIM;
    if (n == "single_dimension_fast") return PROJ_SINGLE_DIM_FAST;
    if (n == "spherical_positionwise") return PROJ_SPH_POSITIONWISE;
    if (n == "spherical_orientationwise") return PROJ_SPH_ORIENTATIONWISE;
    if (n == "spherical_split_orientationwise") return PROJ_SPH_SPLIT_ORIENTATIONWISE;
    if (n == "spherical_orientationwise_ids") return PROJ_SPH_ORIENTATIONWISE_IDS;
    if (n == "spherical_positionwise_ids") return PROJ_SPH_POSITIONWISE_IDS;
    return PROJ_UNKNOWN;
}
}  // namespace

Renderer& renderer()
{
    static Renderer r;
    return r;
}

Renderer::Renderer()
{
    if (const char* e = getenv("CR_ENTRY_FRONTIER")) entryFrontier = atoi(e);
    if (const char* e = getenv("CR_ENTRY_MIN_S")) entryMinSamples = atoi(e);
    if (const char* e = getenv("CR_ENTRY_MIN_RAYS")) entryMinRays = atoll(e);
    if (const char* e = getenv("CR_REDUCE")) fusedReduce = (std::string(e) == "fused" || std::string(e) == "1");
    if (const char* e = getenv("CR_FAST_MATH")) fastMath = atoi(e) != 0;
    if (const char* e = getenv("CR_CANDIDATE_LISTS")) candidateLists = atoi(e);
    if (const char* e = getenv("CR_ZERO_COPY")) zeroCopyFrames = atoi(e) != 0;
    if (const char* e = getenv("CR_DYNAMIC_CHUNKS")) dynamicChunks = atoi(e) != 0;
    if (const char* e = getenv("CR_CHUNK_UNITS")) chunkUnits = atoi(e);
    if (const char* e = getenv("CR_SM_AFFINE")) smAffine = atoi(e);
    if (const char* e = getenv("CR_PDL")) pdl = atoi(e) != 0;
    if (const char* e = getenv("CR_SM_AFFINE_MIN_BLOCKS")) smAffineMinBlocks = atoi(e);
    if (const char* e = getenv("CR_ENTRY_MAX_LEVELS")) entryMaxLevels = atoi(e);
    if (const char* e = getenv("CR_STANDING_FRONTIER")) standingFrontier = atoi(e) != 0;
    if (const char* e = getenv("CR_FRAME_GROUPS")) frameGroups = atoi(e) != 0;
    if (const char* e = getenv("CR_READ_AHEAD")) readAhead = atoi(e) != 0;
    if (const char* e = getenv("CR_READ_AHEAD_MS")) readAheadBudgetMs = atof(e);
    if (const char* e = getenv("CR_SPIN_SYNC")) spinSync = atoi(e) != 0;
    if (const char* e = getenv("CR_NODE_LANES")) nodeLanes = atoi(e);
    if (const char* e = getenv("CR_WAVEFRONT")) wavefront = atoi(e);
    if (const char* e = getenv("CR_WAVEFRONT_REFILL")) wavefrontRefill = atoi(e);
    if (const char* e = getenv("CR_QUEUE_FRACTION")) queueFraction = atof(e);
}
Renderer::~Renderer()
{
    // process teardown: the CUDA context may already be gone; do not touch the device here
}

const std::vector<uint32_t>& Renderer::xorwowTable()
{
    static const std::vector<uint32_t> table = xorwow::buildTable();
    return table;
}

void Renderer::setDevice(int dev)
{
    if (deviceReady_ && dev != device_) throw std::runtime_error("crSetDevice must be called before the first GPU use");
    device_ = dev;
}

void Renderer::ensureDevice()
{
    if (deviceReady_) { CR_CUDA(cudaSetDevice(device_)); return; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw std::runtime_error("no CUDA device available: this renderer has no CPU path (sm_100a kernels only)");
    if (device_ < 0) {
        const char* env = getenv("CR_DEVICE");
        device_ = env ? atoi(env) : 0;
    }
    if (device_ >= count) throw std::runtime_error("requested CUDA device index out of range");
    if (spinSync) { if (cudaSetDeviceFlags(cudaDeviceScheduleSpin) != cudaSuccess) cudaGetLastError(); }   // (before the context exists)
    CR_CUDA(cudaSetDevice(device_));
    CR_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    CR_CUDA(cudaEventCreate(&evA_));
    CR_CUDA(cudaEventCreate(&evB_));
    cudaDeviceProp prop;
    CR_CUDA(cudaGetDeviceProperties(&prop, device_));
    numSMs_ = prop.multiProcessorCount;
    traceOcc_ = traceKernelOccupancy();
    deviceReady_ = true;
    if (verbose)
        std::cout << "[PyEye] CUDA device " << device_ << ": " << prop.name << ", " << numSMs_ << " SMs, trace kernel "
                  << traceOcc_ << " CTAs/SM" << std::endl;
}

DevicePose Renderer::toDevicePose(const Pose& p)
{
    DevicePose d;
    d.px = p.pos.x; d.py = p.pos.y; d.pz = p.pos.z;
    d.xx = p.ax.x; d.xy = p.ax.y; d.xz = p.ax.z;
    d.yx = p.ay.x; d.yy = p.ay.y; d.yz = p.ay.z;
    d.zx = p.az.x; d.zy = p.az.y; d.zz = p.az.z;
    return d;
}

// ------------------------------------------------------------------------------------------
void Renderer::freeCompound(CompoundState& cs)
{
    dfree(cs.dOmm); dfree(cs.dPre); dfree(cs.dRng); dfree(cs.dRngAlt); dfree(cs.dSummed); dfree(cs.dSamples); dfree(cs.dMap);
    cs.rngAltCap = 0;
    dfree(cs.dDumpO); dfree(cs.dDumpD); dfree(cs.dDumpH); dfree(cs.dDumpC);
    dfree(cs.dBatchSamples); dfree(cs.dBatchSummed); dfree(cs.dBatchPoses); dfree(cs.dEntries); dfree(cs.dPartials); dfree(cs.dLists);
    dfree(cs.dQueueRays); dfree(cs.dQueueHits); dfree(cs.dQueueWarps); dfree(cs.dQueueCounters);
    dfree(cs.dSmSeq); dfree(cs.dSmTab);
    cs.smSlots = cs.smCap = 0;
    dfree(cs.dAheadRows);
    if (cs.hAheadRows) cudaFreeHost(cs.hAheadRows);
    cs.hAheadRows = nullptr;
    cs.aheadRowCap = 0;
    cs.ahead = CompoundState::ReadAhead{};
    cs.needRewind = false;
    cs.lastPoseValid = false;
    cs.standingFrames = 0;
    cs.queueCap = 0;
    cs.entryCap = 0;
    cs.listCap = 0;
    cs.entriesValid = false;
    cs.partialCap = 0;
    cs.batchSampleCap = cs.batchSummedCap = cs.batchPoseCap = 0;
    cs.dumpCap = 0;
    cs.rngN = cs.rngS = 0;
    cs.mapMode = -2;
}

void Renderer::freeScene()
{
    if (!deviceReady_) return;
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    for (auto& kv : compound_) freeCompound(kv.second);
    compound_.clear();
    for (auto t : texObjects_) cudaDestroyTextureObject(t);
    for (auto a : texArrays_) cudaFreeArray(a);
    texObjects_.clear();
    texArrays_.clear();
    dfree(bvh_.nodes); dfree(bvh_.tris);
    bvh_ = BvhBuildResult();
    dfree(dPositions_); dfree(dIndices_); dfree(dPrims_); dfree(dUvs_); dfree(dColors_); dfree(dMeshes_);
    dscene_ = DeviceScene();
}

void Renderer::stop()
{
    if (verbose) std::cout << "[PyEye] Cleaning eye renderer resources." << std::endl;
    freeScene();
    if (deviceReady_) {
        dfree(dFrame_);
        if (hFrame_) cudaFreeHost(hFrame_);
        hFrame_ = nullptr;
        frameCap_ = 0;
        frameW_ = frameH_ = 0;
        dfree(dJumpTable_);
    }
    scene_ = HostScene();
    loaded_ = false;
    current_ = 0;
}

void Renderer::loadScene(const std::string& path)
{
    HostScene sc = loadGltfScene(path, verbose);
    freeScene();
    scene_ = std::move(sc);
    loaded_ = true;
    current_ = 0;
    (void)camera();   // finalize() touches getCamera(), which creates the default camera (MulticamScene.cpp:855-857,911-927)
    int count = 0;
    if (cudaGetDeviceCount(&count) == cudaSuccess && count > 0) {
        ensureDevice();
        uploadScene();
    } else if (verbose) {
        std::cout << "[PyEye] no CUDA device visible: scene parsed on the host only; rendering will fail." << std::endl;
    }
}

void Renderer::uploadScene()
{
    const size_t T = scene_.triangleCount(), V = scene_.vertexCount();
    float smin[3] = {0, 0, 0}, smax[3] = {0, 0, 0};
    if (V) {
        for (int a = 0; a < 3; a++) { smin[a] = INFINITY; smax[a] = -INFINITY; }
        for (size_t v = 0; v < V; v++)
            for (int a = 0; a < 3; a++) {
                const float p = scene_.positions[3 * v + a];
                smin[a] = fminf(smin[a], p);
                smax[a] = fmaxf(smax[a], p);
            }
    }
    dPositions_ = dallocT<float>(3 * V);
    dIndices_ = dallocT<uint32_t>(3 * T);
    if (V) CR_CUDA(cudaMemcpyAsync(dPositions_, scene_.positions.data(), sizeof(float) * 3 * V, cudaMemcpyHostToDevice, stream_));
    if (T) CR_CUDA(cudaMemcpyAsync(dIndices_, scene_.indices.data(), sizeof(uint32_t) * 3 * T, cudaMemcpyHostToDevice, stream_));
    int leafSize = 2;   // measured best with the entry frontier (1: 16.2, 2: 17.9, 4: 17.2, 8: 16.0 Grays/s)
    if (const char* env = getenv("CR_LEAF_SIZE")) leafSize = atoi(env);
    bvh_ = buildLbvh(dPositions_, dIndices_, static_cast<int>(T), smin, smax, leafSize, stream_);
    dfree(dPositions_);
    dfree(dIndices_);

    std::vector<uint4> prims(T);
    for (size_t t = 0; t < T; t++)
        prims[t] = make_uint4(scene_.indices[3 * t], scene_.indices[3 * t + 1], scene_.indices[3 * t + 2], scene_.triMesh[t]);
    dPrims_ = dallocT<uint4>(T);
    if (T) CR_CUDA(cudaMemcpyAsync(dPrims_, prims.data(), sizeof(uint4) * T, cudaMemcpyHostToDevice, stream_));
    if (scene_.anyUV) {
        dUvs_ = dallocT<float2>(V);
        CR_CUDA(cudaMemcpyAsync(dUvs_, scene_.uvs.data(), sizeof(float) * 2 * V, cudaMemcpyHostToDevice, stream_));
    }
    if (scene_.anyColor) {
        dColors_ = dallocT<float4>(V);
        CR_CUDA(cudaMemcpyAsync(dColors_, scene_.colors.data(), sizeof(float) * 4 * V, cudaMemcpyHostToDevice, stream_));
    }
    // textures: wrap + bilinear + normalised float reads, the only sampler the reference can create
    // (MulticamScene.cpp:801-834)
    for (const ImageRGBA8& img : scene_.textures) {
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
        cudaArray_t arr = nullptr;
        CR_CUDA(cudaMallocArray(&arr, &desc, static_cast<size_t>(img.width), static_cast<size_t>(img.height)));
        const size_t pitch = static_cast<size_t>(img.width) * 4;
        CR_CUDA(cudaMemcpy2DToArray(arr, 0, 0, img.pixels.data(), pitch, pitch, static_cast<size_t>(img.height), cudaMemcpyHostToDevice));
        cudaResourceDesc res = {};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = arr;
        cudaTextureDesc td = {};
        td.addressMode[0] = cudaAddressModeWrap;
        td.addressMode[1] = cudaAddressModeWrap;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 1;
        td.maxAnisotropy = 1;
        td.sRGB = 0;
        cudaTextureObject_t tex = 0;
        CR_CUDA(cudaCreateTextureObject(&tex, &res, &td, nullptr));
        texArrays_.push_back(arr);
        texObjects_.push_back(tex);
    }
    std::vector<MeshRec> recs(scene_.meshes.size());
    for (size_t i = 0; i < recs.size(); i++) {
        const MeshGroup& m = scene_.meshes[i];
        MeshRec& r = recs[i];
        r.colorType = m.colorType;
        r.hasUV = m.hasUV;
        r.hasTex = (m.texture >= 0 && static_cast<size_t>(m.texture) < texObjects_.size()) ? 1 : 0;
        r.pad = 0;
        r.tex = r.hasTex ? static_cast<unsigned long long>(texObjects_[static_cast<size_t>(m.texture)]) : 0ull;
        memcpy(r.baseColor, m.baseColor, sizeof r.baseColor);
    }
    dMeshes_ = dallocT<MeshRec>(recs.size());
    if (!recs.empty()) CR_CUDA(cudaMemcpyAsync(dMeshes_, recs.data(), sizeof(MeshRec) * recs.size(), cudaMemcpyHostToDevice, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));

    dscene_.nodes = bvh_.nodes;
    dscene_.tris = bvh_.tris;
    dscene_.prims = dPrims_;
    dscene_.uvs = dUvs_;
    dscene_.colors = dColors_;
    dscene_.meshes = dMeshes_;
    dscene_.nNodes = bvh_.nNodes;
    dscene_.nodeVariantStride = static_cast<size_t>(4) * static_cast<size_t>(bvh_.nNodes);
    dscene_.nTris = bvh_.nTris;
    dscene_.missShader = scene_.missShader;
    float absMax = 0.0f;
    for (int a = 0; a < 3; a++) absMax = fmaxf(absMax, fmaxf(fabsf(smin[a]), fabsf(smax[a])));
    dscene_.boundsAbsMax = absMax * 1.001f;            // (leaf boxes are padded by 2^-20 relative: well inside)
    if (verbose)
        std::cout << "[PyEye] scene on device: " << T << " triangles, " << bvh_.nNodes << " BVH nodes, built in "
                  << bvh_.buildMs << " ms" << std::endl;
}

// ------------------------------------------------------------------------------------------
HostCamera& Renderer::camera()
{
    if (scene_.cameras.empty()) {                                     // MulticamScene.cpp:911-927
        std::cerr << "Initializing default camera" << std::endl;
        HostCamera c;
        c.name = "Default Camera";
        c.kind = CAM_PERSPECTIVE;
        scene_.cameras.push_back(c);
    }
    if (current_ >= scene_.cameras.size()) current_ = 0;
    return scene_.cameras[current_];
}
size_t Renderer::cameraCount() { return scene_.cameras.size(); }
void Renderer::setCurrentCamera(int index)
{
    const int s = static_cast<int>(cameraCount());
    if (s == 0) { current_ = 0; return; }
    current_ = static_cast<size_t>((index % s + s) % s);            // MulticamScene.cpp:929-934
}
bool Renderer::compoundActive() { return camera().kind == CAM_COMPOUND; }

CompoundState& Renderer::compoundState(size_t camIdx)
{
    auto it = compound_.find(camIdx);
    if (it == compound_.end()) {
        CompoundState cs;
        cs.N = static_cast<int>(scene_.cameras[camIdx].ommatidia.size());
        it = compound_.emplace(camIdx, cs).first;
    }
    return it->second;
}

void Renderer::setSamples(int s)
{
    if (!compoundActive()) return;
    CompoundState& cs = compoundState(current_);
    cs.S = std::max(1, s);                                            // CompoundEye.cpp:171-183
    cs.randomsConfigured = false;
}
int Renderer::samples()
{
    if (!compoundActive()) return -1;
    return compoundState(current_).S;
}
size_t Renderer::ommatidialCount()
{
    if (!compoundActive()) return 0;
    return camera().ommatidia.size();
}
void Renderer::setOmmatidia(const Ommatidium* omm, size_t count)
{
    if (!compoundActive()) return;
    if (!omm && count) throw std::runtime_error("setOmmatidia: null table with a non-zero count");
    HostCamera& cam = camera();
    CompoundState& cs = compoundState(current_);
    if (count != cam.ommatidia.size()) cs.randomsConfigured = false;  // CompoundEye.cpp:35-48
    cam.ommatidia.assign(omm, omm + count);
    cs.N = static_cast<int>(count);
    cs.ommDirty = true;
    cs.eyeVersion++;
}
void Renderer::setOmmatidialShard(uint64_t globalCount, uint64_t first)
{
    if (!compoundActive()) return;
    CompoundState& cs = compoundState(current_);
    dropReadAhead(cs);
    cs.shardGlobalN = globalCount;
    cs.shardFirst = globalCount ? first : 0;
    cs.randomsConfigured = false;
}
void Renderer::setFirstFrame(uint64_t k)
{
    if (!compoundActive()) return;
    CompoundState& cs = compoundState(current_);
    dropReadAhead(cs);
    cs.firstFrame = k;
    cs.randomsConfigured = false;
}

void Renderer::prepareCompound(CompoundState& cs, HostCamera& cam)
{
    const int N = static_cast<int>(cam.ommatidia.size());
    cs.N = N;
    if (static_cast<long long>(N) * cs.S >= (1ll << 31)) throw std::runtime_error("N*S must stay below 2^31 rays per frame");
    if (cs.shardGlobalN > 0) {
        // stream id = globalN*s + first + o must stay a valid subsequence index (the reference's id is an int,
        // shaders.cu:668-669) and the shard must lie inside the eye, or ids collide across samples
        if (cs.shardFirst + static_cast<uint64_t>(N) > cs.shardGlobalN)
            throw std::runtime_error("crSetOmmatidialShard: firstIndex + ommatidia exceeds globalCount");
        if (cs.shardGlobalN * static_cast<uint64_t>(cs.S) >= (1ull << 31))
            throw std::runtime_error("crSetOmmatidialShard: globalCount * samples must stay below 2^31 sample streams");
    }
    if (cs.ommDirty || !cs.dOmm) {
        dfree(cs.dOmm); dfree(cs.dPre);
        cs.dOmm = dallocT<float4>(2 * static_cast<size_t>(N));
        cs.dPre = dallocT<float4>(kPreStride * static_cast<size_t>(N));
        CR_CUDA(cudaMemcpyAsync(cs.dOmm, cam.ommatidia.data(), sizeof(Ommatidium) * static_cast<size_t>(N), cudaMemcpyHostToDevice, stream_));
        launchPrepOmmatidia(cs.dOmm, N, cs.dPre, stream_);
        launches_++;
        cs.ommDirty = false;
    }
    if (cs.rngN != N || cs.rngS != cs.S || !cs.dRng) {
        dfree(cs.dRng); dfree(cs.dSummed); dfree(cs.dSamples); dfree(cs.dRngAlt);
        cs.rngAltCap = 0;
        cs.dLastSummed = nullptr;
        cs.dRng = dallocT<uint4>(2 * static_cast<size_t>(N) * static_cast<size_t>(cs.S));
        cs.dSamples = dallocT<float>(3 * static_cast<size_t>(N) * static_cast<size_t>(cs.S));
        cs.dSummed = dallocT<float4>(static_cast<size_t>(N));
        CR_CUDA(cudaMemsetAsync(cs.dSummed, 0, sizeof(float4) * static_cast<size_t>(N ? N : 1), stream_));
        cs.rngN = N;
        cs.rngS = cs.S;
        cs.randomsConfigured = false;
    }
    if (cs.needRewind) {                                 // frames rendered ahead were dropped: put the streams back where the caller is
        if (cs.randomsConfigured) {                      // (a pending reset -- new S, new count, crSetFirstFrame -- supersedes the rewind)
            cs.firstFrame = cs.rewindTo;
            cs.randomsConfigured = false;
        }
        cs.needRewind = false;
    }
    if (!cs.randomsConfigured) {
        const bool sharded = cs.shardGlobalN > 0;        // stream ids of an ommatidium-range shard use global indices
        if (!dJumpTable_) {                              // XORWOW jump-ahead matrices: derived once per process (~20 ms), 1.8 MB
            const std::vector<uint32_t>& t = xorwowTable();
            dJumpTable_ = dallocT<uint4>(t.size() / 4);
            CR_CUDA(cudaMemcpyAsync(dJumpTable_, t.data(), sizeof(uint32_t) * t.size(), cudaMemcpyHostToDevice, stream_));
        }
        launchRngInit(cs.dRng, N, cs.S, cs.firstFrame, sharded ? cs.shardGlobalN : static_cast<unsigned long long>(N),
                      sharded ? cs.shardFirst : 0ull, dJumpTable_, stream_);
        launches_++;
        cs.frameIndex = cs.firstFrame;
        cs.firstFrame = 0;                               // consumed: a later re-initialisation (new S, new count) restarts at frame 0,
        cs.randomsConfigured = true;                     // as the reference's curand_init(42, id, 0) always does
    }
}

// Entry frontier of this launch's (frame, ommatidium) cones; leaves ep.entries null when switched off
// or when there are too few samples per ommatidium to amortise the pass.
bool Renderer::entryFrontierActive(const CompoundState& cs, int frames) const
{
    return entryFrontier && cs.N > 0 && cs.S >= entryMinSamples && static_cast<long long>(cs.N) * cs.S * frames >= entryMinRays;
}

void Renderer::buildEntries(CompoundState& cs, EyeParams& ep)
{
    cs.listsLast = 0;
    // Frames too small to amortise the pass get the frontier all the same once the camera has stood still for three
    // frames (the reference's speed-test and variance protocols: hundreds of frames from one pose): built once, then reused.
    // (cs.standingFrames: noteSinglePose, once per renderFrame)
    if (!entryFrontierActive(cs, ep.poses ? ep.nFrames : 1) &&
        !(entryFrontier && standingFrontier && ep.poses == nullptr && cs.N > 0 && cs.standingFrames >= 2)) return;
    const size_t need = static_cast<size_t>(cs.N) * static_cast<size_t>(ep.poses ? ep.nFrames : 1);
    if (cs.entryCap < need) {
        dfree(cs.dEntries);
        cs.dEntries = dallocT<int4>(need);
        cs.entryCap = need;
        cs.entriesValid = false;
    }
    // candidate lists only where K1 can use them: whole warps per ommatidium (S % 32 == 0)
    // ... and where the second stage pays: it lengthens the frontier pass's latency chain, which a batch hides behind the
    // previous frames' tracing but a single synchronous frame does not (measured on the headline workload: +1.7 % per
    // batched frame, -3 % through renderFrame).  candidateLists: 0 never, 1 batches of >= 4 frames, 2 always.
    const int framesNow = ep.poses ? ep.nFrames : 1;
    const bool wantLists = cs.S % 32 == 0 && (candidateLists >= 2 || (candidateLists == 1 && framesNow >= 4));
    if (wantLists && cs.listCap < need) {
        dfree(cs.dLists);
        cs.dLists = dallocT<int>(need * static_cast<size_t>(candidateListStride()));
        cs.listCap = need;
        cs.entriesValid = false;
    }
    // The frontier depends on (scene, eye, pose) only -- not on the frame's random draws: a camera that has not moved
    // since the previous single-frame launch (the reference's speed-test and variance protocols render hundreds of
    // frames from one pose) keeps its entries and lists.
    const bool single = ep.poses == nullptr;
    const bool reuse = single && cs.entriesValid && cs.entriesEyeVersion == cs.eyeVersion && cs.entriesLists == wantLists &&
                       memcmp(&cs.entriesPose, &ep.pose, sizeof(DevicePose)) == 0;
    if (!reuse) {
        launchBuildEntries(dscene_, ep, cs.dEntries, wantLists ? cs.dLists : nullptr, stream_);
        launches_++;
    }
    cs.entriesValid = single;
    cs.entriesEyeVersion = cs.eyeVersion;
    cs.entriesLists = wantLists;
    cs.entriesPose = ep.pose;
    ep.entries = cs.dEntries;
    ep.entryFrameStride = static_cast<unsigned>(cs.N);
    ep.lists = wantLists ? cs.dLists : nullptr;
    cs.listsLast = wantLists ? need : 0;
}

// Once per renderFrame of a compound eye: how many consecutive frames this camera has been rendered from one pose.
void Renderer::noteSinglePose(CompoundState& cs, const DevicePose& pose)
{
    const bool standing = cs.lastPoseValid && cs.lastPoseEyeVersion == cs.eyeVersion && memcmp(&cs.lastPose, &pose, sizeof(DevicePose)) == 0;
    cs.standingFrames = standing ? cs.standingFrames + 1 : 0;
    cs.lastPoseValid = true;
    cs.lastPose = pose;
    cs.lastPoseEyeVersion = cs.eyeVersion;
}

// The in-kernel reduction needs whole warps per ommatidium (S % 32 == 0) and is bypassed where somebody reads the
// per-sample buffer: the raw_ommatidial_samples projection and the per-ray dump.
bool Renderer::fusedActive(const CompoundState& cs, const HostCamera& cam) const
{
    return fusedReduce && cs.S % 32 == 0 && !dumpRays && projectionFromName(cam.projection) != PROJ_RAW_SAMPLES;
}

void Renderer::ensurePartials(CompoundState& cs, size_t frames)
{
    const size_t need = frames * static_cast<size_t>(cs.N) * static_cast<size_t>(cs.S / 32);
    if (cs.partialCap >= need && cs.dPartials) return;
    dfree(cs.dPartials);
    cs.dPartials = dallocT<float4>(need);
    cs.partialCap = need;
}

// Wavefront queue: sized for `queueFraction` of the rays of a launch of `frames` frames (48 B per ray: origin, direction,
// id and the hit record).  Warps that find it full walk their rays inline, so the size is a performance knob only.
size_t Renderer::queueRaysFor(const CompoundState& cs, size_t frames) const
{
    const double rays = static_cast<double>(frames) * static_cast<double>(cs.N) * static_cast<double>(cs.S);
    return static_cast<size_t>(std::min(rays * std::max(0.0, std::min(1.0, queueFraction)) + 32.0, 2.0e9)) & ~size_t(31);
}

void Renderer::ensureQueue(CompoundState& cs, size_t frames)
{
    const size_t need = queueRaysFor(cs, frames);
    if (!cs.dQueueCounters) {
        cs.dQueueCounters = dallocT<unsigned>(4);
        CR_CUDA(cudaMemsetAsync(cs.dQueueCounters, 0, sizeof(unsigned) * 4, stream_));
    }
    if (cs.queueCap >= need && cs.dQueueRays) return;
    dfree(cs.dQueueRays); dfree(cs.dQueueHits); dfree(cs.dQueueWarps);
    cs.dQueueRays = dallocT<float4>(2 * need);
    cs.dQueueHits = dallocT<int4>(need);
    cs.dQueueWarps = dallocT<int>(need / 32 + 1);
    cs.queueCap = need;
}

// Work hand-out of the trace kernel.  counters: [0] rays pushed to the queue, [1] rays handed out by k_traceQueue (zeroed per
// launch that has a queue); [2] work chunks handed out, [3] warps that left the trace kernel (zeroed once: the kernel rearms
// them itself).  SM-affine hand-out on top of the counter when the launch has enough blocks of 32 units to keep every SM
// busy: a ticket counter per slot (rearmed by the kernel) and the slot's block table, whose entries carry the launch's epoch.
void Renderer::attachWorkCounter(CompoundState& cs, EyeParams& ep)
{
    if (!cs.dQueueCounters) {
        cs.dQueueCounters = dallocT<unsigned>(4);
        CR_CUDA(cudaMemsetAsync(cs.dQueueCounters, 0, sizeof(unsigned) * 4, stream_));
    }
    if (!dynamicChunks) return;
    ep.workCounter = cs.dQueueCounters + 2;
    if (smAffine <= 0 || dumpRays) return;
    const unsigned long long rayUnits = (static_cast<unsigned long long>(cs.N) * static_cast<unsigned long long>(cs.S) + 31ull) / 32ull;
    const unsigned long long units = rayUnits * static_cast<unsigned long long>(std::max(1, ep.poses ? ep.frameGroups : 1));
    const unsigned long long blocks = (units + 31ull) / 32ull;
    if (smIdLimit_ == 0u) {
        smIdLimit_ = std::max(static_cast<unsigned>(numSMs_), deviceSmIdLimit(stream_));
        if (getenv("CR_SM_AFFINE_VERBOSE")) std::cerr << "[cr] SM ids below " << smIdLimit_ << " on " << numSMs_ << " SMs" << std::endl;
    }
    const unsigned slots = smIdLimit_;
    const unsigned long long gridWarps = static_cast<unsigned long long>(numSMs_) * traceOcc_ * (kTraceThreads / 32);
    const unsigned long long cap = blocks + gridWarps / 16ull + 2ull;
    // enough blocks for the counter to balance the SMs (a block is what a unit is to a warp), and a table of sane size
    if (blocks < static_cast<unsigned long long>(std::max(0, smAffineMinBlocks)) * static_cast<unsigned>(numSMs_) || cap * slots * 8ull > (128ull << 20)) return;
    if (cs.smSlots != slots || cs.smCap < cap) {
        dfree(cs.dSmSeq); dfree(cs.dSmTab);
        cs.dSmSeq = dallocT<unsigned>(slots);
        cs.dSmTab = dallocT<unsigned long long>(static_cast<size_t>(cap) * slots);
        CR_CUDA(cudaMemsetAsync(cs.dSmSeq, 0, sizeof(unsigned) * slots, stream_));
        CR_CUDA(cudaMemsetAsync(cs.dSmTab, 0, sizeof(unsigned long long) * static_cast<size_t>(cap) * slots, stream_));
        cs.smSlots = slots; cs.smCap = static_cast<unsigned>(cap); cs.smEpoch = 0;
    }
    if (++cs.smEpoch == 0u) {              // epoch wrapped: entries of 2^32 launches ago could look current
        CR_CUDA(cudaMemsetAsync(cs.dSmTab, 0, sizeof(unsigned long long) * static_cast<size_t>(cs.smCap) * cs.smSlots, stream_));
        cs.smEpoch = 1u;
    }
    ep.smSeq = cs.dSmSeq; ep.smTab = cs.dSmTab;
    ep.smSlots = cs.smSlots; ep.smCap = cs.smCap; ep.smEpoch = cs.smEpoch;
}

// The queue takes the warp-frames that have no candidate list, so it exists only in launches that build the lists.
void Renderer::attachQueue(CompoundState& cs, EyeParams& ep)
{
    attachWorkCounter(cs, ep);
    ep.chunkUnits = std::max(1, chunkUnits);
    if (!wavefront || ep.lists == nullptr || dumpRays) return;
    if (static_cast<double>(ep.poses ? ep.nFrames : 1) * cs.N >= 2147483648.0) return;   // frame*N + ommatidium would not fit 31 bits
    ensureQueue(cs, ep.poses ? static_cast<size_t>(ep.nFrames) : 1);
    CR_CUDA(cudaMemsetAsync(cs.dQueueCounters, 0, sizeof(unsigned) * 2, stream_));
    ep.queueRays = cs.dQueueRays;
    ep.queueHits = cs.dQueueHits;
    ep.queueWarps = cs.dQueueWarps;
    ep.queueCounters = cs.dQueueCounters;
    ep.queueCap = static_cast<unsigned>(std::min(cs.queueCap, queueRaysFor(cs, ep.poses ? static_cast<size_t>(ep.nFrames) : 1)));
    lastQueueCap_ = ep.queueCap;
    ep.queueRefillBelow = std::max(1, std::min(32, wavefrontRefill));
    launches_ += 2;   // k_traceQueue + k_shadeQueue
}

// Device-side breakdown of the last renderFrame of a compound eye (crDebugSetFrameProfile(1) first): milliseconds between the
// event marks [frontier pass + counter reset, trace kernel(s), reduction kernel].
void Renderer::debugFrameBreakdown(float* out3)
{
    out3[0] = out3[1] = out3[2] = -1.0f;
    if (!profileFrame || !evMark_[3]) return;
    CR_CUDA(cudaStreamSynchronize(stream_));
    for (int i = 0; i < 3; i++) cudaEventElapsedTime(out3 + i, evMark_[i], evMark_[i + 1]);
}

unsigned long long Renderer::debugLastQueuedRays()
{
    if (!compoundActive()) return 0;
    CompoundState& cs = compoundState(current_);
    if (!cs.dQueueCounters) return 0;
    unsigned c[4] = {0, 0, 0, 0};
    CR_CUDA(cudaMemcpyAsync(c, cs.dQueueCounters, sizeof(c), cudaMemcpyDeviceToHost, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));
    return std::min<unsigned long long>(c[0], lastQueueCap_);
}

void Renderer::launchCompound(CompoundState& cs, const HostCamera& cam, const Pose& pose, uchar4* fastRow, int fastRowCount,
                              uchar4* fastRowHost)
{
    EyeParams ep;
    ep.fastRowHost = fastRowHost;
    ep.fast = fastMath;
    ep.pdl = pdl && !profileFrame;
    ep.nodeLanes = std::max(1, std::min(32, nodeLanes));
    ep.entryMaxLevels = std::max(1, entryMaxLevels);
    if (fusedActive(cs, cam)) {
        ensurePartials(cs, 1);
        ep.fused = true;
        ep.partials = cs.dPartials;
    }
    ep.fastRow = fastRow;
    ep.fastRowCount = fastRowCount;
    ep.pre = cs.dPre;
    ep.rng = cs.dRng;
    ep.summed = cs.dSummed;
    ep.samples = cs.dSamples;
    ep.N = cs.N;
    ep.S = cs.S;
    ep.pose = toDevicePose(pose);
    if (dumpRays) {
        const size_t need = static_cast<size_t>(cs.N) * static_cast<size_t>(cs.S);
        if (cs.dumpCap < need) {
            dfree(cs.dDumpO); dfree(cs.dDumpD); dfree(cs.dDumpH); dfree(cs.dDumpC);
            cs.dDumpO = dallocT<float>(3 * need);
            cs.dDumpD = dallocT<float>(3 * need);
            cs.dDumpH = dallocT<int4>(need);
            cs.dDumpC = dallocT<int2>(need);
            cs.dumpCap = need;
        }
        ep.dumpOrigins = cs.dDumpO; ep.dumpDirs = cs.dDumpD; ep.dumpHits = cs.dDumpH; ep.dumpCounts = cs.dDumpC;
    }
    if (profileFrame) {
        for (auto& e : evMark_) if (!e) CR_CUDA(cudaEventCreate(&e));
        CR_CUDA(cudaEventRecord(evMark_[0], stream_));
    }
    buildEntries(cs, ep);
    attachQueue(cs, ep);
    if (profileFrame) CR_CUDA(cudaEventRecord(evMark_[1], stream_));
    const long long slots = static_cast<long long>(numSMs_) * traceOcc_;   // persistent grid: every SM full
    launchTraceCompound(dscene_, ep, static_cast<int>(slots), stream_, profileFrame ? evMark_[2] : nullptr);
    if (profileFrame) CR_CUDA(cudaEventRecord(evMark_[3], stream_));
    launches_ += 2;   // trace + ordered sum
    cs.frameIndex++;
    cs.dLastSummed = cs.dSummed;
}

void Renderer::launchCompoundBatch(CompoundState& cs, const DevicePose* dPoses, int nFrames, float* dSamples, float4* dSummed,
                                   uchar4* fastRow, const Pose* samePose)
{
    EyeParams ep;
    ep.fast = fastMath;
    ep.pdl = pdl && !profileFrame;
    ep.nodeLanes = std::max(1, std::min(32, nodeLanes));
    ep.entryMaxLevels = std::max(1, entryMaxLevels);
    if (dSamples == nullptr) {             // fused reduction: the caller sized cs.dPartials for nFrames
        ep.fused = true;
        ep.partials = cs.dPartials;
    }
    ep.fastRow = fastRow;
    ep.fastRowCount = fastRow ? cs.N * nFrames : 0;
    ep.pre = cs.dPre;
    ep.rng = cs.dRng;
    ep.rngOut = cs.dRng;
    ep.summed = dSummed;
    ep.samples = dSamples;
    ep.N = cs.N;
    ep.S = cs.S;
    ep.nFrames = nFrames;
    // Small frames: too few units of 32 rays for the work counter to balance (the 6 374 x 64 eye of the toy experiment: 2.7 per
    // warp, sm__warps_active 40 of 50 %; a 1000 x 1 eye: 32 units for 4 736 warps).  Cut the launch's frames into groups
    // (EyeParams::frameGroups) -- possible when the launch starts at an even frame, where no stream holds a cached normal.
    bool grouped = false;
    if (frameGroups && nFrames >= 4 && (cs.frameIndex & 1ull) == 0 && !dumpRays) {
        const long long rayUnits = (static_cast<long long>(cs.N) * cs.S + 31) / 32;
        const long long wantUnits = 8ll * numSMs_ * traceOcc_ * (kTraceThreads / 32);
        long long G = std::min<long long>((wantUnits + rayUnits - 1) / std::max<long long>(1, rayUnits), nFrames / 2);
        if (G > 1) {
            const int Fg = 2 * static_cast<int>((nFrames + 2 * G - 1) / (2 * G));
            G = (nFrames + Fg - 1) / Fg;
            if (G > 1) {
                const size_t words = 2 * static_cast<size_t>(cs.N) * static_cast<size_t>(cs.S);
                if (cs.rngAltCap < words) {
                    dfree(cs.dRngAlt);
                    cs.dRngAlt = dallocT<uint4>(words);
                    cs.rngAltCap = words;
                }
                ep.frameGroups = static_cast<int>(G);
                ep.groupFrames = Fg;
                ep.rngOut = cs.dRngAlt;
                grouped = true;
            }
        }
    }
    if (samePose) {                         // every frame of the launch from one pose (read-ahead): ONE frontier row serves them all
        ep.pose = toDevicePose(*samePose);
        ep.nFrames = 1;
        buildEntries(cs, ep);               // single-frame rules: reuses the standing camera's entries
        ep.nFrames = nFrames;
        ep.entryFrameStride = 0;
        ep.poses = dPoses;
        attachWorkCounter(cs, ep);
        ep.chunkUnits = 1;
    } else {
        ep.poses = dPoses;
        buildEntries(cs, ep);
        attachQueue(cs, ep);
    }
    const long long slots = static_cast<long long>(numSMs_) * traceOcc_;
    launchTraceCompound(dscene_, ep, static_cast<int>(slots), stream_);
    launches_ += 2;
    if (grouped) std::swap(cs.dRng, cs.dRngAlt);         // the last group wrote the states there
    cs.frameIndex += static_cast<uint64_t>(nFrames);
    cs.dLastSummed = dSummed + static_cast<size_t>(nFrames - 1) * static_cast<size_t>(cs.N);
}

void Renderer::project(CompoundState& cs, const HostCamera& cam)
{
    const int mode = projectionFromName(cam.projection);
    switch (mode) {
        case PROJ_RAW_SAMPLES:
            launchProjectRaw(fastMath, cs.dSamples, cs.N, cs.S, dFrame_, W_, H_, stream_);
            launches_++;
            break;
        case PROJ_SINGLE_DIM:
        case PROJ_SINGLE_DIM_FAST:
            launchProjectVector(mode, fastMath, cs.dSummed, cs.N, dFrame_, W_, H_, stream_);
            launches_++;
            break;
        case PROJ_UNKNOWN:
            std::cerr << "[PyEye] ERROR: unknown compound projection shader '__raygen__compound_projection_" << cam.projection
                      << "'; frame left untouched." << std::endl;
            break;
        default: {
            if (!cs.dMap || cs.mapMode != mode || cs.mapW != W_ || cs.mapH != H_ || cs.mapEyeVersion != cs.eyeVersion) {
                dfree(cs.dMap);
                cs.dMap = dallocT<uint32_t>(static_cast<size_t>(W_) * static_cast<size_t>(H_));
                launchBuildProjectionMap(mode, cs.dOmm, cs.N, cs.dMap, W_, H_, stream_);
                launches_++;
                cs.mapMode = mode; cs.mapW = W_; cs.mapH = H_; cs.mapEyeVersion = cs.eyeVersion;
            }
            const bool ids = (mode == PROJ_SPH_ORIENTATIONWISE_IDS || mode == PROJ_SPH_POSITIONWISE_IDS);
            launchProjectMap(ids, fastMath, cs.dMap, cs.dSummed, dFrame_, W_, H_, stream_);
            launches_++;
            break;
        }
    }
}

void Renderer::ensureFrame()
{
    const size_t need = static_cast<size_t>(W_) * static_cast<size_t>(H_);
    if (dFrame_ && frameW_ == W_ && frameH_ == H_) return;
    dfree(dFrame_);
    if (hFrame_) cudaFreeHost(hFrame_);
    hFrame_ = nullptr;
    dFrame_ = dallocT<uchar4>(need);
    CR_CUDA(cudaMemsetAsync(dFrame_, 0, sizeof(uchar4) * (need ? need : 1), stream_));
    CR_CUDA(cudaHostAlloc(&hFrame_, sizeof(uchar4) * (need ? need : 1), cudaHostAllocMapped));
    memset(hFrame_, 0, sizeof(uchar4) * (need ? need : 1));
    hFrameDev_ = nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&hFrameDev_), hFrame_, 0) != cudaSuccess) { hFrameDev_ = nullptr; cudaGetLastError(); }
    hostMirrorsDevice_ = true;                        // both all zero
    frameCap_ = need;
    frameW_ = W_;
    frameH_ = H_;
    hostFrameFresh_ = false;
}

void Renderer::setRenderSize(int w, int h)
{
    W_ = std::max(0, w);
    H_ = std::max(0, h);
    if (verbose) std::cout << "[PyEye] Resizing rendering buffer to (" << w << ", " << h << ")." << std::endl;
}

double Renderer::renderFrame()
{
    if (!loaded_) throw std::runtime_error("renderFrame called before loadGlTFscene");
    ensureDevice();
    if (!dscene_.nodes) uploadScene();
    ensureFrame();
    HostCamera& cam = camera();
    const auto t0 = std::chrono::steady_clock::now();
    bool timedTrace = false, eager = false, zeroCopy = false;
    CompoundState* singleCompound = nullptr;
    if (cam.kind == CAM_COMPOUND) {
        CompoundState& cs = compoundState(current_);
        // standing camera: the next frame may already be there, or this call renders several at once
        const bool aheadOk = readAheadEligible(cs, cam);
        if (consumeReadAhead(cs, cam, aheadOk)) {
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (verbose) std::cout << "[PyEye] Rendered frame in " << ms << "ms." << std::endl;
            return ms;
        }
        prepareCompound(cs, cam);
        noteSinglePose(cs, toDevicePose(cam.pose));
        if (readAheadEligible(cs, cam) && launchReadAhead(cs, cam)) {
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (verbose) std::cout << "[PyEye] Rendered frame in " << ms << "ms." << std::endl;
            return ms;
        }
        singleCompound = &cs;
        if (wantTraceEvents_) CR_CUDA(cudaEventRecord(evA_, stream_));
        // single_dimension_fast: pixel x of row 0 is ommatidium x -- K1b writes the row itself
        const bool fused = projectionFromName(cam.projection) == PROJ_SINGLE_DIM_FAST && W_ > 0 && H_ > 0;
        // ... and when the caller reads every frame (eager copy below) and the pinned host frame already equals the device
        // frame everywhere else, K1b writes the row into the host frame as well: no copy is queued behind it
        eager = frameWasFetched_ && sizeof(uchar4) * frameCap_ <= kEagerFrameBytes && frameCap_ > 0;
        zeroCopy = fused && eager && hostMirrorsDevice_ && hFrameDev_ != nullptr && zeroCopyFrames && fusedActive(cs, cam);   // (k_sumPartials stores 128-byte rows)
        launchCompound(cs, cam, cam.pose, fused ? dFrame_ : nullptr, fused ? std::min(cs.N, W_) : 0,
                       zeroCopy ? reinterpret_cast<uchar4*>(hFrameDev_) : nullptr);
        if (wantTraceEvents_) CR_CUDA(cudaEventRecord(evB_, stream_));
        timedTrace = wantTraceEvents_;
        if (!fused) project(cs, cam);
    } else {
        eager = frameWasFetched_ && sizeof(uchar4) * frameCap_ <= kEagerFrameBytes && frameCap_ > 0;
        launchCamera(dscene_, static_cast<int>(cam.kind), fastMath, toDevicePose(cam.pose), cam.scale[0], cam.scale[1], cam.scale[2], dFrame_,
                     W_, H_, stream_);
        launches_++;
    }
    // Small frames (eye vectors, thumbnails) ride back on the same stream, so the usual
    // renderFrame -> getFramePointer pair costs one synchronisation instead of two.  Callers that
    // did not read the previous frame (render-only timing loops) are not charged for the copy.
    hostFrameFresh_ = eager;
    frameWasFetched_ = false;
    if (hostFrameFresh_ && !zeroCopy) CR_CUDA(cudaMemcpyAsync(hFrame_, dFrame_, sizeof(uchar4) * frameCap_, cudaMemcpyDeviceToHost, stream_));
    hostMirrorsDevice_ = hostFrameFresh_;             // after a copy or a dual write the two frames are equal again
    CR_CUDA(cudaStreamSynchronize(stream_));
    CR_CUDA(cudaGetLastError());
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (singleCompound) singleCompound->lastSingleFrameMs = ms;
    if (timedTrace) {
        float k = 0.0f;
        cudaEventElapsedTime(&k, evA_, evB_);
        lastTraceMs_ = k;
    } else {
        lastTraceMs_ = ms;                    // no event pair requested yet: the host-side frame time
    }
    if (verbose) std::cout << "[PyEye] Rendered frame in " << ms << "ms." << std::endl;
    return ms;
}

unsigned char* Renderer::framePointer()
{
    if (verbose) std::cout << "[PyEye] Retrieving frame pointer..." << std::endl;
    ensureDevice();
    ensureFrame();
    frameWasFetched_ = true;
    if (hostFrameFresh_) return hFrame_;              // already copied by renderFrame
    CR_CUDA(cudaMemcpyAsync(hFrame_, dFrame_, sizeof(uchar4) * frameCap_, cudaMemcpyDeviceToHost, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));
    hostFrameFresh_ = true;
    hostMirrorsDevice_ = true;
    return hFrame_;
}

void Renderer::saveFrame(const std::string& path)
{
    // binary P6, RGB, flipped to top-down (sutil/sutil.cpp:82-102, 387-412)
    const unsigned char* px = framePointer();
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    f << "P6\n" << W_ << " " << H_ << "\n255\n";
    std::vector<unsigned char> row(static_cast<size_t>(W_) * 3);
    for (int y = H_ - 1; y >= 0; y--) {
        for (int x = 0; x < W_; x++) {
            const unsigned char* p = px + 4 * (static_cast<size_t>(y) * W_ + x);
            row[3 * x] = p[0]; row[3 * x + 1] = p[1]; row[3 * x + 2] = p[2];
        }
        f.write(reinterpret_cast<const char*>(row.data()), static_cast<std::streamsize>(row.size()));
    }
    if (verbose) std::cout << "[PyEye] Saved render as '" << path << "'" << std::endl;
}

// ------------------------------------------------------------------------------------------
// additive API
// ------------------------------------------------------------------------------------------
void Renderer::copyOmmatidialData(float* outRgb)
{
    if (!compoundActive() || !outRgb) return;
    CompoundState& cs = compoundState(current_);
    const float4* src = cs.dLastSummed ? cs.dLastSummed : cs.dSummed;
    if (!src) return;
    std::vector<float4> tmp(static_cast<size_t>(cs.N));
    CR_CUDA(cudaMemcpyAsync(tmp.data(), src, sizeof(float4) * tmp.size(), cudaMemcpyDeviceToHost, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));
    for (size_t i = 0; i < tmp.size(); i++) { outRgb[3 * i] = tmp[i].x; outRgb[3 * i + 1] = tmp[i].y; outRgb[3 * i + 2] = tmp[i].z; }
}

// Frames per batched launch: bounded by a sample-buffer budget in the ordered mode (12 B per ray per frame held between K1 and
// K1b; the fused mode holds 0.5 B: one float4 per warp).
size_t Renderer::batchFramesPerLaunch(const CompoundState& cs, size_t count, bool fused) const
{
    const size_t raysPerFrame = static_cast<size_t>(cs.N) * static_cast<size_t>(cs.S);
    size_t budget = size_t(4) << 30;   // 8 -> 16 -> 32 -> 64 frames per launch: 18.8 -> 19.3 -> 19.6 -> 19.9 Grays/s on the headline workload
    if (const char* env = getenv("CR_BATCH_BYTES")) budget = static_cast<size_t>(atoll(env));
    size_t F = std::max<size_t>(1, std::min<size_t>(count, fused ? size_t(256) : budget / std::max<size_t>(1, raysPerFrame * 12)));
    if (const char* env = getenv("CR_BATCH_FRAMES")) F = std::max<size_t>(1, std::min<size_t>(count, static_cast<size_t>(atoll(env))));
    return F;
}

void Renderer::ensureBatchBuffers(CompoundState& cs, size_t F, bool fused)
{
    const size_t N = static_cast<size_t>(cs.N);
    const size_t raysPerFrame = N * static_cast<size_t>(cs.S);
    if (fused) {
        ensurePartials(cs, F);
    } else if (cs.batchSampleCap < F * raysPerFrame * 3) {
        dfree(cs.dBatchSamples);
        cs.dBatchSamples = dallocT<float>(F * raysPerFrame * 3);
        cs.batchSampleCap = F * raysPerFrame * 3;
    }
    if (cs.batchSummedCap < F * N) {
        dfree(cs.dBatchSummed);
        cs.dLastSummed = nullptr;
        cs.dBatchSummed = dallocT<float4>(F * N);
        CR_CUDA(cudaMemsetAsync(cs.dBatchSummed, 0, sizeof(float4) * F * N, stream_));
        cs.batchSummedCap = F * N;
    }
    if (entryFrontierActive(cs, static_cast<int>(F)) && cs.entryCap < F * N) {
        dfree(cs.dEntries);
        cs.dEntries = dallocT<int4>(F * N);
        cs.entryCap = F * N;
        cs.entriesValid = false;
    }
    if (entryFrontierActive(cs, static_cast<int>(F)) && cs.S % 32 == 0 && (candidateLists >= 2 || (candidateLists == 1 && F >= 4)) &&
        cs.listCap < F * N) {
        dfree(cs.dLists);
        cs.dLists = dallocT<int>(F * N * static_cast<size_t>(candidateListStride()));
        cs.listCap = F * N;
        cs.entriesValid = false;
    }
    if (wavefront && entryFrontierActive(cs, static_cast<int>(F)) && cs.S % 32 == 0 && (candidateLists >= 2 || (candidateLists == 1 && F >= 4)))
        ensureQueue(cs, F);
}

// ------------------------------------------------------------------------------------------
// Read-ahead for a standing camera (see cr_renderer.h).  The frames of a batched launch equal the frames of as many
// renderFrame calls bit for bit (same kernel body, same streams), so handing them out one per call changes nothing but
// when they are rendered.
// ------------------------------------------------------------------------------------------
bool Renderer::readAheadEligible(const CompoundState& cs, const HostCamera& cam) const
{
    return readAhead && !dumpRays && !profileFrame && cs.randomsConfigured && W_ > 0 && H_ > 0 &&
           projectionFromName(cam.projection) == PROJ_SINGLE_DIM_FAST && cs.N > 0 &&
           static_cast<long long>(cs.N) * cs.S <= readAheadMaxRays;
}

void Renderer::dropReadAhead(CompoundState& cs)
{
    if (cs.ahead.count > cs.ahead.next) {                // the streams ran ahead of the caller: rewind at the next prepareCompound
        cs.needRewind = true;
        cs.rewindTo = cs.frameIndex;
        // Frames were rendered for nothing and a rewind is due: this caller does not stand still for long.  Ask for a
        // longer standing streak before the next read-ahead (doubling, so a "k frames per pose" pattern stops paying after
        // a few poses) and start again with a short batch.
        cs.aheadStreakNeeded = std::min(4096, 2 * cs.aheadStreakNeeded + 2);
        cs.aheadFrames = 4;
    }
    cs.ahead.count = cs.ahead.next = 0;
}

// Hands out the next frame rendered ahead, if there is one and nothing it depends on has changed.
bool Renderer::consumeReadAhead(CompoundState& cs, const HostCamera& cam, bool eligible)
{
    CompoundState::ReadAhead& a = cs.ahead;
    if (a.count <= a.next) return false;
    const DevicePose pose = toDevicePose(cam.pose);
    const bool same = eligible && a.eyeVersion == cs.eyeVersion && a.S == cs.S && a.N == static_cast<int>(cam.ommatidia.size()) &&
                      a.fused == fusedActive(cs, cam) && a.fast == fastMath && a.rowPixels == std::min(cs.N, W_) &&
                      memcmp(&a.pose, &pose, sizeof(DevicePose)) == 0;
    if (!same) { dropReadAhead(cs); return false; }
    const size_t f = static_cast<size_t>(a.next++);
    if (a.next == a.count) {                             // consumed to the end: the next batch may be twice as long, and the streak rule relaxes
        cs.aheadFrames = std::min(64, 2 * cs.aheadFrames);
        cs.aheadStreakNeeded = std::max(2, cs.aheadStreakNeeded / 2);
    }
    memcpy(hFrame_, cs.hAheadRows + sizeof(uchar4) * f * static_cast<size_t>(cs.N), sizeof(uchar4) * static_cast<size_t>(a.rowPixels));
    hostFrameFresh_ = true;                              // the pinned host frame holds this frame; the device frame does not
    hostMirrorsDevice_ = false;
    frameWasFetched_ = false;
    cs.frameIndex++;
    cs.dLastSummed = cs.dBatchSummed + f * static_cast<size_t>(cs.N);
    lastTraceMs_ = a.traceMsPerFrame;
    noteSinglePose(cs, pose);
    return true;
}

// Renders the next frames of a standing camera in one batched launch and hands out the first.  False: not worth it / not
// possible now (the caller renders one frame the usual way).
bool Renderer::launchReadAhead(CompoundState& cs, const HostCamera& cam)
{
    if (cs.standingFrames < cs.aheadStreakNeeded || cs.lastSingleFrameMs <= 0.0) return false;
    const bool fused = fusedActive(cs, cam);
    // batches start short and double while they are consumed to the end (cs.aheadFrames), up to the GPU-time budget
    const size_t Fbudget = static_cast<size_t>(std::min(64.0, readAheadBudgetMs / std::max(cs.lastSingleFrameMs, 1e-3)));
    size_t F = std::min<size_t>(Fbudget, static_cast<size_t>(cs.aheadFrames));
    F = std::min(F, batchFramesPerLaunch(cs, F, fused));
    if (F < 2) return false;
    const size_t N = static_cast<size_t>(cs.N);
    // Buffers for the longest batch the ramp can reach, not for this one: growing them at every doubling (4, 8, 16, 32, 64
    // frames) put a round of cudaFree / cudaMalloc / cudaFreeHost / cudaMallocHost into frames 10, 26, 58, ... of every standing
    // run -- 1 to 50 ms each on a quiet host, 0.9 s seen on a busy one (benchmarks/speed_probe.py), against 13 us per frame.
    const size_t Fcap = std::max(F, std::min(Fbudget, batchFramesPerLaunch(cs, Fbudget, fused)));
    ensureBatchBuffers(cs, Fcap, fused);
    if (cs.aheadRowCap < Fcap * N) {
        dfree(cs.dAheadRows);
        if (cs.hAheadRows) cudaFreeHost(cs.hAheadRows);
        cs.dAheadRows = dallocT<uchar4>(Fcap * N);
        CR_CUDA(cudaMallocHost(&cs.hAheadRows, sizeof(uchar4) * Fcap * N));
        cs.aheadRowCap = Fcap * N;
    }
    if (cs.batchPoseCap < Fcap) {
        dfree(cs.dBatchPoses);
        cs.dBatchPoses = dallocT<DevicePose>(Fcap);
        cs.batchPoseCap = Fcap;
    }
    const DevicePose dp = toDevicePose(cam.pose);
    std::vector<DevicePose> hPoses(F, dp);
    CR_CUDA(cudaMemcpyAsync(cs.dBatchPoses, hPoses.data(), sizeof(DevicePose) * F, cudaMemcpyHostToDevice, stream_));
    const uint64_t frame0 = cs.frameIndex;
    CR_CUDA(cudaEventRecord(evA_, stream_));
    launchCompoundBatch(cs, cs.dBatchPoses, static_cast<int>(F), fused ? nullptr : cs.dBatchSamples, cs.dBatchSummed, cs.dAheadRows, &cam.pose);
    CR_CUDA(cudaEventRecord(evB_, stream_));
    CR_CUDA(cudaMemcpyAsync(cs.hAheadRows, cs.dAheadRows, sizeof(uchar4) * F * N, cudaMemcpyDeviceToHost, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));             // (hPoses must outlive the copy)
    CR_CUDA(cudaGetLastError());
    cs.frameIndex = frame0;                              // the caller has consumed none of them yet
    float k = 0.0f;
    cudaEventElapsedTime(&k, evA_, evB_);
    CompoundState::ReadAhead& a = cs.ahead;
    a.count = static_cast<int>(F); a.next = 0;
    a.rowPixels = std::min(cs.N, W_); a.S = cs.S; a.N = cs.N;
    a.pose = dp; a.eyeVersion = cs.eyeVersion; a.fused = fused; a.fast = fastMath;
    a.traceMsPerFrame = static_cast<double>(k) / static_cast<double>(F);
    return consumeReadAhead(cs, cam, true);
}

// Renders `count` consecutive frames of the current compound eye, one per pose (12 floats each:
// position, x, y, z axes), exactly as `count` calls of setCameraPosition/LocalSpace + renderFrame
// would, but without per-frame host synchronisation.  Row p of the result is the
// single_dimension_fast row of pose p (uchar4 per ommatidium).  The result goes to `outDevice`
// (device pointer, e.g. a collective's send slot) when non-null, else to host `outRgba`.
double Renderer::renderPoseBatch(const float* poses12, size_t count, unsigned char* outRgba, void* outDevice)
{
    if (!loaded_) throw std::runtime_error("renderPoseBatch called before loadGlTFscene");
    if (!compoundActive()) throw std::runtime_error("renderPoseBatch needs an active compound eye");
    if (!poses12 && count) throw std::runtime_error("renderPoseBatch: null pose array with a non-zero count");
    ensureDevice();
    if (!dscene_.nodes) uploadScene();
    HostCamera& cam = camera();
    CompoundState& cs = compoundState(current_);
    const auto t0 = std::chrono::steady_clock::now();
    dropReadAhead(cs);                                   // frames rendered ahead for renderFrame: the batch continues where the CALLER is
    cs.lastPoseValid = false;
    cs.standingFrames = 0;
    prepareCompound(cs, cam);
    const size_t N = static_cast<size_t>(cs.N);
    uchar4* dOut = static_cast<uchar4*>(outDevice);
    uchar4* dTmp = nullptr;
    if (!dOut) { dTmp = dallocT<uchar4>(N * count); dOut = dTmp; }
    const bool fused = fusedReduce && cs.S % 32 == 0 && !dumpRays;
    size_t F = batchFramesPerLaunch(cs, count, fused);
    if (dumpRays) F = 1;
    lastBatchFrames_ = static_cast<int>(F);
    ensureBatchBuffers(cs, F, fused);                    // (allocations stay out of the timed region)
    if (cs.batchPoseCap < count) {
        dfree(cs.dBatchPoses);
        cs.dBatchPoses = dallocT<DevicePose>(count);
        cs.batchPoseCap = count;
    }
    std::vector<DevicePose> hPoses(count);
    for (size_t p = 0; p < count; p++) {
        Pose pose;
        const float* q = poses12 + 12 * p;
        pose.pos = {q[0], q[1], q[2]};
        pose.ax = {q[3], q[4], q[5]};
        pose.ay = {q[6], q[7], q[8]};
        pose.az = {q[9], q[10], q[11]};
        hPoses[p] = toDevicePose(pose);
    }
    CR_CUDA(cudaMemcpyAsync(cs.dBatchPoses, hPoses.data(), sizeof(DevicePose) * count, cudaMemcpyHostToDevice, stream_));
    CR_CUDA(cudaEventRecord(evA_, stream_));
    for (size_t p0 = 0; p0 < count; p0 += F) {
        const size_t Fc = std::min(F, count - p0);
        if (dumpRays) {                       // the per-ray dump lives in the single-frame path
            Pose pose;
            const float* q = poses12 + 12 * p0;
            pose.pos = {q[0], q[1], q[2]}; pose.ax = {q[3], q[4], q[5]}; pose.ay = {q[6], q[7], q[8]}; pose.az = {q[9], q[10], q[11]};
            launchCompound(cs, cam, pose, dOut + p0 * N, cs.N);
        } else {
            launchCompoundBatch(cs, cs.dBatchPoses + p0, static_cast<int>(Fc), fused ? nullptr : cs.dBatchSamples, cs.dBatchSummed, dOut + p0 * N);
        }
    }
    CR_CUDA(cudaEventRecord(evB_, stream_));
    if (outRgba) CR_CUDA(cudaMemcpyAsync(outRgba, dOut, sizeof(uchar4) * N * count, cudaMemcpyDeviceToHost, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));
    CR_CUDA(cudaGetLastError());
    float k = 0.0f;
    cudaEventElapsedTime(&k, evA_, evB_);
    lastTraceMs_ = k;
    dfree(dTmp);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// ------------------------------------------------------------------------------------------
// debug / parity access
// ------------------------------------------------------------------------------------------
void Renderer::debugCopyBvh(float* nodes, float* tris)
{
    ensureDevice();
    if (!dscene_.nodes) uploadScene();
    if (nodes) CR_CUDA(cudaMemcpy(nodes, bvh_.nodes, sizeof(float4) * 4 * static_cast<size_t>(bvh_.nNodes), cudaMemcpyDeviceToHost));
    if (tris && bvh_.nTris) CR_CUDA(cudaMemcpy(tris, bvh_.tris, sizeof(float4) * 3 * static_cast<size_t>(bvh_.nTris), cudaMemcpyDeviceToHost));
}

void Renderer::debugCopyRngStates(uint32_t* out8)
{
    if (!compoundActive()) return;
    CompoundState& cs = compoundState(current_);
    if (!cs.dRng) return;
    dropReadAhead(cs);                                   // the states the CALLER has reached, not those of frames rendered ahead
    if (cs.needRewind) prepareCompound(cs, camera());
    const size_t N = static_cast<size_t>(cs.rngN), S = static_cast<size_t>(cs.rngS);
    std::vector<uint32_t> tmp(8 * N * S);
    CR_CUDA(cudaMemcpyAsync(tmp.data(), cs.dRng, sizeof(uint32_t) * tmp.size(), cudaMemcpyDeviceToHost, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));
    for (size_t o = 0; o < N; o++)
        for (size_t s = 0; s < S; s++)
            memcpy(out8 + 8 * (N * s + o), tmp.data() + 8 * (o * S + s), 32);
}

size_t Renderer::debugCopyLastRays(float* origins, float* dirs, int32_t* hits4)
{
    if (!compoundActive()) return 0;
    CompoundState& cs = compoundState(current_);
    const size_t n = static_cast<size_t>(cs.N) * static_cast<size_t>(cs.S);
    if (!cs.dDumpH || cs.dumpCap < n) return 0;
    CR_CUDA(cudaMemcpy(origins, cs.dDumpO, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
    CR_CUDA(cudaMemcpy(dirs, cs.dDumpD, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
    CR_CUDA(cudaMemcpy(hits4, cs.dDumpH, sizeof(int4) * n, cudaMemcpyDeviceToHost));
    return n;
}

size_t Renderer::debugCopyCandidateLists(int32_t* out, size_t records)
{
    if (!compoundActive()) return 0;
    CompoundState& cs = compoundState(current_);
    const size_t n = std::min(records, cs.listsLast);
    if (!cs.dLists || n == 0) return 0;
    CR_CUDA(cudaMemcpy(out, cs.dLists, sizeof(int) * n * static_cast<size_t>(candidateListStride()), cudaMemcpyDeviceToHost));
    return n;
}

size_t Renderer::debugCopyLastRayCounts(int32_t* counts2)
{
    if (!compoundActive()) return 0;
    CompoundState& cs = compoundState(current_);
    const size_t n = static_cast<size_t>(cs.N) * static_cast<size_t>(cs.S);
    if (!cs.dDumpC || cs.dumpCap < n) return 0;
    CR_CUDA(cudaMemcpy(counts2, cs.dDumpC, sizeof(int2) * n, cudaMemcpyDeviceToHost));
    return n;
}

void Renderer::debugTraceRays(const float* origins, const float* dirs, const float* tmins, int n, int32_t* hits8)
{
    ensureDevice();
    if (!loaded_) throw std::runtime_error("no scene loaded");
    if (!dscene_.nodes) uploadScene();
    float* dO = dallocT<float>(3 * static_cast<size_t>(n));
    float* dD = dallocT<float>(3 * static_cast<size_t>(n));
    float* dT = dallocT<float>(static_cast<size_t>(n));
    int4* dH = dallocT<int4>(2 * static_cast<size_t>(n));
    CR_CUDA(cudaMemcpyAsync(dO, origins, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, stream_));
    CR_CUDA(cudaMemcpyAsync(dD, dirs, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, stream_));
    CR_CUDA(cudaMemcpyAsync(dT, tmins, sizeof(float) * n, cudaMemcpyHostToDevice, stream_));
    launchTraceRays(dscene_, dO, dD, dT, n, dH, stream_);
    launches_++;
    CR_CUDA(cudaMemcpyAsync(hits8, dH, sizeof(int4) * 2 * static_cast<size_t>(n), cudaMemcpyDeviceToHost, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));
    CR_CUDA(cudaGetLastError());
    dfree(dO); dfree(dD); dfree(dT); dfree(dH);
}

void Renderer::debugCopyProjectionMap(uint32_t* out)
{
    if (!compoundActive()) return;
    CompoundState& cs = compoundState(current_);
    if (!cs.dMap) return;
    CR_CUDA(cudaMemcpy(out, cs.dMap, sizeof(uint32_t) * static_cast<size_t>(cs.mapW) * static_cast<size_t>(cs.mapH), cudaMemcpyDeviceToHost));
}

void Renderer::debugSampleTexture(int index, const float* uv, int n, float* out4)
{
    ensureDevice();
    if (!loaded_) throw std::runtime_error("no scene loaded");
    if (!dscene_.nodes) uploadScene();
    if (index < 0 || static_cast<size_t>(index) >= texObjects_.size()) throw std::runtime_error("texture index out of range");
    float* dUv = dallocT<float>(2 * static_cast<size_t>(n));
    float4* dOut = dallocT<float4>(static_cast<size_t>(n));
    CR_CUDA(cudaMemcpyAsync(dUv, uv, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, stream_));
    launchSampleTexture(static_cast<unsigned long long>(texObjects_[static_cast<size_t>(index)]), dUv, n, dOut, stream_);
    launches_++;
    CR_CUDA(cudaMemcpyAsync(out4, dOut, sizeof(float4) * n, cudaMemcpyDeviceToHost, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));
    CR_CUDA(cudaGetLastError());
    dfree(dUv); dfree(dOut);
}

void Renderer::debugEvalMath(int fn, const float* a, const float* b, float* out, int n)
{
    ensureDevice();
    float* dA = dallocT<float>(static_cast<size_t>(n));
    float* dB = b ? dallocT<float>(static_cast<size_t>(n)) : nullptr;
    float* dO = dallocT<float>(static_cast<size_t>(n));
    CR_CUDA(cudaMemcpyAsync(dA, a, sizeof(float) * n, cudaMemcpyHostToDevice, stream_));
    if (b) CR_CUDA(cudaMemcpyAsync(dB, b, sizeof(float) * n, cudaMemcpyHostToDevice, stream_));
    launchEvalMath(fn, dA, dB, dO, n, stream_);
    launches_++;
    CR_CUDA(cudaMemcpyAsync(out, dO, sizeof(float) * n, cudaMemcpyDeviceToHost, stream_));
    CR_CUDA(cudaStreamSynchronize(stream_));
    CR_CUDA(cudaGetLastError());
    dfree(dA); dfree(dB); dfree(dO);
}

}  // namespace cr
