// cr_abi.cpp -- extern "C" boundary of libEyeRenderer3.so (see include/libEyeRenderer.h).
// Mirrors libEyeRenderer3/libEyeRenderer.cpp:205-481 call for call; exceptions stop here.
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/libEyeRenderer.h"
#include "cr_renderer.h"

using cr::renderer;

#define CR_GUARD_BEGIN try {
#define CR_GUARD_END(ret)                                                       \
    } catch (const std::exception& e) {                                         \
        std::cerr << "[PyEye] ERROR: " << e.what() << std::endl;                \
        return ret;                                                             \
    } catch (...) {                                                             \
        std::cerr << "[PyEye] ERROR: unknown exception" << std::endl;           \
        return ret;                                                             \
    }

extern "C" {

// ---------------------------------------------------------------- configuration / frame loop
void setVerbosity(bool v) { renderer().verbose = v; }

void loadGlTFscene(const char* filepath)
{
    CR_GUARD_BEGIN
    renderer().loadScene(filepath ? filepath : "");
    CR_GUARD_END()
}
void stop(void)
{
    CR_GUARD_BEGIN
    renderer().stop();
    CR_GUARD_END()
}
void setRenderSize(int w, int h)
{
    CR_GUARD_BEGIN
    renderer().setRenderSize(w, h);
    CR_GUARD_END()
}
double renderFrame(void)
{
    CR_GUARD_BEGIN
    return renderer().renderFrame();
    CR_GUARD_END(0.0)
}
void displayFrame(void) {}   // headless: nothing to present to
void saveFrameAs(char* ppmFilename)
{
    CR_GUARD_BEGIN
    renderer().saveFrame(ppmFilename ? ppmFilename : "");
    CR_GUARD_END()
}
unsigned char* getFramePointer(void)
{
    CR_GUARD_BEGIN
    return renderer().framePointer();
    CR_GUARD_END(nullptr)
}

// ---------------------------------------------------------------- camera control
size_t getCameraCount(void) { return renderer().cameraCount(); }
void nextCamera(void) { renderer().setCurrentCamera(static_cast<int>(renderer().cameraIndex()) + 1); }
void previousCamera(void) { renderer().setCurrentCamera(static_cast<int>(renderer().cameraIndex()) - 1); }
size_t getCurrentCameraIndex(void) { return renderer().cameraIndex(); }
const char* getCurrentCameraName(void)
{
    CR_GUARD_BEGIN
    return renderer().camera().name.c_str();
    CR_GUARD_END("")
}
void gotoCamera(int index) { renderer().setCurrentCamera(index); }
bool gotoCameraByName(char* name)
{
    CR_GUARD_BEGIN
    cr::Renderer& r = renderer();
    r.setCurrentCamera(0);
    const size_t n = r.cameraCount();
    for (size_t i = 0; i < n; i++) {
        if (name && strcmp(name, r.camera().name.c_str()) == 0) return true;
        r.setCurrentCamera(static_cast<int>(r.cameraIndex()) + 1);
    }
    return false;
    CR_GUARD_END(false)
}
void setCameraPosition(float x, float y, float z) { renderer().camera().pose.pos = {x, y, z}; }
void getCameraPosition(float* x, float* y, float* z)
{
    const cr::Float3& p = renderer().camera().pose.pos;
    if (x) *x = p.x;
    if (y) *y = p.y;
    if (z) *z = p.z;
}
void setCameraLocalSpace(float lxx, float lxy, float lxz, float lyx, float lyy, float lyz, float lzx, float lzy, float lzz)
{
    cr::Pose& p = renderer().camera().pose;
    p.ax = {lxx, lxy, lxz};
    p.ay = {lyx, lyy, lyz};
    p.az = {lzx, lzy, lzz};
}
void rotateCameraAround(float angle, float x, float y, float z) { cr::poseRotateAround(renderer().camera().pose, angle, {x, y, z}); }
void rotateCameraLocallyAround(float angle, float x, float y, float z)
{ cr::poseRotateLocallyAround(renderer().camera().pose, angle, {x, y, z}); }
void translateCamera(float x, float y, float z) { cr::poseMove(renderer().camera().pose, {x, y, z}); }
void translateCameraLocally(float x, float y, float z) { cr::poseMoveLocally(renderer().camera().pose, {x, y, z}); }
void resetCameraPose(void) { cr::poseReset(renderer().camera().pose); }
void setCameraPose(float posX, float posY, float posZ, float rotX, float rotY, float rotZ)
{
    cr::Pose& p = renderer().camera().pose;
    cr::poseReset(p);
    cr::poseRotateAround(p, rotX, {1.f, 0.f, 0.f});
    cr::poseRotateAround(p, rotY, {0.f, 1.f, 0.f});
    cr::poseRotateAround(p, rotZ, {0.f, 0.f, 1.f});
    cr::poseMove(p, {posX, posY, posZ});
}

// ---------------------------------------------------------------- compound eye
bool isCompoundEyeActive(void) { return renderer().compoundActive(); }
void setCurrentEyeSamplesPerOmmatidium(int s)
{
    CR_GUARD_BEGIN
    renderer().setSamples(s);
    CR_GUARD_END()
}
int getCurrentEyeSamplesPerOmmatidium(void) { return renderer().samples(); }
void changeCurrentEyeSamplesPerOmmatidiumBy(int s)
{
    CR_GUARD_BEGIN
    cr::Renderer& r = renderer();
    if (r.compoundActive()) r.setSamples(r.samples() + s);
    CR_GUARD_END()
}
size_t getCurrentEyeOmmatidialCount(void) { return renderer().ommatidialCount(); }
void setOmmatidia(struct OmmatidiumPacket* omms, size_t count)
{
    CR_GUARD_BEGIN
    static_assert(sizeof(OmmatidiumPacket) == sizeof(cr::Ommatidium), "packet layout");
    renderer().setOmmatidia(reinterpret_cast<const cr::Ommatidium*>(omms), count);
    CR_GUARD_END()
}
const char* getCurrentEyeDataPath(void)
{
    cr::Renderer& r = renderer();
    if (r.compoundActive()) return r.camera().eyePath.c_str();
    return "";
}
void setCurrentEyeShaderName(char* name)
{
    cr::Renderer& r = renderer();
    if (r.compoundActive() && name) r.camera().projection = name;
}

// ---------------------------------------------------------------- scene queries
bool isInsideHitGeometry(float x, float y, float z, char* name)
{
    CR_GUARD_BEGIN
    const std::string n = name ? name : "";
    for (const cr::HitboxMesh& hb : renderer().scene().hitboxes)
        if (hb.name == n) return cr::pointInsideHitbox(hb, {x, y, z});
    std::cerr << "WARNING: No hitbox with the given name \"" << n << "\" is present in the scene." << std::endl;
    return false;
    CR_GUARD_END(false)
}
static crFloat3 geometryBound(const char* name, bool wantMax)
{
    const std::string n = name ? name : "";
    cr::HostScene& sc = renderer().scene();
    for (const cr::HitboxMesh& hb : sc.hitboxes)                        // MulticamScene.cpp:1796-1818
        if (hb.name == n) { const cr::Float3& b = wantMax ? hb.wmax : hb.wmin; return crFloat3{b.x, b.y, b.z}; }
    for (const cr::MeshGroup& m : sc.meshes)
        if (m.name == n) { const cr::Float3& b = wantMax ? m.wmax : m.wmin; return crFloat3{b.x, b.y, b.z}; }
    return crFloat3{0.f, 0.f, 0.f};
}
crFloat3 getGeometryMaxBounds(char* name) { return geometryBound(name, true); }
crFloat3 getGeometryMinBounds(char* name) { return geometryBound(name, false); }

// ================================================================ additions
void crSetDevice(int device)
{
    CR_GUARD_BEGIN
    renderer().setDevice(device);
    CR_GUARD_END()
}
void crGetOmmatidialData(float* outRgb)
{
    CR_GUARD_BEGIN
    renderer().copyOmmatidialData(outRgb);
    CR_GUARD_END()
}
double crRenderPoseBatch(const float* poses, size_t count, unsigned char* outHost, void* outDevice)
{
    CR_GUARD_BEGIN
    return renderer().renderPoseBatch(poses, count, outHost, outDevice);
    CR_GUARD_END(-1.0)
}
void crSetFirstFrame(uint64_t frame)
{
    CR_GUARD_BEGIN
    renderer().setFirstFrame(frame);
    CR_GUARD_END()
}
void crSetOmmatidialShard(uint64_t globalCount, uint64_t firstIndex)
{
    CR_GUARD_BEGIN
    renderer().setOmmatidialShard(globalCount, firstIndex);
    CR_GUARD_END()
}
int crCommGetUniqueId(void* out128)
{
    CR_GUARD_BEGIN
    renderer().commUniqueId(out128);
    return 0;
    CR_GUARD_END(-1)
}
int crCommInit(const void* id128, int nRanks, int rank)
{
    CR_GUARD_BEGIN
    renderer().commInit(id128, nRanks, rank);
    return 0;
    CR_GUARD_END(-1)
}
void crCommDestroy(void)
{
    CR_GUARD_BEGIN
    renderer().commDestroy();
    CR_GUARD_END()
}
int crCommRank(void) { return renderer().commRank(); }
int crCommSize(void) { return renderer().commSize(); }
int crCommNcclVersion(void)
{
    CR_GUARD_BEGIN
    return renderer().commNcclVersion();
    CR_GUARD_END(0)
}
int crAllGatherRows(const void* sendDevice, void* recvDevice, size_t bytesPerRank)
{
    CR_GUARD_BEGIN
    renderer().allGatherRows(sendDevice, recvDevice, bytesPerRank);
    return 0;
    CR_GUARD_END(-1)
}
double crRenderPoseBatchSharded(const float* poses, size_t count, unsigned char* outHost, void* outDevice, size_t chunkPoses,
                                uint64_t firstFrame)
{
    CR_GUARD_BEGIN
    return renderer().renderPoseBatchSharded(poses, count, outHost, outDevice, chunkPoses, firstFrame);
    CR_GUARD_END(-1.0)
}
void crSetRenderMode(int fusedReduction, int fastMath)
{
    if (fusedReduction >= 0) renderer().fusedReduce = fusedReduction != 0;
    if (fastMath >= 0) renderer().fastMath = fastMath != 0;
}
int crGetRenderMode(void) { return (renderer().fusedReduce ? 1 : 0) | (renderer().fastMath ? 2 : 0); }
double crGetLastTraceMs(void) { return renderer().lastTraceMs(); }
unsigned long long crGetLaunchCount(void) { return renderer().launchCount(); }
double crGetBvhBuildMs(void) { return renderer().bvhBuildMs(); }
int crGetLastBatchFrames(void) { return renderer().lastBatchFrames(); }

// ---------------------------------------------------------------- parity / debug access
size_t crDebugGetTriangleCount(void) { return renderer().scene().triangleCount(); }
size_t crDebugGetVertexCount(void) { return renderer().scene().vertexCount(); }
size_t crDebugGetMeshCount(void) { return renderer().scene().meshes.size(); }
void crDebugCopyTriangles(float* out9)
{
    const cr::HostScene& sc = renderer().scene();
    const size_t T = sc.triangleCount();
    for (size_t t = 0; t < T; t++) {
        const float* p0 = &sc.positions[3 * sc.indices[3 * t]];
        const float* p1 = &sc.positions[3 * sc.indices[3 * t + 1]];
        const float* p2 = &sc.positions[3 * sc.indices[3 * t + 2]];
        float* o = out9 + 9 * t;
        for (int a = 0; a < 3; a++) { o[a] = p0[a]; o[3 + a] = p1[a] - p0[a]; o[6 + a] = p2[a] - p0[a]; }
    }
}
void crDebugCopyTriangleMesh(int32_t* out)
{
    const cr::HostScene& sc = renderer().scene();
    for (size_t t = 0; t < sc.triMesh.size(); t++) out[t] = static_cast<int32_t>(sc.triMesh[t]);
}
void crDebugCopyMeshInfo(int32_t* out4, float* baseColor4)
{
    const cr::HostScene& sc = renderer().scene();
    for (size_t i = 0; i < sc.meshes.size(); i++) {
        const cr::MeshGroup& m = sc.meshes[i];
        if (out4) { out4[4 * i] = m.colorType; out4[4 * i + 1] = m.hasUV; out4[4 * i + 2] = m.texture; out4[4 * i + 3] = m.material; }
        if (baseColor4) memcpy(baseColor4 + 4 * i, m.baseColor, 16);
    }
}
void crDebugCopyCornerAttributes(float* uv6, float* col12)
{
    const cr::HostScene& sc = renderer().scene();
    const size_t T = sc.triangleCount();
    for (size_t t = 0; t < T; t++)
        for (int c = 0; c < 3; c++) {
            const uint32_t v = sc.indices[3 * t + c];
            if (uv6) { uv6[6 * t + 2 * c] = sc.uvs[2 * v]; uv6[6 * t + 2 * c + 1] = sc.uvs[2 * v + 1]; }
            if (col12) memcpy(col12 + 12 * t + 4 * c, &sc.colors[4 * v], 16);
        }
}
void crDebugCopyCameraPose(float* out12)
{
    const cr::Pose& p = renderer().camera().pose;
    const float v[12] = {p.pos.x, p.pos.y, p.pos.z, p.ax.x, p.ax.y, p.ax.z, p.ay.x, p.ay.y, p.ay.z, p.az.x, p.az.y, p.az.z};
    memcpy(out12, v, sizeof v);
}
void crDebugCopyCameraScale(float* out3) { memcpy(out3, renderer().camera().scale, 12); }
int crDebugGetCameraKind(void) { return static_cast<int>(renderer().camera().kind); }
void crDebugCopyOmmatidia(float* out8)
{
    const auto& o = renderer().camera().ommatidia;
    if (!o.empty()) memcpy(out8, o.data(), sizeof(cr::Ommatidium) * o.size());
}
static cr::ImageRGBA8 g_debugImage;
bool crDebugDecodeImageFile(const char* path, int* w, int* h)
{
    CR_GUARD_BEGIN
    g_debugImage = cr::decodeImageFile(path ? path : "");
    *w = g_debugImage.width;
    *h = g_debugImage.height;
    return true;
    CR_GUARD_END(false)
}
void crDebugCopyDecodedImage(unsigned char* outRgba)
{ if (!g_debugImage.pixels.empty()) memcpy(outRgba, g_debugImage.pixels.data(), g_debugImage.pixels.size()); }
int crDebugGetMissShader(void) { return renderer().scene().missShader; }
size_t crDebugGetTextureCount(void) { return renderer().scene().textures.size(); }
void crDebugGetTextureSize(int index, int* w, int* h)
{
    const auto& t = renderer().scene().textures;
    if (index < 0 || static_cast<size_t>(index) >= t.size()) { *w = 0; *h = 0; return; }
    *w = t[static_cast<size_t>(index)].width;
    *h = t[static_cast<size_t>(index)].height;
}
void crDebugCopyTexture(int index, unsigned char* outRgba)
{
    const auto& t = renderer().scene().textures;
    if (index < 0 || static_cast<size_t>(index) >= t.size()) return;
    memcpy(outRgba, t[static_cast<size_t>(index)].pixels.data(), t[static_cast<size_t>(index)].pixels.size());
}
size_t crDebugGetBvhNodeCount(void) { return static_cast<size_t>(renderer().debugNodeCount()); }
void crDebugCopyBvh(float* nodes16, float* tris12)
{
    CR_GUARD_BEGIN
    renderer().debugCopyBvh(nodes16, tris12);
    CR_GUARD_END()
}
void crDebugXorwowInit(uint64_t seed, uint64_t subsequence, uint64_t offset, uint32_t* out6)
{
    CR_GUARD_BEGIN
    uint32_t d, v[5];
    cr::xorwow::seedState(seed, d, v);
    cr::xorwow::jumpHost(cr::Renderer::xorwowTable(), v, 0, subsequence);
    cr::xorwow::jumpHost(cr::Renderer::xorwowTable(), v, cr::xorwow::kSeqLevels, offset);
    d += 362437u * static_cast<uint32_t>(offset);
    out6[0] = d;
    for (int k = 0; k < 5; k++) out6[1 + k] = v[k];
    CR_GUARD_END()
}
void crDebugSetRayDump(bool on) { renderer().dumpRays = on; }
size_t crDebugCopyLastRayCounts(int32_t* counts2)
{
    CR_GUARD_BEGIN
    return renderer().debugCopyLastRayCounts(counts2);
    CR_GUARD_END(0)
}
void crDebugSetCandidateLists(int on) { renderer().candidateLists = on; }
void crDebugSetWavefront(int on, int refillBelow, double queueFraction)
{
    renderer().wavefront = on;
    if (refillBelow > 0) renderer().wavefrontRefill = refillBelow;
    if (queueFraction >= 0.0) renderer().queueFraction = queueFraction;
}
void crDebugSetNodeLanes(int lanes) { renderer().nodeLanes = lanes; }
void crDebugSetFrameGroups(int on) { renderer().frameGroups = on != 0; }
void crDebugSetReadAhead(int on, double budgetMs)
{
    renderer().readAhead = on != 0;
    if (budgetMs > 0.0) renderer().readAheadBudgetMs = budgetMs;
}
void crDebugSetFrameProfile(int on) { renderer().profileFrame = on != 0; }
void crDebugFrameBreakdown(float* out3)
{
    CR_GUARD_BEGIN
    if (out3) renderer().debugFrameBreakdown(out3);
    CR_GUARD_END()
}
void crDebugSetDynamicChunks(int on) { renderer().dynamicChunks = on != 0; }
void crDebugSetSmAffine(int on, int minBlocksPerSm) { renderer().smAffine = on != 0; renderer().smAffineMinBlocks = minBlocksPerSm; }
void crDebugSetZeroCopy(int on) { renderer().zeroCopyFrames = on != 0; }
unsigned long long crDebugLastQueuedRays()
{
    CR_GUARD_BEGIN
    return renderer().debugLastQueuedRays();
    CR_GUARD_END(0)
}
size_t crDebugCopyCandidateLists(int32_t* out, size_t records)
{
    CR_GUARD_BEGIN
    return renderer().debugCopyCandidateLists(out, records);
    CR_GUARD_END(0)
}
void crDebugSetEntryFrontier(int on, int minSamples, long long minRays)
{
    renderer().entryFrontier = on;
    if (minSamples >= 0) renderer().entryMinSamples = minSamples;
    if (minRays >= 0) renderer().entryMinRays = minRays;
}
size_t crDebugCopyLastRays(float* origins3, float* dirs3, int32_t* hits4)
{
    CR_GUARD_BEGIN
    return renderer().debugCopyLastRays(origins3, dirs3, hits4);
    CR_GUARD_END(0)
}
void crDebugCopyRngStates(uint32_t* out8)
{
    CR_GUARD_BEGIN
    renderer().debugCopyRngStates(out8);
    CR_GUARD_END()
}
void crDebugTraceRays(const float* origins3, const float* dirs3, const float* tmins, int n, int32_t* hits8)
{
    CR_GUARD_BEGIN
    renderer().debugTraceRays(origins3, dirs3, tmins, n, hits8);
    CR_GUARD_END()
}
void crDebugCopyProjectionMap(uint32_t* out)
{
    CR_GUARD_BEGIN
    renderer().debugCopyProjectionMap(out);
    CR_GUARD_END()
}
void crDebugSampleTexture(int index, const float* uv2, int n, float* outRgba4)
{
    CR_GUARD_BEGIN
    renderer().debugSampleTexture(index, uv2, n, outRgba4);
    CR_GUARD_END()
}
void crDebugEvalMath(int fn, const float* a, const float* b, float* out, int n)
{
    CR_GUARD_BEGIN
    renderer().debugEvalMath(fn, a, b, out, n);
    CR_GUARD_END()
}

}  // extern "C"
