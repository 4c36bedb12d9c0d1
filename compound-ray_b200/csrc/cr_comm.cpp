// cr_comm.cpp -- the multi-GPU data plane of the render path, owned by the C++ library.
//
// The path shards by camera pose: rank r of R renders a contiguous block of a P-pose run on its own GPU (scene, BVH
// and eye replicated), positions its sample streams at its first frame, and the ONLY exchange is the gather of the
// per-pose uchar4 rows -- 4*N bytes per pose -- to every rank.  That gather is issued here, on NCCL, chunk by chunk on
// a second stream while the next chunk is still being traced; K1b / k_sumPartials write each pose's row straight into
// its final place of the gathered buffer, so there is no pack kernel, no staging copy and no reshuffle.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy already loaded into the process, e.g. torch's, wins), so
// the library keeps loading, and single-GPU users keep running, on machines without NCCL.  <nccl.h> supplies types only.
//
// One process per GPU.  The 128-byte unique id travels between the processes by whatever the launcher has
// (torch.distributed broadcast, MPI, a file): crCommGetUniqueId on rank 0, crCommInit on every rank.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "cr_renderer.h"

namespace cr {

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    void* handle = nullptr;
};

NcclApi& nccl()
{
    static NcclApi api;
    if (api.handle) return api;
    const char* names[] = {getenv("CR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)                                       // a copy that is already in the process first
        if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;
    if (!h)
        for (const char* n : names)
            if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) throw std::runtime_error("NCCL not found (libnccl.so.2; set CR_NCCL_LIB): multi-GPU gathers are unavailable");
    auto sym = [&](const char* s) {
        void* p = dlsym(h, s);
        if (!p) throw std::runtime_error(std::string("libnccl lacks ") + s);
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.handle = h;
    return api;
}

void ncclCheck(ncclResult_t r, const char* what)
{
    if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + nccl().GetErrorString(r));
}
void cudaCheck(cudaError_t e, const char* what)
{
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

}  // namespace

// Contiguous pose block of `rank` (blocks differ by at most one pose) -- the rule of compound-ray_b200/sharding.py.
void poseBlock(int rank, int world, size_t count, size_t& lo, size_t& hi)
{
    const size_t base = count / static_cast<size_t>(world), extra = count % static_cast<size_t>(world);
    const size_t r = static_cast<size_t>(rank);
    lo = r * base + std::min(r, extra);
    hi = lo + base + (r < extra ? 1 : 0);
}

void Renderer::commUniqueId(void* out128)
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclCheck(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out128, &id, sizeof(id));
}

void Renderer::commInit(const void* id128, int nRanks, int rank)
{
    if (!id128 || nRanks < 1 || rank < 0 || rank >= nRanks) throw std::runtime_error("crCommInit: bad arguments");
    ensureDevice();
    commDestroy();
    if (nRanks == 1) {                       // one rank gathers nothing: no communicator (and none of NCCL's service threads: a size-1
        commRank_ = 0;                       // communicator alone made a 100 000-pose job 40 % slower on one GPU, 2.2 s vs 1.55 s --
        commSize_ = 1;                       // profiles/r03k_pose_batch_100k_plain_1gpu.json vs r03j_pose_batch_100k_native_1gpu.json)
        return;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    ncclCheck(nccl().CommInitRank(&c, nRanks, id, rank), "ncclCommInitRank");
    comm_ = c;
    commRank_ = rank;
    commSize_ = nRanks;
    cudaCheck(cudaStreamCreateWithFlags(&commStream_, cudaStreamNonBlocking), "cudaStreamCreate");
}

void Renderer::commDestroy()
{
    if (comm_) {
        if (commStream_) cudaStreamSynchronize(commStream_);
        nccl().CommDestroy(static_cast<ncclComm_t>(comm_));
        comm_ = nullptr;
    }
    if (commStream_) { cudaStreamDestroy(commStream_); commStream_ = nullptr; }
    commRank_ = 0;
    commSize_ = 1;
}

int Renderer::commNcclVersion()
{
    int v = 0;
    ncclCheck(nccl().GetVersion(&v), "ncclGetVersion");
    return v;
}

// recvDevice[q * bytesPerRank ...] = rank q's sendDevice, on every rank (sendDevice may be the rank's own slot of
// recvDevice).  Blocking: returns when the gathered buffer is complete.
void Renderer::allGatherRows(const void* sendDevice, void* recvDevice, size_t bytesPerRank)
{
    if (!sendDevice || !recvDevice) throw std::runtime_error("crAllGatherRows: null buffer");
    ensureDevice();
    if (commSize_ == 1 || !comm_) {
        char* dst = static_cast<char*>(recvDevice);
        if (dst != sendDevice) cudaCheck(cudaMemcpyAsync(dst, sendDevice, bytesPerRank, cudaMemcpyDeviceToDevice, stream_), "cudaMemcpyAsync");
        cudaCheck(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
        return;
    }
    cudaCheck(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");      // rows written by the render stream are complete
    ncclCheck(nccl().AllGather(sendDevice, recvDevice, bytesPerRank, ncclUint8, static_cast<ncclComm_t>(comm_), commStream_), "ncclAllGather");
    cudaCheck(cudaStreamSynchronize(commStream_), "cudaStreamSynchronize");
}

// One logical P-pose run across the communicator.  Every rank passes the SAME pose array; rank r renders its block
// [lo_r, hi_r) with its streams positioned at firstFrame + lo_r (so pose k is frame firstFrame + k of every stream,
// whatever the number of ranks) and every rank ends up with all P rows, in pose order, in outDevice (device memory,
// P*N*4 bytes) and/or outHost.  The block is cut into chunks of `chunkPoses` poses; the rows of chunk c are gathered
// (one ncclBroadcast per owning rank inside one group, straight into their final places) on the communication stream
// while chunk c+1 is traced.
double Renderer::renderPoseBatchSharded(const float* poses12, size_t count, unsigned char* outHost, void* outDevice, size_t chunkPoses,
                                        uint64_t firstFrame)
{
    if (!loaded_) throw std::runtime_error("crRenderPoseBatchSharded called before loadGlTFscene");
    if (!compoundActive()) throw std::runtime_error("crRenderPoseBatchSharded needs an active compound eye");
    if (!poses12 && count) throw std::runtime_error("crRenderPoseBatchSharded: null pose array with a non-zero count");
    ensureDevice();
    const auto t0 = std::chrono::steady_clock::now();
    const int R = comm_ ? commSize_ : 1, me = comm_ ? commRank_ : 0;
    const size_t N = ommatidialCount();
    const size_t rowBytes = 4 * N;
    unsigned char* dAll = static_cast<unsigned char*>(outDevice);
    unsigned char* dOwn = nullptr;
    if (!dAll) {
        cudaCheck(cudaMalloc(&dOwn, std::max<size_t>(1, count * rowBytes)), "cudaMalloc");
        dAll = dOwn;
    }
    size_t lo = 0, hi = 0;
    poseBlock(me, R, count, lo, hi);
    size_t maxBlock = 0;
    for (int q = 0; q < R; q++) { size_t a, b; poseBlock(q, R, count, a, b); maxBlock = std::max(maxBlock, b - a); }
    const size_t chunk = chunkPoses ? chunkPoses : std::max<size_t>(1, maxBlock);
    setFirstFrame(firstFrame + lo);
    double traceMs = 0.0;
    for (size_t c0 = 0; c0 < maxBlock; c0 += chunk) {
        // my rows of this chunk: rendered into their final place; returns when the render stream is idle
        const size_t a = std::min(lo + c0, hi), b = std::min(lo + c0 + chunk, hi);
        if (b > a) {
            renderPoseBatch(poses12 + 12 * a, b - a, nullptr, dAll + a * rowBytes);
            traceMs += lastTraceMs_;
        }
        if (R > 1) {                                              // gather chunk c of every rank while the next one is traced
            ncclCheck(nccl().GroupStart(), "ncclGroupStart");
            for (int q = 0; q < R; q++) {
                size_t qlo, qhi;
                poseBlock(q, R, count, qlo, qhi);
                const size_t qa = std::min(qlo + c0, qhi), qb = std::min(qlo + c0 + chunk, qhi);
                if (qb > qa)
                    ncclCheck(nccl().Broadcast(dAll + qa * rowBytes, dAll + qa * rowBytes, (qb - qa) * rowBytes, ncclUint8, q,
                                               static_cast<ncclComm_t>(comm_), commStream_), "ncclBroadcast");
            }
            ncclCheck(nccl().GroupEnd(), "ncclGroupEnd");
        }
    }
    if (R > 1) cudaCheck(cudaStreamSynchronize(commStream_), "cudaStreamSynchronize");
    lastTraceMs_ = traceMs;
    if (outHost && count) cudaCheck(cudaMemcpy(outHost, dAll, count * rowBytes, cudaMemcpyDeviceToHost), "cudaMemcpy");
    if (dOwn) cudaFree(dOwn);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace cr
