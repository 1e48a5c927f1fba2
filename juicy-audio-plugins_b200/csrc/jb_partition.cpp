// jb_partition.cpp -- spatial partitions of the GPU for launches that must not share SMs.
//
// Why: several DIFFERENT latency-bound kernels side by side (the parameter sets of a JuicyTexture engine -- BASELINE config 3:
// five materials -- or the plugins of a chain pipelined across streams) evict each other's unrolled sample loops (~50 KB each)
// from the SMs' instruction caches: five Texture kernels that would all fit the GPU at once took 58 ms against 10 - 15 ms each
// alone (profiles/r02_tma.txt).  Kernels confined to disjoint sets of SMs do not meet in any instruction cache.  CUDA's green
// contexts (driver API, 12.4+) provide exactly that: split the device's SM resource into groups, create a green context per
// group and a stream inside it; kernels launched on such a stream run on that group's SMs only.
//
// Everything is resolved at run time (cudaGetDriverEntryPoint) and every failure simply means "no partitions": the caller
// falls back to ordinary streams.  JB_SM_PARTITIONS=0 disables.
#include "jb_partition.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

namespace jb {
namespace {

template <class Fn>
Fn entry(const char* name)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        p = nullptr;
    cudaGetLastError();
    return reinterpret_cast<Fn>(p);
}

typedef CUresult (*GetDevResourceFn)(CUdevice, CUdevResource*, CUdevResourceType);
typedef CUresult (*SplitByCountFn)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
typedef CUresult (*GenerateDescFn)(CUdevResourceDesc*, CUdevResource*, unsigned int);
typedef CUresult (*GreenCtxCreateFn)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
typedef CUresult (*GreenCtxDestroyFn)(CUgreenCtx);
typedef CUresult (*GreenCtxStreamCreateFn)(CUstream*, CUgreenCtx, unsigned int, int);
typedef CUresult (*DeviceGetFn)(CUdevice*, int);

} // namespace

SmPartitions::~SmPartitions() { release(); }

void SmPartitions::release()
{
    for (int i = 0; i < count; ++i) {
        if (streams[i])
            cudaStreamDestroy(static_cast<cudaStream_t>(streams[i]));
        streams[i] = nullptr;
    }
    GreenCtxDestroyFn destroy = entry<GreenCtxDestroyFn>("cuGreenCtxDestroy");
    for (int i = 0; i < count; ++i) {
        if (contexts[i] && destroy)
            destroy(static_cast<CUgreenCtx>(contexts[i]));
        contexts[i] = nullptr;
    }
    count = 0;
    smsPerGroup = 0;
}

// Split `device` into `groups` partitions of equal size (a multiple of 8 SMs on sm_90+).  Returns the number of partitions
// created (0: unavailable -- use ordinary streams).
int SmPartitions::create(int device, int groups)
{
    release();
    static const bool enabled = [] { const char* v = std::getenv("JB_SM_PARTITIONS"); return v == nullptr || std::atoi(v) != 0; }();
    if (!enabled || groups < 2 || groups > kMax)
        return 0;
    GetDevResourceFn getRes = entry<GetDevResourceFn>("cuDeviceGetDevResource");
    SplitByCountFn split = entry<SplitByCountFn>("cuDevSmResourceSplitByCount");
    GenerateDescFn genDesc = entry<GenerateDescFn>("cuDevResourceGenerateDesc");
    GreenCtxCreateFn ctxCreate = entry<GreenCtxCreateFn>("cuGreenCtxCreate");
    GreenCtxStreamCreateFn streamCreate = entry<GreenCtxStreamCreateFn>("cuGreenCtxStreamCreate");
    DeviceGetFn deviceGet = entry<DeviceGetFn>("cuDeviceGet");
    if (!getRes || !split || !genDesc || !ctxCreate || !streamCreate || !deviceGet)
        return 0;
    CUdevice dev;
    if (deviceGet(&dev, device) != CUDA_SUCCESS)
        return 0;
    CUdevResource all;
    if (getRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS)
        return 0;
    const unsigned int total = all.sm.smCount;
    unsigned int per = (total / (unsigned int) groups) / 8u * 8u;
    if (per < 8u)
        return 0;
    CUdevResource parts[kMax];
    CUdevResource rest;
    unsigned int nb = (unsigned int) groups;
    if (split(parts, &nb, &all, &rest, 0, per) != CUDA_SUCCESS || nb < (unsigned int) groups)
        return 0;
    for (int i = 0; i < groups; ++i) {
        CUdevResourceDesc desc;
        CUgreenCtx ctx = nullptr;
        CUstream st = nullptr;
        if (genDesc(&desc, &parts[i], 1) != CUDA_SUCCESS || ctxCreate(&ctx, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS
            || streamCreate(&st, ctx, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
            contexts[i] = ctx;
            count = i + 1;
            release();
            return 0;
        }
        contexts[i] = ctx;
        streams[i] = st;
        count = i + 1;
    }
    smsPerGroup = (int) parts[0].sm.smCount;
    return count;
}

} // namespace jb
