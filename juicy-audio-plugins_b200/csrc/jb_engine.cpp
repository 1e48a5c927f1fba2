// jb_engine.cpp -- the C ABI declared in include/juicy_batch.h.
//
// An engine owns, for N clips on one GPU: the parameter blocks of its chain
// (jb_params), the structure-of-arrays DSP/analyzer state, the Width delay ring
// and Texture waveguides, and the metrics records.  jb_process derives the
// block-constant coefficients on the host (libm) and launches ONE persistent
// kernel that walks every 512-sample block of every clip (jb_kernels.cu).
// There is no CPU fallback: without a CUDA device every compute call fails.
#include "../../include/juicy_batch.h"
#include "jb_kernels.h"
#include "jb_params.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define JB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return fail(JB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(err__));       \
    } while (0)

int pluginVarCount(int kind)
{
    switch (kind) {
        case jb::kPunch: return PV_COUNT;
        case jb::kSaturator: return SV_COUNT;
        case jb::kWidth: return WV_COUNT;
        case jb::kCohere: return CV_COUNT;
        case jb::kTexture: return TV_COUNT;
        case jb::kMotion: return MV_COUNT;
        default: return 0;
    }
}

struct VarInit {
    int var;          // absolute state variable index
    float value;      // initial value (bit pattern for int-typed variables)
    bool everyPrepare; // false: set when the instance is constructed only
};

// With at least this many clips in a launch, one lane per clip fills the GPU on its own.
constexpr int kLaneKernelMinClips = 148 * 4 * 32 * 4;

float bitsToFloat(uint32_t u)
{
    float f;
    std::memcpy(&f, &u, sizeof f);
    return f;
}

} // namespace

struct jb_engine {
    std::vector<int> chain;
    std::vector<jb::ParamSet> params;
    std::vector<int> stateBase;
    int totalVars = 0;
    int nClips = 0, nCh = 2, device = -1;
    long long clipPitch = 0;
    bool hostOnly = false;
    bool prepared = false;
    double sampleRate = 0.0;
    int blockSize = 0;
    long long blocksDone = 0;

    float* dState = nullptr;
    float* dLatest = nullptr;
    float* dHist = nullptr;
    float* dRing = nullptr;
    float* dWave = nullptr;
    float* dCoopScratch = nullptr; // mono rings of the cooperative kernel's analyzer lanes
    bool ringClipMajor = false;    // Width ring layout: [clip][ringLen] (cooperative-capable chains) or [ringLen][clip]
    bool coopCapable = false;      // chain made of plugins the cooperative kernel implements
    int numSMs = 0;
    int pathMode = 0;              // 0 auto, 1 force lane-per-clip, 2 force cooperative (fails if unsupported)
    long long coopLaunches = 0, laneLaunches = 0;
    int ringLen = 0, waveLen = 0;
    int histMaxBlocks = 0;
    bool stateConstructed = false;

    cudaStream_t stream = nullptr;
    bool ownStream = false;
    // jb_process_host staging
    cudaStream_t copyIn = nullptr, copyOut = nullptr;
    float* dStage[3] = { nullptr, nullptr, nullptr };
    std::vector<cudaEvent_t> sliceEvents;
    cudaEvent_t evIn[3] = {}, evDone[3] = {}, evOut[3] = {};
    size_t stageBytes = 0;

    std::vector<float> hostScratch;

    // CUDA-event pairs around every render-kernel launch (jb_kernel_time_ms)
    std::vector<cudaEvent_t> timingEvents; // start0, stop0, start1, stop1, ...
    size_t timingUsed = 0;
    double kernelMs = 0.0;
    long long kernelLaunches = 0;
};

namespace {

int checkEngine(const jb_engine* e)
{
    if (e == nullptr)
        return fail(JB_ERR_ARG, "null engine");
    return JB_OK;
}

int checkSlot(const jb_engine* e, int slot)
{
    if (checkEngine(e) != JB_OK)
        return JB_ERR_ARG;
    if (slot < 0 || slot >= (int) e->chain.size())
        return fail(JB_ERR_ARG, "slot %d out of range (chain length %d)", slot, (int) e->chain.size());
    return JB_OK;
}

int setDevice(const jb_engine* e)
{
    if (e->hostOnly)
        return fail(JB_ERR_CUDA, "engine was created without a CUDA device (device = -1): parameter logic only");
    JB_CUDA(cudaSetDevice(e->device));
    return JB_OK;
}

void freeDevice(jb_engine* e)
{
    if (e->hostOnly)
        return;
    cudaSetDevice(e->device);
    cudaFree(e->dState);
    cudaFree(e->dLatest);
    cudaFree(e->dHist);
    cudaFree(e->dRing);
    cudaFree(e->dWave);
    cudaFree(e->dCoopScratch);
    e->dState = e->dLatest = e->dHist = e->dRing = e->dWave = e->dCoopScratch = nullptr;
    for (int i = 0; i < 3; ++i) {
        cudaFree(e->dStage[i]);
        e->dStage[i] = nullptr;
        if (e->evIn[i]) cudaEventDestroy(e->evIn[i]);
        if (e->evDone[i]) cudaEventDestroy(e->evDone[i]);
        if (e->evOut[i]) cudaEventDestroy(e->evOut[i]);
        e->evIn[i] = e->evDone[i] = e->evOut[i] = nullptr;
    }
    e->stageBytes = 0;
    if (e->copyIn) cudaStreamDestroy(e->copyIn);
    if (e->copyOut) cudaStreamDestroy(e->copyOut);
    e->copyIn = e->copyOut = nullptr;
    for (cudaEvent_t ev : e->timingEvents)
        cudaEventDestroy(ev);
    e->timingEvents.clear();
    for (cudaEvent_t ev : e->sliceEvents)
        cudaEventDestroy(ev);
    e->sliceEvents.clear();
    e->timingUsed = 0;
    if (e->ownStream && e->stream) cudaStreamDestroy(e->stream);
    e->stream = nullptr;
}

// Initial values of the state variables that are not zero, and whether
// prepareToPlay re-applies them.
std::vector<VarInit> stateInits(const jb_engine* e)
{
    std::vector<VarInit> v;
    for (size_t s = 0; s < e->chain.size(); ++s) {
        const int pb = e->stateBase[s] + AV_COUNT;
        switch (e->chain[s]) {
            case jb::kCohere: // JuicyCohere/PluginProcessor.h:55-57: targets start at 0.2 and survive prepareToPlay
                v.push_back({ pb + CV_TGT_LOW, 0.2f, false });
                v.push_back({ pb + CV_TGT_MID, 0.2f, false });
                v.push_back({ pb + CV_TGT_HIGH, 0.2f, false });
                // compensation gains before the first block pre-pass (never read before it; neutral)
                v.push_back({ pb + CV_COMP_LOW, 1.0f, true });
                v.push_back({ pb + CV_COMP_MID, 1.0f, true });
                v.push_back({ pb + CV_COMP_HIGH, 1.0f, true });
                break;
            case jb::kTexture: // JuicyTexture/PluginProcessor.cpp:16 (rng reseeded in prepareToPlay), .h:65 protectGain = 1
                v.push_back({ pb + TV_PROTECT, 1.0f, true });
                v.push_back({ pb + TV_CH_STRIDE + TV_PROTECT, 1.0f, true });
                v.push_back({ pb + TV_RNG, bitsToFloat(0x12345678u), true });
                break;
            case jb::kMotion: // JuicyMotion/PluginProcessor.h:65: rng seeded at construction only
                v.push_back({ pb + MV_RNG, bitsToFloat(0x93ab12f0u), false });
                v.push_back({ pb + MV_REP_SCALE, 1.0f, true });
                v.push_back({ pb + MV_RECOVERY, 1.0f, true });
                break;
            default:
                break;
        }
    }
    return v;
}

bool persistsAcrossPrepare(const std::vector<VarInit>& inits, int var)
{
    for (const auto& i : inits)
        if (i.var == var && !i.everyPrepare)
            return true;
    return false;
}

int fillVar(jb_engine* e, int var, float value)
{
    if (jbk_launch_fill(e->dState + (long long) var * e->clipPitch, value, e->clipPitch, e->stream) != 0)
        return fail(JB_ERR_CUDA, "%s", jbk_last_cuda_error());
    return JB_OK;
}

int resetState(jb_engine* e)
{
    const auto inits = stateInits(e);
    if (!e->stateConstructed) {
        JB_CUDA(cudaMemsetAsync(e->dState, 0, sizeof(float) * (size_t) e->totalVars * (size_t) e->clipPitch, e->stream));
        for (const auto& i : inits)
            if (int rc = fillVar(e, i.var, i.value))
                return rc;
        e->stateConstructed = true;
    } else {
        // prepareToPlay: clear everything except members the reference only sets at construction
        for (int var = 0; var < e->totalVars; ++var) {
            if (persistsAcrossPrepare(inits, var))
                continue;
            JB_CUDA(cudaMemsetAsync(e->dState + (long long) var * e->clipPitch, 0, sizeof(float) * (size_t) e->clipPitch, e->stream));
        }
        for (const auto& i : inits)
            if (i.everyPrepare)
                if (int rc = fillVar(e, i.var, i.value))
                    return rc;
    }
    if (e->dRing)
        JB_CUDA(cudaMemsetAsync(e->dRing, 0, sizeof(float) * (size_t) e->ringLen * (size_t) e->clipPitch, e->stream));
    if (e->dWave)
        JB_CUDA(cudaMemsetAsync(e->dWave, 0, sizeof(float) * 2 * (size_t) e->waveLen * (size_t) e->clipPitch, e->stream));
    JB_CUDA(cudaMemsetAsync(e->dLatest, 0, sizeof(float) * e->chain.size() * JBK_REC * (size_t) e->clipPitch, e->stream));
    {   // getLatestMetrics() before any block: monoSafety mailbox starts at 1 (e.g. JuicyPunch/PluginProcessor.h:51)
        for (size_t s = 0; s < e->chain.size(); ++s)
            if (e->chain[s] != jb::kInfer || true)
                if (jbk_launch_fill(e->dLatest + ((long long) s * JBK_REC + 12) * e->clipPitch, 1.0f, e->clipPitch, e->stream) != 0)
                    return fail(JB_ERR_CUDA, "%s", jbk_last_cuda_error());
    }
    if (e->dHist)
        JB_CUDA(cudaMemsetAsync(e->dHist, 0, sizeof(float) * (size_t) e->histMaxBlocks * e->chain.size() * JBK_REC * (size_t) e->clipPitch, e->stream));
    e->blocksDone = 0;
    return JB_OK;
}

int buildArgs(jb_engine* e, ProcArgs& a, const float* dIn, float* dOut, int nSamples, int nClips, long long clipOffset,
              long long rowPitch = 0)
{
    std::memset(&a, 0, sizeof a);
    a.in = dIn;
    a.out = dOut;
    a.state = e->dState + clipOffset;
    a.latest = e->dLatest + clipOffset;
    a.hist = e->dHist ? e->dHist + clipOffset : nullptr;
    a.ringClipStride = e->ringClipMajor ? e->ringLen : 1;
    a.ringTimeStride = e->ringClipMajor ? 1 : e->clipPitch;
    a.widthRing = e->dRing ? e->dRing + clipOffset * a.ringClipStride : nullptr;
    a.texWave = e->dWave ? e->dWave + clipOffset : nullptr;
    a.clipPitch = e->clipPitch;
    a.nClips = nClips;
    a.nCh = e->nCh;
    a.nSamples = nSamples;
    a.rowPitch = rowPitch > 0 ? rowPitch : nSamples;
    a.blockSize = e->blockSize;
    a.histFirstBlock = (int) std::min<long long>(e->blocksDone, 0x7fffffff);
    a.histMaxBlocks = e->histMaxBlocks;
    a.chainLen = (int) e->chain.size();
    a.recSlotBase = 0;
    a.recChainLen = a.chainLen;
    const bool aligned = ((reinterpret_cast<uintptr_t>(dIn) | reinterpret_cast<uintptr_t>(dOut)) & 15u) == 0;
    a.vecOk = (aligned && nSamples % 4 == 0 && a.rowPitch % 4 == 0 && e->blockSize % 4 == 0) ? 1 : 0;
    a.octets = nClips >= 32768 ? 1 : 0;
    for (int k : e->chain)
        if (k == jb::kPunch || k == jb::kTexture || k == jb::kMotion)
            a.octets = 0;
    a.ana = jb::makeAnaCoef(e->sampleRate);
    for (size_t s = 0; s < e->chain.size(); ++s) {
        a.slot[s].kind = e->chain[s];
        a.slot[s].stateBase = e->stateBase[s];
        jb::makeSlotCoef(e->params[s], e->sampleRate, &a.slot[s].c);
    }
    return JB_OK;
}

// Fold the recorded event pairs into the running total (synchronises the stream).
int drainTiming(jb_engine* e)
{
    if (e->timingUsed == 0)
        return JB_OK;
    JB_CUDA(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i + 1 < e->timingUsed; i += 2) {
        float ms = 0.0f;
        JB_CUDA(cudaEventElapsedTime(&ms, e->timingEvents[i], e->timingEvents[i + 1]));
        e->kernelMs += (double) ms;
        ++e->kernelLaunches;
    }
    e->timingUsed = 0;
    return JB_OK;
}

// One render-kernel launch bracketed by timing events on the engine's stream.
int launchProcess(jb_engine* e, const ProcArgs& a)
{
    if (e->timingUsed + 2 > e->timingEvents.size()) {
        if (e->timingEvents.size() >= 256) {
            if (int rc = drainTiming(e))
                return rc;
        } else {
            for (int i = 0; i < 2; ++i) {
                cudaEvent_t ev = nullptr;
                JB_CUDA(cudaEventCreate(&ev));
                e->timingEvents.push_back(ev);
            }
        }
    }
    cudaEvent_t start = e->timingEvents[e->timingUsed], stop = e->timingEvents[e->timingUsed + 1];
    // Path choice: the cooperative (time-parallel) kernel when the chain and the call's shape allow it and
    // the batch is too small to fill the GPU with one lane per clip; the lane-per-clip kernel otherwise.
    bool coop = e->coopCapable && e->dCoopScratch != nullptr && jbk_coop_supported(&a) != 0;
    if (e->pathMode == 1)
        coop = false;
    else if (e->pathMode == 2 && !coop)
        return fail(JB_ERR_UNSUPPORTED, "cooperative path forced but this chain / call shape is not supported by it");
    else if (e->pathMode == 0 && coop) {
        // Measured crossover (profiles/README.md): the cooperative kernel renders 32 clips per SM at a time in
        // ~2.7 ms (one plugin) .. 4.1 ms (Punch -> Width) per second of audio; the lane kernel needs the whole
        // GPU's worth of lanes.  Analyzer-only chains (Infer) are cheap enough per lane that it wins from ~3 rounds.
        bool inferOnly = true, hasPunch = false;
        for (int k : e->chain) {
            inferOnly = inferOnly && k == jb::kInfer;
            hasPunch = hasPunch || k == jb::kPunch;
        }
        // rounds of 32 clips per SM up to which the cooperative kernel stays ahead (profiles/r01_survey_cross.txt):
        // Infer alone 2, Punch alone 3, Width (+Infer) 6, Punch -> Width 7
        const int rounds = inferOnly ? 2 : (hasPunch ? (e->chain.size() == 1 ? 3 : 7) : 6);
        const int limit = e->numSMs * 32 * rounds;
        if (a.nClips > limit)
            coop = false;
    }
    JB_CUDA(cudaEventRecord(start, e->stream));
    if (coop) {
        if (jbk_launch_coop(&a, e->dCoopScratch, e->numSMs, e->stream) != 0)
            return fail(JB_ERR_CUDA, "%s", jbk_coop_last_error());
        ++e->coopLaunches;
    } else {
        // Lane kernel, chain of several plugins: one launch per plugin over the whole call (plugin s + 1 only
        // needs plugin s's output of the same block, and every plugin blocks identically, so plugin-by-plugin
        // equals block-by-block).  A single-plugin launch is small code with no spills and picks its own
        // 8-samples-per-trip mode; the fused 8-sweep kernel took 2.5x the sum of its parts.
        // Measured (profiles/r01_survey_chain.txt): 7-plugin chain on 32768 clips 424 -> 161 ms; below ~20k clips
        // the render is latency-bound and the two modes tie.  JB_LANE_SPLIT=0/1 forces a mode.
        const char* splitEnv = std::getenv("JB_LANE_SPLIT");
        const int splitMode = splitEnv == nullptr ? -1 : std::atoi(splitEnv);
        const bool splitChains = splitMode < 0 ? a.nClips >= 20480 : splitMode != 0;
        const int L = a.chainLen;
        if (L > 1 && splitChains) {
            for (int s = 0; s < L; ++s) {
                ProcArgs one = a;
                one.in = s == 0 ? a.in : a.out;
                one.chainLen = 1;
                one.slot[0] = a.slot[s];
                one.recSlotBase = s;
                one.recChainLen = L;
                one.octets = a.nClips >= 32768 && !(one.slot[0].kind == jb::kPunch || one.slot[0].kind == jb::kTexture || one.slot[0].kind == jb::kMotion);
                if (jbk_launch_process(&one, e->stream) != 0)
                    return fail(JB_ERR_CUDA, "%s", jbk_last_cuda_error());
            }
        } else if (jbk_launch_process(&a, e->stream) != 0) {
            return fail(JB_ERR_CUDA, "%s", jbk_last_cuda_error());
        }
        ++e->laneLaunches;
    }
    JB_CUDA(cudaEventRecord(stop, e->stream));
    e->timingUsed += 2;
    return JB_OK;
}

void unpackRecords(const float* soa, long long pitch, int nClips, jb_metrics* out)
{
    for (int c = 0; c < nClips; ++c) {
        float* rec = reinterpret_cast<float*>(&out[c]);
        for (int f = 0; f < JBK_REC; ++f)
            rec[f] = soa[(long long) f * pitch + c];
    }
}

} // namespace

extern "C" {

const char* jb_last_error(void) { return g_error.c_str(); }
int jb_abi_version(void) { return JB_ABI_VERSION; }

int jb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

long long jb_launch_count(void) { return jbk_launch_count(); }

int jb_create(const int* chain, int chain_len, int n_clips, int n_channels, int device, jb_engine** out)
{
    if (out == nullptr || chain == nullptr)
        return fail(JB_ERR_ARG, "jb_create: null argument");
    *out = nullptr;
    if (chain_len < 1 || chain_len > JB_MAX_CHAIN)
        return fail(JB_ERR_ARG, "jb_create: chain length %d outside 1..%d", chain_len, JB_MAX_CHAIN);
    if (n_clips < 1)
        return fail(JB_ERR_ARG, "jb_create: n_clips must be >= 1");
    if (n_channels != 2)
        return fail(JB_ERR_UNSUPPORTED, "jb_create: this build renders stereo buses only (n_channels = %d)", n_channels);
    for (int i = 0; i < chain_len; ++i)
        if (chain[i] < 0 || chain[i] >= JB_NUM_KINDS)
            return fail(JB_ERR_ARG, "jb_create: unknown plugin kind %d at slot %d", chain[i], i);
    int widthSlots = 0, texSlots = 0;
    for (int i = 0; i < chain_len; ++i) {
        widthSlots += chain[i] == JB_WIDTH;
        texSlots += chain[i] == JB_TEXTURE;
    }
    if (widthSlots > 1 || texSlots > 1)
        return fail(JB_ERR_UNSUPPORTED, "jb_create: at most one Width and one Texture instance per chain");

    auto e = std::make_unique<jb_engine>();
    e->chain.assign(chain, chain + chain_len);
    e->nClips = n_clips;
    e->nCh = n_channels;
    e->clipPitch = ((long long) n_clips + 31) / 32 * 32;
    for (int i = 0; i < chain_len; ++i) {
        e->params.emplace_back(chain[i]);
        e->stateBase.push_back(e->totalVars);
        e->totalVars += AV_COUNT + pluginVarCount(chain[i]);
    }
    if (device < 0) {
        e->hostOnly = true; // parameter / program logic only; every compute call fails with JB_ERR_CUDA
    } else {
        int count = jb_device_count();
        if (count <= 0)
            return fail(JB_ERR_CUDA, "jb_create: no CUDA device available (this library has no CPU fallback)");
        if (device >= count)
            return fail(JB_ERR_ARG, "jb_create: device %d out of range (%d visible)", device, count);
        e->device = device;
        JB_CUDA(cudaSetDevice(device));
        JB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        e->ownStream = true;
    }
    *out = e.release();
    return JB_OK;
}

int jb_destroy(jb_engine* e)
{
    if (e == nullptr)
        return JB_OK;
    freeDevice(e);
    delete e;
    return JB_OK;
}

int jb_chain_length(const jb_engine* e) { return e ? (int) e->chain.size() : 0; }
int jb_chain_kind(const jb_engine* e, int slot) { return checkSlot(e, slot) == JB_OK ? e->chain[(size_t) slot] : JB_ERR_ARG; }
int jb_num_clips(const jb_engine* e) { return e ? e->nClips : 0; }

int jb_prepare(jb_engine* e, double sample_rate, int samples_per_block)
{
    if (int rc = checkEngine(e))
        return rc;
    if (!(sample_rate > 0.0) || samples_per_block < 1)
        return fail(JB_ERR_ARG, "jb_prepare: sample_rate %.3f / samples_per_block %d invalid", sample_rate, samples_per_block);
    if (int rc = setDevice(e))
        return rc;
    const int ringLen = jb::widthRingLength(sample_rate);
    const int waveLen = jb::textureWaveLength(sample_rate);
    const bool hasWidth = std::find(e->chain.begin(), e->chain.end(), (int) jb::kWidth) != e->chain.end();
    const bool hasTex = std::find(e->chain.begin(), e->chain.end(), (int) jb::kTexture) != e->chain.end();
    if (e->dState == nullptr) {
        JB_CUDA(cudaMalloc(&e->dState, sizeof(float) * (size_t) e->totalVars * (size_t) e->clipPitch));
        JB_CUDA(cudaMalloc(&e->dLatest, sizeof(float) * e->chain.size() * JBK_REC * (size_t) e->clipPitch));
    }
    e->coopCapable = e->chain.size() <= 3;
    for (size_t i = 0; i < e->chain.size(); ++i) {
        const int k = e->chain[i];
        if (!((k == jb::kPunch && i == 0) || k == jb::kWidth || k == jb::kInfer))
            e->coopCapable = false;
    }
    e->ringClipMajor = true; // both kernels: each clip owns a contiguous ring (16-byte quads)
    if (e->numSMs == 0) {
        JB_CUDA(cudaDeviceGetAttribute(&e->numSMs, cudaDevAttrMultiProcessorCount, e->device));
    }
    if (e->coopCapable && e->dCoopScratch == nullptr)
        JB_CUDA(cudaMalloc(&e->dCoopScratch, jbk_coop_scratch_bytes((int) e->chain.size(), e->numSMs)));
    if (hasWidth && (e->dRing == nullptr || ringLen != e->ringLen)) {
        cudaFree(e->dRing);
        e->dRing = nullptr;
        JB_CUDA(cudaMalloc(&e->dRing, sizeof(float) * (size_t) ringLen * (size_t) e->clipPitch));
    }
    if (hasTex && (e->dWave == nullptr || waveLen != e->waveLen)) {
        cudaFree(e->dWave);
        e->dWave = nullptr;
        JB_CUDA(cudaMalloc(&e->dWave, sizeof(float) * 2 * (size_t) waveLen * (size_t) e->clipPitch));
    }
    e->ringLen = ringLen;
    e->waveLen = waveLen;
    e->sampleRate = sample_rate;
    e->blockSize = samples_per_block;
    if (int rc = resetState(e))
        return rc;
    e->prepared = true;
    return JB_OK;
}

int jb_reset(jb_engine* e)
{
    if (int rc = checkEngine(e))
        return rc;
    if (!e->prepared)
        return fail(JB_ERR_STATE, "jb_reset before jb_prepare");
    if (int rc = setDevice(e))
        return rc;
    return resetState(e);
}

int jb_num_params(const jb_engine* e, int slot) { return checkSlot(e, slot) == JB_OK ? e->params[(size_t) slot].count() : JB_ERR_ARG; }

int jb_param_info_at(const jb_engine* e, int slot, int index, jb_param_info* out)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    const auto& specs = jb::paramSpecs(e->chain[(size_t) slot]);
    if (out == nullptr || index < 0 || index >= (int) specs.size())
        return fail(JB_ERR_ARG, "jb_param_info_at: bad index %d", index);
    const auto& s = specs[(size_t) index];
    out->id = s.id;
    out->name = s.name;
    out->min_value = s.lo;
    out->max_value = s.hi;
    out->interval = s.interval;
    out->default_value = s.def;
    out->is_output = s.isOutput ? 1 : 0;
    return JB_OK;
}

static int findParam(const jb_engine* e, int slot, const char* id, int* index)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    if (id == nullptr)
        return fail(JB_ERR_ARG, "null parameter id");
    const int idx = e->params[(size_t) slot].find(id);
    if (idx < 0)
        return fail(JB_ERR_ARG, "%s has no parameter \"%s\"", jb::kindName(e->chain[(size_t) slot]), id);
    *index = idx;
    return JB_OK;
}

int jb_get_param(const jb_engine* e, int slot, const char* id, float* out)
{
    int idx = -1;
    if (int rc = findParam(e, slot, id, &idx))
        return rc;
    if (out == nullptr)
        return fail(JB_ERR_ARG, "null output");
    *out = e->params[(size_t) slot].raw(idx);
    return JB_OK;
}

int jb_set_param(jb_engine* e, int slot, const char* id, float plain_value)
{
    int idx = -1;
    if (int rc = findParam(e, slot, id, &idx))
        return rc;
    e->params[(size_t) slot].setPlain(idx, plain_value);
    return JB_OK;
}

int jb_set_param_normalised(jb_engine* e, int slot, const char* id, float normalised)
{
    int idx = -1;
    if (int rc = findParam(e, slot, id, &idx))
        return rc;
    e->params[(size_t) slot].setNormalised(idx, normalised);
    return JB_OK;
}

int jb_num_programs(const jb_engine* e, int slot) { return checkSlot(e, slot) == JB_OK ? e->params[(size_t) slot].numPrograms() : JB_ERR_ARG; }
int jb_get_program(const jb_engine* e, int slot) { return checkSlot(e, slot) == JB_OK ? e->params[(size_t) slot].currentProgram() : JB_ERR_ARG; }

int jb_set_program(jb_engine* e, int slot, int index)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    e->params[(size_t) slot].setProgram(index);
    return JB_OK;
}

const char* jb_program_name(const jb_engine* e, int slot, int index)
{
    if (checkSlot(e, slot) != JB_OK)
        return "";
    return e->params[(size_t) slot].programName(index);
}

int jb_set_stream(jb_engine* e, void* cuda_stream)
{
    if (int rc = checkEngine(e))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    if (int rc = drainTiming(e))
        return rc;
    if (e->ownStream && e->stream) {
        cudaStreamSynchronize(e->stream);
        cudaStreamDestroy(e->stream);
    }
    e->stream = (cudaStream_t) cuda_stream;
    e->ownStream = false;
    return JB_OK;
}

int jb_synchronize(jb_engine* e)
{
    if (int rc = checkEngine(e))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    JB_CUDA(cudaStreamSynchronize(e->stream));
    return JB_OK;
}

int jb_process(jb_engine* e, const float* d_in, float* d_out, int n_samples)
{
    if (int rc = checkEngine(e))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    if (!e->prepared)
        return fail(JB_ERR_STATE, "jb_process before jb_prepare (prepareToPlay)");
    if (d_in == nullptr || d_out == nullptr)
        return fail(JB_ERR_ARG, "jb_process: null audio pointer");
    if (n_samples < 0)
        return fail(JB_ERR_ARG, "jb_process: negative n_samples");
    if (n_samples == 0)
        return JB_OK;
    ProcArgs a;
    buildArgs(e, a, d_in, d_out, n_samples, e->nClips, 0);
    if (int rc = launchProcess(e, a))
        return rc;
    e->blocksDone += (n_samples + e->blockSize - 1) / e->blockSize;
    return JB_OK;
}

int jb_process_host(jb_engine* e, const float* h_in, float* h_out, int n_samples)
{
    if (int rc = checkEngine(e))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    if (!e->prepared)
        return fail(JB_ERR_STATE, "jb_process_host before jb_prepare (prepareToPlay)");
    if (h_in == nullptr || h_out == nullptr)
        return fail(JB_ERR_ARG, "jb_process_host: null audio pointer");
    if (n_samples <= 0)
        return n_samples == 0 ? JB_OK : fail(JB_ERR_ARG, "jb_process_host: negative n_samples");

    // Time-sliced streaming.  The batch lives in one device buffer with the caller's layout
    // [clip][channel][sample]; the render is cut along TIME into slices of whole host blocks, and three
    // streams overlap slice i+1's upload, slice i's render and slice i-1's download (2-D copies: one row per
    // (clip, channel), the slice's samples wide).  Cutting along time keeps every launch as wide as the
    // whole batch (all SMs busy, the per-clip sequential recurrences only a slice long per launch) and is
    // exact: state carries across launches like across consecutive host callbacks.  Batches larger than the
    // pass budget are additionally cut into clip ranges ("passes"), two device buffers alternating.
    const size_t rowBytes = sizeof(float) * (size_t) n_samples;
    const size_t clipBytes = rowBytes * (size_t) e->nCh;
    auto envMiB = [](const char* name, size_t dflt) {
        const char* v = std::getenv(name);
        return (size_t) (v ? std::max(1, std::atoi(v)) : (int) dflt) << 20;
    };
    const size_t passBudget = envMiB("JB_HOST_PASS_MIB", 8192), sliceTarget = envMiB("JB_HOST_SLICE_MIB", 96);
    long long passClips = std::max<long long>(32, (long long) (passBudget / clipBytes) / 32 * 32);
    passClips = std::min<long long>(passClips, e->nClips);
    const int nPasses = (int) ((e->nClips + passClips - 1) / passClips);
    const int totalBlocks = (n_samples + e->blockSize - 1) / e->blockSize;
    const size_t blockBytesAllClips = sizeof(float) * (size_t) e->blockSize * (size_t) e->nCh * (size_t) passClips;
    int sliceBlocks = (int) std::max<size_t>(1, sliceTarget / std::max<size_t>(1, blockBytesAllClips));
    sliceBlocks = std::min(sliceBlocks, totalBlocks);
    const int nSlices = (totalBlocks + sliceBlocks - 1) / sliceBlocks;

    if (e->copyIn == nullptr) {
        JB_CUDA(cudaStreamCreateWithFlags(&e->copyIn, cudaStreamNonBlocking));
        JB_CUDA(cudaStreamCreateWithFlags(&e->copyOut, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i)
            JB_CUDA(cudaEventCreateWithFlags(&e->evOut[i], cudaEventDisableTiming));
    }
    const size_t need = clipBytes * (size_t) passClips;
    const int nBuffers = nPasses > 1 ? 2 : 1;
    for (int i = 0; i < nBuffers; ++i) {
        if (e->dStage[i] == nullptr || need > e->stageBytes) {
            cudaFree(e->dStage[i]);
            e->dStage[i] = nullptr;
            JB_CUDA(cudaMalloc(&e->dStage[i], need));
        }
    }
    e->stageBytes = std::max(e->stageBytes, need);
    while (e->sliceEvents.size() < (size_t) 2 * (size_t) nSlices) {
        cudaEvent_t ev = nullptr;
        JB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        e->sliceEvents.push_back(ev);
    }

    const long long blocksBase = e->blocksDone;
    int rcLaunch = JB_OK;
    for (int pass = 0; pass < nPasses && rcLaunch == JB_OK; ++pass) {
        const long long c0 = (long long) pass * passClips;
        const int nc = (int) std::min<long long>(passClips, e->nClips - c0);
        float* dBuf = e->dStage[pass % nBuffers];
        const size_t rows = (size_t) nc * (size_t) e->nCh;
        const float* hIn = h_in + (size_t) c0 * e->nCh * n_samples;
        float* hOut = h_out + (size_t) c0 * e->nCh * n_samples;
        if (pass >= nBuffers) // this buffer's previous pass has been downloaded
            JB_CUDA(cudaStreamWaitEvent(e->copyIn, e->evOut[pass % nBuffers], 0));
        for (int sl = 0; sl < nSlices; ++sl) {
            const int firstBlock = sl * sliceBlocks;
            const int t0 = firstBlock * e->blockSize;
            const int ns = std::min(n_samples - t0, sliceBlocks * e->blockSize);
            cudaEvent_t evIn = e->sliceEvents[(size_t) 2 * sl], evDone = e->sliceEvents[(size_t) 2 * sl + 1];
            JB_CUDA(cudaMemcpy2DAsync(dBuf + t0, rowBytes, hIn + t0, rowBytes, sizeof(float) * (size_t) ns, rows,
                                      cudaMemcpyHostToDevice, e->copyIn));
            JB_CUDA(cudaEventRecord(evIn, e->copyIn));
            JB_CUDA(cudaStreamWaitEvent(e->stream, evIn, 0));
            ProcArgs a;
            e->blocksDone = blocksBase + firstBlock; // history index of the slice's first block
            buildArgs(e, a, dBuf + t0, dBuf + t0, ns, nc, c0, n_samples);
            if ((rcLaunch = launchProcess(e, a)) != JB_OK)
                break;
            JB_CUDA(cudaEventRecord(evDone, e->stream));
            JB_CUDA(cudaStreamWaitEvent(e->copyOut, evDone, 0));
            JB_CUDA(cudaMemcpy2DAsync(hOut + t0, rowBytes, dBuf + t0, rowBytes, sizeof(float) * (size_t) ns, rows,
                                      cudaMemcpyDeviceToHost, e->copyOut));
        }
        JB_CUDA(cudaEventRecord(e->evOut[pass % nBuffers], e->copyOut));
        // (the slice events are re-recorded by the next pass; the waits above captured this pass's records)
    }
    JB_CUDA(cudaStreamSynchronize(e->copyIn));
    JB_CUDA(cudaStreamSynchronize(e->stream));
    JB_CUDA(cudaStreamSynchronize(e->copyOut));
    e->blocksDone = blocksBase;
    if (rcLaunch != JB_OK)
        return rcLaunch;
    e->blocksDone += totalBlocks;
    return JB_OK;
}

int jb_get_metrics(jb_engine* e, int slot, jb_metrics* out)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    if (out == nullptr)
        return fail(JB_ERR_ARG, "jb_get_metrics: null output");
    if (int rc = setDevice(e))
        return rc;
    if (!e->prepared)
        return fail(JB_ERR_STATE, "jb_get_metrics before jb_prepare");
    e->hostScratch.resize((size_t) JBK_REC * (size_t) e->clipPitch);
    JB_CUDA(cudaMemcpyAsync(e->hostScratch.data(), e->dLatest + (long long) slot * JBK_REC * e->clipPitch,
                            sizeof(float) * e->hostScratch.size(), cudaMemcpyDeviceToHost, e->stream));
    JB_CUDA(cudaStreamSynchronize(e->stream));
    unpackRecords(e->hostScratch.data(), e->clipPitch, e->nClips, out);
    return JB_OK;
}

int jb_metrics_device(jb_engine* e, int slot, const float** d_out)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    if (d_out == nullptr)
        return fail(JB_ERR_ARG, "jb_metrics_device: null output");
    if (!e->prepared || e->hostOnly)
        return fail(JB_ERR_STATE, "jb_metrics_device before jb_prepare");
    *d_out = e->dLatest + (long long) slot * JBK_REC * e->clipPitch;
    return JB_OK;
}

int jb_enable_history(jb_engine* e, int max_blocks)
{
    if (int rc = checkEngine(e))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    if (max_blocks < 0)
        return fail(JB_ERR_ARG, "jb_enable_history: negative max_blocks");
    JB_CUDA(cudaStreamSynchronize(e->stream));
    cudaFree(e->dHist);
    e->dHist = nullptr;
    e->histMaxBlocks = 0;
    if (max_blocks > 0) {
        const size_t bytes = sizeof(float) * (size_t) max_blocks * e->chain.size() * JBK_REC * (size_t) e->clipPitch;
        JB_CUDA(cudaMalloc(&e->dHist, bytes));
        JB_CUDA(cudaMemsetAsync(e->dHist, 0, bytes, e->stream));
        e->histMaxBlocks = max_blocks;
    }
    return JB_OK;
}

int jb_history_blocks(const jb_engine* e)
{
    if (e == nullptr)
        return 0;
    return (int) std::min<long long>(e->blocksDone, e->histMaxBlocks);
}

int jb_get_history(jb_engine* e, int slot, int first_block, int n_blocks, jb_metrics* out)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    if (e->dHist == nullptr)
        return fail(JB_ERR_STATE, "jb_get_history: history not enabled");
    if (out == nullptr || first_block < 0 || n_blocks < 0 || first_block + n_blocks > jb_history_blocks(e))
        return fail(JB_ERR_ARG, "jb_get_history: block range [%d, %d) outside the %d recorded blocks", first_block,
                    first_block + n_blocks, jb_history_blocks(e));
    e->hostScratch.resize((size_t) JBK_REC * (size_t) e->clipPitch);
    const long long chainLen = (long long) e->chain.size();
    for (int b = 0; b < n_blocks; ++b) {
        const float* src = e->dHist + (((long long) (first_block + b) * chainLen + slot) * JBK_REC) * e->clipPitch;
        JB_CUDA(cudaMemcpyAsync(e->hostScratch.data(), src, sizeof(float) * e->hostScratch.size(), cudaMemcpyDeviceToHost, e->stream));
        JB_CUDA(cudaStreamSynchronize(e->stream));
        unpackRecords(e->hostScratch.data(), e->clipPitch, e->nClips, out + (size_t) b * (size_t) e->nClips);
    }
    return JB_OK;
}

int jb_meter_statistics(jb_engine* e, int slot, int first_block, int n_blocks, int block_stride, jb_meter_stats* out)
{
    static_assert(sizeof(jb_meter_stats) == sizeof(float) * JBK_METER, "jb_meter_stats is 40 floats");
    if (int rc = checkSlot(e, slot))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    if (e->dHist == nullptr)
        return fail(JB_ERR_STATE, "jb_meter_statistics: history not enabled");
    if (out == nullptr || first_block < 0 || n_blocks < 0 || block_stride < 1 || first_block + n_blocks > jb_history_blocks(e))
        return fail(JB_ERR_ARG, "jb_meter_statistics: block range [%d, %d) stride %d outside the %d recorded blocks", first_block,
                    first_block + n_blocks, block_stride, jb_history_blocks(e));
    float* dOut = nullptr;
    const size_t count = (size_t) JBK_METER * (size_t) e->clipPitch;
    JB_CUDA(cudaMalloc(&dOut, sizeof(float) * count));
    int rc = JB_OK;
    if (jbk_launch_meter(e->dHist, e->clipPitch, (int) e->chain.size(), slot, first_block, n_blocks, block_stride, e->nClips, dOut,
                         e->stream) != 0)
        rc = fail(JB_ERR_CUDA, "jb_meter_statistics: kernel launch failed");
    if (rc == JB_OK) {
        e->hostScratch.resize(count);
        cudaError_t err = cudaMemcpyAsync(e->hostScratch.data(), dOut, sizeof(float) * count, cudaMemcpyDeviceToHost, e->stream);
        if (err == cudaSuccess)
            err = cudaStreamSynchronize(e->stream);
        if (err != cudaSuccess)
            rc = fail(JB_ERR_CUDA, "jb_meter_statistics: %s", cudaGetErrorString(err));
    }
    cudaFree(dOut);
    if (rc != JB_OK)
        return rc;
    float* o = reinterpret_cast<float*>(out);
    for (int c = 0; c < e->nClips; ++c)
        for (int f = 0; f < JBK_METER; ++f)
            o[(size_t) c * JBK_METER + f] = e->hostScratch[(size_t) f * (size_t) e->clipPitch + c];
    return JB_OK;
}

int jb_set_path(jb_engine* e, int mode)
{
    if (int rc = checkEngine(e))
        return rc;
    if (mode < JB_PATH_AUTO || mode > JB_PATH_COOP)
        return fail(JB_ERR_ARG, "jb_set_path: unknown mode %d", mode);
    e->pathMode = mode;
    return JB_OK;
}

int jb_path_launches(const jb_engine* e, long long* cooperative, long long* lane_per_clip)
{
    if (int rc = checkEngine(e))
        return rc;
    if (cooperative)
        *cooperative = e->coopLaunches;
    if (lane_per_clip)
        *lane_per_clip = e->laneLaunches;
    return JB_OK;
}

int jb_kernel_time_ms(jb_engine* e, double* ms, long long* launches)
{
    if (int rc = checkEngine(e))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    if (int rc = drainTiming(e))
        return rc;
    if (ms)
        *ms = e->kernelMs;
    if (launches)
        *launches = e->kernelLaunches;
    e->kernelMs = 0.0;
    e->kernelLaunches = 0;
    return JB_OK;
}

int jb_host_alloc(size_t bytes, void** out)
{
    if (out == nullptr)
        return fail(JB_ERR_ARG, "jb_host_alloc: null output");
    *out = nullptr;
    if (jb_device_count() <= 0)
        return fail(JB_ERR_CUDA, "jb_host_alloc: no CUDA device available");
    JB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return JB_OK;
}

int jb_host_free(void* p)
{
    if (p != nullptr)
        JB_CUDA(cudaFreeHost(p));
    return JB_OK;
}

int jb_device_alloc(int device, size_t bytes, void** out)
{
    if (out == nullptr)
        return fail(JB_ERR_ARG, "jb_device_alloc: null output");
    *out = nullptr;
    if (jb_device_count() <= 0)
        return fail(JB_ERR_CUDA, "jb_device_alloc: no CUDA device available");
    JB_CUDA(cudaSetDevice(device));
    JB_CUDA(cudaMalloc(out, bytes));
    return JB_OK;
}

int jb_device_free(int device, void* p)
{
    if (p == nullptr)
        return JB_OK;
    JB_CUDA(cudaSetDevice(device));
    JB_CUDA(cudaFree(p));
    return JB_OK;
}

int jb_copy_to_device(int device, void* d_dst, const void* h_src, size_t bytes)
{
    if (d_dst == nullptr || h_src == nullptr)
        return fail(JB_ERR_ARG, "jb_copy_to_device: null pointer");
    JB_CUDA(cudaSetDevice(device));
    JB_CUDA(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
    return JB_OK;
}

int jb_copy_to_host(int device, void* h_dst, const void* d_src, size_t bytes)
{
    if (h_dst == nullptr || d_src == nullptr)
        return fail(JB_ERR_ARG, "jb_copy_to_host: null pointer");
    JB_CUDA(cudaSetDevice(device));
    JB_CUDA(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return JB_OK;
}

int jb_synth_fill_host(float* h_audio, int kind, long long first_clip, int n_clips, int n_channels, int n_samples,
                       double sample_rate, unsigned int seed)
{
    if (h_audio == nullptr || kind < 0 || kind > 4 || n_clips < 0 || n_samples < 0 || n_channels < 1 || n_channels > 2)
        return fail(JB_ERR_ARG, "jb_synth_fill_host: bad argument");
    jb::synthFillHost(h_audio, kind, first_clip, n_clips, n_channels, n_samples, sample_rate, seed);
    return JB_OK;
}

int jb_synth_fill(float* d_audio, int kind, long long first_clip, int n_clips, int n_channels, int n_samples,
                  double sample_rate, unsigned int seed, int device, void* cuda_stream)
{
    if (d_audio == nullptr || kind < 0 || kind > 4 || n_clips < 0 || n_samples < 0 || n_channels < 1 || n_channels > 2)
        return fail(JB_ERR_ARG, "jb_synth_fill: bad argument");
    if (jb_device_count() <= 0)
        return fail(JB_ERR_CUDA, "jb_synth_fill: no CUDA device available");
    JB_CUDA(cudaSetDevice(device));
    if (jbk_launch_synth(d_audio, kind, first_clip, n_clips, n_channels, n_samples, sample_rate, seed, cuda_stream) != 0)
        return fail(JB_ERR_CUDA, "%s", jbk_last_cuda_error());
    return JB_OK;
}

} // extern "C"
