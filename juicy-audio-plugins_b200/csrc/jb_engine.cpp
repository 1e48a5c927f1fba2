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
#include "jb_partition.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include <algorithm>
#include <chrono>
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

// NVTX range around a C-ABI call (tracing), closed on every exit path
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

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
    int mathMode = 0;              // 0 auto (exact where a Texture follows a Saturator / Punch), 1 exact, 2 fast
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
    int16_t* dPcm[3] = { nullptr, nullptr, nullptr }; // 16-bit images of the staging buffers (jb_process_host_pcm16)
    size_t pcmBytes[3] = { 0, 0, 0 };
    std::vector<cudaEvent_t> sliceEvents;
    cudaEvent_t evIn[3] = {}, evDone[3] = {}, evOut[3] = {};
    size_t stageBytes[3] = { 0, 0, 0 }; // capacity of each staging buffer (they grow independently)

    std::vector<float> hostScratch;

    // Per-clip parameters.  `params` is parameter set 0; clips given other values use a set of
    // `variants` (set k > 0 = variants[k - 1]); clipSet[clip] = set index (empty: every clip uses set 0).  Clips of one
    // set render in one launch: a contiguous clip range directly, a scattered set through a device clip map.
    struct Group { int set; int first; int count; long long mapOffset; }; // first < 0: mapped, mapOffset into clipMapHost
    std::vector<std::vector<jb::ParamSet>> variants;
    std::vector<int> clipSet;
    std::vector<Group> groups;
    std::vector<int> clipMapHost;
    int* dClipMap = nullptr;
    bool groupsDirty = false;

    // Per-block automation: parameter changes taking effect at the start of an absolute block index
    struct AutoEvent { long long block; int slot; int index; float value; int first; int count; };
    std::vector<AutoEvent> schedule; // kept sorted by block (stable)
    static constexpr int kGroupStreams = 8;
    cudaStream_t groupStream[kGroupStreams] = {};
    cudaEvent_t groupJoin[kGroupStreams] = {};
    cudaEvent_t groupFork = nullptr;
    std::vector<cudaEvent_t> pipeEvents; // [plugin][segment] of a pipelined chain render
    jb::SmPartitions partitions;         // disjoint SM groups (green contexts) for parameter sets rendered side by side
    int partitionsFor = 0;               // number of sets the partitions were made (or found unavailable) for
    int pipelineMaxClips = 16384;        // chains of at most this many clips are pipelined across plugins

    // Score gather across the GPUs of one box: an NCCL communicator, opaque here (jb_comm_* below)
    void* comm = nullptr;
    int commRanks = 0, commRank = -1;
    float* dGather = nullptr;      // [commRanks][JBK_REC][clipPitch] scratch of jb_gather_records_host
    size_t gatherBytes = 0;

    // CUDA-event pairs around every render-kernel launch (jb_kernel_time_ms)
    std::vector<cudaEvent_t> timingEvents; // start0, stop0, start1, stop1, ...
    size_t timingUsed = 0;
    double kernelMs = 0.0;
    long long kernelLaunches = 0;

    // Opt-in per-plugin timing of a chain rendered plugin by plugin (jb_enable_slot_timing / jb_slot_time_ms): one event
    // before the first launch and one after each plugin's launch, on the render's stream
    bool slotTiming = false;
    std::vector<cudaEvent_t> slotEvents;   // pool
    size_t slotEventsUsed = 0;
    struct SlotMark { int slot; size_t evStart, evStop; };
    std::vector<SlotMark> slotMarks;
    double slotMs[JBK_MAX_CHAIN] = {};
    long long slotLaunches[JBK_MAX_CHAIN] = {};
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
    cudaFree(e->dClipMap);
    e->dClipMap = nullptr;
    cudaFree(e->dGather);
    e->dGather = nullptr;
    e->gatherBytes = 0;
    for (int i = 0; i < jb_engine::kGroupStreams; ++i) {
        if (e->groupStream[i]) cudaStreamDestroy(e->groupStream[i]);
        if (e->groupJoin[i]) cudaEventDestroy(e->groupJoin[i]);
        e->groupStream[i] = nullptr;
        e->groupJoin[i] = nullptr;
    }
    if (e->groupFork) cudaEventDestroy(e->groupFork);
    e->groupFork = nullptr;
    e->partitions.release();
    e->partitionsFor = 0;
    for (cudaEvent_t ev : e->pipeEvents)
        cudaEventDestroy(ev);
    e->pipeEvents.clear();
    e->groupsDirty = true;
    e->dState = e->dLatest = e->dHist = e->dRing = e->dWave = e->dCoopScratch = nullptr;
    for (int i = 0; i < 3; ++i) {
        cudaFree(e->dStage[i]);
        e->dStage[i] = nullptr;
        cudaFree(e->dPcm[i]);
        e->dPcm[i] = nullptr;
        e->pcmBytes[i] = 0;
        if (e->evIn[i]) cudaEventDestroy(e->evIn[i]);
        if (e->evDone[i]) cudaEventDestroy(e->evDone[i]);
        if (e->evOut[i]) cudaEventDestroy(e->evOut[i]);
        e->evIn[i] = e->evDone[i] = e->evOut[i] = nullptr;
    }
    e->stageBytes[0] = e->stageBytes[1] = e->stageBytes[2] = 0;
    if (e->copyIn) cudaStreamDestroy(e->copyIn);
    if (e->copyOut) cudaStreamDestroy(e->copyOut);
    e->copyIn = e->copyOut = nullptr;
    for (cudaEvent_t ev : e->timingEvents)
        cudaEventDestroy(ev);
    e->timingEvents.clear();
    for (cudaEvent_t ev : e->slotEvents)
        cudaEventDestroy(ev);
    e->slotEvents.clear();
    e->slotMarks.clear();
    e->slotEventsUsed = 0;
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
    const bool constructed = e->stateConstructed;
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
    if (!constructed) {
        // The latest* mailboxes are constructor-initialised members (monoSafety = 1, the rest 0: e.g.
        // JuicyPunch/PluginProcessor.h:44-51); prepareToPlay never touches them, so getLatestMetrics() after a re-prepare
        // still returns the previous block's values.
        JB_CUDA(cudaMemsetAsync(e->dLatest, 0, sizeof(float) * e->chain.size() * JBK_REC * (size_t) e->clipPitch, e->stream));
        for (size_t s = 0; s < e->chain.size(); ++s)
            if (jbk_launch_fill(e->dLatest + ((long long) s * JBK_REC + 12) * e->clipPitch, 1.0f, e->clipPitch, e->stream) != 0)
                return fail(JB_ERR_CUDA, "%s", jbk_last_cuda_error());
    }
    if (e->dHist)
        JB_CUDA(cudaMemsetAsync(e->dHist, 0, sizeof(float) * (size_t) e->histMaxBlocks * e->chain.size() * JBK_REC * (size_t) e->clipPitch, e->stream));
    e->blocksDone = 0;
    e->schedule.clear(); // scheduled changes are addressed by block index since this prepare / reset
    return JB_OK;
}

// Sample-streaming mode of one plugin's lane launch: 0 = four samples per trip through the lane's cp.async ring; 1 = eight,
// 32-byte stores, L1-cached row pieces (light plugins, >= 32768 clips: L2 sector throughput); 2 = warp-transposed tile
// streaming (jb_lane.cuh).  Measured (profiles/r01_s6_tile.txt): the tile wins where nothing is written back -- Infer,
// 16384 clips 7.0 -> 5.6 ms, 65536 clips out of place 19.3 -> 16.5 ms -- and loses for the plugins that store every sample
// (Saturator 65536 clips 20.0 -> 23.8 ms), so only Infer takes it.  JB_TILE=0 / 1 forces it off / on for light plugins.
int octetsFor(int kind, int nClips, int nSamples, bool mapped, bool exactMath)
{
    static const int tileMode = [] { const char* v = std::getenv("JB_TILE"); return v == nullptr ? -1 : std::atoi(v); }();
    // 3 = tile streaming through TMA (cp.async.bulk.tensor.3d, jb_lane.cuh): the warp's rows arrive without touching the
    // LSU / L1TEX path, which bounds the light plugins on big batches.  It does what it was built for -- l1tex__throughput
    // 82 % -> 25 % on JuicyCohere, 65536 clips -- and is still SLOWER there (20.9 -> 24 - 27 ms; 32768 clips 11.2 -> 16.7 ms;
    // only at <= 16384 clips 8.5 -> 7.9 ms): with one row per clip a tile is 32 separate 64 / 128-byte row fetches, and the
    // copy engine of an SM retires about one box row per 12 cycles whatever its length (the time follows rows / SM x 12
    // cycles at every batch size and both tile widths: profiles/r02_tma.txt).  So it stays opt-in: JB_TMA=1 turns it on for
    // light plugins, JB_TMA_MIN_CLIPS sets the smallest batch.
    static const int tmaMode = [] { const char* v = std::getenv("JB_TMA"); return v == nullptr ? 0 : std::atoi(v); }();
    static const int tmaMinClips = [] { const char* v = std::getenv("JB_TMA_MIN_CLIPS"); return v == nullptr ? 0 : std::atoi(v); }();
    if (mapped || kind == jb::kTexture || kind == jb::kMotion || ((kind == jb::kPunch || kind == jb::kSaturator) && exactMath))
        return 0; // the exact routines' code: four samples per trip is what fits the instruction cache (jb_kernels.cu)
    const bool tileOk = nClips % 32 == 0 && nSamples % 4 == 0;
    const bool tma = tmaMode != 0 && nClips >= tmaMinClips;
    if (tileOk && tma && !(tileMode > 0))
        return 3;
    // Round 2: with the tile loop's body rolled to eight samples per trip (jb_lane.cuh, JB_TILE_UNROLL = 1: the 16-sample body
    // of a writer was 42 KB of code, past the SM's 32 KB instruction cache) the tile also wins for the plugins that store
    // every sample from 16384 clips up: Saturator 65536 clips 20.1 -> 18.3 ms, Punch (fast) 23.1 -> 22.5, Width 28.4 -> 28.0,
    // Cohere 16384 clips 8.5 -> 7.7 ms (profiles/r02_tile_rolled.txt).
    static const int forced = [] { const char* v = std::getenv("JB_OCTETS"); return v == nullptr ? -1 : std::atoi(v); }();
    if (forced >= 0 && (forced < 2 || tileOk))
        return forced; // experiments: force one streaming mode
    const bool tile = tileMode < 0 ? nClips >= (kind == jb::kInfer ? 8192 : 16384) : tileMode != 0;
    if (tileOk && tile)
        return 2;
    // 4 = 32-byte loads into registers two octets ahead (falls back to 1 inside the kernel where a row is not 32-byte
    // aligned): half the L1TEX wavefronts of the cp.async rings -- and slower all the same (JuicyCohere 65536 clips 21.2 ->
    // 27.1 ms, 32768 clips 11.3 -> 16.4; Saturator 20.2 -> 22.2; profiles/r02_tma.txt): a warp issues in order, so the first
    // use of a register still in flight stalls everything behind it, where a cp.async ring only ever blocks at its
    // wait_group.  Opt-in: JB_LDG256=1.
    static const int ldgMode = [] { const char* v = std::getenv("JB_LDG256"); return v == nullptr ? 0 : std::atoi(v); }();
    return nClips >= 32768 ? (ldgMode ? 4 : 1) : 0;
}

int buildArgs(jb_engine* e, ProcArgs& a, const std::vector<jb::ParamSet>& params, const float* dIn, float* dOut, int nSamples,
              int nClips, long long clipOffset, long long rowPitch = 0)
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
    a.laneOnly = e->pathMode == 1 ? 1 : 0;
    a.octets = 0; // single-plugin engines: set below, once the math mode of this launch is known
    if (e->chain.size() > 1) { // fused generic kernel: eight samples per trip only when every plugin is light
        a.octets = nClips >= 32768 ? 1 : 0;
        for (int k : e->chain)
            if (k == jb::kPunch || k == jb::kTexture || k == jb::kMotion)
                a.octets = 0;
    }
    {   // Saturator / Punch transcendentals (JB_MATH_AUTO).  The MUFU-based tanh / pow are within 3e-6 of the reference --
        // inside the sample tolerance for the shaper's own output, but not once another plugin consumes that output:
        //   * Texture's metal / wood / plastic resonators amplify an input difference ~200x (profiles/r01_s6_chain_sensitivity.txt);
        //   * Width's `if (corrProxy < -0.1f) width *= dynamicLimit` (JuicyWidth/PluginProcessor.cpp:109-112), Motion's onset
        //     detector (JuicyMotion/PluginProcessor.cpp:75-95) and every later analyzer's onset threshold
        //     (JuicinessAnalyzer.cpp:69-75) are DISCONTINUOUS in their input: a last-bit difference flips a decision once in
        //     ~10^8 samples, and a flipped Width decision moves the rest of the block by ~2e-2 of peak.  Measured at full
        //     population (profiles/r02_parity_population.json): Saturator -> Width, fast math, 3 of 8192 mixed clips out of
        //     tolerance; the 7-plugin chain, 5 of 32768; with the exact routines 0 of 32768 (bit-identical samples).
        // So: a shaper with ANY plugin behind it runs the C library's own algorithms (jb_libm.h) and feeds its successor the
        // reference's bits; a shaper that ends the chain (or stands alone) keeps the fast routines.  Decided per launch.
        bool shaperFeedsPlugin = false;
        for (size_t s2 = 0; s2 + 1 < e->chain.size(); ++s2)
            if (e->chain[s2] == jb::kPunch || e->chain[s2] == jb::kSaturator)
                shaperFeedsPlugin = true;
        a.exactMath = e->mathMode == 1 || (e->mathMode == 0 && shaperFeedsPlugin) ? 1 : 0;
    }
    if (e->chain.size() == 1)
        a.octets = octetsFor(e->chain[0], nClips, nSamples, false, a.exactMath != 0);
    if (e->nCh == 1) // mono buses run the generic kernel's one-channel instantiations (either math mode), four samples per trip
        a.octets = 0;
    a.ana = jb::makeAnaCoef(e->sampleRate);
    for (size_t s = 0; s < e->chain.size(); ++s) {
        a.slot[s].kind = e->chain[s];
        a.slot[s].stateBase = e->stateBase[s];
        jb::makeSlotCoef(params[s], e->sampleRate, &a.slot[s].c);
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

// Timing-event pair for one render on the engine's stream (jb_kernel_time_ms).  Never blocks: when the pool is in use,
// pairs whose stop event has already completed are folded into the running total (cudaEventQuery) and their events
// recycled; if the device is further behind than that, the pool grows instead of waiting for it (jb_process stays
// asynchronous and jb_process_host's upload / render / download overlap is not stalled).
int timingPair(jb_engine* e, cudaEvent_t* start, cudaEvent_t* stop)
{
    if (e->timingUsed + 2 > e->timingEvents.size() && e->timingUsed >= 64) {
        size_t done = 0;
        while (done + 1 < e->timingUsed && cudaEventQuery(e->timingEvents[done + 1]) == cudaSuccess) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, e->timingEvents[done], e->timingEvents[done + 1]) == cudaSuccess) {
                e->kernelMs += (double) ms;
                ++e->kernelLaunches;
            }
            done += 2;
        }
        cudaGetLastError(); // cudaErrorNotReady from the query is not an error of ours
        if (done > 0) { // recycle: completed pairs go to the back of the pool
            std::rotate(e->timingEvents.begin(), e->timingEvents.begin() + (long) done, e->timingEvents.begin() + (long) e->timingUsed);
            e->timingUsed -= done;
        }
    }
    while (e->timingUsed + 2 > e->timingEvents.size()) {
        cudaEvent_t ev = nullptr;
        JB_CUDA(cudaEventCreate(&ev));
        e->timingEvents.push_back(ev);
    }
    *start = e->timingEvents[e->timingUsed];
    *stop = e->timingEvents[e->timingUsed + 1];
    return JB_OK;
}

// Side streams for launches that may overlap (parameter sets of one call; the plugins of a pipelined chain)
int ensureGroupStreams(jb_engine* e)
{
    if (e->groupStream[0] != nullptr)
        return JB_OK;
    for (int i = 0; i < jb_engine::kGroupStreams; ++i) {
        JB_CUDA(cudaStreamCreateWithFlags(&e->groupStream[i], cudaStreamNonBlocking));
        JB_CUDA(cudaEventCreateWithFlags(&e->groupJoin[i], cudaEventDisableTiming));
    }
    JB_CUDA(cudaEventCreateWithFlags(&e->groupFork, cudaEventDisableTiming));
    return JB_OK;
}

// The render kernel(s) of one ProcArgs on `stream`, untimed.  allowCoop: the cooperative kernel may be chosen (it owns
// engine-wide scratch, so concurrent launches of several parameter sets stay on the lane kernels).
int launchKernels(jb_engine* e, const ProcArgs& a, cudaStream_t stream, bool allowCoop)
{
    // Path choice: the cooperative (time-parallel) kernel when the chain and the call's shape allow it and
    // the batch is too small to fill the GPU with one lane per clip; the lane-per-clip kernels otherwise.
    bool coop = allowCoop && a.nCh == 2 && e->coopCapable && e->dCoopScratch != nullptr && jbk_coop_supported(&a) != 0;
    if (e->pathMode == 1)
        coop = false;
    else if (e->pathMode == 0 && coop && jbk_solo_pick(&a))
        coop = false; // a few clips of one plugin: the clip-per-CTA kernel is ahead of the cooperative one (profiles/r02_solo.txt)
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
    if (coop) {
        if (jbk_launch_coop(&a, e->dCoopScratch, e->numSMs, stream) != 0)
            return fail(JB_ERR_CUDA, "%s", jbk_coop_last_error());
        ++e->coopLaunches;
        return JB_OK;
    }
    // Lane kernels, chain of several plugins: one launch per plugin over the whole call (plugin s + 1 only needs
    // plugin s's output of the same block, and every plugin blocks identically, so plugin-by-plugin equals
    // block-by-block).  Each launch is that plugin's own kernel (jb_single_kernel: coefficients as constant-bank
    // operands, a few thousand instructions, its own register budget); the fused generic kernel spills and re-reads
    // its coefficients per sample.  Measured (profiles/r01_s6_survey_single.txt): 7-plugin chain 32768 clips
    // 424 (fused) -> 125 ms, 4096 clips 83 -> 68 ms.  JB_LANE_SPLIT=0 forces the fused kernel (tests keep it exact).
    const char* splitEnv = std::getenv("JB_LANE_SPLIT");
    const bool splitChains = a.exactMath != 0 || splitEnv == nullptr || std::atoi(splitEnv) != 0;
    const int L = a.chainLen;
    if (L > 1 && splitChains && a.nCh == 2) {
        auto launchOne = [&](int s, int t0, int ns, int firstBlock, cudaStream_t st) -> int {
            ProcArgs one = a;
            one.in = (s == 0 ? a.in : a.out) + t0;
            one.out = a.out + t0;
            one.nSamples = ns;
            one.histFirstBlock = a.histFirstBlock + firstBlock;
            one.chainLen = 1;
            one.slot[0] = a.slot[s];
            one.recSlotBase = s;
            one.recChainLen = L;
            one.octets = octetsFor(one.slot[0].kind, a.nClips, ns, a.clipMap != nullptr, a.exactMath != 0);
            if (jbk_launch_process(&one, st) != 0)
                return fail(JB_ERR_CUDA, "%s", jbk_last_cuda_error());
            return JB_OK;
        };
        // Plugin pipeline across streams.  One launch per plugin over the whole call leaves a small batch's GPU mostly
        // idle: each launch is a latency-bound walk through time on a few hundred warps, and the L launches run one
        // after the other.  Cut the call into segments of whole host blocks instead and give every plugin its own
        // stream: plugin s renders segment k as soon as plugin s - 1 has (an event) and it has finished segment k - 1
        // itself (stream order), so up to L kernels run side by side on different stretches of time -- the same
        // arithmetic in the same order per clip, state carried across launches as between host callbacks.
        // Measured: profiles/r01_s6_pipeline.txt.  JB_CHAIN_PIPELINE=0 / 1 forces it off / on.
        static const int pipeMode = [] { const char* v = std::getenv("JB_CHAIN_PIPELINE"); return v == nullptr ? -1 : std::atoi(v); }();
        const int B = a.blockSize;
        const int nBlocks = (a.nSamples + B - 1) / B;
        const bool onEngineStream = stream == e->stream && L <= jb_engine::kGroupStreams;
        const bool pipeline = onEngineStream && nBlocks >= 4 && (pipeMode < 0 ? a.nClips <= e->pipelineMaxClips : pipeMode != 0);
        if (!pipeline) {
            // opt-in: an event between the plugins' launches (bench.py's per-kernel roofline of a chain)
            const bool marks = e->slotTiming && e->slotMarks.size() < 65536;
            size_t prev = 0;
            auto mark = [&](size_t* index) -> int {
                if (e->slotEventsUsed == e->slotEvents.size()) {
                    cudaEvent_t ev = nullptr;
                    JB_CUDA(cudaEventCreate(&ev));
                    e->slotEvents.push_back(ev);
                }
                *index = e->slotEventsUsed++;
                JB_CUDA(cudaEventRecord(e->slotEvents[*index], stream));
                return JB_OK;
            };
            if (marks)
                if (int rc = mark(&prev))
                    return rc;
            for (int s = 0; s < L; ++s) {
                if (int rc = launchOne(s, 0, a.nSamples, 0, stream))
                    return rc;
                if (marks) {
                    size_t now = 0;
                    if (int rc = mark(&now))
                        return rc;
                    e->slotMarks.push_back({ s, prev, now });
                    prev = now;
                }
            }
        } else {
            static const int segEnv = [] { const char* v = std::getenv("JB_PIPE_SEGMENTS"); return v == nullptr ? 32 : std::max(2, std::atoi(v)); }();
            const int K = std::min(segEnv, nBlocks / 2);
            const int segBlocks = (nBlocks + K - 1) / K;
            if (int rc = ensureGroupStreams(e))
                return rc;
            while (e->pipeEvents.size() < (size_t) L * (size_t) K) {
                cudaEvent_t ev = nullptr;
                JB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                e->pipeEvents.push_back(ev);
            }
            // The L kernels are all different code.  Giving each plugin its own group of SMs (jb_partition.cpp), which doubles
            // the speed of five Texture sets side by side, does NOT pay here: seven groups of 16 SMs leave 36 SMs idle and
            // the plugins' costs are unequal (4096 clips 36.0 -> 42.0 ms, 16384 clips 82.8 -> 154 ms).  Opt-in: JB_PIPE_PARTITIONS=1.
            static const bool pipeParts = [] { const char* v = std::getenv("JB_PIPE_PARTITIONS"); return v != nullptr && std::atoi(v) != 0; }();
            if (pipeParts && e->partitionsFor != L) {
                e->partitions.create(e->device, L);
                e->partitionsFor = L;
            }
            const bool parted = pipeParts && e->partitions.count == L;
            auto pipeStream = [&](int s) { return parted ? static_cast<cudaStream_t>(e->partitions.streams[s]) : e->groupStream[s]; };
            JB_CUDA(cudaEventRecord(e->groupFork, stream));
            for (int s = 0; s < L; ++s)
                JB_CUDA(cudaStreamWaitEvent(pipeStream(s), e->groupFork, 0));
            for (int k = 0; k * segBlocks < nBlocks; ++k) {
                const int b0 = k * segBlocks;
                const int t0 = b0 * B;
                const int ns = std::min(a.nSamples - t0, segBlocks * B);
                for (int s = 0; s < L; ++s) {
                    cudaStream_t st = pipeStream(s);
                    if (s > 0)
                        JB_CUDA(cudaStreamWaitEvent(st, e->pipeEvents[(size_t) (s - 1) * K + k], 0));
                    if (int rc = launchOne(s, t0, ns, b0, st))
                        return rc;
                    JB_CUDA(cudaEventRecord(e->pipeEvents[(size_t) s * K + k], st));
                }
            }
            for (int s = 0; s < L; ++s) {
                JB_CUDA(cudaEventRecord(e->groupJoin[s], pipeStream(s)));
                JB_CUDA(cudaStreamWaitEvent(stream, e->groupJoin[s], 0));
            }
        }
    } else if (jbk_launch_process(&a, stream) != 0) {
        return fail(JB_ERR_CUDA, "%s", jbk_last_cuda_error());
    }
    ++e->laneLaunches;
    return JB_OK;
}

// One render bracketed by timing events on the engine's stream.
int launchProcess(jb_engine* e, const ProcArgs& a)
{
    cudaEvent_t start = nullptr, stop = nullptr;
    if (int rc = timingPair(e, &start, &stop))
        return rc;
    JB_CUDA(cudaEventRecord(start, e->stream));
    if (int rc = launchKernels(e, a, e->stream, true))
        return rc;
    JB_CUDA(cudaEventRecord(stop, e->stream));
    e->timingUsed += 2;
    return JB_OK;
}

const std::vector<jb::ParamSet>& paramsOfSet(const jb_engine* e, int set)
{
    return set == 0 ? e->params : e->variants[(size_t) set - 1];
}

// Merge identical parameter sets, drop unused ones, fall back to the single-set state when every clip uses set 0.
void normalizeSets(jb_engine* e)
{
    e->groupsDirty = true;
    if (e->clipSet.empty()) {
        e->variants.clear();
        return;
    }
    const int nSets = 1 + (int) e->variants.size();
    std::vector<int> canon((size_t) nSets);
    for (int k = 0; k < nSets; ++k) {
        canon[(size_t) k] = k;
        for (int j = 0; j < k; ++j)
            if (canon[(size_t) j] == j && paramsOfSet(e, j) == paramsOfSet(e, k)) {
                canon[(size_t) k] = j;
                break;
            }
    }
    std::vector<char> used((size_t) nSets, 0);
    for (int& c : e->clipSet) {
        c = canon[(size_t) c];
        used[(size_t) c] = 1;
    }
    std::vector<int> renum((size_t) nSets, 0);
    std::vector<std::vector<jb::ParamSet>> kept;
    for (int k = 1; k < nSets; ++k)
        if (used[(size_t) k]) {
            kept.push_back(std::move(e->variants[(size_t) k - 1]));
            renum[(size_t) k] = (int) kept.size();
        }
    e->variants = std::move(kept);
    bool anyVariant = false;
    for (int& c : e->clipSet) {
        c = renum[(size_t) c];
        anyVariant = anyVariant || c != 0;
    }
    if (!anyVariant)
        e->clipSet.clear();
}

// Apply `change` to the parameters of clips [first, first + count) (count < 0: every clip, set 0 included).
template <class F>
void changeParams(jb_engine* e, int first, int count, F change)
{
    if (count < 0 || (first == 0 && count == e->nClips)) {
        change(e->params);
        for (auto& v : e->variants)
            change(v);
        normalizeSets(e);
        return;
    }
    if (e->clipSet.empty())
        e->clipSet.assign((size_t) e->nClips, 0);
    std::vector<int> newOf(1 + e->variants.size(), -1);
    for (int c = first; c < first + count; ++c) {
        const int old = e->clipSet[(size_t) c];
        if (newOf[(size_t) old] < 0) {
            std::vector<jb::ParamSet> copy = paramsOfSet(e, old);
            change(copy);
            e->variants.push_back(std::move(copy));
            newOf[(size_t) old] = (int) e->variants.size();
        }
        e->clipSet[(size_t) c] = newOf[(size_t) old];
    }
    normalizeSets(e);
}

int rebuildGroups(jb_engine* e)
{
    e->groups.clear();
    e->clipMapHost.clear();
    e->groupsDirty = false;
    if (e->clipSet.empty())
        return JB_OK;
    const int nSets = 1 + (int) e->variants.size();
    std::vector<std::vector<int>> members((size_t) nSets);
    for (int c = 0; c < e->nClips; ++c)
        members[(size_t) e->clipSet[(size_t) c]].push_back(c);
    for (int k = 0; k < nSets; ++k) {
        const std::vector<int>& m = members[(size_t) k];
        if (m.empty())
            continue;
        if (m.back() - m.front() + 1 == (int) m.size()) {
            e->groups.push_back({ k, m.front(), (int) m.size(), -1 });
        } else {
            e->groups.push_back({ k, -1, (int) m.size(), (long long) e->clipMapHost.size() });
            e->clipMapHost.insert(e->clipMapHost.end(), m.begin(), m.end());
        }
    }
    if (!e->clipMapHost.empty()) {
        if (e->dClipMap == nullptr)
            JB_CUDA(cudaMalloc(&e->dClipMap, sizeof(int) * (size_t) e->nClips));
        JB_CUDA(cudaMemcpyAsync(e->dClipMap, e->clipMapHost.data(), sizeof(int) * e->clipMapHost.size(), cudaMemcpyHostToDevice, e->stream));
        JB_CUDA(cudaStreamSynchronize(e->stream)); // clipMapHost may change before the copy would otherwise run
    }
    return JB_OK;
}

// Render ns samples of clips [c0, c0 + nc); dIn / dOut point at clip c0's first sample of the range.  One launch per
// parameter set present in the range.
int renderClips(jb_engine* e, const float* dIn, float* dOut, int ns, int nc, long long c0, long long rowPitch)
{
    if (e->clipSet.empty()) {
        ProcArgs a;
        buildArgs(e, a, e->params, dIn, dOut, ns, nc, c0, rowPitch);
        return launchProcess(e, a);
    }
    if (e->groupsDirty)
        if (int rc = rebuildGroups(e))
            return rc;
    // The sets' launches touch disjoint clips, and each is usually far too small to fill the GPU (a latency-bound walk
    // through time), so they go out on a small pool of side streams between a fork and a join on the engine's stream.
    const long long clipStride = (long long) e->nCh * rowPitch;
    constexpr int kPool = jb_engine::kGroupStreams;
    if (int rc = ensureGroupStreams(e))
        return rc;
    static const bool forceSerial = [] { const char* v = std::getenv("JB_GROUP_SERIAL"); return v != nullptr && std::atoi(v) != 0; }();
    // ... as long as all of them fit the GPU together (the heavy kernels hold at most 8 warps per SM): beyond that the
    // launches queue behind each other's tails and one after the other is faster (32768 clips in 5 Texture materials:
    // 141 ms serial, 192 ms concurrent; 8192 clips: 133 ms serial, 80 ms concurrent -- profiles/r01_s6_survey_single.txt)
    const long long warpsInCall = (nc + 31) / 32 + (long long) e->groups.size();
    const bool serial = forceSerial || e->pathMode == 2 || e->groups.size() < 2 || warpsInCall > (long long) e->numSMs * 6;
    cudaEvent_t start = nullptr, stop = nullptr;
    if (int rc = timingPair(e, &start, &stop))
        return rc;
    JB_CUDA(cudaEventRecord(start, e->stream));
    if (!serial)
        JB_CUDA(cudaEventRecord(e->groupFork, e->stream));
    bool usedStream[kPool] = {};
    int next = 0;
    // Three or more Texture sets side by side (BASELINE config 3: material = clip mod 5): every set gets its own stream and
    // the small-code form of the two-lanes-per-clip kernel.  Unrolled, the kernels evict each other's loops from the SMs'
    // instruction caches (1 / 2 / 3 / 5 at a time: 63 / 37 / 37 / 58 ms); rolled, one kernel alone is 2.3x slower but five at
    // once take no longer than one: 24.7 ms (profiles/r02_tma.txt).  JB_SMALL_CODE=0 keeps the unrolled kernels, three at a time.
    static const bool smallCodeOk = [] { const char* v = std::getenv("JB_SMALL_CODE"); return v == nullptr || std::atoi(v) != 0; }();
    const bool manySets = !serial && e->chain.size() == 1 && e->chain[0] == jb::kTexture && e->nCh == 2 && e->groups.size() >= 3
                          && e->groups.size() <= (size_t) kPool && nc <= 24576;
    // Better still: every set on its own disjoint group of SMs (green contexts, jb_partition.cpp) -- then the full-speed
    // unrolled kernels never meet in an instruction cache.  Falls back to the small-code kernels where partitions are
    // unavailable.
    if (manySets && e->partitionsFor != (int) e->groups.size()) {
        e->partitions.create(e->device, (int) e->groups.size());
        e->partitionsFor = (int) e->groups.size();
    }
    const bool partitioned = manySets && e->partitions.count == (int) e->groups.size();
    const bool smallCode = smallCodeOk && manySets && !partitioned;
    for (const jb_engine::Group& g : e->groups) {
        ProcArgs a;
        if (g.first >= 0) {
            const long long lo = std::max<long long>(g.first, c0), hi = std::min<long long>((long long) g.first + g.count, c0 + nc);
            if (lo >= hi)
                continue;
            buildArgs(e, a, paramsOfSet(e, g.set), dIn + (lo - c0) * clipStride, dOut + (lo - c0) * clipStride, ns, (int) (hi - lo), lo,
                      rowPitch);
        } else {
            const int* m = e->clipMapHost.data() + g.mapOffset;
            const int* i0 = std::lower_bound(m, m + g.count, (int) c0);
            const int* i1 = std::lower_bound(m, m + g.count, (int) (c0 + nc));
            if (i0 >= i1)
                continue;
            // absolute clip numbers index the state arrays; the audio pointers are moved back to where clip 0 would be
            buildArgs(e, a, paramsOfSet(e, g.set), dIn - c0 * clipStride, dOut - c0 * clipStride, ns, (int) (i1 - i0), 0, rowPitch);
            a.clipMap = e->dClipMap + g.mapOffset + (i0 - m);
            a.octets = 0;
        }
        cudaStream_t st = e->stream;
        if (!serial) {
            st = partitioned ? static_cast<cudaStream_t>(e->partitions.streams[next]) : e->groupStream[next];
            if (!usedStream[next]) {
                JB_CUDA(cudaStreamWaitEvent(st, e->groupFork, 0));
                usedStream[next] = true;
            }
            // three launches side by side: more DIFFERENT kernels on the SMs at once cost more than they overlap (C3, five
            // Texture materials on 8192 clips: 1 stream 58 ms, 2 30.5, 3 27.1, 5 37.3 -- profiles/r01_s6_survey_single.txt)
            static const int poolEnv = [] { const char* v = std::getenv("JB_GROUP_STREAMS"); return v == nullptr ? 0 : std::min((int) jb_engine::kGroupStreams, std::max(1, std::atoi(v))); }();
            const int poolUse = partitioned ? (int) e->groups.size() : (poolEnv > 0 ? poolEnv : (smallCode ? (int) e->groups.size() : 3));
            next = (next + 1) % poolUse;
        }
        a.smallCode = smallCode ? 1 : 0;
        if (int rc = launchKernels(e, a, st, serial))
            return rc;
    }
    if (!serial)
        for (int i = 0; i < kPool; ++i)
            if (usedStream[i]) {
                JB_CUDA(cudaEventRecord(e->groupJoin[i], partitioned ? static_cast<cudaStream_t>(e->partitions.streams[i]) : e->groupStream[i]));
                JB_CUDA(cudaStreamWaitEvent(e->stream, e->groupJoin[i], 0));
            }
    JB_CUDA(cudaEventRecord(stop, e->stream));
    e->timingUsed += 2;
    return JB_OK;
}

// The same with the automation schedule: the range starts at absolute block firstBlock; parameter changes scheduled for
// a block take effect before that block's processBlock, exactly like a host that calls setValueNotifyingHost between
// callbacks, so the render is cut there.  Consumes the events it applies.
int renderAutomated(jb_engine* e, const float* dIn, float* dOut, int ns, int nc, long long c0, long long rowPitch, long long firstBlock)
{
    const int B = e->blockSize;
    const long long endBlock = firstBlock + (ns + B - 1) / B;
    long long cur = firstBlock;
    while (cur < endBlock) {
        size_t applied = 0;
        while (applied < e->schedule.size() && e->schedule[applied].block <= cur) {
            const jb_engine::AutoEvent& ev = e->schedule[applied];
            changeParams(e, ev.first, ev.count, [&](std::vector<jb::ParamSet>& p) { p[(size_t) ev.slot].setPlain(ev.index, ev.value); });
            ++applied;
        }
        e->schedule.erase(e->schedule.begin(), e->schedule.begin() + (long) applied);
        long long next = endBlock;
        if (!e->schedule.empty() && e->schedule.front().block < next)
            next = e->schedule.front().block;
        const long long t0 = (cur - firstBlock) * B;
        const int len = (int) std::min<long long>(ns - t0, (next - cur) * B);
        e->blocksDone = cur; // history index of this segment's first block
        if (int rc = renderClips(e, dIn + t0, dOut + t0, len, nc, c0, rowPitch))
            return rc;
        cur = next;
    }
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
    if (n_channels != 1 && n_channels != 2) // isBusesLayoutSupported: mono or stereo, in == out
        return fail(JB_ERR_UNSUPPORTED, "jb_create: mono or stereo buses only (n_channels = %d)", n_channels);
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
    if (e->comm != nullptr)
        jb_comm_destroy(e);
    freeDevice(e);
    delete e;
    return JB_OK;
}

int jb_chain_length(const jb_engine* e) { return e ? (int) e->chain.size() : 0; }
int jb_chain_kind(const jb_engine* e, int slot) { return checkSlot(e, slot) == JB_OK ? e->chain[(size_t) slot] : JB_ERR_ARG; }
int jb_num_clips(const jb_engine* e) { return e ? e->nClips : 0; }

int jb_prepare(jb_engine* e, double sample_rate, int samples_per_block)
{
    NvtxRange nvtxRange("jb_prepare");
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
    changeParams(e, 0, -1, [&](std::vector<jb::ParamSet>& p) { p[(size_t) slot].setPlain(idx, plain_value); });
    return JB_OK;
}

int jb_set_param_normalised(jb_engine* e, int slot, const char* id, float normalised)
{
    int idx = -1;
    if (int rc = findParam(e, slot, id, &idx))
        return rc;
    changeParams(e, 0, -1, [&](std::vector<jb::ParamSet>& p) { p[(size_t) slot].setNormalised(idx, normalised); });
    return JB_OK;
}

static int checkClipRange(const jb_engine* e, int first_clip, int n_clips)
{
    if (first_clip == JB_ALL_CLIPS)
        return JB_OK;
    if (first_clip < 0 || n_clips < 1 || (long long) first_clip + n_clips > e->nClips)
        return fail(JB_ERR_ARG, "clip range [%d, %d) outside the engine's %d clips", first_clip, first_clip + n_clips, e->nClips);
    return JB_OK;
}

int jb_set_param_clips(jb_engine* e, int slot, const char* id, float plain_value, int first_clip, int n_clips)
{
    int idx = -1;
    if (int rc = findParam(e, slot, id, &idx))
        return rc;
    if (int rc = checkClipRange(e, first_clip, n_clips))
        return rc;
    changeParams(e, first_clip == JB_ALL_CLIPS ? 0 : first_clip, first_clip == JB_ALL_CLIPS ? -1 : n_clips,
                 [&](std::vector<jb::ParamSet>& p) { p[(size_t) slot].setPlain(idx, plain_value); });
    return JB_OK;
}

int jb_set_program_clips(jb_engine* e, int slot, int index, int first_clip, int n_clips)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    if (int rc = checkClipRange(e, first_clip, n_clips))
        return rc;
    changeParams(e, first_clip == JB_ALL_CLIPS ? 0 : first_clip, first_clip == JB_ALL_CLIPS ? -1 : n_clips,
                 [&](std::vector<jb::ParamSet>& p) { p[(size_t) slot].setProgram(index); });
    return JB_OK;
}

int jb_get_param_clip(const jb_engine* e, int slot, const char* id, int clip, float* out)
{
    int idx = -1;
    if (int rc = findParam(e, slot, id, &idx))
        return rc;
    if (out == nullptr || clip < 0 || clip >= e->nClips)
        return fail(JB_ERR_ARG, "jb_get_param_clip: bad clip %d or null output", clip);
    const int set = e->clipSet.empty() ? 0 : e->clipSet[(size_t) clip];
    *out = paramsOfSet(e, set)[(size_t) slot].raw(idx);
    return JB_OK;
}

int jb_num_param_sets(const jb_engine* e)
{
    if (checkEngine(e) != JB_OK)
        return JB_ERR_ARG;
    if (e->clipSet.empty())
        return 1;
    std::vector<char> used(1 + e->variants.size(), 0);
    int n = 0;
    for (int c : e->clipSet)
        if (!used[(size_t) c]) {
            used[(size_t) c] = 1;
            ++n;
        }
    return n;
}

int jb_schedule_param(jb_engine* e, int slot, const char* id, long long at_block, float plain_value, int first_clip, int n_clips)
{
    int idx = -1;
    if (int rc = findParam(e, slot, id, &idx))
        return rc;
    if (int rc = checkClipRange(e, first_clip, n_clips))
        return rc;
    if (at_block < e->blocksDone)
        return fail(JB_ERR_ARG, "jb_schedule_param: block %lld has already been rendered (%lld done)", at_block, e->blocksDone);
    jb_engine::AutoEvent ev { at_block, slot, idx, plain_value, first_clip == JB_ALL_CLIPS ? 0 : first_clip,
                              first_clip == JB_ALL_CLIPS ? -1 : n_clips };
    auto pos = std::upper_bound(e->schedule.begin(), e->schedule.end(), ev,
                                [](const jb_engine::AutoEvent& a, const jb_engine::AutoEvent& b) { return a.block < b.block; });
    e->schedule.insert(pos, ev);
    return JB_OK;
}

int jb_clear_schedule(jb_engine* e)
{
    if (int rc = checkEngine(e))
        return rc;
    e->schedule.clear();
    return JB_OK;
}

// ---------------------------------------------------------------- state blobs
// get/setStateInformation (e.g. JuicySaturator/PluginProcessor.cpp:117-131): the APVTS state tree -- type "PARAMS", one
// <PARAM id="..." value="..."/> child per parameter -- as the blob AudioProcessor::copyXmlToBinary makes of it: int32 LE magic
// 0x21324356, int32 LE length of the XML text, the single-line UTF-8 XML text, one NUL.  JUCE is not available offline, so
// this follows its documented format; the import side accepts any attribute order / whitespace / XML prolog.
namespace {

std::string formatStateValue(float v)
{
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.15g", (double) v);
    std::string t(buf);
    if (t.find_first_of(".eEn") == std::string::npos) // JUCE writes doubles with a decimal point ("6.0")
        t += ".0";
    return t;
}

std::string stateXml(const jb::ParamSet& p)
{
    const std::vector<jb::ParamSpec>& specs = jb::paramSpecs(p.kind());
    std::string x = "<?xml version=\"1.0\" encoding=\"UTF-8\"?> <PARAMS>";
    for (size_t i = 0; i < specs.size(); ++i)
        x += std::string("<PARAM id=\"") + specs[i].id + "\" value=\"" + formatStateValue(p.raw((int) i)) + "\"/>";
    x += "</PARAMS>";
    return x;
}

// value of attribute `name` inside the tag text [b, e)
bool xmlAttribute(const std::string& s, size_t b, size_t e, const char* name, std::string* out)
{
    const std::string key = std::string(name) + "=";
    size_t pos = b;
    while ((pos = s.find(key, pos)) != std::string::npos && pos < e) {
        const bool boundary = pos == b || s[pos - 1] == ' ' || s[pos - 1] == '\t' || s[pos - 1] == '\n' || s[pos - 1] == '\r';
        const size_t q = pos + key.size();
        if (boundary && q < e && (s[q] == '"' || s[q] == '\'')) {
            const size_t close = s.find(s[q], q + 1);
            if (close == std::string::npos || close > e)
                return false;
            *out = s.substr(q + 1, close - q - 1);
            return true;
        }
        pos = q;
    }
    return false;
}

// replaceState(ValueTree::fromXml(...)): every parameter takes its child's value; one without a child goes back to its
// default (AudioProcessorValueTreeState re-creates the missing child from the parameter's default value)
int applyStateXml(const std::string& xml, jb::ParamSet& p)
{
    const size_t root = xml.find("<PARAMS");
    if (root == std::string::npos)
        return JB_ERR_ARG; // setStateInformation ignores trees of another type; the C ABI reports it
    const std::vector<jb::ParamSpec>& specs = jb::paramSpecs(p.kind());
    std::vector<char> seen(specs.size(), 0);
    size_t pos = root;
    while ((pos = xml.find("<PARAM", pos + 1)) != std::string::npos) {
        const char after = pos + 6 < xml.size() ? xml[pos + 6] : '>';
        if (after != ' ' && after != '\t' && after != '\n' && after != '\r' && after != '/')
            continue; // <PARAMS ...>
        const size_t end = xml.find('>', pos);
        if (end == std::string::npos)
            break;
        std::string id, value;
        if (!xmlAttribute(xml, pos, end, "id", &id) || !xmlAttribute(xml, pos, end, "value", &value))
            continue;
        const int idx = p.find(id.c_str());
        if (idx < 0)
            continue;
        char* stop = nullptr;
        const double v = std::strtod(value.c_str(), &stop);
        if (stop == value.c_str())
            continue;
        p.setPlain(idx, (float) v);
        seen[(size_t) idx] = 1;
    }
    for (size_t i = 0; i < specs.size(); ++i)
        if (!seen[i])
            p.setPlain((int) i, specs[i].def);
    return JB_OK;
}

} // namespace

int jb_get_state(const jb_engine* e, int slot, int clip, void* buffer, size_t capacity, size_t* size_out)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    if (clip != JB_ALL_CLIPS && (clip < 0 || clip >= e->nClips))
        return fail(JB_ERR_ARG, "jb_get_state: clip %d outside the engine's %d clips", clip, e->nClips);
    const int set = clip == JB_ALL_CLIPS || e->clipSet.empty() ? 0 : e->clipSet[(size_t) clip];
    const std::string xml = stateXml(paramsOfSet(e, set)[(size_t) slot]);
    const size_t total = 8 + xml.size() + 1;
    if (size_out != nullptr)
        *size_out = total;
    if (buffer == nullptr)
        return JB_OK; // size query
    if (capacity < total)
        return fail(JB_ERR_ARG, "jb_get_state: buffer of %zu bytes, %zu needed", capacity, total);
    unsigned char* out = static_cast<unsigned char*>(buffer);
    const uint32_t magic = 0x21324356u, len = (uint32_t) xml.size();
    for (int i = 0; i < 4; ++i) {
        out[i] = (unsigned char) (magic >> (8 * i));
        out[4 + i] = (unsigned char) (len >> (8 * i));
    }
    std::memcpy(out + 8, xml.data(), xml.size());
    out[8 + xml.size()] = 0;
    return JB_OK;
}

int jb_set_state(jb_engine* e, int slot, const void* data, size_t size, int first_clip, int n_clips)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    if (int rc = checkClipRange(e, first_clip, n_clips))
        return rc;
    if (data == nullptr || size < 9)
        return fail(JB_ERR_ARG, "jb_set_state: no state blob");
    const unsigned char* in = static_cast<const unsigned char*>(data);
    uint32_t magic = 0, len = 0;
    for (int i = 0; i < 4; ++i) {
        magic |= (uint32_t) in[i] << (8 * i);
        len |= (uint32_t) in[4 + i] << (8 * i);
    }
    if (magic != 0x21324356u) // getXmlFromBinary returns nullptr and setStateInformation does nothing
        return fail(JB_ERR_ARG, "jb_set_state: not a copyXmlToBinary blob (magic %08x)", magic);
    const size_t textLen = std::min<size_t>(len, size - 8);
    const std::string xml(reinterpret_cast<const char*>(in + 8), textLen);
    jb::ParamSet probe = e->params[(size_t) slot];
    if (applyStateXml(xml, probe) != JB_OK)
        return fail(JB_ERR_ARG, "jb_set_state: the blob holds no <PARAMS> tree");
    changeParams(e, first_clip == JB_ALL_CLIPS ? 0 : first_clip, first_clip == JB_ALL_CLIPS ? -1 : n_clips,
                 [&](std::vector<jb::ParamSet>& p) { applyStateXml(xml, p[(size_t) slot]); });
    return JB_OK;
}

int jb_num_programs(const jb_engine* e, int slot) { return checkSlot(e, slot) == JB_OK ? e->params[(size_t) slot].numPrograms() : JB_ERR_ARG; }
int jb_get_program(const jb_engine* e, int slot) { return checkSlot(e, slot) == JB_OK ? e->params[(size_t) slot].currentProgram() : JB_ERR_ARG; }

int jb_set_program(jb_engine* e, int slot, int index)
{
    if (int rc = checkSlot(e, slot))
        return rc;
    changeParams(e, 0, -1, [&](std::vector<jb::ParamSet>& p) { p[(size_t) slot].setProgram(index); });
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
    NvtxRange nvtxRange("jb_process");
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
    const long long blocksBase = e->blocksDone;
    const int rc = renderAutomated(e, d_in, d_out, n_samples, e->nClips, 0, n_samples, blocksBase);
    e->blocksDone = blocksBase;
    if (rc != JB_OK)
        return rc;
    e->blocksDone += (n_samples + e->blockSize - 1) / e->blockSize;
    return JB_OK;
}

} // extern "C"

extern "C" int jbk_launch_pcm16_to_float(const int16_t* src, float* dst, long long rows, int n, long long srcPitch, long long dstPitch, void* stream);
extern "C" int jbk_launch_float_to_pcm16(const float* src, int16_t* dst, long long rows, int n, long long srcPitch, long long dstPitch, void* stream);

namespace {

// Time slices of a host-buffer render (jb_process_host): `sliceBlocks` host blocks each, the last ones halving down to one
// block when `taper` (nothing overlaps the final slice's render and download, so it should be short).  Returns the first
// block of every slice, then totalBlocks.
std::vector<int> planSlices(int totalBlocks, int sliceBlocks, bool taper)
{
    std::vector<int> first;
    sliceBlocks = std::max(1, std::min(sliceBlocks, totalBlocks));
    std::vector<int> tail;                       // 1, 2, 4, ... blocks, walked from the end of the call
    if (taper && totalBlocks >= 4 * sliceBlocks)
        for (int len = 1; len < sliceBlocks; len *= 2)
            tail.push_back(len);
    int tailBlocks = 0;
    for (int len : tail)
        tailBlocks += len;
    int b = 0;
    for (; b + sliceBlocks <= totalBlocks - tailBlocks; b += sliceBlocks)
        first.push_back(b);
    if (b < totalBlocks - tailBlocks) {          // ragged rest of the uniform part
        first.push_back(b);
        b = totalBlocks - tailBlocks;
    }
    for (size_t i = tail.size(); i-- > 0;) {
        first.push_back(b);
        b += tail[i];
    }
    first.push_back(totalBlocks);
    return first;
}

// Does a render leave the audio as it is?  A chain of JuicyInfer instances whose trim is 0 dB everywhere: applyGain(1) is a
// no-op (JuicyInfer/PluginProcessor.cpp:79), the plugin only scores.  (Pending automation: conservatively no.)
bool audioUntouched(const jb_engine* e)
{
    if (!e->schedule.empty())
        return false;
    for (int k : e->chain)
        if (k != jb::kInfer)
            return false;
    const int nSets = 1 + (int) e->variants.size();
    for (int set = 0; set < nSets; ++set)
        for (const jb::ParamSet& p : paramsOfSet(e, set)) {
            SlotCoef c;
            jb::makeSlotCoef(p, e->sampleRate, &c);
            if (c.infer.gainMode != 0)
                return false;
        }
    return true;
}

// jb_process_host / jb_process_host_pcm16: host audio [clip][channel][sample] as fp32 or as 16-bit PCM (`pcm16`).
int processHost(jb_engine* e, const void* h_in_v, void* h_out_v, int n_samples, bool pcm16)
{
    NvtxRange nvtxRange(pcm16 ? "jb_process_host_pcm16" : "jb_process_host");
    const float* h_in = static_cast<const float*>(h_in_v);
    float* h_out = static_cast<float*>(h_out_v);
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
    // Geometry (measured: profiles/r02_e2e_geometry.txt).  ONE pass whenever the batch fits the device beside what the
    // engine already holds: the C5 shard (12.6 GB) in two passes of 8 GiB ran at 35 GB/s per direction, in one pass at 47
    // -- the PCIe full-duplex rate of the box.  Within a pass small slices win (nothing overlaps the first slice's upload
    // and the last slice's render + download): one host block per slice on big batches; small batches keep several blocks
    // per slice (JB_HOST_SLICE_MIB: a launch over one block of a few thousand clips is mostly prologue) and halve the last
    // slices down to one block instead (JB_HOST_TAPER).
    size_t passBudget = envMiB("JB_HOST_PASS_MIB", 32768);
    const size_t sliceTarget = envMiB("JB_HOST_SLICE_MIB", 96);
    {
        size_t freeB = 0, totalB = 0;
        if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
            const size_t held = e->stageBytes[0] + e->stageBytes[1] + e->stageBytes[2];
            const size_t whole = clipBytes * (size_t) e->nClips;
            const size_t avail = (size_t) ((double) (freeB + held) * 0.85);
            if (whole > avail)                       // does not fit in one piece: two alternating buffers must
                passBudget = std::min(passBudget, avail / 2);
        } else {
            cudaGetLastError();
        }
    }
    long long passClips = std::max<long long>(32, (long long) (passBudget / clipBytes) / 32 * 32);
    passClips = std::min<long long>(passClips, e->nClips);
    const int nPasses = (int) ((e->nClips + passClips - 1) / passClips);
    const int totalBlocks = (n_samples + e->blockSize - 1) / e->blockSize;
    const size_t blockBytesAllClips = sizeof(float) * (size_t) e->blockSize * (size_t) e->nCh * (size_t) passClips;
    const int minSliceBlocks = [] { const char* v = std::getenv("JB_HOST_MIN_SLICE_BLOCKS"); return v ? std::max(1, std::atoi(v)) : 1; }();
    const bool taper = [] { const char* v = std::getenv("JB_HOST_TAPER"); return v == nullptr || std::atoi(v) != 0; }();
    // A slice's 2-D copy moves one row per (clip, channel), and the copy engines retire rows at a bounded RATE: 65536 rows
    // take ~2.7 ms upward whether they hold 1 KB or 2 KB (profiles/r02_e2e_timeline.txt), and 1 KB rows downward twice that.
    // Rows of at least 2 KB (fp32) / 4 KB (16-bit PCM, which is render-bound and gains from longer launches as well: C5
    // shard 1 / 3 / 4 / 8 blocks per slice 330 / 210 / 189 / 187 ms) keep the copies on the PCIe rate instead.
    const size_t minRowBytes = pcm16 ? 4096 : 2048;
    const size_t blockRowBytes = (size_t) e->blockSize * (pcm16 ? sizeof(int16_t) : sizeof(float));
    const size_t rowBlocks = (minRowBytes + blockRowBytes - 1) / blockRowBytes;
    int sliceBlocks = (int) std::max({ (size_t) minSliceBlocks, rowBlocks, sliceTarget / std::max<size_t>(1, blockBytesAllClips) });
    sliceBlocks = std::min(sliceBlocks, totalBlocks);
    // slice s covers host blocks [sliceFirst[s], sliceFirst[s + 1])
    const std::vector<int> sliceFirst = planSlices(totalBlocks, sliceBlocks, taper);
    const int nSlices = (int) sliceFirst.size() - 1;

    if (e->copyIn == nullptr) {
        JB_CUDA(cudaStreamCreateWithFlags(&e->copyIn, cudaStreamNonBlocking));
        JB_CUDA(cudaStreamCreateWithFlags(&e->copyOut, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i)
            JB_CUDA(cudaEventCreateWithFlags(&e->evOut[i], cudaEventDisableTiming));
    }
    const size_t need = clipBytes * (size_t) passClips;
    const int nBuffers = nPasses > 1 ? 2 : 1;
    for (int i = 0; i < nBuffers; ++i) {
        if (e->dStage[i] == nullptr || need > e->stageBytes[i]) {
            cudaFree(e->dStage[i]);
            e->dStage[i] = nullptr;
            e->stageBytes[i] = 0;
            JB_CUDA(cudaMalloc(&e->dStage[i], need));
            e->stageBytes[i] = need;
        }
    }
    if (pcm16) { // 16-bit images of the staging buffers: what actually crosses PCIe
        for (int i = 0; i < nBuffers; ++i) {
            if (e->dPcm[i] == nullptr || need / 2 > e->pcmBytes[i]) {
                cudaFree(e->dPcm[i]);
                e->dPcm[i] = nullptr;
                e->pcmBytes[i] = 0;
                JB_CUDA(cudaMalloc(&e->dPcm[i], need / 2));
                e->pcmBytes[i] = need / 2;
            }
        }
    }
    while (e->sliceEvents.size() < (size_t) 2 * (size_t) nSlices) {
        cudaEvent_t ev = nullptr;
        JB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        e->sliceEvents.push_back(ev);
    }

    // In place on the host and nothing in the chain changes the audio (JuicyInfer, trim = 0 dB): the caller's buffer already
    // holds the result, only the records come back -- half the PCIe traffic of a scoring run (BASELINE config 4).
    // (fp32 buffers only: the 16-bit writer's rule turns a -32768 into -32767, so that path always writes back.)
    const bool skipDownload = !pcm16 && h_in_v == h_out_v && audioUntouched(e);
    const long long blocksBase = e->blocksDone;
    // renderAutomated moves e->blocksDone while it walks the slices; whatever exit path is taken (a CUDA error return
    // included) the counter ends at the value the call began with, or past the whole render on success
    struct BlocksGuard {
        jb_engine* e;
        long long value;
        ~BlocksGuard() { e->blocksDone = value; }
    } blocksGuard { e, blocksBase };
    int rcLaunch = JB_OK;
    // JB_HOST_TRACE=file: the pipeline's timeline of this call (one line per slice: when the host thread had issued it,
    // when its upload / render / download finished on the device), appended to `file`.  Diagnostic only.
    const char* tracePath = std::getenv("JB_HOST_TRACE");
    struct TraceRow { int pass, slice, blocks; double issuedMs; cudaEvent_t up, render, down; };
    std::vector<TraceRow> trace;
    cudaEvent_t traceStart = nullptr;
    const auto hostT0 = std::chrono::steady_clock::now();
    if (tracePath != nullptr) {
        JB_CUDA(cudaEventCreate(&traceStart));
        JB_CUDA(cudaEventRecord(traceStart, e->copyIn));
    }
    // every pass walks the same stretch of time, so each starts from the parameters and schedule the call began with
    const bool replay = nPasses > 1 && !e->schedule.empty();
    const auto params0 = replay ? e->params : std::vector<jb::ParamSet>();
    const auto variants0 = replay ? e->variants : std::vector<std::vector<jb::ParamSet>>();
    const auto clipSet0 = replay ? e->clipSet : std::vector<int>();
    const auto schedule0 = replay ? e->schedule : std::vector<jb_engine::AutoEvent>();
    for (int pass = 0; pass < nPasses && rcLaunch == JB_OK; ++pass) {
        if (replay && pass > 0) {
            e->params = params0;
            e->variants = variants0;
            e->clipSet = clipSet0;
            e->schedule = schedule0;
            e->groupsDirty = true;
        }
        const long long c0 = (long long) pass * passClips;
        const int nc = (int) std::min<long long>(passClips, e->nClips - c0);
        float* dBuf = e->dStage[pass % nBuffers];
        int16_t* dPcm = pcm16 ? e->dPcm[pass % nBuffers] : nullptr;
        const size_t rows = (size_t) nc * (size_t) e->nCh;
        const float* hIn = h_in + (size_t) c0 * e->nCh * n_samples;
        float* hOut = h_out + (size_t) c0 * e->nCh * n_samples;
        if (pass >= nBuffers) // this buffer's previous pass has been downloaded
            JB_CUDA(cudaStreamWaitEvent(e->copyIn, e->evOut[pass % nBuffers], 0));
        for (int sl = 0; sl < nSlices; ++sl) {
            const int firstBlock = sliceFirst[(size_t) sl];
            const int t0 = firstBlock * e->blockSize;
            const int ns = std::min(n_samples, sliceFirst[(size_t) sl + 1] * e->blockSize) - t0;
            cudaEvent_t evIn = e->sliceEvents[(size_t) 2 * sl], evDone = e->sliceEvents[(size_t) 2 * sl + 1];
            if (!pcm16) {
                JB_CUDA(cudaMemcpy2DAsync(dBuf + t0, rowBytes, hIn + t0, rowBytes, sizeof(float) * (size_t) ns, rows,
                                          cudaMemcpyHostToDevice, e->copyIn));
            } else {
                const int16_t* hIn16 = static_cast<const int16_t*>(h_in_v) + (size_t) c0 * e->nCh * n_samples;
                JB_CUDA(cudaMemcpy2DAsync(dPcm + t0, rowBytes / 2, hIn16 + t0, rowBytes / 2, sizeof(int16_t) * (size_t) ns, rows,
                                          cudaMemcpyHostToDevice, e->copyIn));
            }
            JB_CUDA(cudaEventRecord(evIn, e->copyIn));
            TraceRow row { pass, sl, sliceFirst[(size_t) sl + 1] - firstBlock, 0.0, nullptr, nullptr, nullptr };
            if (tracePath != nullptr) {
                JB_CUDA(cudaEventCreate(&row.up));
                JB_CUDA(cudaEventCreate(&row.render));
                JB_CUDA(cudaEventCreate(&row.down));
                JB_CUDA(cudaEventRecord(row.up, e->copyIn));
            }
            JB_CUDA(cudaStreamWaitEvent(e->stream, evIn, 0));
            if (pcm16 && jbk_launch_pcm16_to_float(dPcm + t0, dBuf + t0, (long long) rows, ns, n_samples, n_samples, e->stream) != 0)
                return fail(JB_ERR_CUDA, "pcm16 -> float conversion kernel failed to launch");
            if ((rcLaunch = renderAutomated(e, dBuf + t0, dBuf + t0, ns, nc, c0, n_samples, blocksBase + firstBlock)) != JB_OK)
                break;
            if (pcm16 && jbk_launch_float_to_pcm16(dBuf + t0, dPcm + t0, (long long) rows, ns, n_samples, n_samples, e->stream) != 0)
                return fail(JB_ERR_CUDA, "float -> pcm16 conversion kernel failed to launch");
            JB_CUDA(cudaEventRecord(evDone, e->stream));
            if (tracePath != nullptr)
                JB_CUDA(cudaEventRecord(row.render, e->stream));
            JB_CUDA(cudaStreamWaitEvent(e->copyOut, evDone, 0));
            if (skipDownload) {
                // nothing to bring back; copyOut still follows the renders, so the pass-buffer hand-over below holds
            } else if (!pcm16) {
                JB_CUDA(cudaMemcpy2DAsync(hOut + t0, rowBytes, dBuf + t0, rowBytes, sizeof(float) * (size_t) ns, rows,
                                          cudaMemcpyDeviceToHost, e->copyOut));
            } else {
                int16_t* hOut16 = static_cast<int16_t*>(h_out_v) + (size_t) c0 * e->nCh * n_samples;
                JB_CUDA(cudaMemcpy2DAsync(hOut16 + t0, rowBytes / 2, dPcm + t0, rowBytes / 2, sizeof(int16_t) * (size_t) ns, rows,
                                          cudaMemcpyDeviceToHost, e->copyOut));
            }
            if (tracePath != nullptr) {
                JB_CUDA(cudaEventRecord(row.down, e->copyOut));
                row.issuedMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hostT0).count();
                trace.push_back(row);
            }
        }
        JB_CUDA(cudaEventRecord(e->evOut[pass % nBuffers], e->copyOut));
        // (the slice events are re-recorded by the next pass; the waits above captured this pass's records)
    }
    JB_CUDA(cudaStreamSynchronize(e->copyIn));
    JB_CUDA(cudaStreamSynchronize(e->stream));
    JB_CUDA(cudaStreamSynchronize(e->copyOut));
    if (tracePath != nullptr) {
        if (FILE* f = std::fopen(tracePath, "a")) {
            std::fprintf(f, "# call: %d clips, %d samples, %d pass(es), %d slices; pass slice blocks issued_ms up_done_ms render_done_ms down_done_ms\n",
                         e->nClips, n_samples, nPasses, nSlices);
            for (const TraceRow& r : trace) {
                float up = 0.0f, rd = 0.0f, dn = 0.0f;
                cudaEventElapsedTime(&up, traceStart, r.up);
                cudaEventElapsedTime(&rd, traceStart, r.render);
                cudaEventElapsedTime(&dn, traceStart, r.down);
                std::fprintf(f, "%d %d %d %.3f %.3f %.3f %.3f\n", r.pass, r.slice, r.blocks, r.issuedMs, up, rd, dn);
            }
            std::fclose(f);
        }
        for (const TraceRow& r : trace) {
            cudaEventDestroy(r.up);
            cudaEventDestroy(r.render);
            cudaEventDestroy(r.down);
        }
        cudaEventDestroy(traceStart);
    }
    if (rcLaunch != JB_OK)
        return rcLaunch;
    blocksGuard.value = blocksBase + totalBlocks;
    return JB_OK;
}

} // namespace

extern "C" {

int jb_plan_slices(int total_blocks, int slice_blocks, int taper, int* first, int capacity)
{
    if (total_blocks <= 0 || slice_blocks <= 0)
        return fail(JB_ERR_ARG, "jb_plan_slices: block counts must be positive");
    const std::vector<int> plan = planSlices(total_blocks, slice_blocks, taper != 0);
    if (first != nullptr)
        for (int i = 0; i < capacity && i < (int) plan.size(); ++i)
            first[i] = plan[(size_t) i];
    return (int) plan.size() - 1;
}

int jb_process_host(jb_engine* e, const float* h_in, float* h_out, int n_samples)
{
    return processHost(e, h_in, h_out, n_samples, false);
}

int jb_process_host_pcm16(jb_engine* e, const int16_t* h_in, int16_t* h_out, int n_samples)
{
    return processHost(e, h_in, h_out, n_samples, true);
}

int jb_get_metrics(jb_engine* e, int slot, jb_metrics* out)
{
    NvtxRange nvtxRange("jb_get_metrics");
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

int jb_set_math_mode(jb_engine* e, int mode)
{
    if (int rc = checkEngine(e))
        return rc;
    if (mode < 0 || mode > 2)
        return fail(JB_ERR_ARG, "jb_set_math_mode: mode %d (0 auto, 1 exact, 2 fast)", mode);
    e->mathMode = mode;
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

int jb_enable_slot_timing(jb_engine* e, int on)
{
    if (int rc = checkEngine(e))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    JB_CUDA(cudaStreamSynchronize(e->stream));
    e->slotTiming = on != 0;
    e->slotMarks.clear();
    e->slotEventsUsed = 0;
    for (int s = 0; s < JBK_MAX_CHAIN; ++s) {
        e->slotMs[s] = 0.0;
        e->slotLaunches[s] = 0;
    }
    return JB_OK;
}

int jb_slot_time_ms(jb_engine* e, int slot, double* ms, long long* launches)
{
    if (int rc = checkEngine(e))
        return rc;
    if (slot < 0 || slot >= (int) e->chain.size())
        return fail(JB_ERR_ARG, "jb_slot_time_ms: slot %d outside the chain", slot);
    if (int rc = setDevice(e))
        return rc;
    if (!e->slotMarks.empty()) {
        JB_CUDA(cudaStreamSynchronize(e->stream));
        for (const jb_engine::SlotMark& m : e->slotMarks) {
            float t = 0.0f;
            JB_CUDA(cudaEventElapsedTime(&t, e->slotEvents[m.evStart], e->slotEvents[m.evStop]));
            e->slotMs[m.slot] += (double) t;
            ++e->slotLaunches[m.slot];
        }
        e->slotMarks.clear();
        e->slotEventsUsed = 0;
    }
    if (ms)
        *ms = e->slotMs[slot];
    if (launches)
        *launches = e->slotLaunches[slot];
    e->slotMs[slot] = 0.0;
    e->slotLaunches[slot] = 0;
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

// ---------------------------------------------------------------- clip sharding and the NCCL score gather
// Clips are independent plugin-instance chains: rank r of N renders a contiguous clip range with its own engine and
// nothing is exchanged while rendering.  The one collective is ncclAllGather of the per-clip records after the render.
// NCCL is resolved at run time (dlopen) so the library itself has no link-time dependency on it: a host that never
// shards never needs it.  JB_NCCL_LIB overrides the library name; a process that already holds NCCL (e.g. one that
// imported torch) gets that copy.
} // extern "C"

#include <dlfcn.h>

namespace {

struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    void* handle = nullptr;
    std::string error;
};
constexpr int kNcclFloat32 = 7; // ncclFloat32 (nccl.h: ncclDataType_t)

NcclApi* ncclApi()
{
    static NcclApi api;
    static bool tried = false;
    if (tried)
        return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = { std::getenv("JB_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) {
        if (n == nullptr || *n == 0)
            continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle)
            break;
        api.error = dlerror();
    }
    if (!api.handle)
        return nullptr;
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(api.handle, name); ok = ok && p != nullptr; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    if (!ok) {
        api.error = "a required nccl* symbol is missing";
        dlclose(api.handle);
        api.handle = nullptr;
        return nullptr;
    }
    return &api;
}

int needNccl(NcclApi** out)
{
    *out = ncclApi();
    if (*out == nullptr)
        return fail(JB_ERR_UNSUPPORTED, "NCCL is not available (dlopen libnccl.so.2 failed; set JB_NCCL_LIB)");
    return JB_OK;
}

#define JB_NCCL(api, call)                                                                              \
    do {                                                                                                \
        const int rc__ = (call);                                                                        \
        if (rc__ != 0)                                                                                  \
            return fail(JB_ERR_CUDA, "%s failed: %s", #call, (api)->GetErrorString(rc__));              \
    } while (0)

} // namespace

extern "C" {

int jb_shard_range(long long n_clips, int rank, int world, long long* first, long long* count)
{
    if (n_clips < 0 || world < 1 || rank < 0 || rank >= world || first == nullptr || count == nullptr)
        return fail(JB_ERR_ARG, "jb_shard_range: rank %d of %d, %lld clips", rank, world, n_clips);
    const long long base = n_clips / world, extra = n_clips % world;
    *first = rank * base + std::min<long long>(rank, extra);
    *count = base + (rank < extra ? 1 : 0);
    return JB_OK;
}

long long jb_record_pitch(const jb_engine* e) { return e ? e->clipPitch : 0; }

int jb_comm_version(int* version)
{
    NcclApi* api = nullptr;
    if (int rc = needNccl(&api))
        return rc;
    if (version == nullptr)
        return fail(JB_ERR_ARG, "jb_comm_version: null output");
    JB_NCCL(api, api->GetVersion(version));
    return JB_OK;
}

int jb_comm_unique_id(void* id_out)
{
    NcclApi* api = nullptr;
    if (int rc = needNccl(&api))
        return rc;
    if (id_out == nullptr)
        return fail(JB_ERR_ARG, "jb_comm_unique_id: null output");
    NcclApi::UniqueId id;
    JB_NCCL(api, api->GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof id);
    return JB_OK;
}

int jb_comm_init_rank(jb_engine* e, const void* id, int n_ranks, int rank)
{
    if (int rc = checkEngine(e))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    NcclApi* api = nullptr;
    if (int rc = needNccl(&api))
        return rc;
    if (id == nullptr || n_ranks < 1 || rank < 0 || rank >= n_ranks)
        return fail(JB_ERR_ARG, "jb_comm_init_rank: rank %d of %d", rank, n_ranks);
    if (e->comm != nullptr)
        return fail(JB_ERR_STATE, "jb_comm_init_rank: this engine already has a communicator");
    NcclApi::UniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    JB_NCCL(api, api->CommInitRank(&e->comm, n_ranks, uid, rank));
    e->commRanks = n_ranks;
    e->commRank = rank;
    return JB_OK;
}

int jb_comm_init_all(jb_engine* const* engines, int n)
{
    if (engines == nullptr || n < 1)
        return fail(JB_ERR_ARG, "jb_comm_init_all: no engines");
    NcclApi* api = nullptr;
    if (int rc = needNccl(&api))
        return rc;
    std::vector<int> devs;
    for (int i = 0; i < n; ++i) {
        if (int rc = checkEngine(engines[i]))
            return rc;
        if (engines[i]->hostOnly)
            return fail(JB_ERR_CUDA, "jb_comm_init_all: engine %d has no device", i);
        if (engines[i]->comm != nullptr)
            return fail(JB_ERR_STATE, "jb_comm_init_all: engine %d already has a communicator", i);
        if (std::find(devs.begin(), devs.end(), engines[i]->device) != devs.end())
            return fail(JB_ERR_ARG, "jb_comm_init_all: two engines share device %d (one engine per GPU)", engines[i]->device);
        devs.push_back(engines[i]->device);
    }
    std::vector<void*> comms((size_t) n, nullptr);
    JB_NCCL(api, api->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        engines[i]->comm = comms[(size_t) i];
        engines[i]->commRanks = n;
        engines[i]->commRank = i;
    }
    return JB_OK;
}

int jb_comm_destroy(jb_engine* e)
{
    if (int rc = checkEngine(e))
        return rc;
    if (e->comm == nullptr)
        return JB_OK;
    NcclApi* api = nullptr;
    if (int rc = needNccl(&api))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    cudaStreamSynchronize(e->stream);
    api->CommDestroy(e->comm);
    e->comm = nullptr;
    e->commRanks = 0;
    e->commRank = -1;
    return JB_OK;
}

int jb_comm_size(const jb_engine* e) { return e ? e->commRanks : 0; }
int jb_comm_rank(const jb_engine* e) { return e ? e->commRank : -1; }

int jb_gather_records(jb_engine* e, int slot, float* d_out)
{
    NvtxRange nvtxRange("jb_gather_records");
    if (int rc = checkSlot(e, slot))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    if (e->comm == nullptr)
        return fail(JB_ERR_STATE, "jb_gather_records: no communicator (jb_comm_init_rank / jb_comm_init_all)");
    if (!e->prepared)
        return fail(JB_ERR_STATE, "jb_gather_records before jb_prepare");
    if (d_out == nullptr)
        return fail(JB_ERR_ARG, "jb_gather_records: null output");
    NcclApi* api = nullptr;
    if (int rc = needNccl(&api))
        return rc;
    const float* src = e->dLatest + (long long) slot * JBK_REC * e->clipPitch;
    JB_NCCL(api, api->AllGather(src, d_out, (size_t) JBK_REC * (size_t) e->clipPitch, kNcclFloat32, e->comm, e->stream));
    return JB_OK;
}

int jb_gather_records_all(jb_engine* const* engines, int n, int slot, float* const* d_out)
{
    if (engines == nullptr || d_out == nullptr || n < 1)
        return fail(JB_ERR_ARG, "jb_gather_records_all: null argument");
    NcclApi* api = nullptr;
    if (int rc = needNccl(&api))
        return rc;
    for (int i = 0; i < n; ++i) {
        if (int rc = checkSlot(engines[i], slot))
            return rc;
        if (engines[i]->comm == nullptr || engines[i]->commRanks != n || !engines[i]->prepared || d_out[i] == nullptr)
            return fail(JB_ERR_STATE, "jb_gather_records_all: engine %d is not a prepared member of an %d-rank communicator", i, n);
        if (engines[i]->clipPitch != engines[0]->clipPitch)
            return fail(JB_ERR_ARG, "jb_gather_records_all: engines must hold the same number of clips (record pitch %lld vs %lld)",
                        engines[i]->clipPitch, engines[0]->clipPitch);
    }
    JB_NCCL(api, api->GroupStart());
    for (int i = 0; i < n; ++i) {
        jb_engine* e = engines[i];
        cudaSetDevice(e->device);
        const float* src = e->dLatest + (long long) slot * JBK_REC * e->clipPitch;
        const int rc = api->AllGather(src, d_out[i], (size_t) JBK_REC * (size_t) e->clipPitch, kNcclFloat32, e->comm, e->stream);
        if (rc != 0) {
            api->GroupEnd();
            return fail(JB_ERR_CUDA, "ncclAllGather failed: %s", api->GetErrorString(rc));
        }
    }
    JB_NCCL(api, api->GroupEnd());
    return JB_OK;
}

int jb_gather_records_host(jb_engine* e, int slot, jb_metrics* out, int clips_per_rank)
{
    NvtxRange nvtxRange("jb_gather_records_host");
    if (int rc = checkSlot(e, slot))
        return rc;
    if (int rc = setDevice(e))
        return rc;
    if (out == nullptr || clips_per_rank < 0 || clips_per_rank > e->clipPitch)
        return fail(JB_ERR_ARG, "jb_gather_records_host: bad output / clips_per_rank %d", clips_per_rank);
    if (e->comm == nullptr)
        return fail(JB_ERR_STATE, "jb_gather_records_host: no communicator");
    const size_t per = (size_t) JBK_REC * (size_t) e->clipPitch;
    const size_t need = sizeof(float) * per * (size_t) e->commRanks;
    if (need > e->gatherBytes) {
        cudaFree(e->dGather);
        e->dGather = nullptr;
        e->gatherBytes = 0;
        JB_CUDA(cudaMalloc(&e->dGather, need));
        e->gatherBytes = need;
    }
    if (int rc = jb_gather_records(e, slot, e->dGather))
        return rc;
    e->hostScratch.resize(per * (size_t) e->commRanks);
    JB_CUDA(cudaMemcpyAsync(e->hostScratch.data(), e->dGather, need, cudaMemcpyDeviceToHost, e->stream));
    JB_CUDA(cudaStreamSynchronize(e->stream));
    for (int r = 0; r < e->commRanks; ++r)
        unpackRecords(e->hostScratch.data() + (size_t) r * per, e->clipPitch, clips_per_rank, out + (size_t) r * (size_t) clips_per_rank);
    return JB_OK;
}

} // extern "C"
