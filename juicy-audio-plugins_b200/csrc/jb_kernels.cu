// jb_kernels.cu -- sm_100a device code of the JuicySuite batch engine.
//
// Mapping: one lane = one clip (stereo pair); a warp walks 32
// clips through time.  Per 512-sample block the lane makes chainLen+1 "sweeps":
//
//   sweep 0      : pre-analysis of plugin 0 on the block's input
//   sweep s+1    : DSP of plugin s  ->  post-analysis of plugin s  +  pre-analysis
//                  (and block pre-pass) of plugin s+1 on the samples just produced
//
// which reproduces the reference's per-plugin processBlock order
// analyze(in) -> DSP -> analyze(out) with ONE shared analyzer per plugin
// (e.g. JuicyPunch/PluginProcessor.cpp:82,114) while reading
// and writing every sample once per sweep.  All recurrences keep the
// reference's fp32 operand order; this file is compiled with
// -fmad=false -ftz=true -prec-div=true -prec-sqrt=true (see Makefile).
//
// Reference lines are cited per routine, relative to /root/reference.
#include "jb_lane.cuh"

namespace {

using namespace jbdev;

__global__ void jb_fill_kernel(float* dst, float value, long long count)
{
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        dst[i] = value;
}

// ------------------------------------------------------------------ synthetic clips

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// uniform in [-1, 1), exactly representable
__device__ __forceinline__ float unit_noise(uint32_t seed, uint32_t ch, uint32_t n)
{
    const uint32_t h = hash32(seed ^ hash32(n * 2u + ch + 0x9E3779B9u));
    return (float) (h >> 8) * (1.0f / 8388608.0f) - 1.0f;
}

__global__ void jb_synth_kernel(float* audio, int kind, long long firstClip, int nClips, int nCh, int nSamples,
                                float sr, uint32_t baseSeed)
{
    const long long total = (long long) nClips * nSamples;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int clipLocal = (int) (idx / nSamples);
        const int n = (int) (idx - (long long) clipLocal * nSamples);
        const long long clipId = firstClip + clipLocal;
        const uint32_t seed = baseSeed ^ ((uint32_t) clipId * 0x9E3779B9u);
        const int k = kind == 4 ? (int) (clipId & 3) : kind;
        float l = 0.0f, r = 0.0f;
        if (k == 0) { // exponential sine sweep 20 Hz -> 20 kHz over the clip, R 0.3 rad ahead
            const float T = (float) nSamples / sr;
            const float K = logf(1000.0f);
            const float t = (float) n / sr;
            const float ph = 2.0f * PI_F * 20.0f * T / K * (expf(t / T * K) - 1.0f);
            l = 0.5f * sinf(ph);
            r = 0.5f * sinf(ph + 0.3f);
        } else if (k == 1) { // white noise +-0.25, R = 0.5 (L + independent)
            l = 0.25f * unit_noise(seed, 0u, (uint32_t) n);
            r = 0.5f * (l + 0.25f * unit_noise(seed, 1u, (uint32_t) n));
        } else if (k == 2) { // impulse train, period 2400 + 37 (clip mod 64), R 7 samples late
            const int period = 2400 + 37 * (int) (clipId & 63);
            l = (n % period) == 0 ? 0.9f : 0.0f;
            r = (n >= 7 && ((n - 7) % period) == 0) ? 0.9f : 0.0f;
        } else { // drum hit
            const uint32_t h = hash32(seed);
            const int onset = 480 + (int) (h % 4800u);
            const float f0 = 45.0f + 45.0f * (float) ((h >> 13) & 1023u) / 1023.0f;
            if (n >= onset) {
                const float m = (float) (n - onset);
                const float body = 0.8f * expf(-m / 2400.0f) * sinf(2.0f * PI_F * f0 * m / sr);
                const float burst = 0.4f * expf(-m / 600.0f);
                l = body + burst * unit_noise(seed, 0u, (uint32_t) n);
                r = 0.8f * l + 0.2f * burst * unit_noise(seed, 1u, (uint32_t) n);
            }
        }
        float* base = audio + (long long) clipLocal * nCh * nSamples + n;
        base[0] = l;
        if (nCh > 1)
            base[nSamples] = r;
    }
}

thread_local char g_cudaErr[256];
long long g_launches = 0;
// JB_LANE_GENERIC=1: single-plugin launches also take the generic kernel (A/B timing, tests of the generic path)
const int g_pairMode = [] { const char* v = getenv("JB_PAIR"); return v == nullptr ? -1 : atoi(v); }();
int envInt(const char* name, int dflt) { const char* v = getenv(name); return v == nullptr ? dflt : atoi(v); }
// measured in place (profiles/r01_s6_pair.txt, last block): Texture 8192 clips 1.6x faster in pairs, 16384 / 24576 1.07-1.1x,
// 32768 0.87x; Saturator / Punch with the MUFU math 1.45x at 8192, ~1.1x at 16384 / 24576 (0.93x around 12288), 0.7x at
// 32768 where the one-lane kernel switches to eight samples per trip; with the exact routines faster at every size
// Round 2, Saturator with the exact routines: the one-lane kernel on FOUR samples per trip (its eight-sample body is 42 KB
// of code, past the instruction cache: no_instruction led its stalls) is ahead of the pairs from 32768 clips up -- 18.3
// against 21.3 ms there, 35.0 against 36.7 ms at 65536 (profiles/r02_tile_rolled.txt); Punch stays in pairs (35.9 / 37.8 ms).
const int g_pairLimitTexture = envInt("JB_PAIR_LIMIT_TEXTURE", 24576), g_pairLimitExact = envInt("JB_PAIR_LIMIT_EXACT", 1 << 30),
          g_pairLimitExactSat = envInt("JB_PAIR_LIMIT_EXACT_SAT", 24576), g_pairLimitFast = envInt("JB_PAIR_LIMIT_FAST", 24576);
// clip-per-CTA kernel up to this many clips: eight per SM, two rounds of resident CTAs (measured at the end of round 2,
// profiles/r02_solo_clocks.txt: 1184 clips Saturator 3.2 ms against 4.2 on the lane kernels, Infer 2.8 / 4.7, Cohere 4.6 / 6.3;
// 1480 clips 4.6 / 4.2, 4.2 / 4.7, 6.2 / 6.3)
const int g_soloLimit = envInt("JB_SOLO_LIMIT", 1184);
const bool g_forceGeneric = [] { const char* v = getenv("JB_LANE_GENERIC"); return v != nullptr && atoi(v) != 0; }();

int check(cudaError_t e, const char* what)
{
    if (e == cudaSuccess)
        return 0;
    snprintf(g_cudaErr, sizeof g_cudaErr, "%s: %s", what, cudaGetErrorString(e));
    return -1;
}

} // namespace

extern "C" int jbk_launch_single(const ProcArgs* args, int grid, void* stream); // jb_single_light.cu
extern "C" int jbk_launch_mono(const ProcArgs* args, int grid, void* stream);        // jb_mono.cu
extern "C" int jbk_pair_supported(const ProcArgs* args);                          // jb_pair.cu
extern "C" int jbk_solo_supported(const ProcArgs* args);                          // jb_solo.cu
extern "C" int jbk_launch_solo(const ProcArgs* args, void* stream);
extern "C" int jbk_launch_pair(const ProcArgs* args, void* stream);

extern "C" {

const char* jbk_last_cuda_error(void) { return g_cudaErr; }
long long jbk_launch_count(void) { return g_launches; }
void jbk_note_launch(void) { ++g_launches; }

// Would jbk_launch_process render this single-plugin launch with the clip-per-CTA kernel?
int jbk_solo_pick(const ProcArgs* args)
{
    return args->chainLen == 1 && g_soloLimit > 0 && !args->laneOnly && args->nClips <= g_soloLimit && jbk_solo_supported(args) != 0
           && (!g_forceGeneric || args->exactMath);
}

int jbk_launch_process(const ProcArgs* args, void* stream)
{
    if (args->nClips <= 0 || args->nSamples <= 0)
        return 0;
    const int grid = (args->nClips + JB_CTA_THREADS - 1) / JB_CTA_THREADS;
    cudaStream_t st = (cudaStream_t) stream;
    ++g_launches;
    if (args->nCh == 1) // mono bus: its own instantiation of the generic kernel (jb_mono.cu)
        return check((cudaError_t) jbk_launch_mono(args, grid, stream), "jb_process_kernel<mono> launch");
    if (args->chainLen > 1 && args->exactMath)
        return check(cudaErrorInvalidValue, "exact math needs one launch per plugin (the fused kernel has the fast routines only)");
    if (args->chainLen == 1 && (!g_forceGeneric || args->exactMath)) {
        // Few live streams: one CTA per clip, everything but the analyzer's envelope walk taken off the sequential chain
        // (jb_solo.cu).  Up to g_soloLimit clips (measured crossover against the lane kernels: profiles/r02_solo.txt).
        if (jbk_solo_pick(args))
            return check((cudaError_t) jbk_launch_solo(args, stream), "jb_solo_kernel launch");
        // Two lanes per clip (jb_pair.cu) while the batch is too small to keep the schedulers busy with one: measured
        // crossovers in profiles/r01_s6_pair.txt.  JB_PAIR=0 / 1 forces a choice.
        if (jbk_pair_supported(args)) {
            const int kind = args->slot[0].kind;
            const int limit = kind == K_TEXTURE ? g_pairLimitTexture
                                                : (args->exactMath ? (kind == K_SAT ? g_pairLimitExactSat : g_pairLimitExact) : g_pairLimitFast);
            const bool pair = g_pairMode < 0 ? args->nClips <= limit : g_pairMode != 0;
            if (pair)
                return check((cudaError_t) jbk_launch_pair(args, stream), "jb_pair_kernel launch");
        }
        return check((cudaError_t) jbk_launch_single(args, grid, stream), "jb_single_kernel launch");
    }
    jb_process_kernel<false><<<grid, JB_CTA_THREADS, lane_smem_bytes(args->octets), st>>>(*args);
    return check(cudaGetLastError(), "jb_process_kernel launch");
}

int jbk_launch_fill(float* dst, float value, long long count, void* stream)
{
    if (count <= 0)
        return 0;
    long long blocks = (count + 255) / 256;
    if (blocks > 148 * 16)
        blocks = 148 * 16;
    jb_fill_kernel<<<(int) blocks, 256, 0, (cudaStream_t) stream>>>(dst, value, count);
    ++g_launches;
    return check(cudaGetLastError(), "jb_fill_kernel launch");
}

int jbk_launch_synth(float* dAudio, int kind, long long firstClip, int nClips, int nCh, int nSamples,
                     double sampleRate, unsigned int seed, void* stream)
{
    if (nClips <= 0 || nSamples <= 0)
        return 0;
    jb_synth_kernel<<<148 * 8, 256, 0, (cudaStream_t) stream>>>(dAudio, kind, firstClip, nClips, nCh, nSamples,
                                                               (float) sampleRate, seed);
    ++g_launches;
    return check(cudaGetLastError(), "jb_synth_kernel launch");
}

} // extern "C"
