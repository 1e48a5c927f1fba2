// jb_meter.cu -- meter-panel statistics of a whole render, sm_100a.
//
// The reference's editor hands getLatestMetrics() to JuicyMeterPanel::setMetrics on a timer
// (src/shared/JuicyPluginEditor.cpp:85-89); the panel smooths the bars (smoothValue,
// src/shared/JuicyMeterPanel.cpp:3-7) and keeps min / max / running-average "ghost" statistics
// (updateStats, :54-71).  Here the per-block record history the render kernels leave in HBM
// ([block][slot][16][clipPitch]) is walked once per clip, in block order, with the panel's exact
// arithmetic (sequential running average, IEEE division, no FMA -- see the Makefile flags), so the
// 40 numbers per clip are what the panel of that clip's plugin instance would hold after the render.
//
// One lane per clip: a warp reads one 128-byte line per record field per block (the history is
// structure-of-arrays over clips), 13 fields of 16 are touched.  The walk is a 6-deep dependent
// chain per block, so lanes are plentiful and cheap; the kernel is a latency-bound reduction over
// blocks * 52 B per clip and runs once per render, not per block.
#include "jb_kernels.h"

#include <cuda_runtime.h>

namespace {

__device__ __forceinline__ float smooth_value(float current, float target) // JuicyMeterPanel.cpp:3-7
{
    const float alpha = target > current ? 0.28f : 0.12f;
    return current + (target - current) * alpha;
}

struct Stat { float mn, mx, avg; };

__global__ void __launch_bounds__(128) jb_meter_kernel(const float* __restrict__ hist, long long clipPitch, int chainLen, int slot,
                                                        int firstBlock, int nBlocks, int stride, int nClips,
                                                        float* __restrict__ out)
{
    const int clip = blockIdx.x * blockDim.x + threadIdx.x;
    if (clip >= nClips)
        return;
    // `JuicinessMetrics metrics;` in the panel starts from the struct's defaults
    // (src/shared/JuicinessAnalyzer.h:6-21): monoSafety 1, everything else 0
    float preScore = 0.0f, postScore = 0.0f, score = 0.0f, punch = 0.0f, richness = 0.0f, clarity = 0.0f, width = 0.0f,
          monoSafety = 1.0f;
    Stat st[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        st[k].mn = 1.0f; // MetricStats defaults, src/shared/JuicyMeterPanel.h:16-22
        st[k].mx = 0.0f;
        st[k].avg = 0.0f;
    }
    int count = 0;
    // statistics order of setMetrics (:16-25): punch, richness, clarity, width, monoSafety, emphasis, coherence,
    // synesthesia, fatigue, repetition -> record fields
    const int statField[10] = { 8, 9, 10, 11, 12, 3, 4, 5, 6, 7 };
    for (int b = 0; b < nBlocks; b += stride) {
        const float* r = hist + ((long long) (firstBlock + b) * chainLen + slot) * JBK_REC * clipPitch + clip;
        float f[13];
#pragma unroll
        for (int i = 0; i < 13; ++i)
            f[i] = __ldg(r + (long long) i * clipPitch);
        const float newPre = f[1] > 0.0f ? f[1] : f[0];
        const float newPost = f[2] > 0.0f ? f[2] : f[0];
        preScore = smooth_value(preScore, newPre);
        postScore = smooth_value(postScore, newPost);
        ++count;
        const float n = (float) count;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const float v = fminf(fmaxf(f[statField[k]], 0.0f), 1.0f);
            if (count == 1) {
                st[k].mn = v;
                st[k].mx = v;
                st[k].avg = v;
            } else {
                st[k].mn = fminf(st[k].mn, v);
                st[k].mx = fmaxf(st[k].mx, v);
                st[k].avg += (v - st[k].avg) / n;
            }
        }
        score = smooth_value(score, newPost);
        punch = smooth_value(punch, f[8]);
        richness = smooth_value(richness, f[9]);
        clarity = smooth_value(clarity, f[10]);
        width = smooth_value(width, f[11]);
        monoSafety = smooth_value(monoSafety, f[12]);
    }
    // out: [JBK_METER][clipPitch] structure-of-arrays, unpacked to [clip][40] on the host
    float* o = out + clip;
    o[0 * clipPitch] = preScore;
    o[1 * clipPitch] = postScore;
    o[2 * clipPitch] = score;
    o[3 * clipPitch] = punch;
    o[4 * clipPitch] = richness;
    o[5 * clipPitch] = clarity;
    o[6 * clipPitch] = width;
    o[7 * clipPitch] = monoSafety;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        o[(long long) (8 + 3 * k) * clipPitch] = st[k].mn;
        o[(long long) (9 + 3 * k) * clipPitch] = st[k].mx;
        o[(long long) (10 + 3 * k) * clipPitch] = st[k].avg;
    }
    o[38 * clipPitch] = (float) count;
    o[39 * clipPitch] = 0.0f;
}

} // namespace

extern "C" int jbk_launch_meter(const float* hist, long long clipPitch, int chainLen, int slot, int firstBlock, int nBlocks,
                                int stride, int nClips, float* out, void* stream)
{
    if (nClips <= 0)
        return 0;
    const int threads = 128;
    jb_meter_kernel<<<(nClips + threads - 1) / threads, threads, 0, (cudaStream_t) stream>>>(hist, clipPitch, chainLen, slot, firstBlock,
                                                                                           nBlocks, stride, nClips, out);
    jbk_note_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
