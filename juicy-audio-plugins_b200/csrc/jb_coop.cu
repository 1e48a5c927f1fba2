// jb_coop.cu -- block-cooperative, time-parallel render kernel for sm_100a.
//
// The lane-per-clip kernel (jb_kernels.cu) needs tens of thousands of clips to fill a B200.
// This kernel is for batches that do not offer that: one persistent CTA per SM owns up to
// CO_GMAX clips and walks them through time in steps of CO_T samples, with warp roles
//
//   producer  : 1-D bulk async copies (cp.async.bulk + mbarrier) global -> shared tile and back;
//   scouts    : one lane per (clip, channel) row runs the CHEAP exact recurrences of the next
//               step sequentially (JuicyPunch's fast/slow envelopes) and drops a checkpoint of the
//               state every CO_CH samples;
//   bulk      : one warp per clip-step, one lane per CO_CH-sample chunk: restarts the recurrence
//               from the scout's checkpoint (bit-identical, same operand order) and does the
//               expensive pointwise work (pow, tanh, mid/side, gains) for all chunks in parallel,
//               plus the order-independent sums of the analyzer (tree-reduced);
//   analyzers : one lane per (clip, plugin) walks JuicinessAnalyzer's nonlinear state machine
//               (attack/release envelopes, onset cooldown, band splits) strictly in sample order,
//               one host block behind the bulk warps, reading the mono sums the bulk warps left
//               in an L2-resident scratch ring.
//
// What is exact and what is reassociated: every recurrence that feeds a discontinuity or a
// singular function (analyzer envelopes -> onset threshold, Punch envelopes -> pow at 0, Width's
// in-block width decay) is evaluated with the reference's own fp32 operation order; only the plain
// sums of analyze() (rms, side, corr, getRMSLevel) are tree-reduced, and per-sample pow/tanh use
// MUFU-based evaluations accurate to ~1e-6 relative (the tests hold 1e-5 of clip peak).
//
// Reference lines are cited per routine, relative to /root/reference.
#include "jb_device.cuh"
#include "jb_libm.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace {

using namespace jbdev;

constexpr int CO_T = 256;            // samples per step
constexpr int CO_CH = 8;             // samples per bulk lane per step (32 lanes x 8 = CO_T)
constexpr int CO_GMAX = 32;          // clips per CTA group
constexpr int CO_ROWS = 2 * CO_GMAX; // (clip, channel) rows of a tile
constexpr int CO_PITCH = CO_T + 4;   // floats per tile row (+16 B: row-walking lanes hit distinct banks)
constexpr int CO_BLOCKMAX = 512;     // largest host block size this path takes
constexpr int CO_MAXCHAIN = 3;       // plugins per chain on this path
constexpr int CO_NSIG = CO_MAXCHAIN + 1;
// Warp roles (the SM's arbiter favours high warp ids, so the longest sequential chains sit on top):
// [0] producer, [1 .. nBulk] bulk, then 2 scouts, nAna "band" analyzer warps and, highest, nAna
// "envelope" analyzer warps (nAna = ceil(chainLen * groupClips / 32), decided per launch).
constexpr int CO_NSCOUT = CO_ROWS / 32;                 // 2 warps
constexpr int CO_ANA_LANES = CO_GMAX * CO_MAXCHAIN;     // 96
#ifndef JB_CO_WARPS
#define JB_CO_WARPS 16
#endif
#ifndef JB_CO_ISOLATE
#define JB_CO_ISOLATE 1   // 1: envelope warps alone on SM sub-partition 3; 0: roles by plain warp index
#endif
#ifndef JB_CO_BANDARRIVE
#define JB_CO_BANDARRIVE 1 // band analyzer warps arrive at the hand-off barrier instead of waiting on it
#endif
#ifndef JB_CO_ROLLCH
#define JB_CO_ROLLCH -1    // Punch's two channels through ONE copy of the chunk code (rolled loop of two trips): 1 always, 0 never,
#endif                     // -1 with the exact routines only (their code is 3x the fast ones'; the fast kernel gains nothing)
#ifndef JB_CO_COLDROLL
#define JB_CO_COLDROLL 1   // unroll factor of the exact routines' general-form loops (8 = fully unrolled): code size beats ILP here
#endif
#ifndef JB_CO_HOTROLL
#define JB_CO_HOTROLL 2    // unroll factor of the k = 0 tanhf loop (8 / 4 / 2 / 1 measured alike on drum hits; mixed clips 8.1 / 7.2 / 7.0 / 6.9 ms)
#endif
constexpr int CO_COLDROLL = JB_CO_COLDROLL;
constexpr int CO_HOTROLL = JB_CO_HOTROLL;
constexpr int CO_WARPS = JB_CO_WARPS;
constexpr int CO_THREADS = CO_WARPS * 32;               // 672
constexpr int CO_BAR_ANA = 1;                           // named barriers 1, 2 of the analyzer warps (by call parity)

struct CoopSmem {
    // first, so that every 256-byte row is 256-byte aligned (the swizzle xors address bits 4..7)
    float4 anaRing[2][CO_ANA_LANES][16];        //  49,152 B  per analyzer lane (envelope / band role): 64 mono samples in flight
    float tile[2][CO_ROWS][CO_PITCH];           // 133,120 B
    float2 ckpt[2][CO_ROWS][CO_T / CO_CH];      //  32,768 B  Punch (fast, slow) at every chunk start
    float stats[2][CO_GMAX][CO_NSIG][8];        //   8,192 B  sum m^2, peak, sum l^2, sum r^2, sum l*r per (block parity, clip, signal)
    float bandAcc[2][CO_ANA_LANES][2];          //   low/high band energies handed from the band lanes to the envelope lanes
    float widthTab[CO_BLOCKMAX + 1];            //   width * dyn^k, k multiplies applied in sample order
    int widthPos[CO_GMAX];                      //   ring write position of the step's first sample
    int widthCount[CO_GMAX];                    //   multiplies so far in the current host block
    unsigned long long bar[2];                  //   mbarriers: tile slot filled
    int claim[2];                               //   next unclaimed clip of the step in tile slot 0 / 1 (bulk work queue)
};

// ---------------------------------------------------------------- PTX wrappers (async proxy)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// L2 policies: the audio streams through once (evict_first) so that it does not push the analyzer
// lanes' mono rings (evict_last), which are re-read one host block later, out to HBM.
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load(void* smemDst, const void* gmemSrc, uint32_t bytes, unsigned long long* bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smemDst)),
                 "l"(gmemSrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmemDst, const void* smemSrc, uint32_t bytes, uint64_t policy)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmemDst),
                 "r"(smem_u32(smemSrc)), "r"(bytes), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void st_hint(float4* dst, float4 v, uint64_t policy)
{
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
                 "l"(policy)
                 : "memory");
}
__device__ __forceinline__ float4 ld_hint(const float4* src, uint64_t policy)
{
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(src), "l"(policy)
                 : "memory");
    return v;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- time cursor over the call's steps
struct Cursor {
    int blk, off, pos, n, nBlk; // host block, offset inside it, absolute sample, step length, block length
    bool valid;
};
__device__ __forceinline__ Cursor cursor_at(const ProcArgs& a, int blk, int off)
{
    Cursor c;
    c.blk = blk;
    c.off = off;
    c.pos = blk * a.blockSize + off;
    c.valid = c.pos < a.nSamples;
    c.nBlk = min(a.blockSize, a.nSamples - blk * a.blockSize);
    c.n = c.valid ? min(CO_T, c.nBlk - off) : 0;
    return c;
}
__device__ __forceinline__ Cursor cursor_next(const ProcArgs& a, const Cursor& c)
{
    if (c.off + c.n < c.nBlk)
        return cursor_at(a, c.blk, c.off + c.n);
    return cursor_at(a, c.blk + 1, 0);
}

// ---------------------------------------------------------------- bulk: per-plugin chunk transforms

// JuicyPunch/PluginProcessor.cpp:94-110 on one channel's chunk, restarting the two linear
// envelopes from the scout's checkpoint (same operand order as the scout => same bits).
// EXACT: the C library's own pow / tanh (jb_libm.h) and the reference's unfused operand order throughout, so the
// samples are the reference's bit for bit (ProcArgs::exactMath: something downstream thresholds or amplifies them);
// otherwise the MUFU-based routines and fused multiply-adds in the pointwise part (within 3e-6 of clip peak).
template <bool EXACT>
__device__ __forceinline__ void punch_chunk(float (&x)[CO_CH], float2 ck, const PunchCoef& c)
{
    float fEnv = ck.x, sEnv = ck.y;
    if (EXACT) {
        // The lanes of a bulk warp hold consecutive chunks of ONE channel of ONE clip, so what a sample needs from the exact
        // routines is nearly always the same for the whole warp, and two votes per chunk pick the cheap form for all of it:
        //   * pow(0, e) = 0 (e > 0): no transient anywhere in the step -- the decay of a hit, silence, steady material --
        //     skips the 8 powf calls;
        //   * |wet * drive| <= 0.25 ln2 on all of them: tanhf's k = 0 form (jb_libm.h), under half the operations.
        // Either form gives the C library's bits; the votes only decide how much is computed.  (Called by whole warps.)
        float tr[CO_CH], sus[CO_CH], trMax = 0.0f;
#pragma unroll
        for (int i = 0; i < CO_CH; ++i) {
            const float adry = fabsf(x[i]);
            fEnv = c.omFast * adry + c.fastCoeff * fEnv;
            sEnv = c.omSlow * adry + c.slowCoeff * sEnv;
            tr[i] = jmaxf(0.0f, fEnv - sEnv);
            sus[i] = jmaxf(0.0f, sEnv - tr[i] * 0.6f);
            trMax = fmaxf(trMax, tr[i]);
        }
        float curve[CO_CH];
        if (__any_sync(0xffffffffu, trMax > 0.0f)) {
#pragma unroll CO_COLDROLL
            for (int i = 0; i < CO_CH; ++i)
                curve[i] = jblibm::powf_glibc_pos(tr[i], c.curveExp);
        } else {
#pragma unroll
            for (int i = 0; i < CO_CH; ++i)
                curve[i] = 0.0f;
        }
        float wet[CO_CH], th[CO_CH];
        uint32_t dMax = 0;
#pragma unroll
        for (int i = 0; i < CO_CH; ++i) {
            const float punchGain = 1.0f + c.punchK * curve[i];
            const float sustainGain = 1.0f + c.sustainK * sus[i];
            wet[i] = x[i] * punchGain * sustainGain;
            th[i] = wet[i] * c.drive;
            dMax = max(dMax, __float_as_uint(th[i]) & 0x7fffffffu);
        }
        if (__all_sync(0xffffffffu, dMax <= jblibm::kTanhSmallMaxBits)) {
#pragma unroll CO_HOTROLL
            for (int i = 0; i < CO_CH; ++i)
                th[i] = jblibm::tanhf_fdlibm_small(th[i]);
        } else {
#pragma unroll CO_COLDROLL
            for (int i = 0; i < CO_CH; ++i)
                th[i] = jblibm::tanhf_fdlibm(th[i]);
        }
#pragma unroll
        for (int i = 0; i < CO_CH; ++i) {
            const float dry = x[i];
            const float soft = jblibm::fdiv(th[i], c.tanhDrive);
            const float hard = jlimitf(-0.95f, 0.95f, wet[i] * c.hardK);
            const float w = soft + c.clipAmt * (hard - soft);
            x[i] = (dry + c.mix * (w - dry)) * c.outGain;
        }
        return;
    }
    const float invTanhDrive = 1.0f / c.tanhDrive;
#pragma unroll
    for (int i = 0; i < CO_CH; ++i) {
        const float dry = x[i];
        const float adry = fabsf(dry);
        fEnv = c.omFast * adry + c.fastCoeff * fEnv;
        sEnv = c.omSlow * adry + c.slowCoeff * sEnv;
        const float transient = jmaxf(0.0f, fEnv - sEnv);
        // from here on pointwise: fused multiply-adds (1e-7 relative, inside the sample tolerance)
        const float transientCurve = pow_unit(transient, c.curveExp);
        const float punchGain = fmaf(c.punchK, transientCurve, 1.0f);
        const float sustainGain = fmaf(c.sustainK, jmaxf(0.0f, fmaf(-0.6f, transient, sEnv)), 1.0f);
        float wet = dry * punchGain * sustainGain;
        const float soft = tanh_fast(wet * c.drive) * invTanhDrive;
        const float hard = jlimitf(-0.95f, 0.95f, wet * c.hardK);
        wet = fmaf(c.clipAmt, hard - soft, soft);
        x[i] = fmaf(c.mix, wet - dry, dry) * c.outGain;
    }
}

// JuicyWidth/PluginProcessor.cpp:106-137 on one clip's chunk.  width after k in-block multiplies
// comes from the table (bit-identical to the reference's repeated `width *= dynamicLimit`);
// the right channel's wet signal goes through the 60 ms ring in global memory (clip-major).
// The delayed wet-R samples of a step lie in ring positions written by EARLIER steps whenever the
// delay is at least a step long (and jbk_coop_supported keeps ringLen - delay >= a step, so this
// step's writes cannot land on them either): then they are fetched at the top of the clip-step
// and their L2 latency hides behind the stages in front of Width.
struct WidthPrefetch {
    float4 v0, v1;
    bool have;
};
__device__ __forceinline__ WidthPrefetch width_prefetch(const WidthCoef& c, const float* ring, int wpos0, int lane, int nValid)
{
    WidthPrefetch pf;
    pf.v0 = pf.v1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    pf.have = c.delaySamples >= CO_T && ((wpos0 | c.ringLen | c.delaySamples) & 3) == 0; // warp-uniform
    if (pf.have) {
        int wp = wpos0 + lane * CO_CH;
        wp -= wp >= c.ringLen ? c.ringLen : 0;
        int rp0 = wp - c.delaySamples;
        rp0 += rp0 < 0 ? c.ringLen : 0;
        int p1 = wp + 4;
        p1 -= p1 >= c.ringLen ? c.ringLen : 0;
        int rp1 = p1 - c.delaySamples;
        rp1 += rp1 < 0 ? c.ringLen : 0;
        if (nValid > 0)
            pf.v0 = __ldcg(reinterpret_cast<const float4*>(ring + rp0));
        if (nValid > 4)
            pf.v1 = __ldcg(reinterpret_cast<const float4*>(ring + rp1));
    }
    return pf;
}
template <bool EXACT>
__device__ __forceinline__ void width_chunk(float (&l)[CO_CH], float (&r)[CO_CH], const WidthCoef& c, const float* tab,
                                            int kStart, int& warpTotal, float* ring, int wpos0, int lane, int nValid,
                                            const WidthPrefetch& pf)
{
    int kk[CO_CH];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < CO_CH; ++i) {
        const float corrProxy = jlimitf(-1.0f, 1.0f, l[i] * r[i] * 12.0f);
        cnt += (corrProxy < -0.1f) ? 1 : 0;
        kk[i] = cnt;
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += t;
    }
    warpTotal = __shfl_sync(0xffffffffu, incl, 31);
    const int base = kStart + incl - cnt;

    float wetL[CO_CH], wetR[CO_CH];
#pragma unroll
    for (int i = 0; i < CO_CH; ++i) {
        const float width = tab[base + kk[i]];
        const float mid = 0.5f * (l[i] + r[i]);
        const float side = 0.5f * (l[i] - r[i]) * (1.0f + width);
        wetL[i] = mid + side;
        wetR[i] = mid - side;
    }
    int wp = wpos0 + lane * CO_CH;
    wp -= wp >= c.ringLen ? c.ringLen : 0; // wpos0 < ringLen and lane*CO_CH < CO_T <= ringLen
    // 16-byte ring accesses when write position, delay and ring length are all multiples of 4 samples
    // (the usual case: 2880-sample ring, 576-sample delay, steps of 256); single floats otherwise.
    const bool vec = ((wpos0 | c.ringLen | c.delaySamples) & 3) == 0;
    if (vec) {
#pragma unroll
        for (int g = 0; g < CO_CH / 4; ++g) {
            int p = wp + 4 * g;
            p -= p >= c.ringLen ? c.ringLen : 0;
            if (4 * g < nValid)
                *reinterpret_cast<float4*>(ring + p) = make_float4(wetR[4 * g], wetR[4 * g + 1], wetR[4 * g + 2], wetR[4 * g + 3]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < CO_CH; ++i) {
            int p = wp + i;
            p -= p >= c.ringLen ? c.ringLen : 0;
            if (i < nValid)
                ring[p] = wetR[i];
        }
    }
    if (pf.have) {
        wetR[0] = pf.v0.x; wetR[1] = pf.v0.y; wetR[2] = pf.v0.z; wetR[3] = pf.v0.w;
        wetR[4] = pf.v1.x; wetR[5] = pf.v1.y; wetR[6] = pf.v1.z; wetR[7] = pf.v1.w;
    } else {
        __syncwarp(); // the delayed read below may land on a value another lane wrote in this step
        if (vec) {
#pragma unroll
            for (int g = 0; g < CO_CH / 4; ++g) {
                int p = wp + 4 * g;
                p -= p >= c.ringLen ? c.ringLen : 0;
                int rp = p - c.delaySamples;
                rp += rp < 0 ? c.ringLen : 0;
                float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (4 * g < nValid)
                    v = __ldcg(reinterpret_cast<const float4*>(ring + rp));
                wetR[4 * g] = v.x; wetR[4 * g + 1] = v.y; wetR[4 * g + 2] = v.z; wetR[4 * g + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < CO_CH; ++i) {
                int p = wp + i;
                p -= p >= c.ringLen ? c.ringLen : 0;
                int rp = p - c.delaySamples;
                rp += rp < 0 ? c.ringLen : 0;
                wetR[i] = (i < nValid) ? __ldcg(ring + rp) : 0.0f;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CO_CH; ++i) {
        const float dryL = l[i], dryR = r[i];
        if (EXACT) { // the reference's two roundings (JuicyWidth/PluginProcessor.cpp:134-135)
            l[i] = (dryL + c.mix * (wetL[i] - dryL)) * c.outGain;
            r[i] = (dryR + c.mix * (wetR[i] - dryR)) * c.outGain;
        } else {
            l[i] = fmaf(c.mix, wetL[i] - dryL, dryL) * c.outGain;
            r[i] = fmaf(c.mix, wetR[i] - dryR, dryR) * c.outGain;
        }
    }
}

// JuicyInfer/PluginProcessor.cpp:79: buffer.applyGain(trimGain)
__device__ __forceinline__ void infer_chunk(float (&l)[CO_CH], float (&r)[CO_CH], const InferCoef& c)
{
    if (c.gainMode == 1) {
#pragma unroll
        for (int i = 0; i < CO_CH; ++i) {
            l[i] *= c.trimGain;
            r[i] *= c.trimGain;
        }
    } else if (c.gainMode == 2) {
#pragma unroll
        for (int i = 0; i < CO_CH; ++i) {
            l[i] = 0.0f;
            r[i] = 0.0f;
        }
    }
}

// Samples past the end of a (ragged) step carry nothing into sums, rings or later stages.
__device__ __forceinline__ void mask_chunk(float (&l)[CO_CH], float (&r)[CO_CH], int nValid)
{
#pragma unroll
    for (int i = 0; i < CO_CH; ++i) {
        if (i >= nValid) {
            l[i] = 0.0f;
            r[i] = 0.0f;
        }
    }
}

// Sum of four values over the warp in 6 shuffles: each butterfly stage halves the values a lane carries.
// The totals land on lane 0 (every lane with lane % 4 == 0 ends up holding one of the four).
__device__ __forceinline__ void warp_sum4(float& v0, float& v1, float& v2, float& v3, int lane)
{
    {   // stage 1 (xor 16): lower half keeps (v0, v1), upper half keeps (v2, v3)
        const bool hi = (lane & 16) != 0;
        const float s0 = hi ? v0 : v2, s1 = hi ? v1 : v3;
        const float k0 = hi ? v2 : v0, k1 = hi ? v3 : v1;
        v0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
        v1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
    }
    {   // stage 2 (xor 8): keep one of the two
        const bool hi = (lane & 8) != 0;
        const float send = hi ? v0 : v1, keep = hi ? v1 : v0;
        v0 = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    v0 += __shfl_xor_sync(0xffffffffu, v0, 4);
    v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
    v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
    // lane bits (16, 8) now select which total v0 is: 00 -> v0, 01 -> v1, 10 -> v2, 11 -> v3
    v1 = __shfl_sync(0xffffffffu, v0, 8);
    v2 = __shfl_sync(0xffffffffu, v0, 16);
    v3 = __shfl_sync(0xffffffffu, v0, 24);
}

// The order-independent sums of analyze() over this lane's chunk of one signal
// (JuicinessAnalyzer.cpp:76-77,86-91,105-106), tree-reduced over the warp and added to the
// block's running totals.  Kept: sum mono^2, max |mono|, sum l^2, sum r^2, sum l*r; the side
// energy follows from mid^2 + side^2 = (l^2 + r^2) / 2 (load_stats).  The mono sum of every sample
// goes to the analyzer lanes' scratch ring.
__device__ __forceinline__ void signal_stats(const float (&l)[CO_CH], const float (&r)[CO_CH], float* acc, bool firstStep,
                                             float* monoDst, int lane, int nValid, uint64_t keep)
{
    float mono[CO_CH];
    float rms = 0.0f, peak = 0.0f, corr = 0.0f, l2 = 0.0f, r2 = 0.0f;
#pragma unroll
    for (int i = 0; i < CO_CH; ++i) {
        const float m = 0.5f * (l[i] + r[i]);
        mono[i] = m;
        rms = fmaf(m, m, rms);
        peak = fmaxf(peak, fabsf(m));
        corr = fmaf(l[i], r[i], corr);
        l2 = fmaf(l[i], l[i], l2);
        r2 = fmaf(r[i], r[i], r2);
    }
    if (nValid > 0)
        st_hint(reinterpret_cast<float4*>(monoDst), make_float4(mono[0], mono[1], mono[2], mono[3]), keep);
    if (nValid > 4)
        st_hint(reinterpret_cast<float4*>(monoDst) + 1, make_float4(mono[4], mono[5], mono[6], mono[7]), keep);
    warp_sum4(rms, l2, r2, corr, lane);
    peak = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(peak))); // non-negative floats order like their bits
    if (lane == 0) {
        if (firstStep) {
            acc[0] = rms; acc[1] = peak; acc[2] = l2; acc[3] = r2; acc[4] = corr;
        } else {
            acc[0] += rms; acc[1] = fmaxf(acc[1], peak); acc[2] += l2; acc[3] += r2; acc[4] += corr;
        }
    }
}

// ---------------------------------------------------------------- analyzer lanes
// One lane walks one analyzer's samples in order.  BANDS = false: envelopes + onset machine;
// BANDS = true: the two band-split one-poles.  Mono samples come from the L2-resident scratch ring,
// 16 samples (4 x 16 B) prefetched ahead.
// Envelope lanes: the two attack/release envelopes (updateEnvelope, JuicinessAnalyzer.cpp:24-29,64-65),
// the transient sum (:66-67) and the largest transient of the group, over 4 samples.  The onset
// machine (:69-75) is not stepped per sample: a group of 8 can hold an onset only if its largest
// transient exceeds the threshold while the cooldown allows one, and only then is it replayed
// sample by sample (onset_replay) -- same decisions, a fraction of the instructions.
__device__ __forceinline__ void env_quad(AnaState& s, float& trAcc, float& gmax, const float4 q, const AnaCoef& c)
{
    const float m[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = fabsf(m[j]);
        ana_env_update(s.sEnv, s.lEnv, a, c); // both envelopes' multiplies in packed halves (jb_device.cuh): same bits, 2 issue slots fewer
        const float tr = fmaxf(0.0f, s.sEnv - s.lEnv);
        trAcc += tr;
        gmax = fmaxf(gmax, tr);
    }
}
// `rem` restates onsetCooldown as "samples from the group's first until an onset is allowed again":
// the reference decrements the counter once per sample and accepts an onset when it has reached 0,
// i.e. at the len-th sample after the previous onset.  Replays the group's 8 samples from the
// envelope state at its start (sEnv, lEnv) and applies the onset decisions.  Rare (once per onset).
__device__ __noinline__ void onset_replay(float sEnv, float lEnv, int& rem, int& onsets, const float4 q0, const float4 q1,
                                          const AnaCoef& c)
{
    const float m[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
        const float a = fabsf(m[j]);
        ana_env_update(sEnv, lEnv, a, c);
        const float tr = fmaxf(0.0f, sEnv - lEnv);
        const bool onset = (tr > 0.045f) & (rem <= j);
        onsets += onset ? 1 : 0;
        rem = onset ? c.cooldownLen + j : rem;
    }
}
// Band lanes: the two band-split one-poles in the reference's order (:79-82); their energies (:83-84)
// are plain sums feeding continuous features, accumulated with fused multiply-adds.
__device__ __forceinline__ void band_quad(AnaState& s, AnaAcc& acc, const float4 q, const AnaCoef& c)
{
    const float m[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s.low += c.lowCoeff * (m[j] - s.low);
        s.high += c.highCoeff * (m[j] - s.high);
        const float hi = m[j] - s.high;
        acc.lowAcc = fmaf(s.low, s.low, acc.lowAcc);
        acc.highAcc = fmaf(hi, hi, acc.highAcc);
    }
}
// 8 samples of one analyzer lane
template <bool BANDS>
__device__ __forceinline__ void ana_group(AnaState& st, AnaAcc& acc, int& rem, const float4 q0, const float4 q1, const AnaCoef& c)
{
    if (BANDS) {
        band_quad(st, acc, q0, c); band_quad(st, acc, q1, c);
    } else {
        const float sEnv0 = st.sEnv, lEnv0 = st.lEnv;
        float gmax = 0.0f;
        env_quad(st, acc.trAcc, gmax, q0, c); env_quad(st, acc.trAcc, gmax, q1, c);
        if (gmax > 0.045f && rem <= 7)
            onset_replay(sEnv0, lEnv0, rem, acc.onsets, q0, q1, c);
        rem -= 8;
    }
}

// One lane walks one analyzer's n samples in order.  BANDS = false: envelopes + onset machine;
// BANDS = true: the two band-split one-poles.  The lane's mono stream lies in the L2-resident scratch
// ring; the lane copies it asynchronously (cp.async, 16 B pieces, no register landing, no scoreboard
// wait) into its own 64-sample row of shared memory 48 samples ahead of use, and reads it back one
// 8-sample group ahead -- so neither L2/HBM latency nor the shared-memory load sits on the
// recurrence.  The loop body is kept to 16 samples: with five roles running different code on one SM
// the instruction caches, not the issue slots, were the first limit (profiles/r01_v4_*).
// Piece k of a row is stored at k ^ (lane & 7): the 8 lanes of a quarter-warp reading the same piece
// hit 8 different 16-byte bank groups.
__device__ __forceinline__ void cp_async16(uint32_t smemDst, const void* gmemSrc, uint64_t policy)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smemDst), "l"(gmemSrc), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int CO_FEED_AHEAD = 6; // groups of 8 samples in flight per lane (ring of 8 groups)
constexpr int CO_FEED_PAD = 8 * (CO_FEED_AHEAD + 2); // floats a walk may copy (never use) past its stream's end

__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

struct AnaFeed {
    const float* stream; // global: this lane's mono samples of the call
    uint32_t laneBase;   // shared: address of this lane's 256-byte ring row, xor 16 * (lane & 7) (the swizzle)
    uint64_t policy;
    // group g -> ring pieces 2 (g & 7), +1.  Unconditional: groups past the call's end copy scratch that is never used.
    __device__ __forceinline__ void issue(int g) const
    {
        const float* src = stream + 8 * g;
        const uint32_t off = (uint32_t) (g & 7) << 5;
        cp_async16(laneBase ^ off, src, policy);
        cp_async16(laneBase ^ (off | 16u), src + 4, policy);
        cp_async_commit();
    }
    __device__ __forceinline__ void read(int g, float4& q0, float4& q1) const
    {
        const uint32_t off = (uint32_t) (g & 7) << 5;
        q0 = lds128(laneBase ^ off);
        q1 = lds128(laneBase ^ (off | 16u));
    }
};

template <bool BANDS>
__device__ __forceinline__ void ana_walk(AnaState& st, AnaAcc& acc, const AnaFeed& f, int n, const AnaCoef& c)
{
    const int nGroups = n >> 3;  // the path requires n % 4 == 0
    const int nPairs = nGroups >> 1;
    int rem = st.cool - 1;       // onsetCooldown -> samples until the next onset may fire (cool = 0 or 1: immediately)
    float4 a0, a1, b0, b1;
#pragma unroll
    for (int g = 0; g < CO_FEED_AHEAD; ++g)
        f.issue(g);
    cp_async_wait<CO_FEED_AHEAD - 1>(); // group 0 has landed
    f.read(0, a0, a1);
    int g = 0;
#pragma unroll 1
    for (int pr = 0; pr < nPairs; ++pr, g += 2) { // 16 samples per trip, no conditions but the trip count
        f.issue(g + CO_FEED_AHEAD);          // its slot held group g - 2 (consumed)
        cp_async_wait<CO_FEED_AHEAD - 1>();  // groups <= g + 1 have landed
        f.read(g + 1, b0, b1);
        ana_group<BANDS>(st, acc, rem, a0, a1, c);
        f.issue(g + 1 + CO_FEED_AHEAD);
        cp_async_wait<CO_FEED_AHEAD - 1>();
        f.read(g + 2, a0, a1);
        ana_group<BANDS>(st, acc, rem, b0, b1, c);
    }
    cp_async_wait<0>();
    if (g < nGroups) {           // odd group count
        ana_group<BANDS>(st, acc, rem, a0, a1, c);
        ++g;
    }
    if (!BANDS) // back to the reference's counter for the tail and for the state arrays
        st.cool = max(rem + 1, 0);
    if (n & 4) {                 // n % 8 == 4: one last quad
        const float4 v = ld_hint(reinterpret_cast<const float4*>(f.stream) + 2 * nGroups, f.policy);
        if (BANDS) {
            band_quad(st, acc, v, c);
        } else {
            ana_step_env(st, acc, v.x, c); ana_step_env(st, acc, v.y, c);
            ana_step_env(st, acc, v.z, c); ana_step_env(st, acc, v.w, c);
        }
    }
}

__device__ __forceinline__ StatSums load_stats(const float* s)
{
    StatSums t;
    t.rms = s[0]; t.peak = s[1];
    t.l2 = (double) s[2]; t.r2 = (double) s[3];
    t.corr = s[4];
    t.side = jmaxf(0.0f, 0.5f * (s[2] + s[3]) - s[0]); // sum side^2 = sum (l^2 + r^2) / 2 - sum mid^2
    return t;
}

__device__ __forceinline__ void named_barrier(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct CoopArgs {
    ProcArgs p;
    float* monoScratch; // [grid][CO_GMAX][chainLen + 1][2][CO_BLOCKMAX]
    int groupClips;     // clips per group (<= CO_GMAX)
    int numGroups;
    int debugSkip;      // profiling builds (-DJB_COOP_DEBUG, JB_COOP_DEBUG_SKIP): bit 0 envelope walk, 1 band walk, 2 bulk math, 3 scout
};

template <bool EXACT, bool ISOLATE>
__global__ void __launch_bounds__(CO_THREADS, 1) jb_coop_kernel(const __grid_constant__ CoopArgs ca)
{
    extern __shared__ __align__(1024) unsigned char smemRaw[];
    CoopSmem& sm = *reinterpret_cast<CoopSmem*>(smemRaw);
    const ProcArgs& a = ca.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = a.chainLen;
    const int nSig = L + 1;
    const AnaCoef& ana = a.ana; // stays in the kernel-parameter constant bank
#ifdef JB_COOP_DEBUG
    const int dbgSkip = ca.debugSkip; // profiling builds only (-DJB_COOP_DEBUG): skip roles to see what the others cost
#else
    constexpr int dbgSkip = 0;        // release builds: no environment variable can change what is rendered
#endif

    // roles (uniform over the launch)
    const int nAna = (L * ca.groupClips + 31) / 32;          // envelope-analyzer warps; as many band-analyzer warps
    // A warp's scheduler (SM sub-partition) is warp % 4.  The envelope warps carry the kernel's critical
    // sequential chain and every one of their instructions is on it, so they get sub-partition 3 to
    // themselves (its other warps idle); everybody else shares sub-partitions 0..2.
    // ISOLATE = false (exact math: the bulk warps' pow / tanh dominate and want all four schedulers): roles by plain warp
    // index, envelope warps highest.
    const bool onSeqPartition = ISOLATE ? (warp & 3) == 3 : warp >= CO_WARPS - nAna;
    const int seqIdx = ISOLATE ? warp >> 2 : warp - (CO_WARPS - nAna);  // index among the warps of sub-partition 3
    const int parIdx = ISOLATE ? warp - ((warp + 1) >> 2) : warp;       // index among the others
    const int wScout = 1, wBand = wScout + CO_NSCOUT, wBulk = wBand + nAna;
    const bool isProducerWarp = !onSeqPartition && parIdx == 0;
    const bool isScoutWarp = !onSeqPartition && parIdx >= wScout && parIdx < wBand;
    const bool isBandWarp = !onSeqPartition && parIdx >= wBand && parIdx < wBulk;
    const bool isBulkWarp = !onSeqPartition && parIdx >= wBulk;
    const bool isEnvWarp = onSeqPartition && seqIdx < nAna;
    const int anaThreads = 2 * nAna * 32;
    const uint64_t polStream = policy_evict_first(), polKeep = policy_evict_last();

    if (threadIdx.x == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // which slot (if any) is Width; Punch may only lead the chain (its scout reads the raw input)
    int widthSlot = -1;
    for (int s = 0; s < L; ++s)
        if (a.slot[s].kind == K_WIDTH)
            widthSlot = s;
    const bool punchFirst = a.slot[0].kind == K_PUNCH;
    if (widthSlot >= 0 && threadIdx.x == 32) { // width after k multiplies, in the reference's order (:93,:112)
        const WidthCoef& w = a.slot[widthSlot].c.width;
        float v = w.width;
        sm.widthTab[0] = v;
        for (int k = 1; k <= CO_BLOCKMAX; ++k) {
            v *= w.dynamicLimit;
            sm.widthTab[k] = v;
        }
    }
    __syncthreads();

    unsigned gstep = 0; // steps loaded so far by this CTA (tile slot = gstep & 1, mbarrier parity = (gstep >> 1) & 1)
    float* const monoCta = ca.monoScratch + (size_t) blockIdx.x * CO_GMAX * nSig * 2 * CO_BLOCKMAX;

    for (int group = blockIdx.x; group < ca.numGroups; group += gridDim.x) {
        const int clip0 = group * ca.groupClips;
        const int G = min(ca.groupClips, a.nClips - clip0);
        const int rows = 2 * G;

        // ---- per-role state for this group
        // scout: Punch envelopes of row (clip, ch)
        float scF = 0.0f, scS = 0.0f;
        const int scRow = (parIdx - wScout) * 32 + lane;
        const bool isScout = isScoutWarp && punchFirst && scRow < rows;
        if (isScout) {
            const long long clip = clip0 + (scRow >> 1);
            const int b = a.slot[0].stateBase + AV_COUNT;
            scF = a.state[(long long) (b + PV_FAST0 + (scRow & 1)) * a.clipPitch + clip];
            scS = a.state[(long long) (b + PV_SLOW0 + (scRow & 1)) * a.clipPitch + clip];
        }
        // analyzers: lane <-> (plugin slot, clip), once in an envelope warp and once in a band warp
        const int anaIdx = (isBandWarp ? parIdx - wBand : seqIdx) * 32 + lane;
        const int anaSlot = anaIdx / ca.groupClips, anaClip = anaIdx % ca.groupClips;
        const bool isAna = (isEnvWarp || isBandWarp) && anaSlot < L && anaClip < G;
        // shared-memory ring row of this analyzer lane, swizzle folded in; pinned to a register
        uint32_t anaLaneBase = smem_u32(&sm.anaRing[isBandWarp ? 1 : 0][isAna ? anaIdx : 0][0]) ^ ((uint32_t) (lane & 7) << 4);
        asm volatile("" : "+r"(anaLaneBase));
        AnaState ast {};
        float preScore = 0.0f;
        if (isAna) {
            const long long clip = clip0 + anaClip;
            const int b = a.slot[anaSlot].stateBase;
            auto ld = [&](int v) { return a.state[(long long) (b + v) * a.clipPitch + clip]; };
            ast.sEnv = ld(AV_SHORT); ast.lEnv = ld(AV_LONG); ast.low = ld(AV_LOW); ast.high = ld(AV_HIGH);
            ast.repEma = ld(AV_REP_EMA); ast.fatEma = ld(AV_FAT_EMA);
            ast.cool = __float_as_int(ld(AV_COOLDOWN));
            preScore = ld(AV_PRE_SCORE);
        }
        if (widthSlot >= 0 && threadIdx.x < G) {
            const int b = a.slot[widthSlot].stateBase + AV_COUNT;
            sm.widthPos[threadIdx.x] = __float_as_int(a.state[(long long) (b + WV_WPOS) * a.clipPitch + clip0 + threadIdx.x]);
            sm.widthCount[threadIdx.x] = 0;
        }
        if (threadIdx.x == 0)
            sm.claim[gstep & 1] = 0;

        Cursor cur = cursor_at(a, 0, 0);
        auto issue_load = [&](const Cursor& c, unsigned step) {
            // producer warp: one bulk copy per (clip, channel) row
            const int slot = step & 1;
            if (lane == 0)
                mbar_expect_tx(&sm.bar[slot], (uint32_t) (rows * c.n * 4));
            __syncwarp();
            for (int row = lane; row < rows; row += 32) {
                const float* src = a.in + ((long long) (clip0 + (row >> 1)) * 2 + (row & 1)) * a.rowPitch + c.pos;
                bulk_load(&sm.tile[slot][row][0], src, (uint32_t) (c.n * 4), &sm.bar[slot], polStream);
            }
        };
        auto scout_step = [&](const Cursor& c, unsigned step) {
            const int slot = step & 1;
            mbar_wait(&sm.bar[slot], (step >> 1) & 1);
            if (!isScout || (dbgSkip & 8))
                return;
            const PunchCoef& pc = a.slot[0].c.punch;
            const float4* src = reinterpret_cast<const float4*>(&sm.tile[slot][scRow][0]);
            float2* ck = &sm.ckpt[slot][scRow][0];
            for (int i = 0; i < c.n; i += CO_CH) {
                ck[i / CO_CH] = make_float2(scF, scS);
                const float4 v0 = src[i / 4], v1 = src[i / 4 + 1];
                const float x[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
#pragma unroll
                for (int k = 0; k < CO_CH; ++k) {
                    if (i + k < c.n) { // JuicyPunch/PluginProcessor.cpp:94-96
                        const float adry = fabsf(x[k]);
                        scF = pc.omFast * adry + pc.fastCoeff * scF;
                        scS = pc.omSlow * adry + pc.slowCoeff * scS;
                    }
                }
            }
        };
        // One analyze() call of host block `blk` on signal `sig` (its mono samples and plain sums were left by
        // the bulk warps): both analyzer roles walk, the band lanes hand their energies over, the envelope
        // lanes finish.  Called by all lanes of the analyzer warps (named barrier inside).
        int anaCalls = 0;
        auto analyze = [&](int blk, int n, int sig) -> Metrics {
            const int par = blk & 1, hand = anaCalls & 1;
            ++anaCalls;
            AnaAcc acc;
            const float* stream = monoCta + ((size_t) anaClip * nSig + sig) * 2 * CO_BLOCKMAX + par * CO_BLOCKMAX;
            if (dbgSkip & 16) // probe: every walk re-reads one L1-resident kilobyte per lane
                stream = monoCta + (size_t) anaIdx * 256;
            AnaFeed feed;
            feed.stream = stream;
            feed.laneBase = anaLaneBase;
            feed.policy = polKeep;
            if (isAna && !(dbgSkip & (isBandWarp ? 2 : 1))) {
                if (isBandWarp) {
                    ana_walk<true>(ast, acc, feed, n, ana);
                    sm.bandAcc[hand][anaIdx][0] = acc.lowAcc;
                    sm.bandAcc[hand][anaIdx][1] = acc.highAcc;
                } else {
                    ana_walk<false>(ast, acc, feed, n, ana);
                }
            }
            // Band -> envelope hand-off of the call's two band energies.  The band lanes' walk is a third of the envelope
            // lanes' (a 3-link chain against 4 links twice), so they only ARRIVE (bar.arrive, the PTX producer / consumer
            // idiom) and go back to the bulk queue instead of sitting out the envelope walk; the envelope lanes wait.  Two
            // barrier ids by call parity: at most two analyze() calls run between two __syncthreads, so a band warp that is
            // a call ahead arrives on the other id, and bandAcc[hand] is not rewritten before the next __syncthreads.
#if JB_CO_BANDARRIVE
            if (isBandWarp) {
                __threadfence_block();
                asm volatile("bar.arrive %0, %1;" ::"r"(CO_BAR_ANA + hand), "r"(anaThreads) : "memory");
            } else {
                named_barrier(CO_BAR_ANA + hand, anaThreads);
            }
#else
            named_barrier(CO_BAR_ANA, anaThreads);
#endif
            Metrics m {};
            if (isAna && isEnvWarp && !(dbgSkip & 32)) {
                acc.lowAcc = sm.bandAcc[hand][anaIdx][0];
                acc.highAcc = sm.bandAcc[hand][anaIdx][1];
                m = ana_finish(ast, acc, load_stats(sm.stats[par][anaClip][sig]), n, ana);
            }
            return m;
        };
        // phase 0: analyze(buffer) before the plugin's DSP (e.g. JuicyPunch/PluginProcessor.cpp:82); phase 1: after it (:114),
        // then the mailboxes (:115-123).  ONE call site for every analyze() of the kernel (the loop in the step below): the
        // walks are ~1100 instructions per inlined copy, there used to be four copies, and the exact-math instantiation's
        // hot code had outgrown the SM's instruction cache (no_instruction was 25 % of its stall samples,
        // profiles/r02_coop_icache.txt).
        auto analyze_phase = [&](int blk, int n, int ph) {
            const Metrics m = analyze(blk, n, anaSlot + ph);
            if (isAna && isEnvWarp) {
                if (ph == 0)
                    preScore = m.score;
                else
                    publish_record(a, anaSlot, clip0 + anaClip, a.histFirstBlock + blk, m, preScore, 0.0f);
            }
        };
        const int lastBlk = (a.nSamples + a.blockSize - 1) / a.blockSize - 1;
        const int lastN = a.nSamples - lastBlk * a.blockSize;

        // ---- prologue: load step 0 and scout it
        __syncthreads(); // widthPos/widthCount visible; previous group's tiles are drained
        if (isProducerWarp) {
            bulk_wait_read();
            issue_load(cur, gstep);
        }
        if (isScoutWarp)
            scout_step(cur, gstep);
        __syncthreads();

        // ---- main loop over steps, plus one trailing trip in which only the analyzers work (the last host block)
        bool tailDone = false;
        while (cur.valid || !tailDone) {
            const bool tail = !cur.valid;
            tailDone = tail;
            const Cursor nxt = tail ? cur : cursor_next(a, cur);
            const unsigned step = gstep, slot = gstep & 1;
            const int blockPar = cur.blk & 1;

            if (isProducerWarp) {
                if (lane == 0)
                    sm.claim[slot ^ 1] = 0; // the next step's work queue (nobody reads it before the barrier below)
                if (!tail && nxt.valid) {
                    bulk_wait_read(); // the store that last read the other slot has drained
                    issue_load(nxt, step + 1);
                }
            } else {
                if (isScoutWarp) {
                    if (!tail && nxt.valid)
                        scout_step(nxt, step + 1);
                } else if (isEnvWarp || isBandWarp) {
                    // analyzers run one host block behind: pre-analysis of block blk-1 during the first
                    // step of block blk, post-analysis during the second (or both, if blk is one step long);
                    // both calls of the last block in the trailing trip
                    int blk = lastBlk, n = lastN, ph0 = 0, ph1 = 2;
                    if (!tail) {
                        blk = cur.blk - 1;
                        n = a.blockSize;
                        ph0 = cur.off == 0 ? 0 : 1;
                        ph1 = (cur.nBlk <= CO_T || cur.off == CO_T) ? 2 : 1;
                        if (cur.blk == 0)
                            ph1 = ph0;
                    }
#pragma unroll 1
                    for (int ph = ph0; ph < ph1; ++ph)
                        analyze_phase(blk, n, ph);
                }
                // ---- bulk work: the step's clips are a queue in shared memory, claimed one at a time.  The bulk warps
                // live on it; scout and band warps join once their own (short) work of the step is done.  With the exact
                // routines the bulk math is the larger part (116 / 93 issue cycles per tanhf / powf against 20 / 16 for the
                // MUFU forms, profiles/microbench/exact_math.cu), so there the envelope warps and the otherwise idle warps of
                // their scheduler take clips too; with fast math the envelope walk is the critical chain and its warps keep
                // to it.  (Measured: the queue balances the step but the exact kernel stays dependency-bound -- 0.84 eligible
                // warps per scheduler at 16 warps per SM, profiles/r02_coop_exact_ncu.json; 20 / 24 / 32 warps with fewer
                // registers gave 7.0 / 6.9 / 7.6 ms against 7.3, and cost the fast instantiation 0.4 - 1.1 ms.)
#ifndef JB_CO_ENVJOIN
#define JB_CO_ENVJOIN 1 // exact math: the envelope warps take bulk clips after their walk (1), never (0), or only while more than JB_CO_ENVJOIN clips are unclaimed (>1)
#endif
                bool joins = (isBulkWarp || isScoutWarp || isBandWarp || EXACT) && !tail;
                if (JB_CO_ENVJOIN == 0 && isEnvWarp)
                    joins = false;
                if (JB_CO_ENVJOIN > 1 && isEnvWarp && joins)
                    joins = G - *reinterpret_cast<volatile int*>(&sm.claim[slot]) > JB_CO_ENVJOIN;
                if (joins) {
                mbar_wait(&sm.bar[slot], (step >> 1) & 1);
                const int nValid = max(0, min(CO_CH, cur.n - lane * CO_CH));
                const bool firstStep = cur.off == 0;
                const bool ragged = cur.n < CO_T; // warp-uniform
                while (!(dbgSkip & 4)) {
                    int ci = 0;
                    if (lane == 0)
                        ci = atomicAdd(&sm.claim[slot], 1);
                    ci = __shfl_sync(0xffffffffu, ci, 0);
                    if (ci >= G)
                        break;
                    float* rowL = &sm.tile[slot][2 * ci][lane * CO_CH];
                    float* rowR = &sm.tile[slot][2 * ci + 1][lane * CO_CH];
                    float l[CO_CH], r[CO_CH];
                    {
                        const float4 a0 = *reinterpret_cast<const float4*>(rowL), a1 = *reinterpret_cast<const float4*>(rowL + 4);
                        const float4 b0 = *reinterpret_cast<const float4*>(rowR), b1 = *reinterpret_cast<const float4*>(rowR + 4);
                        l[0] = a0.x; l[1] = a0.y; l[2] = a0.z; l[3] = a0.w; l[4] = a1.x; l[5] = a1.y; l[6] = a1.z; l[7] = a1.w;
                        r[0] = b0.x; r[1] = b0.y; r[2] = b0.z; r[3] = b0.w; r[4] = b1.x; r[5] = b1.y; r[6] = b1.z; r[7] = b1.w;
                    }
                    WidthPrefetch widthPf;
                    widthPf.have = false;
                    if (widthSlot >= 0)
                        widthPf = width_prefetch(a.slot[widthSlot].c.width, a.widthRing + (long long) (clip0 + ci) * a.ringClipStride,
                                                 sm.widthPos[ci], lane, nValid);
                    if (ragged)
                        mask_chunk(l, r, nValid);
                    float* monoClip = monoCta + ((size_t) ci * nSig) * 2 * CO_BLOCKMAX + blockPar * CO_BLOCKMAX + cur.off + lane * CO_CH;
                    signal_stats(l, r, sm.stats[blockPar][ci][0], firstStep, monoClip, lane, nValid, polKeep);
                    for (int s = 0; s < L; ++s) {
                        const SlotDesc& d = a.slot[s];
                        if (d.kind == K_PUNCH) {
                            const float2 zero = make_float2(0.0f, 0.0f); // lanes past the step's end have no checkpoint
                            const float2 ckL = nValid > 0 ? sm.ckpt[slot][2 * ci][lane] : zero;
                            const float2 ckR = nValid > 0 ? sm.ckpt[slot][2 * ci + 1][lane] : zero;
                            constexpr bool rollCh = JB_CO_ROLLCH < 0 ? EXACT : JB_CO_ROLLCH != 0;
                            if (rollCh) {
#pragma unroll 1
                            for (int ch = 0; ch < 2; ++ch) { // one copy of the code: left, swap, right, swap back
                                punch_chunk<EXACT>(l, ch == 0 ? ckL : ckR, d.c.punch);
#pragma unroll
                                for (int i = 0; i < CO_CH; ++i) {
                                    const float t = l[i];
                                    l[i] = r[i];
                                    r[i] = t;
                                }
                            }
                            } else {
                            punch_chunk<EXACT>(l, ckL, d.c.punch);
                            punch_chunk<EXACT>(r, ckR, d.c.punch);
                            }
                        } else if (d.kind == K_WIDTH) {
                            int total = 0;
                            const int kStart = firstStep ? 0 : sm.widthCount[ci];
                            const int wpos0 = sm.widthPos[ci];
                            __syncwarp();
                            float* ring = a.widthRing + (long long) (clip0 + ci) * a.ringClipStride;
                            width_chunk<EXACT>(l, r, d.c.width, sm.widthTab, kStart, total, ring, wpos0, lane, nValid, widthPf);
                            if (lane == 0) {
                                sm.widthCount[ci] = kStart + total;
                                int np = wpos0 + cur.n;
                                np -= np >= d.c.width.ringLen ? d.c.width.ringLen : 0;
                                sm.widthPos[ci] = np;
                            }
                        } else if (d.kind == K_INFER) {
                            infer_chunk(l, r, d.c.infer);
                        }
                        if (ragged)
                            mask_chunk(l, r, nValid);
                        signal_stats(l, r, sm.stats[blockPar][ci][s + 1], firstStep, monoClip + (size_t) (s + 1) * 2 * CO_BLOCKMAX,
                                     lane, nValid, polKeep);
                    }
                    *reinterpret_cast<float4*>(rowL) = make_float4(l[0], l[1], l[2], l[3]);
                    *reinterpret_cast<float4*>(rowL + 4) = make_float4(l[4], l[5], l[6], l[7]);
                    *reinterpret_cast<float4*>(rowR) = make_float4(r[0], r[1], r[2], r[3]);
                    *reinterpret_cast<float4*>(rowR + 4) = make_float4(r[4], r[5], r[6], r[7]);
                }
                fence_async_smem(); // tile writes -> visible to the bulk store issued after the barrier
                }
            }
            __syncthreads();
            if (tail)
                break;
            if (isProducerWarp) {
                for (int row = lane; row < rows; row += 32) {
                    float* dst = a.out + ((long long) (clip0 + (row >> 1)) * 2 + (row & 1)) * a.rowPitch + cur.pos;
                    bulk_store(dst, &sm.tile[slot][row][0], (uint32_t) (cur.n * 4), polStream);
                }
                bulk_commit();
            }
            ++gstep;
            cur = nxt;
        }

        // ---- epilogue: analyzers finish the last host block; everyone stores state
        if (isEnvWarp || isBandWarp) {
            if (isAna) {
                const long long clip = clip0 + anaClip;
                const int b = a.slot[anaSlot].stateBase;
                auto stv = [&](int v, float x) { a.state[(long long) (b + v) * a.clipPitch + clip] = x; };
                if (isBandWarp) {
                    stv(AV_LOW, ast.low); stv(AV_HIGH, ast.high);
                } else {
                    stv(AV_SHORT, ast.sEnv); stv(AV_LONG, ast.lEnv);
                    stv(AV_REP_EMA, ast.repEma); stv(AV_FAT_EMA, ast.fatEma);
                    stv(AV_COOLDOWN, __int_as_float(ast.cool));
                    stv(AV_PRE_SCORE, preScore);
                }
            }
        }
        if (isScout) {
            const long long clip = clip0 + (scRow >> 1);
            const int b = a.slot[0].stateBase + AV_COUNT;
            a.state[(long long) (b + PV_FAST0 + (scRow & 1)) * a.clipPitch + clip] = scF;
            a.state[(long long) (b + PV_SLOW0 + (scRow & 1)) * a.clipPitch + clip] = scS;
        }
        __syncthreads(); // analyzers are done with stats/scratch of this group; widthPos is final
        if (widthSlot >= 0 && threadIdx.x < G) {
            const int b = a.slot[widthSlot].stateBase + AV_COUNT;
            a.state[(long long) (b + WV_WPOS) * a.clipPitch + clip0 + threadIdx.x] = __int_as_float(sm.widthPos[threadIdx.x]);
        }
    }
    if (isProducerWarp)
        bulk_wait_all();
}

thread_local char g_coopErr[256];

} // namespace

extern "C" {

const char* jbk_coop_last_error(void) { return g_coopErr; }

// Scratch the coop kernel needs (mono rings of the analyzer lanes), sized for a full grid.
size_t jbk_coop_scratch_bytes(int chainLen, int numSMs)
{
    return sizeof(float) * ((size_t) numSMs * CO_GMAX * (size_t) (chainLen + 1) * 2 * CO_BLOCKMAX + CO_FEED_PAD);
}

// Can this call take the cooperative path?  (Chain made of Punch / Width / Infer with Punch only
// in front, host block <= 512, everything 16-byte aligned, the Width ring long enough for a step.)
int jbk_coop_supported(const ProcArgs* a)
{
    if (a->chainLen < 1 || a->chainLen > CO_MAXCHAIN || a->nCh != 2 || a->clipMap != nullptr)
        return 0;
    if (a->blockSize > CO_BLOCKMAX || a->blockSize % 4 != 0 || a->nSamples % 4 != 0 || a->rowPitch % 4 != 0)
        return 0;
    if (((uintptr_t) a->in | (uintptr_t) a->out) & 15u)
        return 0;
    for (int s = 0; s < a->chainLen; ++s) {
        const int k = a->slot[s].kind;
        if (k == K_PUNCH) {
            if (s != 0)
                return 0;
        } else if (k == K_WIDTH) {
            const WidthCoef& w = a->slot[s].c.width;
            if (a->ringTimeStride != 1 || w.ringLen - w.delaySamples < CO_T || w.ringLen < CO_T)
                return 0;
        } else if (k != K_INFER) {
            return 0;
        }
    }
    return 1;
}

int jbk_launch_coop(const ProcArgs* args, float* monoScratch, int numSMs, void* stream)
{
    if (args->nClips <= 0 || args->nSamples <= 0)
        return 0;
    // Envelope warps alone on one scheduler pays while their sequential walk is the kernel's critical path (fast math);
    // with the exact routines the bulk warps are, and they get all four (profiles/r02_coop_variants.txt).  JB_CO_ISOLATE=0/1 forces.
    static const int isoEnv = [] { const char* v = getenv("JB_CO_ISOLATE"); return v == nullptr ? -1 : atoi(v); }();
    const bool exact = args->exactMath != 0;
    const bool isolate = isoEnv < 0 ? !exact : isoEnv != 0;
    void (*kernel)(const CoopArgs) = exact ? (isolate ? jb_coop_kernel<true, true> : jb_coop_kernel<true, false>)
                                           : (isolate ? jb_coop_kernel<false, true> : jb_coop_kernel<false, false>);
    {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(CoopSmem));
        if (e != cudaSuccess) {
            snprintf(g_coopErr, sizeof g_coopErr, "cudaFuncSetAttribute(jb_coop_kernel): %s", cudaGetErrorString(e));
            return -1;
        }
    }
    CoopArgs ca;
    ca.p = *args;
    ca.monoScratch = monoScratch;
    int g = (args->nClips + numSMs - 1) / numSMs;
    if (g > CO_GMAX)
        g = CO_GMAX;
    ca.groupClips = g;
    ca.numGroups = (args->nClips + g - 1) / g;
#ifdef JB_COOP_DEBUG
    const char* dbg = getenv("JB_COOP_DEBUG_SKIP");
    ca.debugSkip = dbg ? atoi(dbg) : 0;
#else
    ca.debugSkip = 0;
#endif
    const int grid = ca.numGroups < numSMs ? ca.numGroups : numSMs;
    // (A persisting-L2 access window over the mono scratch ring was tried: it cut the ring's write-back -- dram__bytes_write
    // 2.9 -> 2.4 GB against 1.6 GB of audio -- without changing the kernel's time, but the L2 set-aside it needs is a
    // device-wide limit that stays in force for every later kernel of the process: JuicyInfer on 65536 clips went 16.5 ->
    // 44 ms behind it.  Not worth it; the per-access evict_last hints stay.)
    kernel<<<grid, CO_THREADS, sizeof(CoopSmem), (cudaStream_t) stream>>>(ca);
    jbk_note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_coopErr, sizeof g_coopErr, "jb_coop_kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

} // extern "C"
