// jb_lane.cuh -- the lane-per-clip render code shared by the generic multi-plugin kernel (jb_kernels.cu) and the
// single-plugin kernels (jb_single_*.cu): per-plugin DSP ("Main"), block pre-passes ("Pre"), sample streaming and the
// sweep over one block of one clip.  Everything lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once
#include "jb_device.cuh"
#include "jb_libm.h"

#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

namespace {

using namespace jbdev;

#define JB_CTA_THREADS 32
#define JB_LANE_CTA_THREADS JB_CTA_THREADS

// Per-sample transcendentals.  The oracle uses glibc's (nearly correctly rounded)
// float functions; block-rate pow/log10 are evaluated in fp64 and rounded once,
// Texture-metal's cos restates glibc's own algorithm (below), the others use
// CUDA's <= 2 ulp float versions.
// std::cos(float) as glibc >= 2.28 computes it (sysdeps/ieee754/flt-32/s_cosf.c + sincosf.h,
// from ARM's optimized-routines): reduce by pi/2 in double, a degree-8 (cos) or degree-7 (sin)
// polynomial in double, one rounding to float.  Restating the published algorithm with its
// published coefficients makes the device bit-identical to the oracle's libm for Texture-metal's
// per-sample pole angle, the one transcendental whose last bit the recurrences amplify
// (SURVEY.md Appendix D.3).  glibc selects its FMA build on every x86-64 CPU with FMA, hence
// the explicit fma() here (a double-rounding difference would be ~1e-16 relative anyway).
__device__ __forceinline__ float cosf_glibc(float y)
{
    const double hpiInv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const double C0 = 1.0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    const unsigned top = (__float_as_uint(y) >> 20) & 0x7ffu; // abstop12
    double x = (double) y;
    int n = 0;
    double sg = 1.0; // the second table row is the first with the cosine coefficients negated
    if (top < 0x3f4u) {                 // |y| < 0.75 (abstop12(pi/4))
        if (top < 0x398u)               // |y| < 2^-12
            return 1.0f;
        n = 1;
    } else if (top < 0x42fu) {          // |y| < 120
        const double r = x * hpiInv;
        const int q = ((int) r + 0x800000) >> 24;
        x = fma(-(double) q, hpi, x);
        if (q & 2)
            sg = -1.0;
        if (((q + 1) & 2) != 0)         // sign[q & 3] = {1, -1, -1, 1}
            x = -x;
        n = q ^ 1;
    } else {
        return (float) cos((double) y); // outside the range any caller here produces
    }
    const double x2 = x * x;
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double s1 = fma(x2, S3, S2);
        const double x7 = x3 * x2;
        const double s = fma(x3, S1, x);
        return (float) fma(x7, s1, s);
    }
    const double x4 = x2 * x2;
    const double c2 = sg * fma(x2, C4, C3);
    const double c1 = sg * fma(x2, C1, C0);
    const double x6 = x4 * x2;
    const double c = fma(x4, sg * C2, c1);
    return (float) fma(x6, c2, c);
}
// The same for 2^-12 <= |y| < 120 only, without a branch: below 0.75 glibc merely skips the reduction (q = 0 gives the
// same x, n and sign), and both polynomials are evaluated and selected, so the loop that calls it stays one basic block.
// Texture-metal's pole angle is 2 pi f / sr with f in [20 Hz, 0.45 sr], i.e. within [2.6e-3, 2.83] at any rate <= 500 kHz.
__device__ __forceinline__ float cosf_glibc_mid(float y)
{
    const double hpiInv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const double C0 = 1.0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    double x = (double) y;
    const double r = x * hpiInv;
    const int q = ((int) r + 0x800000) >> 24;
    x = fma(-(double) q, hpi, x);
    const double sg = (q & 2) ? -1.0 : 1.0;
    x = ((q + 1) & 2) ? -x : x;
    const double x2 = x * x;
    // n = q ^ 1: sine polynomial when q is odd
    const double x3 = x * x2;
    const double s1 = fma(x2, S3, S2);
    const double x7 = x3 * x2;
    const double sn = fma(x7, s1, fma(x3, S1, x));
    const double x4 = x2 * x2;
    const double c2 = sg * fma(x2, C4, C3);
    const double c1 = sg * fma(x2, C1, C0);
    const double x6 = x4 * x2;
    const double cs = fma(x6, c2, fma(x4, sg * C2, c1));
    return (float) ((q & 1) ? sn : cs);
}
__device__ __forceinline__ float pow_exact(float x, float y) { return (float) pow((double) x, (double) y); }
__device__ __forceinline__ float log10_exact(float x) { return (float) log10((double) x); }

struct Lane {
    const ProcArgs& a;
    long long clip;
    __device__ __forceinline__ float* sp(int var) const { return a.state + (long long) var * a.clipPitch + clip; }
    __device__ __forceinline__ float ld(int var) const { return *sp(var); }
    __device__ __forceinline__ void st(int var, float v) const { *sp(var) = v; }
    __device__ __forceinline__ int ldi(int var) const { return __float_as_int(*sp(var)); }
    __device__ __forceinline__ void sti(int var, int v) const { *sp(var) = __int_as_float(v); }
};

// ------------------------------------------------------------------ analyzer
// JuicinessAnalyzer::analyze, src/shared/JuicinessAnalyzer.cpp:31-155.

// Sums that depend only on the block's samples (not on analyzer state), shared by
// the post-analysis of plugin s and the pre-analysis of plugin s+1 (:76-77, :86-91,
// and getRMSLevel :105-106, which accumulates in double).
struct BlockStats {
    float rms = 0.0f, peak = 0.0f, side = 0.0f, corr = 0.0f;
    double l2 = 0.0, r2 = 0.0;
    // Plain sums that feed the metrics record only (tolerance 0.01 absolute): fused multiply-adds, one instruction per
    // sum.  getRMSLevel's two sums of squares stay in fp64 like the shim's: fp32 would do for the tolerance, but measured
    // on 65536 clips (profiles/r01_s6_light_variants.txt) the fp32 form made the Infer kernel 1.6x SLOWER (21.3 -> 35.3 ms,
    // stalls move to the cp.async ring: long_scoreboard 32 %, mio 26 %) and nothing else faster, so the DFMAs stay.
    __device__ __forceinline__ void step(float l, float r, float mono)
    {
        rms = fmaf(mono, mono, rms);        // rmsAccum; midAccum is the same expression (:62, :86, :88)
        peak = fmaxf(peak, fabsf(mono));
        const float s = 0.5f * (l - r);
        side = fmaf(s, s, side);
        corr = fmaf(l, r, corr);
        const double dl = (double) l, dr = (double) r;
        l2 = fma(dl, dl, l2);               // dl*dl is exact in fp64, so the fused form rounds identically
        r2 = fma(dr, dr, r2);
    }
    __device__ __forceinline__ StatSums sums() const { return StatSums { rms, peak, side, corr, l2, r2 }; }
};

// One analyzer's walk over one block (jb_device.cuh: ana_step / ana_finish) with its state in the SoA arrays.
struct AnaWalk {
    AnaState st;
    AnaAcc acc;
    __device__ __forceinline__ void load(const Lane& L, int base)
    {
        st.sEnv = L.ld(base + AV_SHORT);
        st.lEnv = L.ld(base + AV_LONG);
        st.low = L.ld(base + AV_LOW);
        st.high = L.ld(base + AV_HIGH);
        st.cool = L.ldi(base + AV_COOLDOWN);
    }
    // The walk goes in groups of up to four samples (a quad of the sweep): st.cool holds `rem` while a block is walked
    // (AnaGroup, jb_device.cuh) and is turned back into the reference's counter by finish().
    AnaGroup grp;
    __device__ __forceinline__ void begin_block() { st.cool = st.cool - 1; }
    __device__ __forceinline__ void group_begin() { grp.begin(st); }
    __device__ __forceinline__ void step(int k, float mono, const AnaCoef& c)
    {
        grp.step(st, acc, k, mono, c);
        ana_step_bands(st, acc, mono, c);
    }
    __device__ __forceinline__ void group_end(int count, const AnaCoef& c) { grp.end(st.cool, acc, count, c); }
    __device__ Metrics finish(const Lane& L, int base, const BlockStats& s, int n, const AnaCoef& c)
    {
        st.cool = max(st.cool + 1, 0);
        st.repEma = L.ld(base + AV_REP_EMA);
        st.fatEma = L.ld(base + AV_FAT_EMA);
        const Metrics m = ana_finish(st, acc, s.sums(), n, c);
        L.st(base + AV_SHORT, st.sEnv);
        L.st(base + AV_LONG, st.lEnv);
        L.st(base + AV_LOW, st.low);
        L.st(base + AV_HIGH, st.high);
        L.sti(base + AV_COOLDOWN, st.cool);
        L.st(base + AV_REP_EMA, st.repEma);
        L.st(base + AV_FAT_EMA, st.fatEma);
        return m;
    }
};

// ------------------------------------------------------------------ plugin DSP ("main" part of a sweep)
// Interface: writes() (must the sweep store samples), kSeqChannels (channel 0's whole block
// must precede channel 1's), load/store of per-clip state, step(l, r) in place.

struct MainBase {
    static constexpr bool kMonoBypass = false; // the plugin returns before its DSP on a mono bus
    __device__ __forceinline__ void quad_begin() {} // around every group of four whole samples (vector path only)
    __device__ __forceinline__ void quad_end() {}
    __device__ __forceinline__ float stepCh0(float l) { return l; }
    __device__ __forceinline__ bool writes(bool outOfPlace) const { return true; }
};

struct MainNone : MainBase { // sweep 0: samples pass through untouched, nothing to publish
    static constexpr bool kHas = false, kSeqChannels = false;
    static constexpr bool kHeavy = false; // heavy per-sample state: 4 samples per trip (registers); light: 8 (one sector per store)
    __device__ __forceinline__ void load(const Lane&, const SlotDesc&, int) {}
    __device__ __forceinline__ void step(float&, float&) {}
    __device__ __forceinline__ void store(const Lane&, const SlotDesc&) {}
};

// JuicyInfer/PluginProcessor.cpp:78-81: buffer.applyGain(trimGain) between the two analyses
struct MainInfer : MainBase {
    static constexpr bool kHas = true, kSeqChannels = false;
    static constexpr bool kHeavy = false; // heavy per-sample state: 4 samples per trip (registers); light: 8 (one sector per store)
    float g;
    int mode;
    __device__ __forceinline__ void load(const Lane&, const SlotDesc& d, int)
    {
        g = d.c.infer.trimGain;
        mode = d.c.infer.gainMode;
    }
    __device__ __forceinline__ void step(float& l, float& r)
    {
        // selects, not branches: the quad stays one basic block (mode 0 keeps the sample's bits, like applyGain(1))
        const float gl = l * g, gr = r * g;
        l = mode == 1 ? gl : (mode == 2 ? 0.0f : l);
        r = mode == 1 ? gr : (mode == 2 ? 0.0f : r);
    }
    __device__ __forceinline__ void store(const Lane&, const SlotDesc&) {}
    __device__ __forceinline__ bool writes(bool outOfPlace) const { return mode != 0 || outOfPlace; }
};

// JuicySaturator/PluginProcessor.cpp:83-98
template <bool EXACT>
struct MainSat : MainBase {
    static constexpr bool kHas = true, kSeqChannels = false;
    static constexpr bool kHeavy = false; // heavy per-sample state: 4 samples per trip (registers); light: 8 (one sector per store)
    float s0, s1;
    SatCoef c;
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d, int)
    {
        c = d.c.sat;
        const int b = d.stateBase + AV_COUNT;
        s0 = L.ld(b + SV_TONE0);
        s1 = L.ld(b + SV_TONE1);
    }
    __device__ __forceinline__ float one(float dry, float& state) const
    {
        const float driven = dry * c.inGain;
        const float skewed = driven + c.asym * driven * driven;
        // MUFU-based (<= 3e-6 relative, jb_device.cuh), or the C library's own result when a resonator follows (jb_libm.h)
        const float soft = EXACT ? jblibm::tanhf_fdlibm(skewed) : tanh_fast(skewed);
        state += c.toneCoeff * (soft - state);
        const float wet = state * c.outGain;
        return dry + c.mix * (wet - dry);
    }
    // both channels in packed halves (jb_device.cuh: F2): the same operations in the same order as one()
    __device__ __forceinline__ void step(float& l, float& r)
    {
        const F2 dry = f2(l, r);
        const F2 driven = mul2(dry, f2(c.inGain));
        const F2 skewed = mul_add2(mul2(f2(c.asym), driven), driven, driven);
        const F2 soft = EXACT ? f2(jblibm::tanhf_fdlibm(skewed.x), jblibm::tanhf_fdlibm(skewed.y)) : tanh_fast2(skewed);
        F2 st = f2(s0, s1);
        st = mul_add2(f2(c.toneCoeff), sub2(soft, st), st);
        s0 = st.x;
        s1 = st.y;
        const F2 wet = mul2(st, f2(c.outGain));
        const F2 out = mul_add2(f2(c.mix), f2(wet.x - dry.x, wet.y - dry.y), dry);
        l = out.x;
        r = out.y;
    }
    __device__ __forceinline__ void store(const Lane& L, const SlotDesc& d)
    {
        const int b = d.stateBase + AV_COUNT;
        L.st(b + SV_TONE0, s0);
        L.st(b + SV_TONE1, s1);
    }
};

// JuicyPunch/PluginProcessor.cpp:86-112
template <bool EXACT>
struct MainPunch : MainBase {
    static constexpr bool kHas = true, kSeqChannels = false;
    static constexpr bool kHeavy = EXACT; // heavy per-sample state: 4 samples per trip (registers); light: 8 (one sector per store)
    float f0, f1, sl0, sl1, invTanhDrive;
    PunchCoef c;
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d, int)
    {
        c = d.c.punch;
        invTanhDrive = 1.0f / c.tanhDrive;
        const int b = d.stateBase + AV_COUNT;
        f0 = L.ld(b + PV_FAST0);
        f1 = L.ld(b + PV_FAST1);
        sl0 = L.ld(b + PV_SLOW0);
        sl1 = L.ld(b + PV_SLOW1);
    }
    __device__ __forceinline__ float one(float dry, float& fEnv, float& sEnv) const
    {
        const float adry = fabsf(dry);
        fEnv = c.omFast * adry + c.fastCoeff * fEnv;
        sEnv = c.omSlow * adry + c.slowCoeff * sEnv;
        const float transient = jmaxf(0.0f, fEnv - sEnv);
        // MUFU-based pow / tanh (jb_device.cuh), like the cooperative kernel -- or, when a resonator follows in the chain,
        // the C library's own results and the reference's division (jb_libm.h)
        const float transientCurve = EXACT ? jblibm::powf_glibc_pos(transient, c.curveExp) : pow_unit(transient, c.curveExp);
        const float punchGain = 1.0f + c.punchK * transientCurve;
        const float sustainGain = 1.0f + c.sustainK * jmaxf(0.0f, sEnv - transient * 0.6f);
        float wet = dry * punchGain * sustainGain;
        const float soft = EXACT ? jblibm::fdiv(jblibm::tanhf_fdlibm(wet * c.drive), c.tanhDrive) : tanh_fast(wet * c.drive) * invTanhDrive;
        const float hard = jlimitf(-0.95f, 0.95f, wet * c.hardK);
        wet = soft + c.clipAmt * (hard - soft);
        return (dry + c.mix * (wet - dry)) * c.outGain;
    }
    // both channels in packed halves: the operations of one() in the same order (see MainCohere::step)
    __device__ __forceinline__ void step(float& l, float& r)
    {
        const F2 dry = f2(l, r);
        const F2 adry = f2(fabsf(l), fabsf(r));
        const F2 fin = mul2(f2(c.omFast), adry), fkeep = mul2(f2(c.fastCoeff), f2(f0, f1));
        const F2 sin_ = mul2(f2(c.omSlow), adry), skeep = mul2(f2(c.slowCoeff), f2(sl0, sl1));
        f0 = fin.x + fkeep.x;
        f1 = fin.y + fkeep.y;
        sl0 = sin_.x + skeep.x;
        sl1 = sin_.y + skeep.y;
        const F2 S = f2(sl0, sl1);
        const F2 d = sub2(f2(f0, f1), S);
        const F2 tr = f2(jmaxf(0.0f, d.x), jmaxf(0.0f, d.y));
        const F2 curve = EXACT ? f2(jblibm::powf_glibc_pos(tr.x, c.curveExp), jblibm::powf_glibc_pos(tr.y, c.curveExp))
                               : f2(pow_unit(tr.x, c.curveExp), pow_unit(tr.y, c.curveExp));
        const F2 punchGain = mul_add2(f2(c.punchK), curve, f2(1.0f));
        const F2 t06 = mul2(tr, f2(0.6f));
        const F2 sus = f2(jmaxf(0.0f, S.x - t06.x), jmaxf(0.0f, S.y - t06.y));
        const F2 sustainGain = mul_add2(f2(c.sustainK), sus, f2(1.0f));
        const F2 wet0 = mul2(mul2(dry, punchGain), sustainGain);
        const F2 driven = mul2(wet0, f2(c.drive));
        F2 soft;
        if (EXACT)
            soft = f2(jblibm::fdiv(jblibm::tanhf_fdlibm(driven.x), c.tanhDrive), jblibm::fdiv(jblibm::tanhf_fdlibm(driven.y), c.tanhDrive));
        else
            soft = mul2(tanh_fast2(driven), f2(invTanhDrive));
        const F2 hk = mul2(wet0, f2(c.hardK));
        const F2 hard = f2(jlimitf(-0.95f, 0.95f, hk.x), jlimitf(-0.95f, 0.95f, hk.y));
        const F2 wet = mul_add2(f2(c.clipAmt), f2(hard.x - soft.x, hard.y - soft.y), soft);
        const F2 out = mul2(mul_add2(f2(c.mix), sub2(wet, dry), dry), f2(c.outGain));
        l = out.x;
        r = out.y;
    }
    __device__ __forceinline__ void store(const Lane& L, const SlotDesc& d)
    {
        const int b = d.stateBase + AV_COUNT;
        L.st(b + PV_FAST0, f0);
        L.st(b + PV_FAST1, f1);
        L.st(b + PV_SLOW0, sl0);
        L.st(b + PV_SLOW1, sl1);
    }
};

// JuicyWidth/PluginProcessor.cpp:91-137.  The delay line keeps only the right
// channel's wet signal (the left ring is written but never read, :122,:131).
// Ring layouts: clip-major [clip][ringLen] (pitch 1; what the engine uses) or time-major
// [ringLen][clip].  With the clip-major ring and everything a multiple of four samples (ring
// length, delay, write position: the usual 2880 / 576 case) a quad of wet samples is stored with
// one 16-byte store and the delayed quad is fetched with one 16-byte load issued a quad EARLY --
// the delayed samples were written delaySamples ago -- so the ring's L2 / HBM latency no longer
// sits in front of every sample (it was 55 % of all stall samples, profiles/r01_width_lane_*).
struct MainWidth : MainBase {
    static constexpr bool kMonoBypass = true; // JuicyWidth/PluginProcessor.cpp:76-89
    static constexpr bool kHas = true, kSeqChannels = false;
    static constexpr bool kHeavy = false; // heavy per-sample state: 4 samples per trip (registers); light: 8 (one sector per store)
    float width;
    int wpos;
    WidthCoef c;
    float* ring;
    long long pitch;
    bool quads;          // vector ring path
    static constexpr int kDepth = 1; // quads fetched ahead of use (needs delay >= 4 * kDepth + 4)
    float4 ahead[kDepth]; // delayed quads of the next kDepth quad_begin calls
    float d0, d1, d2, d3; // delayed samples of the current quad, next first
    float w0, w1, w2, w3; // wet samples of the current quad, oldest first (static shifts keep both in registers)
    __device__ __forceinline__ int delayed_pos(int p) const
    {
        int rp = p - c.delaySamples;
        if (rp < 0)
            rp += c.ringLen;
        return rp;
    }
    __device__ __forceinline__ int wrap(int p) const { return p >= c.ringLen ? p - c.ringLen : p; }
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d, int)
    {
        c = d.c.width;
        width = c.width; // re-read from the parameter every block (:93)
        wpos = L.ldi(d.stateBase + AV_COUNT + WV_WPOS);
        ring = L.a.widthRing + L.clip * L.a.ringClipStride;
        pitch = L.a.ringTimeStride;
        quads = L.a.vecOk != 0 && pitch == 1 && ((c.ringLen | c.delaySamples | wpos | (int) L.a.ringClipStride) & 3) == 0
                && c.delaySamples >= 4 * kDepth + 4 && c.ringLen >= 8 * kDepth + 8;
        d0 = d1 = d2 = d3 = w0 = w1 = w2 = w3 = 0.0f;
        if (quads) {
#pragma unroll
            for (int j = 0; j < kDepth; ++j)
                ahead[j] = *reinterpret_cast<const float4*>(ring + delayed_pos(wrap(wpos + 4 * j)));
        }
    }
    __device__ __forceinline__ void quad_begin()
    {
        if (!quads)
            return;
        d0 = ahead[0].x; d1 = ahead[0].y; d2 = ahead[0].z; d3 = ahead[0].w;
#pragma unroll
        for (int j = 0; j + 1 < kDepth; ++j)
            ahead[j] = ahead[j + 1];
        // the quad used kDepth calls from now was written >= 4 samples ago (delay >= 4 kDepth + 4)
        ahead[kDepth - 1] = *reinterpret_cast<const float4*>(ring + delayed_pos(wrap(wpos + 4 * kDepth)));
    }
    __device__ __forceinline__ void quad_end()
    {
        if (!quads)
            return;
        *reinterpret_cast<float4*>(ring + wpos) = make_float4(w0, w1, w2, w3);
        wpos += 4;
        if (wpos >= c.ringLen)
            wpos = 0;
    }
    __device__ __forceinline__ void step(float& l, float& r)
    {
        const float dryL = l, dryR = r;
        const float corrProxy = jlimitf(-1.0f, 1.0f, dryL * dryR * 12.0f);
        if (corrProxy < -0.1f)
            width *= c.dynamicLimit;
        const float mid = 0.5f * (dryL + dryR);
        const float side = 0.5f * (dryL - dryR) * (1.0f + width);
        const float wetL = mid + side;
        float wetR = mid - side;
        if (quads) {
            w0 = w1; w1 = w2; w2 = w3; w3 = wetR;
            wetR = d0;
            d0 = d1; d1 = d2; d2 = d3;
        } else {
            ring[(long long) wpos * pitch] = wetR;
            wetR = ring[(long long) delayed_pos(wpos) * pitch];
            if (++wpos >= c.ringLen)
                wpos = 0;
        }
        l = (dryL + c.mix * (wetL - dryL)) * c.outGain;
        r = (dryR + c.mix * (wetR - dryR)) * c.outGain;
    }
    __device__ __forceinline__ void store(const Lane& L, const SlotDesc& d) { L.sti(d.stateBase + AV_COUNT + WV_WPOS, wpos); }
};

// JuicyCohere/PluginProcessor.cpp:99-119 (lpA/lpB restart at 0 every block, :103-104)
struct MainCohere : MainBase {
    static constexpr bool kHas = true, kSeqChannels = false;
    static constexpr bool kHeavy = false; // heavy per-sample state: 4 samples per trip (registers); light: 8 (one sector per store)
    float t0, t1, a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
    float lowComp, midComp, highComp;
    CohereCoef c;
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d, int)
    {
        c = d.c.cohere;
        const int b = d.stateBase + AV_COUNT;
        t0 = L.ld(b + CV_TAIL0);
        t1 = L.ld(b + CV_TAIL1);
        lowComp = L.ld(b + CV_COMP_LOW);
        midComp = L.ld(b + CV_COMP_MID);
        highComp = L.ld(b + CV_COMP_HIGH);
    }
    __device__ __forceinline__ float one(float dry, float& lpA, float& lpB, float& tail) const
    {
        lpA += c.lowCoeff * (dry - lpA);
        lpB += c.highCoeff * (dry - lpB);
        const float low = lpA * lowComp;
        const float high = (dry - lpB) * highComp;
        const float mid = (dry - lpA - (dry - lpB)) * midComp;
        const float matched = low + mid + high;
        tail = matched + tail * c.fb;
        const float wet = matched + c.tailK * tail;
        return (dry + c.mix * (wet - dry)) * c.outGain;
    }
    // both channels in packed halves: the operations of one() in the same order (adds that consume a packed product stay
    // scalar, see the caution at F2 in jb_device.cuh)
    __device__ __forceinline__ void step(float& l, float& r)
    {
        const F2 dry = f2(l, r);
        F2 A = f2(a0, a1), B = f2(b0, b1), T = f2(t0, t1);
        A = mul_add2(f2(c.lowCoeff), sub2(dry, A), A);
        B = mul_add2(f2(c.highCoeff), sub2(dry, B), B);
        const F2 low = mul2(A, f2(lowComp));
        const F2 hb = sub2(dry, B);
        const F2 high = mul2(hb, f2(highComp));
        const F2 mid = mul2(sub2(sub2(dry, A), hb), f2(midComp));
        const F2 matched = f2((low.x + mid.x) + high.x, (low.y + mid.y) + high.y);
        T = mul_add2(T, f2(c.fb), matched);
        const F2 wet = mul_add2(f2(c.tailK), T, matched);
        const F2 out = mul2(mul_add2(f2(c.mix), sub2(wet, dry), dry), f2(c.outGain));
        a0 = A.x; a1 = A.y; b0 = B.x; b1 = B.y; t0 = T.x; t1 = T.y;
        l = out.x;
        r = out.y;
    }
    __device__ __forceinline__ void store(const Lane& L, const SlotDesc& d)
    {
        const int b = d.stateBase + AV_COUNT;
        L.st(b + CV_TAIL0, t0);
        L.st(b + CV_TAIL1, t1);
    }
};

// JuicyTexture ChannelState (JuicyTexture/PluginProcessor.h:55-77) held in registers
struct TexChan {
    float tail, lp, hp, env, wetEnv, noiseHp, dcIn, dcOut, protectGain, springPos, springVel;
    float fleshPosA, fleshVelA, fleshPosB, fleshVelB, prevWave;
    float y1[4], y2[4];
    __device__ __forceinline__ void load(const Lane& L, int b)
    {
        tail = L.ld(b + TV_TAIL); lp = L.ld(b + TV_LP); hp = L.ld(b + TV_HP); env = L.ld(b + TV_ENV);
        wetEnv = L.ld(b + TV_WETENV); noiseHp = L.ld(b + TV_NOISEHP); dcIn = L.ld(b + TV_DCIN); dcOut = L.ld(b + TV_DCOUT);
        protectGain = L.ld(b + TV_PROTECT); springPos = L.ld(b + TV_SPRING_POS); springVel = L.ld(b + TV_SPRING_VEL);
        fleshPosA = L.ld(b + TV_FLESH_PA); fleshVelA = L.ld(b + TV_FLESH_VA); fleshPosB = L.ld(b + TV_FLESH_PB);
        fleshVelB = L.ld(b + TV_FLESH_VB); prevWave = L.ld(b + TV_PREVWAVE);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            y1[k] = L.ld(b + TV_Y1_0 + k);
            y2[k] = L.ld(b + TV_Y2_0 + k);
        }
    }
    __device__ __forceinline__ void store(const Lane& L, int b) const
    {
        L.st(b + TV_TAIL, tail); L.st(b + TV_LP, lp); L.st(b + TV_HP, hp); L.st(b + TV_ENV, env);
        L.st(b + TV_WETENV, wetEnv); L.st(b + TV_NOISEHP, noiseHp); L.st(b + TV_DCIN, dcIn); L.st(b + TV_DCOUT, dcOut);
        L.st(b + TV_PROTECT, protectGain); L.st(b + TV_SPRING_POS, springPos); L.st(b + TV_SPRING_VEL, springVel);
        L.st(b + TV_FLESH_PA, fleshPosA); L.st(b + TV_FLESH_VA, fleshVelA); L.st(b + TV_FLESH_PB, fleshPosB);
        L.st(b + TV_FLESH_VB, fleshVelB); L.st(b + TV_PREVWAVE, prevWave);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            L.st(b + TV_Y1_0 + k, y1[k]);
            L.st(b + TV_Y2_0 + k, y2[k]);
        }
    }
};

// JuicyTexture/PluginProcessor.cpp:107-278, one material per instantiation.
// The LCG `rng` is one instance member advanced by channel 0's whole block and
// then by channel 1's (:107,:114,:239); channel 1 therefore starts n draws ahead
// (affine skip-ahead), which lets both channels run sample-interleaved.
template <int MAT>
struct MainTexture : MainBase {
    static constexpr bool kHas = true, kSeqChannels = false;
    static constexpr bool kHeavy = true; // heavy per-sample state: 4 samples per trip (registers); light: 8 (one sector per store)
    TexChan ch0, ch1;
    float a1Rest[4]; // metal: 2 r cos(theta) of the unbent modes (impact == 0)
    uint32_t rng0, rng1;
    int waveIdx;
    const TexCoef* c;
    float* wave;
    long long pitch;
    // MAT >= 0: the material is a compile-time constant (one kernel per material: every `mat() == k` below folds away);
    // MAT < 0 would take it from the coefficients at run time (tried: all parameter sets of a Texture engine in ONE launch
    // with a warp-uniform material; its coefficients then come through LDC with a register index instead of constant-bank
    // operands, and the launch took 35 ms against 24.7 for five small-code kernels side by side -- profiles/r02_tma.txt).
    __device__ __forceinline__ int mat() const { return MAT >= 0 ? MAT : c->material; }

    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d, int n)
    {
        c = &d.c.tex;
        const int b = d.stateBase + AV_COUNT;
        ch0.load(L, b);
        ch1.load(L, b + TV_CH_STRIDE);
        waveIdx = L.ldi(b + TV_WAVEIDX);
        rng0 = (uint32_t) L.ldi(b + TV_RNG);
        uint32_t A = 1664525u, C = 1013904223u, accA = 1u, accC = 0u; // x -> A^n x + C_n
        for (int k = n; k > 0; k >>= 1) {
            if (k & 1) {
                accA *= A;
                accC = accC * A + C;
            }
            C = (A + 1u) * C;
            A *= A;
        }
        rng1 = accA * rng0 + accC;
        wave = L.a.texWave + L.clip;
        pitch = L.a.clipPitch;
        if (mat() == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                a1Rest[k] = metalA1(k, 1.0f); // bend = 1 + 0.09 * 0
        }
    }

    // modeStep with block-constant pole radius (:77-89); a1 is passed in
    __device__ __forceinline__ float mode(TexChan& st, int k, float exc, float a1) const
    {
        const float y = exc * c->modeGain[k] + a1 * st.y1[k] + c->modeA2[k] * st.y2[k];
        st.y2[k] = st.y1[k];
        st.y1[k] = y;
        return y;
    }
    __device__ __forceinline__ float metalA1(int k, float bend) const
    {
        const float f = jlimitf(20.0f, c->fMax, c->modeF[k] * bend);
        const float theta = div_rn_mid(2.0f * PI_F * f, c->srf);
        return c->modeTwoR[k] * cosf_glibc_mid(theta);
    }
    // waveguideRead (:91-105) on the time-major ring of this channel: tap positions and interpolation weight for the
    // write index w (a pure function of w and the block-constant delay)
    __device__ __forceinline__ void wavePos(int w, int& i0, int& i1, float& frac) const
    {
        const int size = c->waveSize;
        float pos = (float) w - c->delaySamp;
        // the reference's two while loops; 8 <= delaySamp <= size - 2 (jb_params.cpp), so each runs at most once
        pos = pos < 0.0f ? pos + (float) size : pos;
        pos = pos >= (float) size ? pos - (float) size : pos;
        i0 = (int) pos;
        i1 = i0 + 1 == size ? 0 : i0 + 1;
        frac = pos - (float) i0;
    }
    __device__ __forceinline__ float waveRead(const float* line) const
    {
        int i0, i1;
        float frac;
        wavePos(waveIdx, i0, i1, frac);
        return jmap3(frac, line[(long long) i0 * pitch], line[(long long) i1 * pitch]);
    }
    // Wood / plastic, vector path: the delayed samples of a whole quad are fetched together at its top.  They were written
    // at least delaySamp - 1 >= 7 samples before the quad's first write position, so the quad's own four writes cannot
    // touch them, and the 16 loads are in flight at once instead of one L2 round trip per sample and channel
    // (long_scoreboard was 38 % of wood's stall samples, profiles/r01_s6_tex_lane_before.txt).
    float dq0[4], dq1[4]; // delayed sample of channel 0 / 1 for the quad's samples, next first
    bool havePref = false;
    __device__ __forceinline__ void quad_begin()
    {
        if (mat() != 2 && mat() != 3)
            return;
        const float* line0 = wave;
        const float* line1 = wave + (long long) c->waveSize * pitch;
        float a0[4], a1[4], b0[4], b1[4], fr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int w = waveIdx + k;
            w = w >= c->waveSize ? w - c->waveSize : w;
            int i0, i1;
            wavePos(w, i0, i1, fr[k]);
            a0[k] = line0[(long long) i0 * pitch];
            a1[k] = line0[(long long) i1 * pitch];
            b0[k] = line1[(long long) i0 * pitch];
            b1[k] = line1[(long long) i1 * pitch];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dq0[k] = jmap3(fr[k], a0[k], a1[k]);
            dq1[k] = jmap3(fr[k], b0[k], b1[k]);
        }
        havePref = true;
    }
    __device__ __forceinline__ void quad_end() { havePref = false; }

    // Per-sample front part (:116-134): transient envelope -> impact / body / trail, three-band split -> core
    struct Front { float impact, body, trail, core; };
    __device__ __forceinline__ Front front(float dry, TexChan& st) const
    {
        const TexCoef& k = *c;
        const float driven = dry * k.inTrim;
        const float adry = fabsf(dry);
        {
            const bool up = adry > st.env;
            st.env = (up ? k.envAtk : k.envRel) * st.env + (up ? k.omEnvAtk : k.omEnvRel) * adry;
        }
        const float impact = jlimitf(0.0f, 1.0f, jmaxf(0.0f, adry - st.env) * 10.0f);
        const float body = jlimitf(0.0f, 1.0f, st.env * 3.2f);
        const float trail = jlimitf(0.0f, 1.0f, 1.0f - impact) * k.tailShape;

        st.lp += k.splitLow * (driven - st.lp);
        st.hp += k.splitHigh * (driven - st.hp);
        const float low = st.lp * k.lowBoost;
        const float high = (driven - st.hp);
        const float mid = driven - st.lp - high;
        const float core = low + mid + high * k.highTilt;
        return Front { impact, body, trail, core };
    }
    // a1: the four modes' 2 r cos(theta) of this sample (metal only)
    __device__ __forceinline__ float back(float dry, TexChan& st, uint32_t& rng, float* line, float delayed, const Front& f,
                                          const float* a1) const
    {
        const TexCoef& k = *c;
        const float impact = f.impact, body = f.body, trail = f.trail, core = f.core;
        float shaped;
        if (mat() == 0) { // gel :137-151
            const float zeta = jmap3(trail, 0.62f, 1.45f);
            const float cc = 2.0f * zeta * k.gelOmega;
            const float force = core * (0.52f + 0.62f * body);
            const float acc = k.gelK * (force - st.springPos) - cc * st.springVel;
            st.springVel += acc;
            st.springPos += st.springVel;
            shaped = 0.48f * core + 1.85f * st.springPos;
            shaped = jblibm::tanhf_fdlibm(shaped * k.shapeGain); // the C library's own std::tanh: Texture's output feeds Width's threshold
        } else if (mat() == 1) { // metal :152-169
            const float exc = core * (0.19f + 0.52f * impact);
            const float m0 = mode(st, 0, exc, a1[0]);
            const float m1 = mode(st, 1, exc, a1[1]);
            const float m2 = mode(st, 2, exc, a1[2]);
            const float m3 = mode(st, 3, exc, a1[3]);
            const float modes = m0 + m1 + m2 + m3;
            const float brightExcite = 0.03f * impact * (core - st.hp);
            shaped = (0.44f * core + 0.42f * modes + brightExcite) * k.shapeGain;
        } else if (mat() == 2 || mat() == 3) { // wood :170-192, plastic :193-213
            const float exc = core * (k.excA + k.excB * impact);
            float newWave;
            if (mat() == 2)
                newWave = k.waveDamp * (0.62f * delayed + 0.38f * st.prevWave) + exc * (0.09f + 0.04f * body);
            else
                newWave = k.waveDamp * (0.76f * delayed + 0.24f * st.prevWave) + 0.14f * exc;
            line[(long long) waveIdx * pitch] = newWave;
            st.prevWave = delayed;
            const float w0 = mode(st, 0, exc, k.modeA1[0]);
            const float w1 = mode(st, 1, exc, k.modeA1[1]);
            const float w2 = mode(st, 2, exc, k.modeA1[2]);
            const float w3 = mode(st, 3, exc, k.modeA1[3]);
            shaped = (k.waveMixA * core + k.waveMixB * delayed + k.waveOut * (w0 + w1 + w2 + w3)) * k.shapeGain;
        } else { // flesh :214-236
            const float force = core * (0.55f + 0.65f * body);
            const float accA = k.kA * (force - st.fleshPosA) - k.cA * st.fleshVelA - k.kCouple * (st.fleshPosA - st.fleshPosB);
            const float accB = k.kB * (st.fleshPosA - st.fleshPosB) - k.cB * st.fleshVelB;
            st.fleshVelA += accA;
            st.fleshVelB += accB;
            st.fleshPosA += st.fleshVelA;
            st.fleshPosB += st.fleshVelB;
            const float tissue = 0.92f * st.fleshPosA + 0.58f * st.fleshPosB;
            const float nl = tissue - 0.19f * tissue * tissue * tissue;
            shaped = jblibm::tanhf_fdlibm((0.50f * core + 1.34f * nl) * k.shapeGain);
        }

        rng = 1664525u * rng + 1013904223u; // :239-243
        const float white = ((float) ((rng >> 8) & 0xFFFFu) / 32768.0f - 1.0f);
        st.noiseHp += 0.08f * (white - st.noiseHp);
        const float rough = white - st.noiseHp;
        shaped += rough * k.noiseAmt * (0.14f + 0.64f * impact);

        const float dynamics = 1.0f + impact * k.dynK + body * 0.06f;
        shaped *= dynamics * k.matTrim;

        const float tailInput = jlimitf(-2.0f, 2.0f, shaped) * (0.45f + 0.55f * trail);
        st.tail = tailInput + st.tail * k.decay;
        float wet = shaped + st.tail * (0.30f + 0.45f * trail);

        const float wetAbs = fabsf(wet); // :253-257
        {
            const bool up = wetAbs > st.wetEnv;
            st.wetEnv = (up ? k.wetAtk : k.wetRel) * st.wetEnv + (up ? k.omWetAtk : k.omWetRel) * wetAbs;
        }
        const float autoComp = div_rn_mid(k.autoGainBase, 1.0f + 1.8f * st.wetEnv);
        wet *= jlimitf(0.18f, 1.0f, autoComp);

        const float mixed = dry + k.mix * (wet - dry);
        float out = mixed * k.outGain;

        const float dcBlocked = out - st.dcIn + 0.995f * st.dcOut; // :263-265
        st.dcIn = out;
        st.dcOut = dcBlocked;

        const float peak = fabsf(dcBlocked); // :268-276
        {   // both arms, then a select (no branch inside the sample loop); the division's result is only used for peak > 0.88
            const float limited = jminf(st.protectGain, div_rn_mid(0.88f, jmaxf(peak, 0.5f)) * 0.98f);
            const float recovered = st.protectGain + (1.0f - st.protectGain) * 0.0028f;
            st.protectGain = peak > 0.88f ? limited : recovered;
        }
        out = dcBlocked * jlimitf(0.2f, 1.0f, st.protectGain);
        return jlimitf(-0.98f, 0.98f, out);
    }

    __device__ __forceinline__ void step(float& l, float& r)
    {
        float* line0 = wave;
        float* line1 = wave + (long long) c->waveSize * pitch;
        float d0 = 0.0f, d1 = 0.0f;
        if (mat() == 2 || mat() == 3) {
            if (havePref) {
                d0 = dq0[0]; dq0[0] = dq0[1]; dq0[1] = dq0[2]; dq0[2] = dq0[3];
                d1 = dq1[0]; dq1[0] = dq1[1]; dq1[1] = dq1[2]; dq1[2] = dq1[3];
            } else {
                d0 = waveRead(line0);
                d1 = waveRead(line1);
            }
        }
        const Front f0 = front(l, ch0), f1 = front(r, ch1);
        float a1L[4] = { 0.0f, 0.0f, 0.0f, 0.0f }, a1R[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
        if (mat() == 1) {
            // The poles bend with the transient (`bend = 1 + 0.09 impact`, :157-158), which costs a std::cos per mode,
            // channel and sample -- but impact is exactly 0 whenever the sample does not exceed its envelope, and then
            // the angle is the block-constant one of load().  Taken only when that holds for every lane of the warp
            // (a warp-uniform branch; same expression, same bits): between the hits of an impulse train or a drum tail
            // that is nearly always.
            const bool rest = __all_sync(__activemask(), f0.impact == 0.0f && f1.impact == 0.0f);
            if (rest) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    a1L[k] = a1R[k] = a1Rest[k];
            } else {
                const float bendL = 1.0f + 0.09f * f0.impact, bendR = 1.0f + 0.09f * f1.impact;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    a1L[k] = metalA1(k, bendL);
                    a1R[k] = metalA1(k, bendR);
                }
            }
        }
        l = back(l, ch0, rng0, line0, d0, f0, a1L);
        r = back(r, ch1, rng1, line1, d1, f1, a1R);
        if (mat() == 2 || mat() == 3)
            waveIdx = waveIdx + 1 == c->waveSize ? 0 : waveIdx + 1;
    }
    __device__ __forceinline__ void store(const Lane& L, const SlotDesc& d)
    {
        const int b = d.stateBase + AV_COUNT;
        ch0.store(L, b);
        ch1.store(L, b + TV_CH_STRIDE);
        L.sti(b + TV_WAVEIDX, waveIdx);
        L.sti(b + TV_RNG, (int) (L.a.nCh < 2 ? rng0 : rng1)); // state after both channels' draws (mono: channel 0's)
    }
};

// JuicyMotion/PluginProcessor.cpp:101-142.  variation*, motionPhase and budgetEnv
// are instance members walked by channel 0's whole block and then channel 1's, so
// the main part runs as two sequential channel passes (kSeqChannels).
struct MainMotion : MainBase {
    static constexpr bool kHas = true, kSeqChannels = true;
    static constexpr bool kHeavy = true; // heavy per-sample state: 4 samples per trip (registers); light: 8 (one sector per store)
    float vTone, vTrans, vTail, tTone, tTrans, tTail, phase, budget;
    float tail0, tail1, lp0, lp1, prev0, prev1, repScale, recovery;
    MotionCoef c;
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d, int)
    {
        c = d.c.motion;
        const int b = d.stateBase + AV_COUNT;
        vTone = L.ld(b + MV_VTONE); vTrans = L.ld(b + MV_VTRANS); vTail = L.ld(b + MV_VTAIL);
        tTone = L.ld(b + MV_TTONE); tTrans = L.ld(b + MV_TTRANS); tTail = L.ld(b + MV_TTAIL);
        phase = L.ld(b + MV_PHASE); budget = L.ld(b + MV_BUDGET);
        tail0 = L.ld(b + MV_TAIL0); tail1 = L.ld(b + MV_TAIL1);
        lp0 = L.ld(b + MV_LP0); lp1 = L.ld(b + MV_LP1);
        prev0 = L.ld(b + MV_PREV0); prev1 = L.ld(b + MV_PREV1);
        repScale = L.ld(b + MV_REP_SCALE); recovery = L.ld(b + MV_RECOVERY);
    }
    __device__ __forceinline__ float one(float dry, float lfoOffset, float& tail, float& lp, float& prev)
    {
        vTone = c.varSlew * vTone + c.omVarSlew * tTone;
        vTrans = c.varSlew * vTrans + c.omVarSlew * tTrans;
        vTail = c.varSlew * vTail + c.omVarSlew * tTail;
        phase += c.motionInc;
        phase = phase > 2.0f * PI_F ? phase - 2.0f * TWO_PI_F : phase; // sic (:114-115)

        const float motionLfo = sin_mid(phase + lfoOffset);
        const float cutoff = jlimitf(120.0f, 4200.0f, 900.0f + vTone * 1100.0f * c.d06 + motionLfo * c.lfoDepth);
        const float lpCoeff = 1.0f - expf(div_rn_mid(-2.0f * PI_F * cutoff, c.srf));
        lp += lpCoeff * (dry - lp);
        const float hp = dry - lp;
        const float transient = dry - prev;
        prev = dry;

        const float transientBoost = 1.0f + vTrans * 1.2f * c.d07 + c.mv035 * motionLfo * c.d08;
        const float toneShift = lp * (1.0f + vTone * 0.65f * c.d0507) + hp * transientBoost + transient * c.mvT * c.d0508;
        tail = toneShift + tail * jlimitf(0.0f, 0.93f, c.tailFeedback + vTail * 0.06f);

        float wet = toneShift * repScale * recovery + c.tailMix * tail;
        budget = c.budgetCoeff * budget + c.omBudget * fabsf(wet);
        const float limiterGain = budget > c.budgetTarget ? div_rn_mid(c.budgetTarget, budget + 1.0e-5f) : 1.0f;
        wet *= limiterGain;
        return (dry + c.mix * (wet * c.wetBoost - dry)) * c.outGain;
    }
    __device__ __forceinline__ float stepCh0(float l) { return one(l, 0.0f, tail0, lp0, prev0); }
    __device__ __forceinline__ void step(float&, float& r) { r = one(r, 0.85f, tail1, lp1, prev1); }
    __device__ __forceinline__ void store(const Lane& L, const SlotDesc& d)
    {
        const int b = d.stateBase + AV_COUNT;
        L.st(b + MV_VTONE, vTone); L.st(b + MV_VTRANS, vTrans); L.st(b + MV_VTAIL, vTail);
        L.st(b + MV_PHASE, phase); L.st(b + MV_BUDGET, budget);
        L.st(b + MV_TAIL0, tail0); L.st(b + MV_TAIL1, tail1);
        L.st(b + MV_LP0, lp0); L.st(b + MV_LP1, lp1);
        L.st(b + MV_PREV0, prev0); L.st(b + MV_PREV1, prev1);
    }
};

// ------------------------------------------------------------------ block pre-passes of the NEXT plugin

struct PreNone { // no next plugin
    static constexpr bool kHas = false;
    __device__ __forceinline__ void load(const Lane&, const SlotDesc&) {}
    __device__ __forceinline__ void step(float, float, float) {}
    __device__ __forceinline__ void finish(const Lane&, const SlotDesc&, int) {}
};
struct PreAna { // next plugin needs only its analyzer's pre pass
    static constexpr bool kHas = true;
    __device__ __forceinline__ void load(const Lane&, const SlotDesc&) {}
    __device__ __forceinline__ void step(float, float, float) {}
    __device__ __forceinline__ void finish(const Lane&, const SlotDesc&, int) {}
};

// JuicyCohere/PluginProcessor.cpp:62-96: band energies of the block's mono sum through
// the persistent 220/2400 Hz one-poles, optional target learning, context fit and
// the three compensation gains used by the main loop.
struct PreCohere {
    static constexpr bool kHas = true;
    float lowLp, highLp, eLow = 0.0f, eMid = 0.0f, eHigh = 0.0f;
    CohereCoef c;
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d)
    {
        c = d.c.cohere;
        const int b = d.stateBase + AV_COUNT;
        lowLp = L.ld(b + CV_LOWLP);
        highLp = L.ld(b + CV_HIGHLP);
    }
    __device__ __forceinline__ void step(float, float, float mono)
    {
        lowLp += c.lowCoeff * (mono - lowLp);
        highLp += c.highCoeff * (mono - highLp);
        const float low = lowLp;
        const float high = mono - highLp;
        const float mid = mono - low - high;
        eLow += low * low;
        eMid += mid * mid;
        eHigh += high * high;
    }
    static __device__ __forceinline__ float gainToDb(float g) { return g > 0.0f ? jmaxf(-100.0f, log10_exact(g) * 20.0f) : -100.0f; }
    __device__ void finish(const Lane& L, const SlotDesc& d, int n)
    {
        const int b = d.stateBase + AV_COUNT;
        const float inv = 1.0f / (float) (n > 1 ? n : 1);
        eLow *= inv;
        eMid *= inv;
        eHigh *= inv;
        float tLow = L.ld(b + CV_TGT_LOW), tMid = L.ld(b + CV_TGT_MID), tHigh = L.ld(b + CV_TGT_HIGH);
        if (c.learn) {
            tLow += (eLow - tLow) * 0.02f;
            tMid += (eMid - tMid) * 0.02f;
            tHigh += (eHigh - tHigh) * 0.02f;
            L.st(b + CV_TGT_LOW, tLow);
            L.st(b + CV_TGT_MID, tMid);
            L.st(b + CV_TGT_HIGH, tHigh);
        }
        const float lowErr = fabsf(gainToDb((eLow + 1.0e-6f) / (tLow + 1.0e-6f)));
        const float midErr = fabsf(gainToDb((eMid + 1.0e-6f) / (tMid + 1.0e-6f)));
        const float highErr = fabsf(gainToDb((eHigh + 1.0e-6f) / (tHigh + 1.0e-6f)));
        const float deviation = (lowErr + midErr + highErr) / 3.0f;
        const float contextFit = jlimitf(0.0f, 100.0f, 100.0f - deviation * 10.0f);
        L.st(b + CV_FIT, output_param(contextFit, 0.0f, 100.0f));
        L.st(b + CV_COMP_LOW, jlimitf(0.5f, 1.8f, pow_exact((tLow + 1.0e-6f) / (eLow + 1.0e-6f), c.matchQ)));
        L.st(b + CV_COMP_MID, jlimitf(0.5f, 1.8f, pow_exact((tMid + 1.0e-6f) / (eMid + 1.0e-6f), c.matchQ)));
        L.st(b + CV_COMP_HIGH, jlimitf(0.5f, 1.8f, pow_exact((tHigh + 1.0e-6f) / (eHigh + 1.0e-6f), c.matchQ)));
        L.st(b + CV_LOWLP, lowLp);
        L.st(b + CV_HIGHLP, highLp);
    }
};

// JuicyMotion/PluginProcessor.cpp:75-99: onset detector over the block's mono sum,
// LCG-drawn variation targets, repetition counter and the two block scalars.
struct PreMotion {
    static constexpr bool kHas = true;
    float env, repetition, tTone, tTrans, tTail;
    int cool;
    uint32_t rng;
    MotionCoef c;
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d)
    {
        c = d.c.motion;
        const int b = d.stateBase + AV_COUNT;
        env = L.ld(b + MV_ENV);
        repetition = L.ld(b + MV_REPETITION);
        tTone = L.ld(b + MV_TTONE);
        tTrans = L.ld(b + MV_TTRANS);
        tTail = L.ld(b + MV_TTAIL);
        cool = L.ldi(b + MV_COOLDOWN);
        rng = (uint32_t) L.ldi(b + MV_RNG);
    }
    __device__ __forceinline__ void step(float, float, float mono)
    {
        const float absMono = fabsf(mono);
        env = c.envCoeff * env + c.omEnv * absMono;
        if (cool > 0)
            --cool;
        if (absMono > env * 1.35f + 0.02f && cool <= 0) {
            cool = c.cooldownLen;
            repetition += 1.0f;
            rng = 1664525u * rng + 1013904223u;
            tTone = (((float) ((rng >> 7) & 0x7FFFu) / 16384.0f) - 1.0f) * c.microVar * 0.9f;
            rng = 1664525u * rng + 1013904223u;
            tTrans = (((float) ((rng >> 9) & 0x7FFFu) / 16384.0f) - 1.0f) * c.microVar * 0.8f;
            rng = 1664525u * rng + 1013904223u;
            tTail = (((float) ((rng >> 11) & 0x7FFFu) / 16384.0f) - 1.0f) * c.microVar * 0.8f;
        }
        repetition *= 0.997f;
    }
    __device__ void finish(const Lane& L, const SlotDesc& d, int)
    {
        const int b = d.stateBase + AV_COUNT;
        const float repNorm = jlimitf(0.0f, 1.0f, repetition * 0.08f);
        L.st(b + MV_REP_SCALE, 1.0f - c.repeatCtrl * repNorm * 0.65f);
        L.st(b + MV_RECOVERY, 1.0f + c.repeatCtrl * (1.0f - repNorm) * 0.25f);
        L.st(b + MV_ENV, env);
        L.st(b + MV_REPETITION, repetition);
        L.st(b + MV_TTONE, tTone);
        L.st(b + MV_TTRANS, tTrans);
        L.st(b + MV_TTAIL, tTail);
        L.sti(b + MV_COOLDOWN, cool);
        L.sti(b + MV_RNG, (int) rng);
    }
};

// ------------------------------------------------------------------ sample access (v1: per-lane row streaming)

struct Quad { float v[4]; };

__device__ __forceinline__ Quad load4(const float* p, int i, int n, bool vec)
{
    Quad q;
    if (vec && i + 4 <= n) {
        const float4 t = *reinterpret_cast<const float4*>(p + i);
        q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            q.v[k] = (i + k < n) ? p[i + k] : 0.0f;
    }
    return q;
}
__device__ __forceinline__ void store4(float* p, int i, int n, bool vec, const Quad& q)
{
    if (vec && i + 4 <= n) {
        *reinterpret_cast<float4*>(p + i) = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k < n)
                p[i + k] = q.v[k];
    }
}

// ---- sample access (v2): lane-local asynchronous prefetch ring.
// A lane streams its own two rows, so a warp's loads are 32 separate 16-byte pieces and their L2 /
// HBM latency (hundreds of cycles) used to sit in front of every four samples.  Each lane now copies
// its rows with cp.async (no register landing, no scoreboard wait) into a private 2 x 32-sample ring
// in shared memory 24 samples ahead of use and reads them back one quad ahead.  Piece k of a row is
// stored at k ^ (lane & 7) so that the 8 lanes of a quarter-warp hit 8 different bank groups.
#ifndef JB_LF_RING
#define JB_LF_RING 8   // quads per row in a lane's ring (8: 128 B per row, 8 KB per warp, 16+ warps per SM stay resident)
#endif
#ifndef JB_LF_AHEAD
#define JB_LF_AHEAD 6  // quads in flight per row (<= JB_LF_RING - 2: the octet loop issues two before it waits)
#endif
constexpr int LF_RING = JB_LF_RING, LF_AHEAD = JB_LF_AHEAD;
// The four-samples-per-trip path (uncached row pieces) looks further ahead: in place (out == in -- every host-buffer render,
// every plugin of a chain after the first) a load that lands in the 128-byte line the lane is currently storing to waits
// for those stores in L2; 10 quads (160 bytes) ahead always reach past that line (profiles/r01_s6_inplace.txt: the
// channel-per-lane kernel ran 2.2x slower in place with 6 ahead, the one-lane kernels 1.2x).  16-quad ring: 16 KB per warp.
constexpr int LF4_RING = 16, LF4_AHEAD = 10;

__device__ __forceinline__ uint32_t lf_smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
// CACHE_L1: the 32-byte sector a piece belongs to is kept in L1, so the row's next piece does not go to L2 again
// (big light batches, where L2 sector throughput is the bound); otherwise the copies bypass L1.
template <bool CACHE_L1>
__device__ __forceinline__ void lf_cp_async16(uint32_t dst, const void* src)
{
    if (CACHE_L1)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void lf_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lf_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ Quad lf_lds(uint32_t addr)
{
    Quad q;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.v[0]), "=f"(q.v[1]), "=f"(q.v[2]), "=f"(q.v[3]) : "r"(addr) : "memory");
    return q;
}
// One static shared buffer per CTA (= warp), used either as the lanes' private rings (LaneFeed, 8 KB) or as the warp's
// three-stage tile (tile streaming below, 15 KB)
constexpr int TILE_PITCH = 20;                          // floats per (row, stage): 16 samples + 4 pad -> conflict-free LDS.128
constexpr int TILE_STAGE_BYTES = 64 * TILE_PITCH * 4;   // 64 rows (2 channels x 32 clips)
constexpr int TILE_STAGES = 3;
__device__ __forceinline__ float4* lane_smem()
{
    extern __shared__ __align__(256) float4 jb_lane_dynamic_smem[]; // lane_smem_bytes(octets) at launch
    return jb_lane_dynamic_smem;
}
// TMA tile streaming (octets == 3): three stages x {left, right} tiles of 32 rows x 16 samples (2 KB each, the first one
// 1024-byte aligned for the swizzle), then three mbarriers.  13 KB per warp: 16 warps per SM stay resident, like the
// cp.async rings (with 32-sample tiles, 17 KB, only 12 fitted and a 65536-clip batch needed a second wave: +15 % time).
constexpr int TMA_S = 16;                                        // samples per stage
constexpr int TMA_STAGES = 3;
constexpr int TMA_TILE_BYTES = 32 * TMA_S * 4;
constexpr int TMA_SMEM_BYTES = 2 * TMA_STAGES * TMA_TILE_BYTES + 64 + 1024; // + slack to align the first tile
inline size_t lane_smem_bytes(int octets)
{
    if (octets == 3)
        return (size_t) TMA_SMEM_BYTES;
    return octets == 2 ? (size_t) TILE_STAGES * TILE_STAGE_BYTES
                       : (size_t) JB_LANE_CTA_THREADS * 2 * ((octets == 1 || octets == 4) ? LF_RING : LF4_RING) * 16;
}
__device__ __forceinline__ void lf_sts(uint32_t addr, const Quad& q)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(q.v[0]), "f"(q.v[1]), "f"(q.v[2]), "f"(q.v[3]) : "memory");
}

template <int RING>
struct LaneFeedT {
    uint32_t base;           // shared address of this lane's 256-byte ring (row L, then row R), swizzle folded in
    const float *srcL, *srcR;
    int nQuads;
    __device__ __forceinline__ void init(const float* l, const float* r, int n)
    {
        float4* ring = lane_smem();
        base = lf_smem_u32(&ring[threadIdx.x * 2 * RING]) ^ ((uint32_t) (threadIdx.x & 7) << 4);
        asm volatile("" : "+r"(base));
        srcL = l;
        srcR = r;
        nQuads = n >> 2;
    }
    template <bool CACHE_L1>
    __device__ __forceinline__ void issue(int q) const // quad q of both rows -> ring piece q & 7; always commits a group
    {
        if (q < nQuads) {
            const uint32_t off = (uint32_t) (q & (RING - 1)) << 4;
            lf_cp_async16<CACHE_L1>(base ^ off, srcL + 4 * q);
            lf_cp_async16<CACHE_L1>((base ^ off) + 16u * RING, srcR + 4 * q);
        }
        lf_commit();
    }
    __device__ __forceinline__ void read(int q, Quad& l, Quad& r) const
    {
        const uint32_t off = (uint32_t) (q & (RING - 1)) << 4;
        l = lf_lds(base ^ off);
        r = lf_lds((base ^ off) + 16u * RING);
    }
};

// Eight samples = one 32-byte sector per store (STG.256, sm_100): half the L2 write requests of two 16-byte stores.
__device__ __forceinline__ void store8(float* p, const Quad& a, const Quad& b)
{
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.v[0]), "f"(a.v[1]), "f"(a.v[2]), "f"(a.v[3]),
                 "f"(b.v[0]), "f"(b.v[1]), "f"(b.v[2]), "f"(b.v[3])
                 : "memory");
}

// ---- sample access (v4): TMA tile streaming.
// One row per lane means every 16-byte request of a warp touches 32 different lines; with cp.async rings that is 32 L1TEX
// wavefronts per LDGSTS, and the L1TEX pipe -- not HBM, not instruction issue -- bounded the light kernels on big batches
// (l1tex__throughput 78 - 82 % for Width / Cohere / Saturator, profiles/r02_c5_kernels_sections.txt).  Here the warp's 32
// left rows (and its 32 right rows) of a 16-sample stretch arrive with ONE cp.async.bulk.tensor.3d each: the tensor map
// views the audio as {sample, channel, clip}, the box is {16, 1, 32}, and the copy engine writes the tile
// [clip][16 samples] into shared memory without touching the LSU / L1TEX path at all.  64-byte swizzle: the 16-byte
// piece j of row r lands at r * 64 + ((j ^ ((r >> 1) & 3)) << 4), so the lanes of a quarter-warp, each reading piece j of
// its own row, hit eight different bank groups.  Three stages, one mbarrier each; lane 0 issues two stages ahead, every
// lane waits on the barrier's phase.  Stage g (counted over the whole kernel) uses buffer and barrier g % 3 in phase
// (g / 3) & 1.
__device__ __forceinline__ void tma_bar_init(uint32_t bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); }
__device__ __forceinline__ void tma_bar_expect(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t tma_tile_base() { return (lf_smem_u32(lane_smem()) + 1023u) & ~1023u; }
// once per kernel, by the whole warp, before the first sweep
__device__ __forceinline__ void tma_tiles_init()
{
    const uint32_t bars = tma_tile_base() + 2 * TMA_STAGES * TMA_TILE_BYTES;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < TMA_STAGES; ++i)
            tma_bar_init(bars + 8 * i);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}

// Eight samples = one 32-byte sector per LOAD (LDG.256, sm_100), straight into registers, not allocated in L1 (each sample is
// used once).  Counted in L1TEX wavefronts -- what bounds the light kernels on big batches -- a lane's eight samples cost
// one request here against two 16-byte cp.async pieces plus two shared-memory reads through the lane ring.
struct Oct { Quad a, b; };
__device__ __forceinline__ Oct load8(const float* p)
{
    Oct o;
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(o.a.v[0]), "=f"(o.a.v[1]), "=f"(o.a.v[2]), "=f"(o.a.v[3]), "=f"(o.b.v[0]), "=f"(o.b.v[1]), "=f"(o.b.v[2]), "=f"(o.b.v[3])
                 : "l"(p) : "memory");
    return o;
}

// One sweep over one block of one clip.  mainSlot < 0 for sweep 0.
// MONO: one-channel bus (isBusesLayoutSupported allows mono == mono, e.g. JuicyPunch/PluginProcessor.cpp:48-54).  The
// reference then loops over one channel only, its analyzer reads the right sample as the left one
// (src/shared/JuicinessAnalyzer.cpp:55-60), and Width does no DSP at all (JuicyWidth/PluginProcessor.cpp:76-89).  A
// compile-time variant so that the stereo kernels' code is untouched; only the generic kernel is instantiated for it.
// REUSE_STATS: the block's state-independent sums are taken from *carry instead of being accumulated again (JuicyInfer with
// trim = 0 dB leaves the buffer untouched between its two analyze() calls: JuicyInfer/PluginProcessor.cpp:78-80); otherwise
// a non-null carry receives this sweep's sums.
// Unroll factor of the tile-streaming loop's two 8-sample halves: 2 = both inline (16 samples per body), 1 = rolled (8 per
// body), 0 = one quad per trip.  Measured on B200 (profiles/r02_tile_rolled.txt): 1 is ahead for every light plugin -- the
// body is what has to stay inside the instruction caches.
#ifndef JB_TILE_UNROLL
#define JB_TILE_UNROLL 1
#endif
constexpr int JB_TILE_UNROLL_N = JB_TILE_UNROLL;
template <class Main, class Pre, bool MONO = false, bool REUSE_STATS = false>
__device__ __forceinline__ void sweep(const ProcArgs& a, long long clip, int mainSlot, int pos, int n, int blockAbs,
                                      BlockStats* carry = nullptr, unsigned* tmaCount = nullptr)
{
    const Lane L { a, clip };
    const int preSlot = mainSlot + 1;
    const bool firstRead = mainSlot <= 0; // sweep 0 and plugin 0's sweep read the caller's input
    const long long rowL = (clip * a.nCh) * a.rowPitch + pos;
    const long long rowR = MONO ? rowL : rowL + a.rowPitch;
    const float* srcL = (firstRead ? a.in : a.out) + rowL;
    const float* srcR = (firstRead ? a.in : a.out) + rowR;
    float* dstL = a.out + rowL;
    float* dstR = a.out + rowR;
    const bool vec = a.vecOk != 0;

    Main mainPart;
    AnaWalk post, pre;
    Pre prePart;
    BlockStats stats;
    const AnaCoef ana = a.ana;
    bool mustWrite = false;
    if constexpr (Main::kHas) {
        mainPart.load(L, a.slot[mainSlot], n);
        post.load(L, a.slot[mainSlot].stateBase);
        post.begin_block();
        mustWrite = mainPart.writes(a.in != a.out);
        if constexpr (MONO && Main::kMonoBypass)
            mustWrite = a.in != a.out;
    }
    if constexpr (Pre::kHas) {
        pre.load(L, a.slot[preSlot].stateBase);
        pre.begin_block();
        prePart.load(L, a.slot[preSlot]);
    }
    if constexpr (REUSE_STATS)
        stats = *carry;

    if constexpr (Main::kSeqChannels) { // Motion: channel 0's block first
        Quad q = load4(srcL, 0, n, vec);
        for (int i = 0; i < n; i += 4) {
            const Quad next = i + 4 < n ? load4(srcL, i + 4, n, vec) : q; // one quad ahead: its L2 latency hides behind this one
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k < n)
                    q.v[k] = mainPart.stepCh0(q.v[k]);
            store4(dstL, i, n, vec, q);
            q = next;
        }
        srcL = dstL;
    }

    // `whole`: all four samples exist (always so on the vector path, where every block is a multiple of four) -- no
    // per-sample bounds test, so the quad is ONE basic block and the scheduler can interleave the samples' and the two
    // channels' independent recurrences.
    auto quad_math = [&](Quad& ql, Quad& qr, int i, auto whole) {
        constexpr bool kSkipMain = MONO && (Main::kMonoBypass || Main::kSeqChannels); // Width: no DSP; Motion: no channel 1
        if (vec && !kSkipMain)
            mainPart.quad_begin();
        if constexpr (Main::kHas)
            post.group_begin();
        if constexpr (Pre::kHas)
            pre.group_begin();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (decltype(whole)::value || i + k < n) {
                float l = ql.v[k], r = qr.v[k];
                if constexpr (!kSkipMain)
                    mainPart.step(l, r);
                if constexpr (MONO)
                    r = l; // `right != nullptr ? right[i] : l`
                const float mono = 0.5f * (l + r);
                if constexpr (!REUSE_STATS)
                    stats.step(l, r, mono);
                if constexpr (Main::kHas)
                    post.step(k, mono, ana);
                if constexpr (Pre::kHas) {
                    pre.step(k, mono, ana);
                    prePart.step(l, r, mono);
                }
                ql.v[k] = l;
                qr.v[k] = r;
            }
        }
        {
            const int count = decltype(whole)::value ? 4 : min(4, n - i);
            if constexpr (Main::kHas)
                post.group_end(count, ana);
            if constexpr (Pre::kHas)
                pre.group_end(count, ana);
        }
        if (vec && !kSkipMain)
            mainPart.quad_end();
    };
    auto quad = [&](Quad& ql, Quad& qr, int i, auto whole) {
        quad_math(ql, qr, i, whole);
        if (mustWrite) {
            if (!Main::kSeqChannels)
                store4(dstL, i, n, vec, ql);
            if constexpr (!MONO)
                store4(dstR, i, n, vec, qr);
        }
    };
    constexpr std::true_type kWhole {};
    constexpr std::false_type kRagged {};
    if (vec && a.octets == 3 && !Main::kHeavy && tmaCount != nullptr) {
        // TMA tile streaming (see above): 16 samples per stage, eight at a time through the lane's math, 32-byte stores
        const uint32_t tiles = tma_tile_base();
        const uint32_t bars = tiles + 2 * TMA_STAGES * TMA_TILE_BYTES;
        const int lane = threadIdx.x;
        const int clip0 = (int) (blockIdx.x * blockDim.x);         // first clip of this warp within the launch
        const void* tmap = &a.tmapIn;
        const int nStages = (n + TMA_S - 1) / TMA_S;
        const unsigned g0 = *tmaCount;
        const uint32_t rowOff = (uint32_t) lane * (TMA_S * 4u), sw = (uint32_t) ((lane >> 1) & 3);
        const bool wide = ((reinterpret_cast<uintptr_t>(dstL) | reinterpret_cast<uintptr_t>(dstR)) & 31u) == 0;
        unsigned buf = g0 % TMA_STAGES, par = (g0 / TMA_STAGES) & 1u;   // buffer / barrier and phase of the stage being consumed
        unsigned ibuf = buf;                                        // ... of the stage being issued
        auto issue = [&](int k) {
            if (k < nStages && lane == 0) {
                const uint32_t bar = bars + 8u * ibuf, dst = tiles + 2u * TMA_TILE_BYTES * ibuf;
                tma_bar_expect(bar, 2u * TMA_TILE_BYTES);
                tma_load_3d(dst, tmap, pos + TMA_S * k, 0, clip0, bar);
                tma_load_3d(dst + TMA_TILE_BYTES, tmap, pos + TMA_S * k, 1, clip0, bar);
            }
            ibuf = ibuf + 1 == TMA_STAGES ? 0 : ibuf + 1;
        };
        issue(0);
        issue(1);
#pragma unroll 1
        for (int k = 0; k < nStages; ++k) {
            issue(k + 2);                                           // into the buffer stage k - 1 left (everybody is past it)
            tma_bar_wait(bars + 8u * buf, par);
            const uint32_t tL = tiles + 2u * TMA_TILE_BYTES * buf + rowOff, tR = tL + TMA_TILE_BYTES;
            const int cnt = min(TMA_S, n - TMA_S * k);            // a multiple of 4
#pragma unroll
            for (int h = 0; h < TMA_S; h += 8) {
                if (h + 8 <= cnt) {
                    const uint32_t j0 = (uint32_t) (h >> 2);
                    Quad l0 = lf_lds(tL + ((j0 ^ sw) << 4)), r0 = lf_lds(tR + ((j0 ^ sw) << 4));
                    Quad l1 = lf_lds(tL + (((j0 + 1u) ^ sw) << 4)), r1 = lf_lds(tR + (((j0 + 1u) ^ sw) << 4));
                    const int i = TMA_S * k + h;
                    quad_math(l0, r0, i, kWhole);
                    quad_math(l1, r1, i + 4, kWhole);
                    if (mustWrite) {
                        if (wide) {
                            if (!Main::kSeqChannels)
                                store8(dstL + i, l0, l1);
                            store8(dstR + i, r0, r1);
                        } else {
                            if (!Main::kSeqChannels) {
                                store4(dstL, i, n, vec, l0);
                                store4(dstL, i + 4, n, vec, l1);
                            }
                            store4(dstR, i, n, vec, r0);
                            store4(dstR, i + 4, n, vec, r1);
                        }
                    }
                }
            }
            if (cnt & 4) { // one last quad of a ragged block
                const uint32_t j0 = (uint32_t) ((cnt & ~7) >> 2);
                Quad ql = lf_lds(tL + ((j0 ^ sw) << 4)), qr = lf_lds(tR + ((j0 ^ sw) << 4));
                quad(ql, qr, TMA_S * k + (cnt & ~7), kWhole);
            }
            __syncwarp();   // every lane has read this stage's tiles
            if (++buf == TMA_STAGES) {
                buf = 0;
                par ^= 1u;
            }
        }
        *tmaCount = g0 + (unsigned) nStages;
    } else if (vec && a.octets == 2 && !Main::kHeavy) {
        // Warp-transposed tile streaming (big batches of light plugins).  With one row per lane every 16-byte request of a
        // warp touches 32 different lines: 32 L1TEX wavefronts per LDGSTS / per store, and the L1TEX pipe, not HBM and
        // not instruction issue, was the bound (l1tex__throughput 80 %, profiles/r01_s6_single_ncu.md).  Here the warp loads
        // its 64 rows 16 samples at a time with requests that cover 8 rows x 64 contiguous bytes (8 lines per request),
        // into a three-stage tile [channel][clip][16 + 4 pad] in shared memory; each lane then walks its own two rows
        // out of the tile (conflict-free LDS.128) and stores its results directly, 32 bytes per row at a time (sending
        // them back through the tile for 8-rows-per-request stores was measured slower: profiles/r01_s6_tile.txt).
        // All synchronisation is intra-warp (cp.async groups + __syncwarp).
        const uint32_t tileS = lf_smem_u32(lane_smem());
        const int lane = threadIdx.x;
        const uint32_t laneOff = (uint32_t) ((lane >> 2) * (TILE_PITCH * 4) + (lane & 3) * 16);
        const uint32_t myL = (uint32_t) (lane * TILE_PITCH * 4), myR = (uint32_t) ((32 + lane) * TILE_PITCH * 4);
        const float* srcBase = firstRead ? a.in : a.out;
        long long rowOff[8]; // element offset of the 16-byte piece this lane moves in request j
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int srow = j * 8 + (lane >> 2);
            const long long cj = __shfl_sync(0xffffffffu, clip, srow & 31);
            rowOff[j] = (cj * a.nCh + (srow >> 5)) * a.rowPitch + pos + (lane & 3) * 4;
        }
        const int nStages = n >> 4;
        const bool wide = ((reinterpret_cast<uintptr_t>(dstL) | reinterpret_cast<uintptr_t>(dstR)) & 31u) == 0;
        auto issue = [&](int k, int slot) {
            if (k < nStages) {
                const uint32_t st = tileS + (uint32_t) slot * TILE_STAGE_BYTES + laneOff;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    lf_cp_async16<false>(st + (uint32_t) j * (8 * TILE_PITCH * 4), srcBase + rowOff[j] + 16 * k);
            }
            lf_commit();
        };
        issue(0, 0);
        issue(1, 1);
        int slot = 0;
#pragma unroll 1
        for (int k = 0; k < nStages; ++k) {
            issue(k + 2, slot == 0 ? 2 : slot - 1); // (k + 2) % 3; that slot was drained by the previous iteration
            lf_wait<2>();                            // stage k has landed
            __syncwarp();
            const uint32_t st = tileS + (uint32_t) slot * TILE_STAGE_BYTES;
            if constexpr (JB_TILE_UNROLL_N == 0) {
                // one quad per trip of a rolled loop (the smallest body: it is the instruction caches, not the issue slots, that
                // the long unrolled bodies run out of); the first quad of an octet waits in registers for the second, so rows
                // still leave with 32-byte stores
                Quad sl, sr;
#pragma unroll 1
                for (int h = 0; h < 4; ++h) {
                    Quad l0 = lf_lds(st + myL + h * 16), r0 = lf_lds(st + myR + h * 16);
                    const int i = 16 * k + 4 * h;
                    quad_math(l0, r0, i, kWhole);
                    if (mustWrite) {
                        if (!wide) {
                            store4(dstL, i, n, vec, l0);
                            store4(dstR, i, n, vec, r0);
                        } else if (h & 1) {
                            store8(dstL + i - 4, sl, l0);
                            store8(dstR + i - 4, sr, r0);
                        } else {
                            sl = l0;
                            sr = r0;
                        }
                    }
                }
            } else
#pragma unroll JB_TILE_UNROLL_N
            for (int h = 0; h < 2; ++h) { // eight samples at a time; results leave with one 32-byte store per row
                Quad l0 = lf_lds(st + myL + h * 32), r0 = lf_lds(st + myR + h * 32);
                Quad l1 = lf_lds(st + myL + h * 32 + 16), r1 = lf_lds(st + myR + h * 32 + 16);
                const int i = 16 * k + 8 * h;
                quad_math(l0, r0, i, kWhole);
                quad_math(l1, r1, i + 4, kWhole);
                if (mustWrite) {
                    if (wide) {
                        store8(dstL + i, l0, l1);
                        store8(dstR + i, r0, r1);
                    } else {
                        store4(dstL, i, n, vec, l0);
                        store4(dstL, i + 4, n, vec, l1);
                        store4(dstR, i, n, vec, r0);
                        store4(dstR, i + 4, n, vec, r1);
                    }
                }
            }
            __syncwarp(); // before this slot is refilled
            slot = slot == 2 ? 0 : slot + 1;
        }
        lf_wait<0>();
        for (int i = nStages << 4; i < n; i += 4) { // the block's last 4 .. 12 samples, straight from global memory
            Quad ql = load4(srcL, i, n, vec);
            Quad qr = load4(srcR, i, n, vec);
            quad(ql, qr, i, kWhole);
        }
    } else if (vec) { // every quad is whole (n % 4 == 0): rows come through the lane's prefetch ring
        if (Main::kHeavy || !a.octets) {
            LaneFeedT<LF4_RING> feed;
            feed.init(srcL, srcR, n);
#pragma unroll
            for (int q = 0; q < LF4_AHEAD; ++q)
                feed.template issue<false>(q);
            lf_wait<LF4_AHEAD - 1>();
            Quad ql, qr, nl, nr;
            feed.read(0, ql, qr);
#pragma unroll 1
            for (int i = 0, q = 0; i < n; i += 4, ++q) {
                feed.template issue<false>(q + LF4_AHEAD);
                lf_wait<LF4_AHEAD - 1>(); // quads <= q + 1 have landed
                feed.read(q + 1, nl, nr);
                quad(ql, qr, i, kWhole);
                ql = nl;
                qr = nr;
            }
            lf_wait<0>();
        } else if (a.octets == 4 && (n & 7) == 0
                   && ((reinterpret_cast<uintptr_t>(srcL) | reinterpret_cast<uintptr_t>(srcR) | reinterpret_cast<uintptr_t>(dstL)
                        | reinterpret_cast<uintptr_t>(dstR)) & 31u) == 0) {
            // Register prefetch, 32 bytes per load and per store: octet o + 2 of both rows is requested before octet o is
            // worked on (two octets ~ 1000 instructions of this warp ahead of use, further than an HBM round trip at 16
            // warps per SM), the loop is unrolled by two so that the two octets in flight sit in fixed registers.
            const int nOct = n >> 3;
            Oct l0 = load8(srcL), r0 = load8(srcR);
            Oct l1 = l0, r1 = r0;
            if (nOct > 1) {
                l1 = load8(srcL + 8);
                r1 = load8(srcR + 8);
            }
            auto work = [&](Oct& l, Oct& r, int o) {
                quad_math(l.a, r.a, 8 * o, kWhole);
                quad_math(l.b, r.b, 8 * o + 4, kWhole);
                if (mustWrite) {
                    if (!Main::kSeqChannels)
                        store8(dstL + 8 * o, l.a, l.b);
                    store8(dstR + 8 * o, r.a, r.b);
                }
            };
#pragma unroll 1
            for (int o = 0; o < nOct; o += 2) {
                Oct cl = l0, cr = r0;
                if (o + 2 < nOct) {
                    l0 = load8(srcL + 8 * (o + 2));
                    r0 = load8(srcR + 8 * (o + 2));
                }
                work(cl, cr, o);
                if (o + 1 < nOct) {
                    cl = l1;
                    cr = r1;
                    if (o + 3 < nOct) {
                        l1 = load8(srcL + 8 * (o + 3));
                        r1 = load8(srcR + 8 * (o + 3));
                    }
                    work(cl, cr, o + 1);
                }
            }
        } else {
            LaneFeedT<LF_RING> feed;
            feed.init(srcL, srcR, n);
#pragma unroll
            for (int q = 0; q < LF_AHEAD; ++q)
                feed.issue<true>(q);
            const bool wide = ((reinterpret_cast<uintptr_t>(dstL) | reinterpret_cast<uintptr_t>(dstR)) & 31u) == 0;
            const int nOct = n >> 3;
            int q = 0;
#pragma unroll 1
            for (int o = 0; o < nOct; ++o, q += 2) {
                Quad l0, r0, l1, r1;
                feed.issue<true>(q + LF_AHEAD);
                feed.issue<true>(q + LF_AHEAD + 1);
                lf_wait<LF_AHEAD>();     // quads <= q + 1 have landed
                feed.read(q, l0, r0);
                feed.read(q + 1, l1, r1);
                quad_math(l0, r0, 4 * q, kWhole);
                quad_math(l1, r1, 4 * q + 4, kWhole);
                if (mustWrite) {
                    if (wide) {
                        if (!Main::kSeqChannels)
                            store8(dstL + 4 * q, l0, l1);
                        store8(dstR + 4 * q, r0, r1);
                    } else {
                        if (!Main::kSeqChannels) {
                            store4(dstL, 4 * q, n, vec, l0);
                            store4(dstL, 4 * q + 4, n, vec, l1);
                        }
                        store4(dstR, 4 * q, n, vec, r0);
                        store4(dstR, 4 * q + 4, n, vec, r1);
                    }
                }
            }
            lf_wait<0>();
            if (n & 4) { // one last quad
                Quad ql, qr;
                feed.read(q, ql, qr);
                quad(ql, qr, 4 * q, kWhole);
            }
        }
    } else {
        for (int i = 0; i < n; i += 4) {
            Quad ql = load4(srcL, i, n, vec);
            Quad qr = load4(srcR, i, n, vec);
            quad(ql, qr, i, kRagged);
        }
    }

    if constexpr (Main::kHas) {
        const SlotDesc& d = a.slot[mainSlot];
        mainPart.store(L, d);
        Metrics m = post.finish(L, d.stateBase, stats, n, ana);
        const float preScore = L.ld(d.stateBase + AV_PRE_SCORE);
        const float aux = d.kind == K_COHERE ? L.ld(d.stateBase + AV_COUNT + CV_FIT) : 0.0f;
        publish_record(a, mainSlot, clip, blockAbs, m, preScore, aux);
    }
    if constexpr (Pre::kHas) {
        const SlotDesc& d = a.slot[preSlot];
        const Metrics m = pre.finish(L, d.stateBase, stats, n, ana);
        L.st(d.stateBase + AV_PRE_SCORE, m.score);
        prePart.finish(L, d, n);
    }
    if constexpr (!REUSE_STATS)
        if (carry != nullptr)
            *carry = stats;
}


// ------------------------------------------------------------------ generic multi-plugin kernel

template <class Main, bool MONO>
__device__ __forceinline__ void sweep_pre_dispatch(const ProcArgs& a, long long clip, int mainSlot, int pos, int n, int blockAbs)
{
    const int preSlot = mainSlot + 1;
    if (preSlot >= a.chainLen) {
        sweep<Main, PreNone, MONO>(a, clip, mainSlot, pos, n, blockAbs);
        return;
    }
    switch (a.slot[preSlot].kind) {
        case K_COHERE: sweep<Main, PreCohere, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
        case K_MOTION: sweep<Main, PreMotion, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
        default: sweep<Main, PreAna, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
    }
}

// EXACT: Saturator / Punch with the C library's own tanh / pow (jb_libm.h); only the mono kernel instantiates it for a
// whole chain (the stereo engines render exact-math chains plugin by plugin with the jb_single kernels).
template <bool MONO, bool EXACT = false>
__device__ void sweep_dispatch(const ProcArgs& a, long long clip, int mainSlot, int pos, int n, int blockAbs)
{
    if (mainSlot < 0) {
        sweep_pre_dispatch<MainNone, MONO>(a, clip, mainSlot, pos, n, blockAbs);
        return;
    }
    const SlotDesc& d = a.slot[mainSlot];
    switch (d.kind) {
        case K_INFER: sweep_pre_dispatch<MainInfer, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
        case K_PUNCH: sweep_pre_dispatch<MainPunch<EXACT>, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
        case K_SAT: sweep_pre_dispatch<MainSat<EXACT>, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
        case K_WIDTH: sweep_pre_dispatch<MainWidth, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
        case K_COHERE: sweep_pre_dispatch<MainCohere, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
        case K_MOTION: sweep_pre_dispatch<MainMotion, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
        case K_TEXTURE:
            switch (d.c.tex.material) {
                case 0: sweep_pre_dispatch<MainTexture<0>, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
                case 1: sweep_pre_dispatch<MainTexture<1>, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
                case 2: sweep_pre_dispatch<MainTexture<2>, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
                case 3: sweep_pre_dispatch<MainTexture<3>, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
                default: sweep_pre_dispatch<MainTexture<4>, MONO>(a, clip, mainSlot, pos, n, blockAbs); break;
            }
            break;
        default: break;
    }
}


template <bool MONO, bool EXACT = false>
__global__ void __launch_bounds__(JB_CTA_THREADS, 16) jb_process_kernel(const __grid_constant__ ProcArgs a)
{
    const long long lane = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= a.nClips)
        return;
    const long long clip = a.clipMap != nullptr ? (long long) a.clipMap[lane] : lane;
    int blockAbs = a.histFirstBlock;
    for (int pos = 0; pos < a.nSamples; pos += a.blockSize, ++blockAbs) {
        const int n = min(a.blockSize, a.nSamples - pos);
        for (int s = -1; s < a.chainLen; ++s)
            sweep_dispatch<MONO, EXACT>(a, clip, s, pos, n, blockAbs);
    }
}

} // namespace
