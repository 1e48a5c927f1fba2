// jb_solo.cu -- few live streams: one CTA per clip (BASELINE.json configs[0], "a chunked ... scan over the
// block when few streams are live").
//
// With a handful of clips there is nothing to fill the GPU with, and the render time IS the longest dependent chain per
// clip.  In the lane kernels that chain holds everything -- loads, std::tanh, the tone one-pole, both analyzer passes --
// so one clip costs ~80 cycles per sample and sweep (JuicySaturator on the 10 s sweep of configs[0]: 39 ms, slower than
// one CPU core).  The only part that has to be sequential in sample order and cannot be re-associated is the analyzer's pair
// of attack / release envelopes with the onset machine (JuicinessAnalyzer.cpp:24-29, :64-75; a max of two affine maps has no
// fixed-size summary to scan), ~26 cycles per sample with one envelope per lane (profiles/microbench/env_chain2.cu).  So this
// kernel takes everything else OFF that chain:
//
//   helper warps (2, 3)  run one host block AHEAD: TMA bulk load of the block (cp.async.bulk + mbarrier), the plugin's
//                        elementwise math for all samples in parallel (std::tanh ...), its short linear recurrence on one
//                        lane per channel in the reference's operand order, the order-independent sums of analyze() by
//                        tree reduction, the mono signals before / after the plugin into shared memory, TMA bulk store;
//   warp 1, lanes 0-1    the analyzer's two band one-poles over the mono signals (linear, 8 cycles per sample);
//   warp 0, lanes 0-1    the short / long envelope, one per lane, exchanged by shuffle for the transient and the onset
//                        machine; lane 0 maps the features and publishes the record.
//
// Per block the envelope lanes walk 2 x 512 samples and nothing else sits in front of them.  Samples are the lane kernels'
// bit for bit (same operations per sample); records agree to rounding (the plain sums are tree-reduced, like the
// cooperative kernel's).  Plugins: JuicySaturator (either math mode), JuicyInfer, JuicyCohere (its block pre-pass and its
// per-channel recurrences are linear and short: one helper lane each, a block ahead of the walkers).  Reference lines per
// routine, relative to /root/reference.
#include "jb_device.cuh"
#include "jb_libm.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace {

using namespace jbdev;

// Profiling builds only (tools/runs/lib_variant.sh clocks jb_solo.cu -DJB_SOLO_CLOCKS): cycle counts of the helpers' stages and
// of the envelope walker's waits, printed by CTA 0 at the end of the launch.
#ifdef JB_SOLO_CLOCKS
#define SO_CLK(i) do { const long long now_ = clock64(); clk[i] += now_ - last_; last_ = now_; } while (0)
#else
#define SO_CLK(i) do { } while (0)
#endif
constexpr int SO_NMAX = 512;          // largest host block
constexpr int SO_THREADS = 128;       // warps 0, 1: walkers; warps 2, 3: helpers
constexpr int SO_HELPERS = 64;
// named barriers (0 is __syncthreads)
constexpr int SO_BAR_FULL = 1;        // +parity: helpers arrive, the envelope warp waits (block's signals and sums are in shared memory)
constexpr int SO_BAR_FULL1 = 10;      // +parity: the same hand-over to the band / finishing warp -- a barrier of its own: on a shared one
                                      // the envelope warp could not start a block before the finishing warp had ended the previous one
constexpr int SO_BAR_FREE = 3;        // +parity: walkers arrive, helpers wait   (walkers are done with that parity's buffers)
constexpr int SO_BAR_HELP = 5;        // helpers among themselves
constexpr int SO_BAR_TFULL = 6;       // +pre/post: envelope warp arrives, finishing warp waits (the walk's traces are complete)
constexpr int SO_BAR_TFREE = 8;       // +pre/post: finishing warp arrives, envelope warp waits (traces consumed)

struct SoloSums { // analyze()'s sums that do not depend on analyzer state (JuicinessAnalyzer.cpp:76-77, :86-91, :105-106)
    float rms, peak, side, corr;
    double l2, r2;
};

struct SoloNoExtra {};
struct SoloCohereExtra {
    float tr[6][SO_NMAX];         // one-pole traces: lowLp, highLp (mono), lpA / lpB of L, lpA / lpB of R; then matched / tail
    float e[3];                   // band energies of the block (low, mid, high), summed in sample order
    float err[3];                 // |dB| error of each band against its target
};
template <class Extra>
struct SoloSmemT {
    float xin[2][2][SO_NMAX];     // input block, by parity and channel (TMA destination)
    float out[2][2][SO_NMAX];     // output block (TMA source)
    float work[2][SO_NMAX];       // per channel: shaped sample, then the one-pole state per sample
    float monoIn[2][SO_NMAX];     // 0.5 (l + r) before / after the plugin, by parity
    float monoOut[2][SO_NMAX];
    SoloSums sums[2][2];          // [parity][before / after]
    SoloSums part[2][2];          // per helper warp partials [warp][before / after]
    float envTrace[2][2][SO_NMAX]; // [pre / post walk][short / long] envelope after every sample
    float bandAcc[2][2];          // [pre / post][low, high] energies handed from the band lanes to lane 0
    float blk[2][4];              // by parity: a plugin's block scalars (JuicyCohere: lowComp, midComp, highComp, contextfit)
    unsigned long long bar[2];    // mbarriers: input block of that parity has landed
    Extra extra;                  // the plugin's own scratch (JuicyCohere: filter traces)
};

__device__ __forceinline__ uint32_t so_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void so_mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(so_u32(bar)), "r"(count));
}
__device__ __forceinline__ void so_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(so_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void so_mbar_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(so_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void so_bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(so_u32(dst)), "l"(src),
                 "r"(bytes), "r"(so_u32(bar)) : "memory");
}
__device__ __forceinline__ void so_bulk_store(void* dst, const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(so_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void so_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void so_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void so_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void so_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void so_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void so_bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// ---------------------------------------------------------------- plugins: parallel part, sequential part, parallel part
// shape(): per-sample work that needs no state; scan(): the plugin's recurrence over the block, one lane per channel, in the
// reference's operand order; mix(): per-sample work after it.

template <bool EXACT>
struct SoloSat { // JuicySaturator/PluginProcessor.cpp:87-97
    static constexpr bool kScan = true, kBlockPre = false;
    static constexpr int kMinCtas = 5;
    using Extra = SoloNoExtra;
    SatCoef c;
    __device__ __forceinline__ void init(const SlotDesc& d) { c = d.c.sat; }
    __device__ __forceinline__ bool writes(bool) const { return true; }
    __device__ __forceinline__ float shape(float dry) const
    {
        const float driven = dry * c.inGain;
        const float skewed = driven + c.asym * driven * driven;
        return EXACT ? jblibm::tanhf_fdlibm(skewed) : tanh_fast(skewed);
    }
    __device__ __forceinline__ float scan(float soft, float& state) const
    {
        state += c.toneCoeff * (soft - state);
        return state;
    }
    __device__ __forceinline__ float mix(float dry, float state) const
    {
        const float wet = state * c.outGain;
        return dry + c.mix * (wet - dry);
    }
    static __device__ __forceinline__ int stateVar(int ch) { return AV_COUNT + SV_TONE0 + ch; }
};

struct SoloInfer { // JuicyInfer/PluginProcessor.cpp:79: buffer.applyGain(trimGain)
    static constexpr bool kScan = false, kBlockPre = false;
    static constexpr int kMinCtas = 5;
    using Extra = SoloNoExtra;
    InferCoef c;
    __device__ __forceinline__ void init(const SlotDesc& d) { c = d.c.infer; }
    __device__ __forceinline__ bool writes(bool outOfPlace) const { return c.gainMode != 0 || outOfPlace; }
    __device__ __forceinline__ float shape(float dry) const { return c.gainMode == 1 ? dry * c.trimGain : (c.gainMode == 2 ? 0.0f : dry); }
    __device__ __forceinline__ float scan(float v, float&) const { return v; }
    __device__ __forceinline__ float mix(float, float v) const { return v; }
    static __device__ __forceinline__ int stateVar(int) { return -1; }
};

// JuicyCohere/PluginProcessor.cpp:62-118.  Per block: a pre-pass over the mono sum (two one-poles carried across blocks, three
// band energies summed in sample order) that ends in the three compensation gains and the contextfit output; then per
// channel two one-poles restarted at 0 and the feedback tail.  Every recurrence is linear and takes two or three dependent
// operations per sample, so each gets a helper lane of its own and leaves its value after every sample in shared memory
// (stages below, a barrier between them); everything between the recurrences is done by all helpers in parallel.  Same
// operations per value as the lane kernels' PreCohere / MainCohere, hence the same bits.
struct SoloCohere {
    static constexpr bool kScan = false, kBlockPre = true;
    static constexpr int kMinCtas = 4;
    using Extra = SoloCohereExtra;
    CohereCoef c;
    float pole = 0.0f;                            // helper threads 0 / 1: lowLp / highLp of the pre-pass (carried across blocks)
    float target = 0.0f;                          // helper threads 0 .. 2 and 32 .. 34: the learnt target of band t & 31
    float tail = 0.0f;                            // helper threads 0 / 1: that channel's feedback tail
    __device__ __forceinline__ void init(const SlotDesc& d) { c = d.c.cohere; }
    __device__ __forceinline__ bool writes(bool) const { return true; }
    __device__ __forceinline__ float scan(float v, float&) const { return v; }
    __device__ __forceinline__ float mix(float, float v) const { return v; }   // stage_out leaves the finished sample
    static __device__ __forceinline__ int stateVar(int) { return -1; }
    template <class F>
    __device__ __forceinline__ void load(F stateAt, int t)
    {
        if (t < 2) {
            pole = *stateAt(AV_COUNT + CV_LOWLP + t);
            tail = *stateAt(AV_COUNT + CV_TAIL0 + t);
        }
        if ((t & 31) < 3)
            target = *stateAt(AV_COUNT + CV_TGT_LOW + (t & 31));
    }
    template <class F>
    __device__ __forceinline__ void store(F stateAt, int t, const float* lastBlk)
    {
        if (t < 2) {
            *stateAt(AV_COUNT + CV_LOWLP + t) = pole;
            *stateAt(AV_COUNT + CV_TAIL0 + t) = tail;
        }
        if (t < 3 && c.learn)
            *stateAt(AV_COUNT + CV_TGT_LOW + t) = target;
        if (t == 0) {
            *stateAt(AV_COUNT + CV_COMP_LOW) = lastBlk[0];
            *stateAt(AV_COUNT + CV_COMP_MID) = lastBlk[1];
            *stateAt(AV_COUNT + CV_COMP_HIGH) = lastBlk[2];
            *stateAt(AV_COUNT + CV_FIT) = lastBlk[3];
        }
    }
    static __device__ __forceinline__ float gainToDb(float g)
    {
        return g > 0.0f ? jmaxf(-100.0f, (float) log10((double) g) * 20.0f) : -100.0f;
    }
    static __device__ __forceinline__ float powx(float x, float y) { return (float) pow((double) x, (double) y); }
    // One one-pole `s += coef * (x - s)` over the block, its value after every sample into `trace` (:66-67, :108-109)
    static __device__ __forceinline__ float pole_walk(const float* x, float* trace, int n, float coef, float st)
    {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const int nq = n >> 2;
        float4 v = x4[0];
        for (int q = 0; q < nq; ++q) {
            const float4 nx = q + 1 < nq ? x4[q + 1] : v;
            float4 o;
            st += coef * (v.x - st); o.x = st;
            st += coef * (v.y - st); o.y = st;
            st += coef * (v.z - st); o.z = st;
            st += coef * (v.w - st); o.w = st;
            reinterpret_cast<float4*>(trace)[q] = o;
            v = nx;
        }
        return st;
    }
    // stage 1, helper threads 0 .. 5: the six one-poles of the block
    __device__ __forceinline__ void stage_poles(int t, const float* mono, const float* xl, const float* xr, SoloCohereExtra& x, int n)
    {
        if (t >= 6)
            return;
        const float coef = (t & 1) ? c.highCoeff : c.lowCoeff;
        const float* src = t < 2 ? mono : (t < 4 ? xl : xr);
        // one call for all six lanes (no divergence): lowLp / highLp are carried across blocks, lpA / lpB restart at 0 (:105-106)
        const float st = pole_walk(src, x.tr[t], n, coef, t < 2 ? pole : 0.0f);
        if (t < 2)
            pole = st;
    }
    // stage 2, helper threads 0 .. 2: one band energy each, summed in sample order (:68-73).  Quads, the next one fetched
    // while this one is summed: what is sequential here is only the chain of additions.
    __device__ __forceinline__ void stage_energy(int t, const float* mono, SoloCohereExtra& x, int n)
    {
        const float4* m4 = reinterpret_cast<const float4*>(mono);
        const float4* l4 = reinterpret_cast<const float4*>(x.tr[0]);
        const float4* h4 = reinterpret_cast<const float4*>(x.tr[1]);
        const int nq = n >> 2;
        float e = 0.0f;
        float4 m = m4[0], lo = l4[0], hp = h4[0];
        // the lane's band is chosen with selects: as `t == 0 ? low : ...` nvcc made a divergent branch per sample of it
        // (BSSY / BSYNC around every choice, 280 cycles per quad instead of 30)
        const int isLow = t == 0, isMid = t == 1;
        auto pick = [&](float low, float mid, float high) {
            float r;
            asm("{ .reg .pred p, q; setp.ne.s32 p, %4, 0; setp.ne.s32 q, %5, 0; selp.f32 %0, %2, %3, q; selp.f32 %0, %1, %0, p; }"
                : "=f"(r) : "f"(low), "f"(mid), "f"(high), "r"(isLow), "r"(isMid));
            return r;
        };
        for (int q = 0; q < nq; ++q) {
            const int qn = q + 1 < nq ? q + 1 : q;
            const float4 mN = m4[qn], loN = l4[qn], hpN = h4[qn];
            const float ms[4] = { m.x, m.y, m.z, m.w }, ls[4] = { lo.x, lo.y, lo.z, lo.w }, hs[4] = { hp.x, hp.y, hp.z, hp.w };
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float low = ls[k];
                const float high = ms[k] - hs[k];
                const float mid = ms[k] - low - high;
                const float v = pick(low, mid, high);
                e += v * v;
            }
            m = mN; lo = loN; hp = hpN;
        }
        x.e[t] = e;
    }
    // stage 3, six helper threads: the block scalars (:75-99).  The three gainToDecibels and the three std::pow calls are the
    // expensive part (double-precision routines, a few thousand cycles each) and independent of one another: threads 0 .. 2
    // (first helper warp) take the error in dB of band t, threads 32 .. 34 (second helper warp, so that the two routines do
    // not serialise as a divergent branch) the compensation gain of band t - 32; all six keep their band's target.
    __device__ void stage_scalars(int t, SoloCohereExtra& x, int n, float* blk)
    {
        const int band = t & 31;
        const float inv = 1.0f / (float) (n > 1 ? n : 1);
        const float e = x.e[band] * inv;
        if (c.learn)
            target += (e - target) * 0.02f;
        if (t < 32)
            x.err[band] = fabsf(gainToDb((e + 1.0e-6f) / (target + 1.0e-6f)));
        else
            blk[band] = jlimitf(0.5f, 1.8f, powx((target + 1.0e-6f) / (e + 1.0e-6f), c.matchQ));
    }
    // ... and helper thread 0: contextfit from the three errors (:86-90)
    __device__ __forceinline__ void stage_fit(const SoloCohereExtra& x, float* blk)
    {
        const float deviation = (x.err[0] + x.err[1] + x.err[2]) / 3.0f;
        const float contextFit = jlimitf(0.0f, 100.0f, 100.0f - deviation * 10.0f);
        blk[3] = output_param(contextFit, 0.0f, 100.0f);
    }
    // stage 4, all helpers: `matched` per sample and channel (:110-113), over the lpA trace; a quad per trip
    __device__ __forceinline__ void stage_matched(int t, const float* xl, const float* xr, SoloCohereExtra& x, int n, const float* blk)
    {
        const float lowComp = blk[0], midComp = blk[1], highComp = blk[2];
        const int nq = n >> 2;
        for (int i = t; i < 2 * nq; i += SO_HELPERS) {
            const int ch = i >= nq ? 1 : 0, q = i - ch * nq;
            const float4 d = reinterpret_cast<const float4*>(ch ? xr : xl)[q];
            const float4 a4 = reinterpret_cast<const float4*>(x.tr[2 + 2 * ch])[q], b4 = reinterpret_cast<const float4*>(x.tr[3 + 2 * ch])[q];
            const float dry[4] = { d.x, d.y, d.z, d.w }, lpA[4] = { a4.x, a4.y, a4.z, a4.w }, lpB[4] = { b4.x, b4.y, b4.z, b4.w };
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float low = lpA[k] * lowComp;
                const float high = (dry[k] - lpB[k]) * highComp;
                const float mid = (dry[k] - lpA[k] - (dry[k] - lpB[k])) * midComp;
                o[k] = low + mid + high;
            }
            reinterpret_cast<float4*>(x.tr[2 + 2 * ch])[q] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    // stage 5, helper threads 0 / 1: the feedback tail of one channel (:115), over the lpB trace
    __device__ __forceinline__ void stage_tail(int t, SoloCohereExtra& x, int n)
    {
        const float4* m4 = reinterpret_cast<const float4*>(x.tr[2 + 2 * t]);
        float4* t4 = reinterpret_cast<float4*>(x.tr[3 + 2 * t]);
        const int nq = n >> 2;
        float4 v = m4[0];
        for (int q = 0; q < nq; ++q) {
            const float4 nx = q + 1 < nq ? m4[q + 1] : v;
            float4 o;
            tail = v.x + tail * c.fb; o.x = tail;
            tail = v.y + tail * c.fb; o.y = tail;
            tail = v.z + tail * c.fb; o.z = tail;
            tail = v.w + tail * c.fb; o.w = tail;
            t4[q] = o;
            v = nx;
        }
    }
    // stage 6, all helpers: the finished sample (:116-117) into work[ch]; a quad per trip
    __device__ __forceinline__ void stage_out(int t, const float* xl, const float* xr, const SoloCohereExtra& x, float* workL, float* workR, int n)
    {
        const int nq = n >> 2;
        for (int i = t; i < 2 * nq; i += SO_HELPERS) {
            const int ch = i >= nq ? 1 : 0, q = i - ch * nq;
            const float4 d = reinterpret_cast<const float4*>(ch ? xr : xl)[q];
            const float4 m4 = reinterpret_cast<const float4*>(x.tr[2 + 2 * ch])[q], t4 = reinterpret_cast<const float4*>(x.tr[3 + 2 * ch])[q];
            const float dry[4] = { d.x, d.y, d.z, d.w }, matched[4] = { m4.x, m4.y, m4.z, m4.w }, tl[4] = { t4.x, t4.y, t4.z, t4.w };
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float wet = matched[k] + c.tailK * tl[k];
                o[k] = (dry[k] + c.mix * (wet - dry[k])) * c.outGain;
            }
            reinterpret_cast<float4*>(ch ? workR : workL)[q] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
};

__device__ __forceinline__ void sums_clear(SoloSums& s) { s.rms = s.peak = s.side = s.corr = 0.0f; s.l2 = s.r2 = 0.0; }
__device__ __forceinline__ void sums_step(SoloSums& s, float l, float r)
{
    const float mono = 0.5f * (l + r);
    s.rms = fmaf(mono, mono, s.rms);
    s.peak = fmaxf(s.peak, fabsf(mono));
    const float sd = 0.5f * (l - r);
    s.side = fmaf(sd, sd, s.side);
    s.corr = fmaf(l, r, s.corr);
    const double dl = (double) l, dr = (double) r;
    s.l2 = fma(dl, dl, s.l2);
    s.r2 = fma(dr, dr, s.r2);
}
__device__ __forceinline__ void sums_warp_reduce(SoloSums& s)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s.rms += __shfl_xor_sync(0xffffffffu, s.rms, d);
        s.peak = fmaxf(s.peak, __shfl_xor_sync(0xffffffffu, s.peak, d));
        s.side += __shfl_xor_sync(0xffffffffu, s.side, d);
        s.corr += __shfl_xor_sync(0xffffffffu, s.corr, d);
        s.l2 += __shfl_xor_sync(0xffffffffu, s.l2, d);
        s.r2 += __shfl_xor_sync(0xffffffffu, s.r2, d);
    }
}

// Resident CTAs per SM are set by shared memory (36.6 KB: six; JuicyCohere's traces, 48.6 KB: four), and the register budget
// follows that, not more: under a cap of 64 registers the JuicyCohere instantiation spilled inside the envelope walkers'
// loop -- a local-memory load on the kernel's one critical chain -- and ran at 3.7 ms per second of audio instead of 1.8.
template <class P>
__global__ void __launch_bounds__(SO_THREADS, P::kMinCtas) jb_solo_kernel(const __grid_constant__ ProcArgs a)
{
    extern __shared__ __align__(128) unsigned char soloRaw[];
    using SoloSmem = SoloSmemT<typename P::Extra>;
    SoloSmem& sm = *reinterpret_cast<SoloSmem*>(soloRaw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long clip = a.clipMap != nullptr ? (long long) a.clipMap[blockIdx.x] : (long long) blockIdx.x;
    const SlotDesc& d = a.slot[0];
    const AnaCoef& ana = a.ana;
    const int B = a.blockSize;
    const int nBlocks = (a.nSamples + B - 1) / B;
    const long long rowL = (clip * 2) * a.rowPitch, rowR = rowL + a.rowPitch;
    auto stateAt = [&](int var) -> float* { return a.state + (long long) (d.stateBase + var) * a.clipPitch + clip; };

    P plug;
    plug.init(d);
    const bool mustWrite = plug.writes(a.in != a.out);

    if (tid == 0) {
        so_mbar_init(&sm.bar[0], 1);
        so_mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= 2) {
        // ------------------------------------------------------------ helpers: everything that needs no analyzer state
        const int t = tid - 2 * 32;      // 0 .. 63
        const int hw = warp - 2;
        float scanState = 0.0f;          // the plugin's recurrence state of channel `t` (lanes 0, 1 of the first helper warp)
        if (P::kScan && t < 2)
            scanState = *stateAt(P::stateVar(t));
        if constexpr (P::kBlockPre)
            plug.load(stateAt, t);
        auto issue_load = [&](int b) { // thread 0 of the helpers
            const int pos = b * B, n = min(B, a.nSamples - pos), p = b & 1;
            so_expect_tx(&sm.bar[p], (uint32_t) (2 * n * 4));
            so_bulk_load(&sm.xin[p][0][0], a.in + rowL + pos, (uint32_t) (n * 4), &sm.bar[p]);
            so_bulk_load(&sm.xin[p][1][0], a.in + rowR + pos, (uint32_t) (n * 4), &sm.bar[p]);
        };
        if (t == 0) {
            issue_load(0);
            if (nBlocks > 1)
                issue_load(1);
        }
#ifdef JB_SOLO_CLOCKS
        long long clk[12] = {};
        long long last_ = clock64();
#endif
        for (int b = 0; b < nBlocks; ++b) {
            const int p = b & 1, pos = b * B, n = min(B, a.nSamples - pos), nq = n >> 2;
            SO_CLK(11);
            if (b >= 2)
                so_bar_sync(SO_BAR_FREE + p, SO_THREADS);   // the walkers have finished block b - 2
            SO_CLK(0);
            if (t == 0)
                so_bulk_wait_read();                        // ... and its output has left out[p]
            so_mbar_wait(&sm.bar[p], (uint32_t) ((b >> 1) & 1));
            so_bar_sync(SO_BAR_HELP, SO_HELPERS);           // nobody writes out[p] before thread 0's wait above is over
            SO_CLK(1);
            SoloSums before, after;
            sums_clear(before);
            sums_clear(after);
            const float4* xl = reinterpret_cast<const float4*>(&sm.xin[p][0][0]);
            const float4* xr = reinterpret_cast<const float4*>(&sm.xin[p][1][0]);
            for (int q = t; q < nq; q += SO_HELPERS) {
                const float4 l = xl[q], r = xr[q];
                sums_step(before, l.x, r.x); sums_step(before, l.y, r.y); sums_step(before, l.z, r.z); sums_step(before, l.w, r.w);
                reinterpret_cast<float4*>(&sm.monoIn[p][0])[q] =
                    make_float4(0.5f * (l.x + r.x), 0.5f * (l.y + r.y), 0.5f * (l.z + r.z), 0.5f * (l.w + r.w));
                if constexpr (!P::kBlockPre) {
                    reinterpret_cast<float4*>(&sm.work[0][0])[q] = make_float4(plug.shape(l.x), plug.shape(l.y), plug.shape(l.z), plug.shape(l.w));
                    reinterpret_cast<float4*>(&sm.work[1][0])[q] = make_float4(plug.shape(r.x), plug.shape(r.y), plug.shape(r.z), plug.shape(r.w));
                }
            }
            SO_CLK(2);
            if constexpr (P::kBlockPre) {
                const float *mono = &sm.monoIn[p][0], *dl = &sm.xin[p][0][0], *dr = &sm.xin[p][1][0];
                so_bar_sync(SO_BAR_HELP, SO_HELPERS);       // monoIn[p] is complete
                plug.stage_poles(t, mono, dl, dr, sm.extra, n);
                so_bar_sync(SO_BAR_HELP, SO_HELPERS);
                SO_CLK(3);
                if (t < 3)
                    plug.stage_energy(t, mono, sm.extra, n);
                so_bar_sync(SO_BAR_HELP, SO_HELPERS);
                SO_CLK(4);
                if (t < 3 || (t >= 32 && t < 35))
                    plug.stage_scalars(t, sm.extra, n, &sm.blk[p][0]);
                so_bar_sync(SO_BAR_HELP, SO_HELPERS);       // the compensation gains are in blk[p]
                SO_CLK(5);
                if (t == 0)
                    plug.stage_fit(sm.extra, &sm.blk[p][0]);                       // (read by the walkers after FULL, by store() at the end)
                plug.stage_matched(t, dl, dr, sm.extra, n, &sm.blk[p][0]);
                so_bar_sync(SO_BAR_HELP, SO_HELPERS);
                SO_CLK(6);
                if (t < 2)
                    plug.stage_tail(t, sm.extra, n);
                so_bar_sync(SO_BAR_HELP, SO_HELPERS);
                SO_CLK(7);
                plug.stage_out(t, dl, dr, sm.extra, &sm.work[0][0], &sm.work[1][0], n);
                so_bar_sync(SO_BAR_HELP, SO_HELPERS);
                SO_CLK(8);
            }
            if (P::kScan) {
                so_bar_sync(SO_BAR_HELP, SO_HELPERS);
                if (t < 2) { // one lane per channel, sample order, the reference's operands
                    float* w = &sm.work[t][0];
                    float4 v = *reinterpret_cast<const float4*>(w);
                    for (int q = 0; q < nq; ++q) {
                        const float4 nx = q + 1 < nq ? reinterpret_cast<const float4*>(w)[q + 1] : v;
                        float4 o;
                        o.x = plug.scan(v.x, scanState); o.y = plug.scan(v.y, scanState);
                        o.z = plug.scan(v.z, scanState); o.w = plug.scan(v.w, scanState);
                        reinterpret_cast<float4*>(w)[q] = o;
                        v = nx;
                    }
                }
                so_bar_sync(SO_BAR_HELP, SO_HELPERS);
            }
            for (int q = t; q < nq; q += SO_HELPERS) {
                const float4 l = xl[q], r = xr[q];
                const float4 sl = reinterpret_cast<const float4*>(&sm.work[0][0])[q], sr = reinterpret_cast<const float4*>(&sm.work[1][0])[q];
                const float4 yl = make_float4(plug.mix(l.x, sl.x), plug.mix(l.y, sl.y), plug.mix(l.z, sl.z), plug.mix(l.w, sl.w));
                const float4 yr = make_float4(plug.mix(r.x, sr.x), plug.mix(r.y, sr.y), plug.mix(r.z, sr.z), plug.mix(r.w, sr.w));
                sums_step(after, yl.x, yr.x); sums_step(after, yl.y, yr.y); sums_step(after, yl.z, yr.z); sums_step(after, yl.w, yr.w);
                reinterpret_cast<float4*>(&sm.monoOut[p][0])[q] =
                    make_float4(0.5f * (yl.x + yr.x), 0.5f * (yl.y + yr.y), 0.5f * (yl.z + yr.z), 0.5f * (yl.w + yr.w));
                reinterpret_cast<float4*>(&sm.out[p][0][0])[q] = yl;
                reinterpret_cast<float4*>(&sm.out[p][1][0])[q] = yr;
            }
            sums_warp_reduce(before);
            sums_warp_reduce(after);
            if (lane == 0) {
                sm.part[hw][0] = before;
                sm.part[hw][1] = after;
            }
            so_fence_async();                               // out[p] (generic proxy) -> visible to the bulk store
            so_bar_sync(SO_BAR_HELP, SO_HELPERS);
            if (t == 0) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const SoloSums &x = sm.part[0][k], &y = sm.part[1][k];
                    SoloSums s;
                    s.rms = x.rms + y.rms; s.peak = fmaxf(x.peak, y.peak); s.side = x.side + y.side; s.corr = x.corr + y.corr;
                    s.l2 = x.l2 + y.l2; s.r2 = x.r2 + y.r2;
                    sm.sums[p][k] = s;
                }
                if (mustWrite) {
                    so_bulk_store(a.out + rowL + pos, &sm.out[p][0][0], (uint32_t) (n * 4));
                    so_bulk_store(a.out + rowR + pos, &sm.out[p][1][0], (uint32_t) (n * 4));
                    so_bulk_commit();
                }
                if (b + 2 < nBlocks)
                    issue_load(b + 2);                      // xin[p] is free: every helper passed the barrier above
            }
            __threadfence_block();
            so_bar_arrive(SO_BAR_FULL + p, SO_HELPERS + 32);
            so_bar_arrive(SO_BAR_FULL1 + p, SO_HELPERS + 32);
            SO_CLK(9);
        }
#ifdef JB_SOLO_CLOCKS
        if (blockIdx.x == 0 && t == 0)
            printf("helper cycles per block: wait FREE %lld | load wait %lld | loop A %lld | poles %lld | energy %lld | scalars %lld | matched %lld | tail %lld | out %lld | loop B + store %lld | (loop top %lld)\n",
                   clk[0] / nBlocks, clk[1] / nBlocks, clk[2] / nBlocks, clk[3] / nBlocks, clk[4] / nBlocks, clk[5] / nBlocks, clk[6] / nBlocks,
                   clk[7] / nBlocks, clk[8] / nBlocks, clk[9] / nBlocks, clk[11] / nBlocks);
#endif
        if (P::kScan && t < 2)
            *stateAt(P::stateVar(t)) = scanState;
        if constexpr (P::kBlockPre)
            if (nBlocks > 0)
                plug.store(stateAt, t, &sm.blk[(nBlocks - 1) & 1][0]);
        if (t == 0)
            so_bulk_wait_all();
    } else {
        // ------------------------------------------------------------ walkers: the analyzer's state machine
        // Warp 0, lanes 0 / 1: the short / long envelope (updateEnvelope, JuicinessAnalyzer.cpp:24-29) and NOTHING else: each
        // leaves its value after every sample in a trace in shared memory, so the loop runs at the latency of its
        // FSETP -> FSEL -> FMUL -> FADD chain.  Warp 1: lanes 0 / 1 walk the two band one-poles (:79-84, the lane kernels' fused
        // form) over both signals of the block, then the whole warp finishes the analyze() calls from the traces: transient
        // = max(0, short - long) and its sum (tree-reduced), the onset machine (:69-75) restated on the set of samples whose
        // transient exceeds the threshold -- the reference decrements `onsetCooldown` once per sample and accepts an onset when
        // it has reached 0, i.e. an onset may fire at sample i iff i >= (cooldown at the start) - 1, and after one at i the
        // next may fire at i + len: same decisions -- then the feature mapping (:94-141) and the record.
        const bool act = lane < 2;
        const int ch = lane & 1;
        if (warp == 0) {
            float env = act ? *stateAt(AV_SHORT + ch) : 0.0f;
            const float cA = ch ? ana.aL : ana.aS, cR = ch ? ana.rL : ana.rS, cOmA = ch ? ana.omaL : ana.omaS, cOmR = ch ? ana.omrL : ana.omrS;
            auto walk = [&](const float* mono, int n, int which) {
                if (!act)
                    return;
                const int nq = n >> 2;
                const float4* m4 = reinterpret_cast<const float4*>(mono);
                float4* tr4 = reinterpret_cast<float4*>(&sm.envTrace[which][ch][0]);
                float4 v = m4[0];
                for (int q = 0; q < nq; ++q) {
                    const float4 nx = q + 1 < nq ? m4[q + 1] : v;
                    const float m[4] = { v.x, v.y, v.z, v.w };
                    float o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float av = fabsf(m[k]);
                        const bool up = av > env;
                        env = (up ? cOmA : cOmR) * av + (up ? cA : cR) * env;
                        o[k] = env;
                    }
                    tr4[q] = make_float4(o[0], o[1], o[2], o[3]);
                    v = nx;
                }
            };
#ifdef JB_SOLO_CLOCKS
            long long clk[12] = {};
            long long last_ = clock64();
#endif
            for (int b = 0; b < nBlocks; ++b) {
                const int p = b & 1, pos = b * B, n = min(B, a.nSamples - pos);
                SO_CLK(11);
                so_bar_sync(SO_BAR_FULL + p, SO_HELPERS + 32);
                SO_CLK(0);
                if (b > 0)
                    so_bar_sync(SO_BAR_TFREE, 64);                 // warp 1 is done with the previous block's pre trace
                SO_CLK(1);
                walk(&sm.monoIn[p][0], n, 0);                      // analyze(buffer) before the DSP (e.g. JuicySaturator/PluginProcessor.cpp:72)
                __threadfence_block();
                so_bar_arrive(SO_BAR_TFULL, 64);
                SO_CLK(2);
                if (b > 0)
                    so_bar_sync(SO_BAR_TFREE + 1, 64);
                SO_CLK(3);
                walk(&sm.monoOut[p][0], n, 1);                     // ... and after it (:100)
                __threadfence_block();
                so_bar_arrive(SO_BAR_TFULL + 1, 64);
                so_bar_arrive(SO_BAR_FREE + p, SO_THREADS);
                SO_CLK(4);
            }
#ifdef JB_SOLO_CLOCKS
            if (blockIdx.x == 0 && lane == 0)
                printf("envelope walker cycles per block: wait FULL %lld | wait TFREE pre %lld | pre walk %lld | wait TFREE post %lld | post walk %lld\n",
                       clk[0] / nBlocks, clk[1] / nBlocks, clk[2] / nBlocks, clk[3] / nBlocks, clk[4] / nBlocks);
#endif
            if (act)
                *stateAt(AV_SHORT + ch) = env;
        } else {
            float band = act ? *stateAt(AV_LOW + ch) : 0.0f;
            const float cBand = ch ? ana.highCoeff : ana.lowCoeff;
            int cool = __float_as_int(*stateAt(AV_COOLDOWN));      // every lane steps the onset machine alike
            float repEma = *stateAt(AV_REP_EMA), fatEma = *stateAt(AV_FAT_EMA), preScore = *stateAt(AV_PRE_SCORE);
            auto band_walk = [&](const float* mono, int n, int which) {
                if (!act)
                    return;
                const int nq = n >> 2;
                const float4* m4 = reinterpret_cast<const float4*>(mono);
                float acc = 0.0f;
                float4 v = m4[0];
                for (int q = 0; q < nq; ++q) {
                    const float4 nx = q + 1 < nq ? m4[q + 1] : v;
                    const float m[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        band = fmaf(cBand, m[k] - band, band);
                        const float e = ch ? m[k] - band : band;
                        acc = fmaf(e, e, acc);
                    }
                    v = nx;
                }
                sm.bandAcc[which][ch] = acc;
            };
            // transient sum and onset machine over the traces of one walk (all 32 lanes: lane L takes samples [16 L, 16 L + 16))
            auto transients = [&](int which, int n, float& trAcc, int& onsets) {
                const int nq = n >> 2;
                unsigned hot = 0u;
                float part = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = lane * 4 + j;
                    if (q < nq) {
                        const float4 sv = reinterpret_cast<const float4*>(&sm.envTrace[which][0][0])[q];
                        const float4 lv = reinterpret_cast<const float4*>(&sm.envTrace[which][1][0])[q];
                        const float t0 = fmaxf(0.0f, sv.x - lv.x), t1 = fmaxf(0.0f, sv.y - lv.y), t2 = fmaxf(0.0f, sv.z - lv.z),
                                    t3 = fmaxf(0.0f, sv.w - lv.w);
                        part += (t0 + t1) + (t2 + t3);
                        hot |= ((t0 > 0.045f ? 1u : 0u) | (t1 > 0.045f ? 2u : 0u) | (t2 > 0.045f ? 4u : 0u) | (t3 > 0.045f ? 8u : 0u)) << (4 * j);
                    }
                }
#pragma unroll
                for (int dd = 16; dd > 0; dd >>= 1)
                    part += __shfl_xor_sync(0xffffffffu, part, dd);
                trAcc = part;
                onsets = 0;
                int next = max(cool - 1, 0);      // first sample at which an onset may fire
                bool fired = false;
                int last = 0;
                // 16-sample groups holding a candidate, in order.  Candidates below `next` can never fire, and the cooldown
                // (35 ms) is longer than a block at any usual sample rate, so the walk is over the groups from `next` on and
                // ends with the first onset -- not over every sample above the threshold (on loud material that is most of
                // the block, and this warp, not the envelope warp, was what the kernel waited for).
                auto dropBelow = [](unsigned gr, int nx) { const int g0 = nx >> 4; return g0 >= 32 ? 0u : (gr & ~((1u << g0) - 1u)); };
                unsigned groups = dropBelow(__ballot_sync(0xffffffffu, hot != 0u), next);
                while (groups != 0u) {
                    const int g = __ffs((int) groups) - 1;
                    unsigned hmask = __shfl_sync(0xffffffffu, hot, g);
                    const int lo = next - 16 * g;                  // < 16: groups below `next` are gone
                    if (lo > 0)
                        hmask &= ~((1u << lo) - 1u);
                    if (hmask == 0u) {
                        groups &= groups - 1u;
                        continue;
                    }
                    const int i = 16 * g + (__ffs((int) hmask) - 1);
                    ++onsets;
                    next = i + max(ana.cooldownLen, 1);
                    fired = true;
                    last = i;
                    groups = dropBelow(groups, next);              // keeps group g while `next` still falls inside it
                }
                cool = fired ? max(ana.cooldownLen - (n - 1 - last), 0) : max(cool - n, 0);   // the counter after the last sample
            };
            auto finish = [&](const SoloSums& s, int which, int n, float trAcc, int onsets) -> Metrics {
                AnaState st;
                st.sEnv = 0.0f; st.lEnv = 0.0f; st.low = 0.0f; st.high = 0.0f;   // envelope / band states are not read by ana_finish
                st.repEma = repEma; st.fatEma = fatEma; st.cool = cool;
                AnaAcc acc;
                acc.trAcc = trAcc; acc.onsets = onsets;
                acc.lowAcc = sm.bandAcc[which][0]; acc.highAcc = sm.bandAcc[which][1];
                const StatSums ss { s.rms, s.peak, s.side, s.corr, s.l2, s.r2 };
                const Metrics m = ana_finish(st, acc, ss, n, ana);
                repEma = st.repEma;
                fatEma = st.fatEma;
                return m;
            };
#ifdef JB_SOLO_CLOCKS
            long long clk[12] = {};
            long long last_ = clock64();
#endif
            for (int b = 0; b < nBlocks; ++b) {
                const int p = b & 1, pos = b * B, n = min(B, a.nSamples - pos);
                SO_CLK(11);
                so_bar_sync(SO_BAR_FULL1 + p, SO_HELPERS + 32);
                SO_CLK(0);
                // Order matters for the envelope warp: it may start the next block's pre walk only once this warp has consumed
                // this block's pre trace, so that comes as early as possible -- between the two band walks.
                band_walk(&sm.monoIn[p][0], n, 0);
                SO_CLK(1);
                float trAcc, trAccPre;
                int onsets, onsetsPre;
                so_bar_sync(SO_BAR_TFULL, 64);                     // the pre walk's traces are complete
                SO_CLK(2);
                transients(0, n, trAccPre, onsetsPre);
                so_bar_arrive(SO_BAR_TFREE, 64);
                SO_CLK(3);
                band_walk(&sm.monoOut[p][0], n, 1);
                __syncwarp();                                      // both walks' band energies are in shared memory
                if (lane == 0)
                    preScore = finish(sm.sums[p][0], 0, n, trAccPre, onsetsPre).score;
                SO_CLK(4);
                so_bar_sync(SO_BAR_TFULL + 1, 64);
                SO_CLK(5);
                transients(1, n, trAcc, onsets);
                so_bar_arrive(SO_BAR_TFREE + 1, 64);
                SO_CLK(6);
                if (lane == 0) {
                    const Metrics m = finish(sm.sums[p][1], 1, n, trAcc, onsets);
                    publish_record(a, 0, clip, a.histFirstBlock + b, m, preScore, P::kBlockPre ? sm.blk[p][3] : 0.0f);
                }
                __syncwarp();
                so_bar_arrive(SO_BAR_FREE + p, SO_THREADS);
                SO_CLK(7);
            }
#ifdef JB_SOLO_CLOCKS
            if (blockIdx.x == 0 && lane == 0)
                printf("band / finishing warp cycles per block: wait FULL %lld | band walk pre %lld | wait pre trace %lld | transients pre %lld | band walk post + finish pre %lld | wait post trace %lld | transients post %lld | finish post + record %lld\n",
                       clk[0] / nBlocks, clk[1] / nBlocks, clk[2] / nBlocks, clk[3] / nBlocks, clk[4] / nBlocks, clk[5] / nBlocks, clk[6] / nBlocks, clk[7] / nBlocks);
#endif
            if (act)
                *stateAt(AV_LOW + ch) = band;
            if (lane == 0) {
                *stateAt(AV_COOLDOWN) = __int_as_float(cool);
                *stateAt(AV_REP_EMA) = repEma;
                *stateAt(AV_FAT_EMA) = fatEma;
                *stateAt(AV_PRE_SCORE) = preScore;
            }
        }
    }
}

template <class P>
int launch_solo(const ProcArgs& a, cudaStream_t stream)
{
    using SoloSmem = SoloSmemT<typename P::Extra>;
    cudaError_t e = cudaFuncSetAttribute(jb_solo_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(SoloSmem));
    if (e == cudaSuccess) {
        jb_solo_kernel<P><<<a.nClips, SO_THREADS, sizeof(SoloSmem), stream>>>(a);
        e = cudaGetLastError();
    }
    return (int) e;
}

} // namespace

// Can this launch (one plugin) take the clip-per-CTA kernel?  Stereo, whole quads, 16-byte aligned rows, host block <= 512.
extern "C" int jbk_solo_supported(const ProcArgs* a)
{
    if (a->chainLen != 1 || a->nCh != 2 || !a->vecOk || a->blockSize > SO_NMAX || a->blockSize % 4 != 0)
        return 0;
    const int k = a->slot[0].kind;
    return k == K_SAT || k == K_INFER || k == K_COHERE;
}

// Returns a cudaError_t.
extern "C" int jbk_launch_solo(const ProcArgs* args, void* stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    switch (args->slot[0].kind) {
        case K_SAT: return args->exactMath ? launch_solo<SoloSat<true>>(*args, st) : launch_solo<SoloSat<false>>(*args, st);
        case K_INFER: return launch_solo<SoloInfer>(*args, st);
        case K_COHERE: return launch_solo<SoloCohere>(*args, st);
        default: return (int) cudaErrorInvalidValue;
    }
}
