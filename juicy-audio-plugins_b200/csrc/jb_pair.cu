// jb_pair.cu -- two lanes per clip: one lane per (clip, channel), single-plugin launches of Texture / Saturator / Punch.
//
// Why: with one lane per clip a batch of 8192 clips is 256 warps on the GPU's 592 warp schedulers, each walking the two
// channels' recurrences AND both analyzer passes one dependent chain after the other (profiles/r01_s6_single_ncu.md:
// Texture 0.35 eligible warps per scheduler, `stall_wait` + `selected` = 65 %).  These plugins' channels are independent
// (JuicyTexture/PluginProcessor.cpp:107-278, JuicySaturator/...:83-98, JuicyPunch/...:86-112 loop over channels outside
// the sample loop), so the lane pair (2c, 2c + 1) of a warp takes clip c's left and right channel and exchanges samples
// with one shuffle; the analyzer (src/shared/JuicinessAnalyzer.cpp:57-92) is split SIMD-style so that both lanes run the
// SAME instructions on different data -- lane 0 the short envelope, the 250 Hz band, sum mono^2 and sum L^2; lane 1 the
// long envelope, the 2500 Hz band, sum side^2 and sum R^2 -- and the halves are joined with shuffles once per sample
// (the transient needs both envelopes) and once per block (feature mapping).  Same operations in the same order as the
// one-lane kernels, hence bit-identical samples and records (tests/test_gpu_parity.py::test_pair_kernel_*).
// Twice the warps, about half the instructions per lane: ~1.9x on latency-bound batches.
#include "jb_lane.cuh"

namespace {

using namespace jbdev;

// Quads in flight per lane, and the ring that holds them.  In place (out == in: every host-buffer render, every plugin of a
// chain after the first) a load that lands in the 128-byte line the lane is currently storing to waits for those stores in
// L2 -- with 6 quads (96 bytes) ahead that was most loads, and the kernel ran 2.2x slower in place than out of place
// (profiles/r01_s6_inplace.txt).  Ten quads ahead always reach past the line being written.
#ifndef JB_PF_AHEAD
#define JB_PF_AHEAD 10
#endif
#ifndef JB_PF_RING
#define JB_PF_RING 16
#endif
constexpr int PF_AHEAD = JB_PF_AHEAD, PF_RING = JB_PF_RING; // ring: 16 quads = 256 B per lane, 8 KB per warp // quads in flight (ring: 8 quads = 128 B per lane, 4 KB per warp)

struct LaneFeed1 {
    uint32_t base;
    const float* src;
    int nQuads;
    __device__ __forceinline__ void init(const float* p, int n)
    {
        static_assert(PF_AHEAD + 2 <= PF_RING && (PF_RING & (PF_RING - 1)) == 0, "ring must hold the quads in flight");
        __shared__ __align__(256) float4 ring[JB_LANE_CTA_THREADS * PF_RING];
        base = lf_smem_u32(&ring[threadIdx.x * PF_RING]) ^ ((uint32_t) (threadIdx.x & 7) << 4);
        asm volatile("" : "+r"(base));
        src = p;
        nQuads = n >> 2;
    }
    __device__ __forceinline__ void issue(int q) const
    {
        if (q < nQuads)
#ifdef JB_PAIR_CA
            lf_cp_async16<true>(base ^ ((uint32_t) (q & (PF_RING - 1)) << 4), src + 4 * q);
#else
            lf_cp_async16<false>(base ^ ((uint32_t) (q & (PF_RING - 1)) << 4), src + 4 * q);
#endif
        lf_commit();
    }
    __device__ __forceinline__ Quad read(int q) const { return lf_lds(base ^ ((uint32_t) (q & (PF_RING - 1)) << 4)); }
};

__device__ __forceinline__ float xchg(unsigned mask, float v) { return __shfl_xor_sync(mask, v, 1); }

// The analyzer of one clip spread over its two lanes (ch = 0 / 1); see the file header.
struct AnaPair {
    float env, band;           // short / long envelope, low / high one-pole
    float repEma, fatEma;      // both lanes (identical)
    int cool;                  // both lanes (identical)
    float cA, cR, cOmA, cOmR, cBand;
    // per-call accumulators
    float trAcc, bandAcc, accA, peak, corr;
    double own2;
    int onsets;
    __device__ __forceinline__ void load(const Lane& L, int base, int ch, const AnaCoef& c)
    {
        env = L.ld(base + AV_SHORT + ch);
        band = L.ld(base + AV_LOW + ch);
        cool = L.ldi(base + AV_COOLDOWN);
        repEma = L.ld(base + AV_REP_EMA);
        fatEma = L.ld(base + AV_FAT_EMA);
        cA = ch ? c.aL : c.aS;
        cR = ch ? c.rL : c.rS;
        cOmA = ch ? c.omaL : c.omaS;
        cOmR = ch ? c.omrL : c.omrS;
        cBand = ch ? c.highCoeff : c.lowCoeff;
    }
    __device__ __forceinline__ void begin()
    {
        trAcc = bandAcc = accA = peak = corr = 0.0f;
        own2 = 0.0;
        onsets = 0;
    }
    // ana_step_env + ana_step_bands + BlockStats::step (jb_device.cuh, jb_lane.cuh), one half per lane
    __device__ __forceinline__ void step(unsigned mask, int ch, float l, float r, float mono, float own, const AnaCoef& c)
    {
        const float a = fabsf(mono);
        {
            const bool up = a > env;
            env = (up ? cOmA : cOmR) * a + (up ? cA : cR) * env;
        }
        const float other = xchg(mask, env);
        const float tr = fmaxf(0.0f, (ch ? other : env) - (ch ? env : other));
        trAcc += tr;
        cool = max(cool - 1, 0);
        const bool onset = (tr > 0.045f) & (cool <= 0);
        onsets += onset ? 1 : 0;
        cool = onset ? c.cooldownLen : cool;

        band = fmaf(cBand, mono - band, band);
        const float v = ch ? mono - band : band;
        bandAcc = fmaf(v, v, bandAcc);

        const float x = ch ? 0.5f * (l - r) : mono;
        accA = fmaf(x, x, accA);
        peak = fmaxf(peak, a);
        corr = fmaf(l, r, corr);
        const double d = (double) own;
        own2 = fma(d, d, own2);
    }
    // joins the halves and maps the features (both lanes compute the same record)
    __device__ __forceinline__ Metrics finish(unsigned mask, int ch, int n, const AnaCoef& c)
    {
        const float envO = xchg(mask, env), bandO = xchg(mask, band), bandAccO = xchg(mask, bandAcc), accAO = xchg(mask, accA);
        const double own2O = __shfl_xor_sync(mask, own2, 1);
        AnaState st;
        st.sEnv = ch ? envO : env;
        st.lEnv = ch ? env : envO;
        st.low = ch ? bandO : band;
        st.high = ch ? band : bandO;
        st.repEma = repEma;
        st.fatEma = fatEma;
        st.cool = cool;
        AnaAcc acc;
        acc.trAcc = trAcc;
        acc.lowAcc = ch ? bandAccO : bandAcc;
        acc.highAcc = ch ? bandAcc : bandAccO;
        acc.onsets = onsets;
        const StatSums s { ch ? accAO : accA, peak, ch ? accA : accAO, corr, ch ? own2O : own2, ch ? own2 : own2O };
        const Metrics m = ana_finish(st, acc, s, n, c);
        repEma = st.repEma;
        fatEma = st.fatEma;
        return m;
    }
    __device__ __forceinline__ void store(const Lane& L, int base, int ch) const
    {
        L.st(base + AV_SHORT + ch, env);
        L.st(base + AV_LOW + ch, band);
        if (ch == 0) {
            L.sti(base + AV_COOLDOWN, cool);
            L.st(base + AV_REP_EMA, repEma);
            L.st(base + AV_FAT_EMA, fatEma);
        }
    }
};

// ------------------------------------------------------------------ per-channel plugin DSP

template <bool EXACT>
struct SatPair {
    static constexpr bool kHeavy = EXACT;
    static constexpr bool kRolled = false;
    MainSat<EXACT> M;
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d, int ch)
    {
        M.c = d.c.sat;
        M.s0 = L.ld(d.stateBase + AV_COUNT + SV_TONE0 + ch);
    }
    __device__ __forceinline__ void block_begin(int, int) {}
    __device__ __forceinline__ void block_end(unsigned, int) {}
    __device__ __forceinline__ void quad_begin() {}
    __device__ __forceinline__ float step(unsigned, float x) { return M.one(x, M.s0); }
    __device__ __forceinline__ void store(const Lane& L, const SlotDesc& d, int ch) { L.st(d.stateBase + AV_COUNT + SV_TONE0 + ch, M.s0); }
};

template <bool EXACT>
struct PunchPair {
    static constexpr bool kHeavy = true;
    static constexpr bool kRolled = false;
    MainPunch<EXACT> M;
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d, int ch)
    {
        M.c = d.c.punch;
        M.invTanhDrive = 1.0f / M.c.tanhDrive;
        M.f0 = L.ld(d.stateBase + AV_COUNT + PV_FAST0 + ch);
        M.sl0 = L.ld(d.stateBase + AV_COUNT + PV_SLOW0 + ch);
    }
    __device__ __forceinline__ void block_begin(int, int) {}
    __device__ __forceinline__ void block_end(unsigned, int) {}
    __device__ __forceinline__ void quad_begin() {}
    __device__ __forceinline__ float step(unsigned, float x) { return M.one(x, M.f0, M.sl0); }
    __device__ __forceinline__ void store(const Lane& L, const SlotDesc& d, int ch)
    {
        L.st(d.stateBase + AV_COUNT + PV_FAST0 + ch, M.f0);
        L.st(d.stateBase + AV_COUNT + PV_SLOW0 + ch, M.sl0);
    }
};

// x -> A^n x + C_n for the LCG of JuicyTexture/PluginProcessor.cpp:239 (channel 1 starts n draws ahead of channel 0)
__device__ __forceinline__ uint32_t lcg_skip(uint32_t x, int n)
{
    uint32_t A = 1664525u, C = 1013904223u, accA = 1u, accC = 0u;
    for (int k = n; k > 0; k >>= 1) {
        if (k & 1) {
            accA *= A;
            accC = accC * A + C;
        }
        C = (A + 1u) * C;
        A *= A;
    }
    return accA * x + accC;
}

template <int MAT, bool ROLLED = false>
struct TexPair {
    static constexpr bool kHeavy = true;
    // ROLLED: the quad's four samples go through ONE copy of the sample code instead of four.  Alone that is 2.3x slower
    // (gel 10.7 -> 24.2 ms: the unrolled quad overlaps a sample's analyzer with the next sample's DSP) -- but several
    // DIFFERENT Texture kernels side by side evict each other's unrolled loops (~50 KB each) from the SM's instruction
    // caches: five materials at once took 58 ms unrolled, 24.7 ms rolled, i.e. as long as one rolled kernel alone
    // (profiles/r02_tma.txt).  Used for launches of three or more parameter sets rendered concurrently.
    static constexpr bool kRolled = ROLLED;
    MainTexture<MAT> T; // this lane's channel lives in T.ch0 / T.rng0; T.waveIdx is kept in step by both lanes
    float* line;
    uint32_t rngStart;  // the instance's LCG state at the top of the block
    float dq[4];
    bool havePref;
    __device__ __forceinline__ void load(const Lane& L, const SlotDesc& d, int ch)
    {
        T.c = &d.c.tex;
        const int b = d.stateBase + AV_COUNT;
        T.ch0.load(L, b + ch * TV_CH_STRIDE);
        T.waveIdx = L.ldi(b + TV_WAVEIDX);
        rngStart = (uint32_t) L.ldi(b + TV_RNG);
        T.wave = L.a.texWave + L.clip;
        T.pitch = L.a.clipPitch;
        line = T.wave + (long long) ch * T.c->waveSize * T.pitch;
        havePref = false;
        if (T.mat() == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                T.a1Rest[k] = T.metalA1(k, 1.0f);
        }
    }
    __device__ __forceinline__ void block_begin(int n, int ch) { T.rng0 = ch ? lcg_skip(rngStart, n) : rngStart; }
    // the instance's state after the block is channel 1's (it drew last)
    __device__ __forceinline__ void block_end(unsigned mask, int ch)
    {
        const uint32_t other = __shfl_xor_sync(mask, T.rng0, 1);
        rngStart = ch ? T.rng0 : other;
    }
    __device__ __forceinline__ void quad_begin()
    {
        if (T.mat() != 2 && T.mat() != 3)
            return;
        float a0[4], a1[4], fr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int w = T.waveIdx + k;
            w = w >= T.c->waveSize ? w - T.c->waveSize : w;
            int i0, i1;
            T.wavePos(w, i0, i1, fr[k]);
            a0[k] = line[(long long) i0 * T.pitch];
            a1[k] = line[(long long) i1 * T.pitch];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            dq[k] = jmap3(fr[k], a0[k], a1[k]);
        havePref = true;
    }
    __device__ __forceinline__ float step(unsigned mask, float x)
    {
        float delayed = 0.0f;
        if (T.mat() == 2 || T.mat() == 3) {
            delayed = dq[0];
            dq[0] = dq[1]; dq[1] = dq[2]; dq[2] = dq[3];
        }
        const typename MainTexture<MAT>::Front f = T.front(x, T.ch0);
        float a1[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
        if (T.mat() == 1) {
            const bool rest = __all_sync(mask, f.impact == 0.0f); // see MainTexture::step
            if (rest) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    a1[k] = T.a1Rest[k];
            } else {
                const float bend = 1.0f + 0.09f * f.impact;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    a1[k] = T.metalA1(k, bend);
            }
        }
        const float y = T.back(x, T.ch0, T.rng0, line, delayed, f, a1);
        if (T.mat() == 2 || T.mat() == 3)
            T.waveIdx = T.waveIdx + 1 == T.c->waveSize ? 0 : T.waveIdx + 1;
        return y;
    }
    __device__ __forceinline__ void store(const Lane& L, const SlotDesc& d, int ch)
    {
        const int b = d.stateBase + AV_COUNT;
        T.ch0.store(L, b + ch * TV_CH_STRIDE);
        if (ch == 0) {
            L.sti(b + TV_WAVEIDX, T.waveIdx);
            L.sti(b + TV_RNG, (int) rngStart);
        }
    }
};

// ------------------------------------------------------------------ kernel

template <class PM, int MIN_CTAS>
__global__ void __launch_bounds__(JB_CTA_THREADS, MIN_CTAS) jb_pair_kernel(const __grid_constant__ ProcArgs a)
{
    const long long gl = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long pair = gl >> 1;
    const int ch = (int) (gl & 1);
    if (pair >= a.nClips)
        return;
    const unsigned mask = __activemask(); // both lanes of a pair are in or out together
    const long long clip = a.clipMap != nullptr ? (long long) a.clipMap[pair] : pair;
    const Lane L { a, clip };
    const SlotDesc& d = a.slot[0];
    const AnaCoef ana = a.ana;
    const long long row = (clip * a.nCh + ch) * a.rowPitch;

    AnaPair an;
    an.load(L, d.stateBase, ch, ana);
    PM pm;
    pm.load(L, d, ch);

    int blockAbs = a.histFirstBlock;
    float lastPreScore = 0.0f;
    for (int pos = 0; pos < a.nSamples; pos += a.blockSize, ++blockAbs) {
        const int n = min(a.blockSize, a.nSamples - pos); // a multiple of 4 (launcher)
        const float* src = a.in + row + pos;
        float* dst = a.out + row + pos;
        LaneFeed1 feed;

        // analyze(buffer) before the DSP (e.g. JuicyTexture/PluginProcessor.cpp:58)
        an.begin();
        feed.init(src, n);
#pragma unroll
        for (int q = 0; q < PF_AHEAD; ++q)
            feed.issue(q);
        lf_wait<PF_AHEAD - 1>();
        Quad cur = feed.read(0), nxt;
#pragma unroll 1
        for (int i = 0, q = 0; i < n; i += 4, ++q) {
            feed.issue(q + PF_AHEAD);
            lf_wait<PF_AHEAD - 1>();
            nxt = feed.read(q + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float x = cur.v[k];
                const float o = xchg(mask, x);
                const float l = ch ? o : x, r = ch ? x : o;
                an.step(mask, ch, l, r, 0.5f * (l + r), x, ana);
            }
            cur = nxt;
        }
        lf_wait<0>();
        const float preScore = an.finish(mask, ch, n, ana).score;
        lastPreScore = preScore;

        // the plugin's per-sample loop, then analyze(buffer) on its output
        an.begin();
        pm.block_begin(n, ch);
        feed.init(src, n);
#pragma unroll
        for (int q = 0; q < PF_AHEAD; ++q)
            feed.issue(q);
        lf_wait<PF_AHEAD - 1>();
        cur = feed.read(0);
#pragma unroll 1
        for (int i = 0, q = 0; i < n; i += 4, ++q) {
            feed.issue(q + PF_AHEAD);
            lf_wait<PF_AHEAD - 1>();
            nxt = feed.read(q + 1);
            pm.quad_begin();
            if constexpr (PM::kRolled) {
#pragma unroll 1
                for (int k = 0; k < 4; ++k) { // the quad rotates through cur.v[0]: no run-time register indexing
                    const float y = pm.step(mask, cur.v[0]);
                    const float o = xchg(mask, y);
                    const float l = ch ? o : y, r = ch ? y : o;
                    an.step(mask, ch, l, r, 0.5f * (l + r), y, ana);
                    cur.v[0] = cur.v[1]; cur.v[1] = cur.v[2]; cur.v[2] = cur.v[3]; cur.v[3] = y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float y = pm.step(mask, cur.v[k]);
                    const float o = xchg(mask, y);
                    const float l = ch ? o : y, r = ch ? y : o;
                    an.step(mask, ch, l, r, 0.5f * (l + r), y, ana);
                    cur.v[k] = y;
                }
            }
#if defined(JB_PAIR_STORE_CS)
            asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(cur.v[0]), "f"(cur.v[1]), "f"(cur.v[2]), "f"(cur.v[3]) : "memory");
#elif defined(JB_PAIR_STORE_CG)
            asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(cur.v[0]), "f"(cur.v[1]), "f"(cur.v[2]), "f"(cur.v[3]) : "memory");
#else
            *reinterpret_cast<float4*>(dst + i) = make_float4(cur.v[0], cur.v[1], cur.v[2], cur.v[3]);
#endif
            cur = nxt;
        }
        lf_wait<0>();
        pm.block_end(mask, ch);
        const Metrics m = an.finish(mask, ch, n, ana);
        if (ch == 0)
            publish_record(a, 0, clip, blockAbs, m, preScore, 0.0f);
    }
    an.store(L, d.stateBase, ch);
    if (ch == 0)
        L.st(d.stateBase + AV_PRE_SCORE, lastPreScore); // what the one-lane kernels leave there
    pm.store(L, d, ch);
}

template <class PM>
cudaError_t launch_pair(const ProcArgs& a, cudaStream_t stream)
{
    const int grid = (int) (((long long) a.nClips * 2 + JB_CTA_THREADS - 1) / JB_CTA_THREADS);
    if (PM::kHeavy && grid <= 8 * 148)
        jb_pair_kernel<PM, 8><<<grid, JB_CTA_THREADS, 0, stream>>>(a);
    else
        jb_pair_kernel<PM, 16><<<grid, JB_CTA_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

} // namespace

// Can this launch (args->chainLen == 1) run two lanes per clip?  Plugin kinds with independent channels, whole quads,
// 16-byte aligned rows.
extern "C" int jbk_pair_supported(const ProcArgs* a)
{
    if (a->chainLen != 1 || a->nCh != 2 || !a->vecOk)
        return 0;
    const int k = a->slot[0].kind;
    return k == K_TEXTURE || k == K_SAT || k == K_PUNCH;
}

// Returns a cudaError_t.
extern "C" int jbk_launch_pair(const ProcArgs* args, void* stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    const SlotDesc& d = args->slot[0];
    switch (d.kind) {
        case K_SAT: return (int) (args->exactMath ? launch_pair<SatPair<true>>(*args, st) : launch_pair<SatPair<false>>(*args, st));
        case K_PUNCH: return (int) (args->exactMath ? launch_pair<PunchPair<true>>(*args, st) : launch_pair<PunchPair<false>>(*args, st));
        case K_TEXTURE:
            if (args->smallCode) {
                switch (d.c.tex.material) {
                    case 0: return (int) launch_pair<TexPair<0, true>>(*args, st);
                    case 1: return (int) launch_pair<TexPair<1, true>>(*args, st);
                    case 2: return (int) launch_pair<TexPair<2, true>>(*args, st);
                    case 3: return (int) launch_pair<TexPair<3, true>>(*args, st);
                    default: return (int) launch_pair<TexPair<4, true>>(*args, st);
                }
            }
            switch (d.c.tex.material) {
                case 0: return (int) launch_pair<TexPair<0>>(*args, st);
                case 1: return (int) launch_pair<TexPair<1>>(*args, st);
                case 2: return (int) launch_pair<TexPair<2>>(*args, st);
                case 3: return (int) launch_pair<TexPair<3>>(*args, st);
                default: return (int) launch_pair<TexPair<4>>(*args, st);
            }
        default: return (int) cudaErrorInvalidValue;
    }
}
