// jb_kernels.h -- plain-old-data shared by the host side (jb_engine.cpp, jb_params.cpp)
// and the device code (jb_kernels.cu).  Everything a kernel needs arrives in one
// ProcArgs value (kernel parameter -> constant bank, i.e. warp-uniform operands).
//
// Division of labour: every block-constant coefficient is
// computed on the HOST with the same glibc libm calls and the same fp32 operand
// order as the reference's processBlock prologue (files cited per struct), so the
// recurrences run on bit-identical coefficients; the device only evaluates the
// per-sample arithmetic (compiled with -fmad=false -ftz=true -prec-div=true
// -prec-sqrt=true) and the per-sample / per-block transcendentals.
#pragma once
#include <stddef.h>
#include <stdint.h>

#define JBK_MAX_CHAIN 8
#define JBK_REC 16 // floats per metrics record (jb_metrics)

// ---- state variables, structure-of-arrays: state[(slotBase + var) * clipPitch + clip]
enum AnaVar { AV_SHORT = 0, AV_LONG, AV_LOW, AV_HIGH, AV_REP_EMA, AV_FAT_EMA, AV_COOLDOWN, AV_PRE_SCORE, AV_COUNT };

enum PunchVar { PV_FAST0 = 0, PV_FAST1, PV_SLOW0, PV_SLOW1, PV_COUNT };
enum SatVar { SV_TONE0 = 0, SV_TONE1, SV_COUNT };
enum WidthVar { WV_WPOS = 0, WV_COUNT };
enum CohereVar {
    CV_LOWLP = 0, CV_HIGHLP, CV_TAIL0, CV_TAIL1, CV_TGT_LOW, CV_TGT_MID, CV_TGT_HIGH,
    CV_COMP_LOW, CV_COMP_MID, CV_COMP_HIGH, CV_FIT, CV_COUNT
};
// Texture: TV_CH_STRIDE variables per channel, then the shared ones
enum TexVar {
    TV_TAIL = 0, TV_LP, TV_HP, TV_ENV, TV_WETENV, TV_NOISEHP, TV_DCIN, TV_DCOUT, TV_PROTECT,
    TV_SPRING_POS, TV_SPRING_VEL, TV_FLESH_PA, TV_FLESH_VA, TV_FLESH_PB, TV_FLESH_VB, TV_PREVWAVE,
    TV_Y1_0, TV_Y1_1, TV_Y1_2, TV_Y1_3, TV_Y2_0, TV_Y2_1, TV_Y2_2, TV_Y2_3, TV_CH_STRIDE,
    TV_WAVEIDX = 2 * TV_CH_STRIDE, TV_RNG, TV_COUNT
};
enum MotionVar {
    MV_ENV = 0, MV_REPETITION, MV_BUDGET, MV_VTONE, MV_VTRANS, MV_VTAIL, MV_TTONE, MV_TTRANS, MV_TTAIL,
    MV_TAIL0, MV_TAIL1, MV_LP0, MV_LP1, MV_PREV0, MV_PREV1, MV_PHASE, MV_COOLDOWN, MV_RNG,
    MV_REP_SCALE, MV_RECOVERY, MV_COUNT
};

// ---- coefficients

// JuicinessAnalyzer::prepare + the per-call coefficients of analyze()
// (src/shared/JuicinessAnalyzer.cpp:3-11, :38-41)
struct AnaCoef {
    float aS, rS, aL, rL;             // attackShort, releaseShort, attackLong, releaseLong
    float omaS, omrS, omaL, omrL;     // (1.0f - coeff) for each, rounded like the reference's expression
    float lowCoeff, highCoeff;
    float srf;                        // static_cast<float>(sr)
    int cooldownLen;                  // static_cast<int>(sr * 0.035)
};

// JuicySaturator/PluginProcessor.cpp:74-81
struct SatCoef { float inGain, outGain, asym, toneCoeff, mix; };

// JuicyPunch/PluginProcessor.cpp:74-84 and the loop-invariant sub-expressions of :94-110
struct PunchCoef {
    float fastCoeff, omFast, slowCoeff, omSlow;
    float curveExp;      // jmap(slam, 0,1, 0.95, 0.55)
    float punchK;        // punch*12 + slam*22
    float sustainK;      // sustain*4 + slam*1.5
    float drive;         // 1 + clip*8 + slam*4
    float tanhDrive;     // std::tanh(drive)
    float hardK;         // 1 + clip*2
    float clipAmt, mix, outGain;
};

// JuicyWidth/PluginProcessor.cpp:91-97, :110
struct WidthCoef { float width, dynamicLimit, mix, outGain; int delaySamples, ringLen; };

// JuicyCohere/PluginProcessor.cpp:13-22, :54-60, :97
struct CohereCoef { float lowCoeff, highCoeff, matchQ /*0.25*match*/, tailK /*tail*0.35*/, fb, mix, outGain; int learn; };

// JuicyInfer/PluginProcessor.cpp:74-81
struct InferCoef { float trimGain, sensitivity; int gainMode; /*0 skip (gain ~ 1), 1 multiply, 2 clear (gain == 0)*/ };

// JuicyTexture/PluginProcessor.cpp:55-75 plus loop-invariant sub-expressions of :116-276
struct TexCoef {
    int material;                     // static_cast<int>(raw "material")
    float inTrim, matTrim;            // materialInputTrim, materialTrim
    float tailShape, decay, lowBoost, splitLow, splitHigh;
    float envAtk, envRel, omEnvAtk, omEnvRel, wetAtk, wetRel, omWetAtk, omWetRel;
    float autoGainBase, mix, outGain;
    float highTilt;                   // 0.9 + texture*1.3
    float noiseAmt;                   // 0.004 + 0.022*texture
    float dynK;                       // 0.18 + texture*0.12
    float shapeGain;                  // per material: trailing (a + b*texture) factor
    // gel
    float gelOmega, gelK;
    // metal / wood / plastic modes
    float modeGain[4];
    float modeF[4];                   // metal: f0*ratio (pre-bend)
    float modeTwoR[4], modeA2[4];     // 2r and -r*r
    float modeA1[4];                  // wood/plastic: 2r*cos(theta) (block constant)
    float fMax;                       // 0.45*sr
    // wood / plastic waveguide
    float delaySamp, waveDamp, waveMixA, waveMixB, excA, excB, waveOut;
    int waveSize;
    // flesh
    float kA, kB, cA, cB, kCouple;
    float srf;
};

// JuicyMotion/PluginProcessor.cpp:59-73 plus loop-invariant sub-expressions of :105-141
struct MotionCoef {
    float envCoeff, omEnv, budgetCoeff, omBudget, tailFeedback, depth, motionInc, varSlew, omVarSlew;
    float microVar, repeatCtrl, mix, outGain;
    float lfoDepth;        // (250 + 550*microVar) * (0.5 + 0.9*depth)
    float d06;             // 0.6 + 0.6*depth
    float d07;             // 0.6 + 0.7*depth
    float d08;             // 0.6 + 0.8*depth
    float d0507;           // 0.55 + 0.7*depth
    float d0508;           // 0.5 + 0.8*depth
    float mv035;           // 0.35*microVar
    float mvT;             // 0.12 + 0.30*microVar
    float tailMix;         // (0.26 + 0.24*microVar) * (0.6 + 0.7*depth)
    float wetBoost;        // 1 + 0.9*microVar*(0.55 + 0.9*depth)
    float budgetTarget;    // jmap(budget, 0,1, 0.8, 0.25)
    float srf;
    int cooldownLen;       // static_cast<int>(sr * 0.04)
};

union SlotCoef {
    SatCoef sat;
    PunchCoef punch;
    WidthCoef width;
    CohereCoef cohere;
    InferCoef infer;
    TexCoef tex;
    MotionCoef motion;
};

struct SlotDesc {
    int kind;        // jb_plugin_kind
    int stateBase;   // first state variable of this slot: AV_COUNT analyzer vars, then the plugin's
    SlotCoef c;
};

struct ProcArgs {
    const float* in;       // [clip][ch][nSamples]
    float* out;            // may alias in
    float* state;          // SoA state, clipPitch floats per variable
    float* latest;         // [slot][JBK_REC][clipPitch]
    float* hist;           // optional [block][slot][JBK_REC][clipPitch], or null
    float* widthRing;      // wetR history of the Width slot: element (clip, t) at clip*ringClipStride + t*ringTimeStride
    long long ringClipStride, ringTimeStride; // time-major [ringLen][clipPitch] (1, clipPitch) or clip-major [clip][ringLen] (ringLen, 1)
    float* texWave;        // [2][waveSize][clipPitch]  (Texture waveguides)
    const int* clipMap;    // lane kernels: lane i of the launch renders clip clipMap[i] (clips sharing one parameter set
                           // that are not a contiguous range); null = clip i
    long long clipPitch;
    long long rowPitch;    // samples between consecutive (clip, channel) rows of in / out (>= nSamples; a call may
                           // render a time slice [t0, t0 + nSamples) of longer clips, in/out already offset by t0)
    int nClips, nCh, nSamples, blockSize;
    int histFirstBlock, histMaxBlocks;
    int chainLen;
    int recSlotBase, recChainLen; // a launch may render one plugin of a longer chain: its records go to slot recSlotBase + s
                                  // of a record block laid out for recChainLen plugins
    int exactMath;         // Saturator / Punch: glibc-exact tanh / pow (jb_libm.h) instead of the MUFU-based ones
    int vecOk;             // 16-byte vector path legal (alignment + sizes)
    int laneOnly;          // jb_set_path(JB_PATH_LANE): single-plugin launches take the lane kernels, not the clip-per-CTA one
    int octets;            // lane kernel: 1 = 8 samples per trip + 32-byte stores (big batches of light chains, where L2 sector
                           // throughput is the bound; costs registers, so not for Punch / Texture / Motion chains); 2 = warp-
                           // transposed tile streaming through cp.async; 3 = tile streaming through TMA (tmapIn below); 4 = 32-byte register loads
    int smallCode;         // two-lanes-per-clip Texture kernel: the quad's samples go through ONE copy of the sample code (launches of
                           // several parameter sets side by side: different kernels on one SM evict each other's unrolled loops
                           // from the instruction caches, profiles/r02_tma.txt)
    AnaCoef ana;
    SlotDesc slot[JBK_MAX_CHAIN];
    // octets == 3: CUtensorMap (128 bytes, opaque here) over the launch's input rows as a 3-D tensor {sample, channel, clip},
    // box {16 samples, 1 channel, 32 clips}, 64-byte swizzle -- the warp's 32 rows of one channel arrive with ONE
    // cp.async.bulk.tensor instruction (jb_lane.cuh: TMA tile streaming).  Filled by the launcher (jb_single_light.cu).
    alignas(64) unsigned long long tmapIn[16];
};

#ifdef __cplusplus
extern "C" {
#endif
// jb_kernels.cu
int jbk_launch_process(const ProcArgs* args, void* stream);
int jbk_solo_pick(const ProcArgs* args); // few clips of one plugin: the clip-per-CTA kernel (jb_solo.cu) would take this launch
int jbk_launch_fill(float* dst, float value, long long count, void* stream);
int jbk_launch_synth(float* dAudio, int kind, long long firstClip, int nClips, int nCh, int nSamples,
                     double sampleRate, unsigned int seed, void* stream);
const char* jbk_last_cuda_error(void);
long long jbk_launch_count(void);
void jbk_note_launch(void);
// jb_meter.cu
#define JBK_METER 40
int jbk_launch_meter(const float* hist, long long clipPitch, int chainLen, int slot, int firstBlock, int nBlocks,
                     int stride, int nClips, float* out, void* stream);
// jb_coop.cu
int jbk_coop_supported(const ProcArgs* args);
size_t jbk_coop_scratch_bytes(int chainLen, int numSMs);
int jbk_launch_coop(const ProcArgs* args, float* monoScratch, int numSMs, void* stream);
const char* jbk_coop_last_error(void);
#ifdef __cplusplus
}
#endif
