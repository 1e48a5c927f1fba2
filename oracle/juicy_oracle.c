/* TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") for the JuicySuite hot path.
 * See juicy_oracle.h.  Every routine cites the reference lines it restates
 * (paths relative to /root/reference).  Arithmetic is single precision with the
 * reference's operand order, no FMA contraction (-ffp-contract=off), FTZ+DAZ
 * inside the block loop like juce::ScopedNoDenormals.  JUCE helper semantics
 * (jmap, jlimit, Decibels, getRMSLevel, NormalisableRange) follow SURVEY.md
 * Appendix C; JUCE itself is an un-vendored dependency of the reference
 * (CMakeLists.txt:7-11, no version pin; artefacts indicate JUCE 8.x).
 */
#include "juicy_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <xmmintrin.h>

#define PI_F 3.14159265358979323846f
#define TWO_PI_F 6.28318530717958647692f

static inline float fmax2(float a, float b) { return a < b ? b : a; }          /* juce::jmax */
static inline float fmin2(float a, float b) { return b < a ? b : a; }          /* juce::jmin */
static inline float clampf(float lo, float hi, float v) { return v < lo ? lo : (hi < v ? hi : v); } /* jlimit */
static inline float map01(float v, float tmin, float tmax) { return tmin + v * (tmax - tmin); }   /* jmap 3-arg */
static inline float map5(float v, float smin, float smax, float tmin, float tmax)                /* jmap 5-arg */
{
    return tmin + ((tmax - tmin) * (v - smin)) / (smax - smin);
}
static inline float db_to_gain(float db) { return db > -100.0f ? powf(10.0f, db * 0.05f) : 0.0f; }
static inline float gain_to_db(float g) { return g > 0.0f ? fmax2(-100.0f, log10f(g) * 20.0f) : -100.0f; }

/* ------------------------------------------------------------------ parameters */

typedef struct { const char* id; float lo, hi, interval, def; } ParamSpec;

/* ids / ranges / defaults: createParameterLayout() of each plugin
 * (JuicyInfer/PluginProcessor.cpp:183-195, JuicyPunch:204-215, JuicySaturator:189-199,
 *  JuicyWidth:229-239, JuicyCohere:166-178, JuicyTexture:325-337, JuicyMotion:189-200) */
static const ParamSpec kInferParams[] = {
    {"trim", -18.0f, 18.0f, 0.0f, 0.0f}, {"sensitivity", 0.5f, 2.0f, 0.0f, 1.0f},
    {"juiciness", 0.0f, 100.0f, 0.0f, 0.0f}, {"emphasis", 0.0f, 1.0f, 0.0f, 0.0f},
    {"coherence", 0.0f, 1.0f, 0.0f, 0.0f}, {"synesthesia", 0.0f, 1.0f, 0.0f, 0.0f},
    {"fatigue", 0.0f, 1.0f, 0.0f, 0.0f}, {"repetition", 0.0f, 1.0f, 0.0f, 0.0f}};
static const ParamSpec kPunchParams[] = {
    {"punch", 0.0f, 1.5f, 0.0f, 0.9f}, {"sustain", 0.0f, 1.5f, 0.0f, 0.35f}, {"slam", 0.0f, 1.0f, 0.0f, 0.65f},
    {"clip", 0.0f, 1.0f, 0.0f, 0.25f}, {"mix", 0.0f, 1.0f, 0.0f, 1.0f}, {"output", -24.0f, 18.0f, 0.0f, -4.0f},
    {"juiciness", 0.0f, 100.0f, 0.0f, 0.0f}};
static const ParamSpec kSaturatorParams[] = {
    {"drive", 0.0f, 24.0f, 0.0f, 6.0f}, {"asymmetry", -0.5f, 0.5f, 0.0f, 0.1f}, {"tone", 0.0f, 1.0f, 0.0f, 0.55f},
    {"mix", 0.0f, 1.0f, 0.0f, 1.0f}, {"output", -18.0f, 18.0f, 0.0f, -3.0f}, {"juiciness", 0.0f, 100.0f, 0.0f, 0.0f}};
static const ParamSpec kWidthParams[] = {
    {"width", 0.0f, 1.0f, 0.0f, 0.45f}, {"haasMs", 0.0f, 35.0f, 0.0f, 12.0f}, {"monoSafe", 0.0f, 1.0f, 0.0f, 0.7f},
    {"mix", 0.0f, 1.0f, 0.0f, 1.0f}, {"output", -18.0f, 18.0f, 0.0f, 0.0f}, {"juiciness", 0.0f, 100.0f, 0.0f, 0.0f}};
static const ParamSpec kCohereParams[] = {
    {"match", 0.0f, 1.0f, 0.0f, 0.65f}, {"learn", 0.0f, 1.0f, 1.0f, 0.0f}, {"tail", 0.0f, 1.0f, 0.0f, 0.45f},
    {"decay", 0.1f, 0.95f, 0.0f, 0.65f}, {"mix", 0.0f, 1.0f, 0.0f, 1.0f}, {"output", -18.0f, 18.0f, 0.0f, 0.0f},
    {"contextfit", 0.0f, 100.0f, 0.0f, 0.0f}, {"juiciness", 0.0f, 100.0f, 0.0f, 0.0f}};
static const ParamSpec kTextureParams[] = {
    {"material", 0.0f, 4.0f, 1.0f, 0.0f}, {"tailshape", 0.0f, 1.0f, 0.0f, 0.55f}, {"damping", 0.0f, 1.0f, 0.0f, 0.5f},
    {"weight", 0.0f, 1.0f, 0.0f, 0.45f}, {"texture", 0.0f, 1.0f, 0.0f, 0.5f}, {"mix", 0.0f, 1.0f, 0.0f, 1.0f},
    {"output", -18.0f, 18.0f, 0.0f, -2.0f}, {"juiciness", 0.0f, 100.0f, 0.0f, 0.0f}};
static const ParamSpec kMotionParams[] = {
    {"microvar", 0.0f, 1.0f, 0.0f, 0.55f}, {"motiondepth", 0.0f, 2.0f, 0.0f, 1.0f}, {"repeatctrl", 0.0f, 1.0f, 0.0f, 0.65f},
    {"budget", 0.0f, 1.0f, 0.0f, 0.5f}, {"mix", 0.0f, 1.0f, 0.0f, 1.0f}, {"output", -18.0f, 18.0f, 0.0f, -2.0f},
    {"juiciness", 0.0f, 100.0f, 0.0f, 0.0f}};

#define COUNT(a) ((int) (sizeof(a) / sizeof((a)[0])))
static const ParamSpec* const kSpecs[JO_NUM_KINDS] = {kInferParams, kPunchParams, kSaturatorParams, kWidthParams,
                                                      kCohereParams, kTextureParams, kMotionParams};
static const int kSpecCounts[JO_NUM_KINDS] = {COUNT(kInferParams), COUNT(kPunchParams), COUNT(kSaturatorParams),
                                              COUNT(kWidthParams), COUNT(kCohereParams), COUNT(kTextureParams),
                                              COUNT(kMotionParams)};

/* factory presets: JuicyInfer:14-20, JuicyPunch:18-24, JuicySaturator:17-23, JuicyWidth:17-23 */
typedef struct { const char* name; int n; const char* ids[6]; float v[6]; } Preset;
static const Preset kInferPresets[5] = {
    {"Reference Lens", 2, {"trim", "sensitivity"}, {0.0f, 1.0f}},
    {"Detail Hunter", 2, {"trim", "sensitivity"}, {0.0f, 1.45f}},
    {"Macro Meter", 2, {"trim", "sensitivity"}, {-6.0f, 1.7f}},
    {"Subtle Scout", 2, {"trim", "sensitivity"}, {0.0f, 0.75f}},
    {"Overdrive Audit", 2, {"trim", "sensitivity"}, {-9.0f, 2.0f}}};
#define PUNCH_IDS {"punch", "sustain", "slam", "clip", "mix", "output"}
static const Preset kPunchPresets[5] = {
    {"Solar Snap", 6, PUNCH_IDS, {0.9f, 0.35f, 0.65f, 0.25f, 1.0f, -4.0f}},
    {"Crater Impact", 6, PUNCH_IDS, {1.4f, 0.2f, 0.95f, 0.65f, 1.0f, -8.0f}},
    {"Elastic Slam", 6, PUNCH_IDS, {1.1f, 0.8f, 0.8f, 0.4f, 0.85f, -6.0f}},
    {"Steel Bounce", 6, PUNCH_IDS, {0.7f, 0.55f, 0.45f, 0.1f, 0.75f, -2.0f}},
    {"Apocalypse Tap", 6, PUNCH_IDS, {1.5f, 1.1f, 1.0f, 1.0f, 1.0f, -12.0f}}};
#define SAT_IDS {"drive", "asymmetry", "tone", "mix", "output"}
static const Preset kSaturatorPresets[5] = {
    {"Amber Heat", 5, SAT_IDS, {6.0f, 0.1f, 0.55f, 1.0f, -3.0f}},
    {"Velvet Burn", 5, SAT_IDS, {11.0f, 0.2f, 0.4f, 0.85f, -6.0f}},
    {"Mirror Glow", 5, SAT_IDS, {8.0f, -0.15f, 0.75f, 0.7f, -4.0f}},
    {"Grain Reactor", 5, SAT_IDS, {18.0f, 0.35f, 0.32f, 1.0f, -10.0f}},
    {"Crystal Edge", 5, SAT_IDS, {4.0f, -0.05f, 0.9f, 0.55f, -1.0f}}};
#define WIDTH_IDS {"width", "haasMs", "monoSafe", "mix", "output"}
static const Preset kWidthPresets[5] = {
    {"Prism Arc", 5, WIDTH_IDS, {0.45f, 12.0f, 0.7f, 1.0f, 0.0f}},
    {"Outer Halo", 5, WIDTH_IDS, {0.9f, 22.0f, 0.35f, 1.0f, -1.5f}},
    {"Studio Spine", 5, WIDTH_IDS, {0.35f, 8.0f, 0.95f, 0.8f, 0.0f}},
    {"Ribbon Drift", 5, WIDTH_IDS, {0.7f, 16.0f, 0.55f, 0.65f, -0.5f}},
    {"Monolith Wide", 5, WIDTH_IDS, {1.0f, 30.0f, 0.2f, 1.0f, -3.0f}}};

static const Preset* presets_of(int kind)
{
    switch (kind) {
        case JO_INFER: return kInferPresets;
        case JO_PUNCH: return kPunchPresets;
        case JO_SATURATOR: return kSaturatorPresets;
        case JO_WIDTH: return kWidthPresets;
        default: return NULL;
    }
}

/* juce::NormalisableRange<float> (linear, optional interval) */
static float range_to01(const ParamSpec* s, float v) { return clampf(0.0f, 1.0f, (v - s->lo) / (s->hi - s->lo)); }
static float range_from01(const ParamSpec* s, float p) { return s->lo + (s->hi - s->lo) * clampf(0.0f, 1.0f, p); }
static float range_snap(const ParamSpec* s, float v)
{
    if (s->interval > 0.0f)
        v = s->lo + s->interval * floorf((v - s->lo) / s->interval + 0.5f);
    return clampf(s->lo, s->hi, v);
}
/* RangedAudioParameter::convertFrom0to1 */
static float param_denorm(const ParamSpec* s, float n) { return range_snap(s, range_from01(s, clampf(0.0f, 1.0f, n))); }

/* ------------------------------------------------------------------ analyzer */

/* JuicinessAnalyzer private state (src/shared/JuicinessAnalyzer.h:33-43) */
typedef struct {
    double sr;
    int channels;
    float shortEnv, longEnv, lowBandState, highBandState, lowCoeff, highCoeff, repetitionEma, fatigueEma;
    int onsetCooldown;
} Analyzer;

/* JuicinessMetrics (src/shared/JuicinessAnalyzer.h:6-21) */
typedef struct {
    float score, preScore, postScore, emphasis, coherence, synesthesia, fatigueRisk, repetitionDensity,
        punch, richness, clarity, width, monoSafety;
} Metrics;

static Metrics metrics_default(void)
{
    Metrics m;
    memset(&m, 0, sizeof m);
    m.monoSafety = 1.0f;
    return m;
}

/* JuicinessAnalyzer::reset (JuicinessAnalyzer.cpp:13-22) */
static void analyzer_reset(Analyzer* a)
{
    a->shortEnv = a->longEnv = a->lowBandState = a->highBandState = 0.0f;
    a->repetitionEma = a->fatigueEma = 0.0f;
    a->onsetCooldown = 0;
}

/* JuicinessAnalyzer::prepare (JuicinessAnalyzer.cpp:3-11) */
static void analyzer_prepare(Analyzer* a, double sampleRate, int numChannels)
{
    a->sr = sampleRate;
    a->channels = numChannels > 1 ? numChannels : 1;
    a->lowCoeff = 1.0f - expf(-2.0f * PI_F * 250.0f / (float) sampleRate);
    a->highCoeff = 1.0f - expf(-2.0f * PI_F * 2500.0f / (float) sampleRate);
    analyzer_reset(a);
}

/* JuicinessAnalyzer::updateEnvelope (JuicinessAnalyzer.cpp:24-29) */
static inline void env_follow(float in, float attack, float release, float* env)
{
    const float c = in > *env ? attack : release;
    *env = (1.0f - c) * in + c * *env;
}

/* AudioBuffer<float>::getRMSLevel: double accumulation */
static float rms_level(const float* d, int n)
{
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const double s = (double) d[i];
        sum += s * s;
    }
    return (float) sqrt(sum / n);
}

/* JuicinessAnalyzer::analyze (JuicinessAnalyzer.cpp:31-155).  left/right planar; right may be NULL. */
static Metrics analyzer_run(Analyzer* a, const float* left, const float* rightIn, int n)
{
    Metrics m = metrics_default();
    if (n <= 0)
        return m;

    const float attackShort = expf(-1.0f / (float) (a->sr * 0.003));   /* :38-41 */
    const float releaseShort = expf(-1.0f / (float) (a->sr * 0.030));
    const float attackLong = expf(-1.0f / (float) (a->sr * 0.050));
    const float releaseLong = expf(-1.0f / (float) (a->sr * 0.300));

    float transientAccum = 0.0f, rmsAccum = 0.0f, peak = 0.0f, lowAccum = 0.0f, highAccum = 0.0f;
    float sideAccum = 0.0f, midAccum = 0.0f, corrAccum = 0.0f;
    int onsetCount = 0;
    const float* right = a->channels > 1 ? rightIn : NULL;

    for (int i = 0; i < n; ++i) {                                      /* :57-92 */
        const float l = left[i];
        const float r = right != NULL ? right[i] : l;
        const float mono = 0.5f * (l + r);
        const float absMono = fabsf(mono);

        env_follow(absMono, attackShort, releaseShort, &a->shortEnv);
        env_follow(absMono, attackLong, releaseLong, &a->longEnv);

        const float transient = fmax2(0.0f, a->shortEnv - a->longEnv);
        transientAccum += transient;
        if (a->onsetCooldown > 0)
            --a->onsetCooldown;
        if (transient > 0.045f && a->onsetCooldown <= 0) {
            ++onsetCount;
            a->onsetCooldown = (int) (a->sr * 0.035);
        }
        rmsAccum += mono * mono;
        peak = fmax2(peak, fabsf(mono));

        a->lowBandState += a->lowCoeff * (mono - a->lowBandState);
        a->highBandState += a->highCoeff * (mono - a->highBandState);
        const float low = a->lowBandState;
        const float high = mono - a->highBandState;
        lowAccum += low * low;
        highAccum += high * high;

        const float mid = 0.5f * (l + r);
        const float side = 0.5f * (l - r);
        midAccum += mid * mid;
        sideAccum += side * side;
        corrAccum += l * r;
    }

    const float invN = 1.0f / (float) n;                               /* :94-141 */
    const float rms = sqrtf(rmsAccum * invN + 1.0e-12f);
    const float crest = peak / (rms + 1.0e-6f);
    const float lowEnergy = lowAccum * invN;
    const float highEnergy = highAccum * invN;
    const float lowHighRatio = lowEnergy / (highEnergy + 1.0e-8f);
    const float widthRatio = sideAccum / (midAccum + sideAccum + 1.0e-8f);

    const float lEnergy = rms_level(left, n);
    const float rEnergy = a->channels > 1 ? rms_level(rightIn, n) : lEnergy;
    float corr = corrAccum * invN / (lEnergy * rEnergy + 1.0e-6f);
    corr = clampf(-1.0f, 1.0f, corr);

    const float punch = clampf(0.0f, 1.0f, 6.0f * transientAccum * invN / (rms + 1.0e-5f));
    const float richness = clampf(0.0f, 1.0f, (2.3f - crest) * 0.65f + (rms * 2.0f));

    float clarity = 1.0f;
    if (lowHighRatio > 2.5f)
        clarity -= clampf(0.0f, 0.6f, (lowHighRatio - 2.5f) * 0.15f);
    if (highEnergy > 0.03f)
        clarity -= clampf(0.0f, 0.5f, (highEnergy - 0.03f) * 8.0f);
    clarity = clampf(0.0f, 1.0f, clarity);

    const float width = clampf(0.0f, 1.0f, widthRatio * 2.0f);
    const float monoSafety = clampf(0.0f, 1.0f, 0.5f * (corr + 1.0f));

    const float blockSeconds = (float) n / (float) a->sr;
    const float onsetRate = blockSeconds > 0.0f ? (float) onsetCount / blockSeconds : 0.0f;
    a->repetitionEma += (onsetRate - a->repetitionEma) * 0.08f;
    const float repetitionDensity = clampf(0.0f, 1.0f, a->repetitionEma / 12.0f);

    const float emphasis = clampf(0.0f, 1.0f, 0.62f * punch + 0.38f * clampf(0.0f, 1.0f, transientAccum * invN * 8.5f));
    const float coherence = clampf(0.0f, 1.0f, 0.50f * clarity + 0.30f * monoSafety + 0.20f * (1.0f - fabsf(width - 0.45f)));
    const float synesthesia = clampf(0.0f, 1.0f, 0.45f * richness + 0.30f * clampf(0.0f, 1.0f, lowHighRatio / 3.5f)
                                                     + 0.25f * clampf(0.0f, 1.0f, transientAccum * invN * 5.0f));

    const float crestPenalty = clampf(0.0f, 1.0f, (1.8f - crest) * 1.1f);
    const float harshPenalty = clampf(0.0f, 1.0f, highEnergy * 12.0f);
    const float instantFatigue = clampf(0.0f, 1.0f, 0.35f * crestPenalty + 0.35f * harshPenalty + 0.30f * repetitionDensity);
    a->fatigueEma += (instantFatigue - a->fatigueEma) * 0.06f;
    const float fatigueRisk = clampf(0.0f, 1.0f, a->fatigueEma);

    float score = 100.0f * (0.30f * punch + 0.25f * richness + 0.25f * clarity + 0.20f * width);
    score *= (0.6f + 0.4f * monoSafety);
    score = clampf(0.0f, 100.0f, score);

    m.score = score;                                                   /* :143-154 */
    m.emphasis = emphasis;
    m.coherence = coherence;
    m.synesthesia = synesthesia;
    m.fatigueRisk = fatigueRisk;
    m.repetitionDensity = repetitionDensity;
    m.punch = punch;
    m.richness = richness;
    m.clarity = clarity;
    m.width = width;
    m.monoSafety = monoSafety;
    return m;
}

/* ------------------------------------------------------------------ plugin instance */

#define TEX_MODES 4
/* JuicyTextureAudioProcessor::ChannelState (JuicyTexture/PluginProcessor.h:55-77) */
typedef struct {
    float tail, lp, hp, env, wetEnv, noiseHp, dcIn, dcOut, protectGain, springPos, springVel;
    float fleshPosA, fleshVelA, fleshPosB, fleshVelB, prevWave;
    float modalY1[TEX_MODES], modalY2[TEX_MODES];
    float* waveguide;
    int waveSize, waveIdx;
} TexChannel;

#define MAX_PARAMS 8
typedef struct {
    int kind, channels;
    double sr; /* AudioProcessor::getSampleRate() as set by the host */
    int blockSize;
    float stored[MAX_PARAMS]; /* RangedAudioParameter value */
    float raw[MAX_PARAMS];    /* APVTS adapter value = *getRawParameterValue(id) */
    int program;
    Analyzer analyzer;
    /* the 8 relaxed-atomic mailboxes every plugin keeps (e.g. JuicyPunch/PluginProcessor.h:44-51) */
    float latestPre, latestPost, latestScore, latestPunch, latestRichness, latestClarity, latestWidth, latestMono;
    /* Saturator (JuicySaturator/PluginProcessor.h:44) */
    float toneState[2];
    /* Punch (JuicyPunch/PluginProcessor.h:53-55) */
    float fastEnv[2], slowEnv[2];
    double punchSr;
    /* Width (JuicyWidth/PluginProcessor.h:54-55) */
    float* delayL;
    float* delayR;
    int delaySize, delayWritePosition;
    /* Cohere (JuicyCohere/PluginProcessor.h:55-63) */
    float targetLow, targetMid, targetHigh, cohTailL, cohTailR, lowLp, highLp, cohLowCoeff, cohHighCoeff;
    /* Texture (JuicyTexture/PluginProcessor.h:79-81) */
    TexChannel tex[2];
    double texSr;
    uint32_t texRng;
    /* Motion (JuicyMotion/PluginProcessor.h:54-72) */
    double motSr;
    float motEnv, repetition, budgetEnv, varTone, varTransient, varTail, varToneTarget, varTransientTarget,
        varTailTarget, motTailL, motTailR, lpL, lpR, prevL, prevR, motionPhase;
    int motCooldown;
    uint32_t motRng;
} Plugin;

static int find_param(const Plugin* p, const char* id)
{
    for (int i = 0; i < kSpecCounts[p->kind]; ++i)
        if (strcmp(kSpecs[p->kind][i].id, id) == 0)
            return i;
    return -1;
}

/* AudioProcessorParameter::setValueNotifyingHost(n) followed by the APVTS adapter update */
static void set_normalised(Plugin* p, int idx, float n)
{
    const ParamSpec* s = &kSpecs[p->kind][idx];
    if (s->interval > 0.0f && s->hi == 1.0f && s->lo == 0.0f)
        p->stored[idx] = n >= 0.5f ? 1.0f : 0.0f; /* AudioParameterBool::setValue */
    else
        p->stored[idx] = param_denorm(s, n);
    const float norm = (s->interval > 0.0f && s->hi == 1.0f && s->lo == 0.0f) ? p->stored[idx] : range_to01(s, p->stored[idx]);
    p->raw[idx] = param_denorm(s, norm);
}

static void set_plain(Plugin* p, const char* id, float v)
{
    const int idx = find_param(p, id);
    if (idx >= 0)
        set_normalised(p, idx, range_to01(&kSpecs[p->kind][idx], v));
}

static float rawv(const Plugin* p, const char* id)
{
    const int idx = find_param(p, id);
    return idx >= 0 ? p->raw[idx] : 0.0f;
}

static void apply_program(Plugin* p, int index)
{
    const Preset* ps = presets_of(p->kind);
    if (ps == NULL) {
        p->program = 0;
        return;
    }
    p->program = index < 0 ? 0 : (index > 4 ? 4 : index);
    const Preset* q = &ps[p->program];
    for (int i = 0; i < q->n; ++i)
        set_plain(p, q->ids[i], q->v[i]);
}

/* pushJuicinessToHost / setOut: range.convertTo0to1 then setValueNotifyingHost (e.g. JuicyPunch:56-62) */
static void push_output(Plugin* p, const char* id, float v) { set_plain(p, id, v); }

static void store_mailboxes(Plugin* p, const Metrics* pre, const Metrics* post)
{
    p->latestPre = pre->score;            /* e.g. JuicyPunch/PluginProcessor.cpp:115-122 */
    p->latestPost = post->score;
    p->latestScore = post->score;
    p->latestPunch = post->punch;
    p->latestRichness = post->richness;
    p->latestClarity = post->clarity;
    p->latestWidth = post->width;
    p->latestMono = post->monoSafety;
    push_output(p, "juiciness", post->score);
}

/* ------------------------------------------------------------------ prepareToPlay */

static void plugin_prepare(Plugin* p, double sampleRate, int blockSize)
{
    p->sr = sampleRate;
    p->blockSize = blockSize;
    analyzer_prepare(&p->analyzer, sampleRate, p->channels);
    switch (p->kind) {
        case JO_SATURATOR: /* JuicySaturator/PluginProcessor.cpp:35-39 */
            p->toneState[0] = p->toneState[1] = 0.0f;
            break;
        case JO_PUNCH: /* JuicyPunch/PluginProcessor.cpp:36-42 */
            p->punchSr = sampleRate;
            p->fastEnv[0] = p->fastEnv[1] = p->slowEnv[0] = p->slowEnv[1] = 0.0f;
            break;
        case JO_WIDTH: { /* JuicyWidth/PluginProcessor.cpp:35-42 */
            const int delaySamples = (int) (sampleRate * 0.060);
            p->delaySize = delaySamples > 1 ? delaySamples : 1;
            free(p->delayL);
            free(p->delayR);
            p->delayL = (float*) calloc((size_t) p->delaySize, sizeof(float));
            p->delayR = (float*) calloc((size_t) p->delaySize, sizeof(float));
            p->delayWritePosition = 0;
            break;
        }
        case JO_COHERE: /* JuicyCohere/PluginProcessor.cpp:13-22 (targets are NOT reset) */
            p->cohLowCoeff = 1.0f - expf(-2.0f * PI_F * 220.0f / (float) sampleRate);
            p->cohHighCoeff = 1.0f - expf(-2.0f * PI_F * 2400.0f / (float) sampleRate);
            p->cohTailL = p->cohTailR = p->lowLp = p->highLp = 0.0f;
            break;
        case JO_TEXTURE: { /* JuicyTexture/PluginProcessor.cpp:12-26 */
            p->texSr = sampleRate;
            p->texRng = 0x12345678u;
            int maxDelay = (int) (p->texSr * 0.08);
            if (maxDelay < 2048)
                maxDelay = 2048;
            for (int c = 0; c < 2; ++c) {
                free(p->tex[c].waveguide);
                memset(&p->tex[c], 0, sizeof(TexChannel));
                p->tex[c].protectGain = 1.0f;
                p->tex[c].waveguide = (float*) calloc((size_t) maxDelay, sizeof(float));
                p->tex[c].waveSize = maxDelay;
            }
            break;
        }
        case JO_MOTION: /* JuicyMotion/PluginProcessor.cpp:12-29 (rng is NOT reseeded) */
            p->motSr = sampleRate;
            p->motEnv = p->repetition = p->budgetEnv = 0.0f;
            p->motCooldown = 0;
            p->motTailL = p->motTailR = p->lpL = p->lpR = p->prevL = p->prevR = 0.0f;
            p->varTone = p->varTransient = p->varTail = 0.0f;
            p->varToneTarget = p->varTransientTarget = p->varTailTarget = 0.0f;
            p->motionPhase = 0.0f;
            break;
        default:
            break;
    }
}

/* ------------------------------------------------------------------ processBlock bodies */

/* JuicySaturatorAudioProcessor::processBlock (JuicySaturator/PluginProcessor.cpp:61-110) */
static void saturator_block(Plugin* p, float* const* ch, int n)
{
    const float driveDb = rawv(p, "drive"), asym = rawv(p, "asymmetry"), tone = rawv(p, "tone");
    const float mix = rawv(p, "mix"), outputDb = rawv(p, "output");
    const Metrics pre = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    const float inGain = db_to_gain(driveDb);
    const float outGain = db_to_gain(outputDb);
    const float cutoff = map5(tone, 0.0f, 1.0f, 2500.0f, 16000.0f);
    const float toneCoeff = 1.0f - expf(-2.0f * PI_F * cutoff / (float) p->sr);
    for (int c = 0; c < p->channels; ++c) {
        float* x = ch[c];
        float state = p->toneState[c];
        for (int i = 0; i < n; ++i) {
            const float dry = x[i];
            const float driven = dry * inGain;
            const float skewed = driven + asym * driven * driven;
            const float soft = tanhf(skewed);
            state += toneCoeff * (soft - state);
            const float wet = state * outGain;
            x[i] = dry + mix * (wet - dry);
        }
        p->toneState[c] = state;
    }
    const Metrics post = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    store_mailboxes(p, &pre, &post);
}

/* JuicyPunchAudioProcessor::processBlock (JuicyPunch/PluginProcessor.cpp:64-124) */
static void punch_block(Plugin* p, float* const* ch, int n)
{
    const float punchAmt = rawv(p, "punch"), sustainAmt = rawv(p, "sustain"), slamAmt = rawv(p, "slam");
    const float clipAmt = rawv(p, "clip"), mix = rawv(p, "mix");
    const float outGain = db_to_gain(rawv(p, "output"));
    const Metrics pre = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    const float fastCoeff = expf(-1.0f / (float) (p->punchSr * 0.0015));
    const float slowCoeff = expf(-1.0f / (float) (p->punchSr * 0.110));
    for (int c = 0; c < p->channels; ++c) {
        float* x = ch[c];
        float fEnv = p->fastEnv[c], sEnv = p->slowEnv[c];
        for (int i = 0; i < n; ++i) {
            const float dry = x[i];
            const float adry = fabsf(dry);
            fEnv = (1.0f - fastCoeff) * adry + fastCoeff * fEnv;
            sEnv = (1.0f - slowCoeff) * adry + slowCoeff * sEnv;
            const float transient = fmax2(0.0f, fEnv - sEnv);
            const float transientCurve = powf(transient, map5(slamAmt, 0.0f, 1.0f, 0.95f, 0.55f));
            const float punchGain = 1.0f + (punchAmt * 12.0f + slamAmt * 22.0f) * transientCurve;
            const float sustainGain = 1.0f + (sustainAmt * 4.0f + slamAmt * 1.5f) * fmax2(0.0f, sEnv - transient * 0.6f);
            float wet = dry * punchGain * sustainGain;
            const float drive = 1.0f + clipAmt * 8.0f + slamAmt * 4.0f;
            const float soft = tanhf(wet * drive) / tanhf(drive);
            const float hard = clampf(-0.95f, 0.95f, wet * (1.0f + clipAmt * 2.0f));
            wet = soft + clipAmt * (hard - soft);
            x[i] = (dry + mix * (wet - dry)) * outGain;
        }
        p->fastEnv[c] = fEnv;
        p->slowEnv[c] = sEnv;
    }
    const Metrics post = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    store_mailboxes(p, &pre, &post);
}

/* JuicyWidthAudioProcessor::processBlock (JuicyWidth/PluginProcessor.cpp:64-150) */
static void width_block(Plugin* p, float* const* ch, int n)
{
    const Metrics pre = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    if (p->channels < 2) { /* :76-89 mono early-out: analyze twice, no DSP */
        const Metrics post = analyzer_run(&p->analyzer, ch[0], ch[1], n);
        store_mailboxes(p, &pre, &post);
        return;
    }
    const int delayBufferSize = p->delaySize;
    const int delaySamples = (int) (p->sr * (rawv(p, "haasMs") * 0.001f));
    float width = rawv(p, "width");
    const float monoSafe = rawv(p, "monoSafe"), mix = rawv(p, "mix");
    const float outputGain = db_to_gain(rawv(p, "output"));
    float* left = ch[0];
    float* right = ch[1];
    for (int i = 0; i < n; ++i) {
        const float dryL = left[i], dryR = right[i];
        const float corrProxy = clampf(-1.0f, 1.0f, dryL * dryR * 12.0f);
        const float dynamicLimit = map5(monoSafe, 0.0f, 1.0f, 1.0f, 0.35f);
        if (corrProxy < -0.1f)
            width *= dynamicLimit;
        const float mid = 0.5f * (dryL + dryR);
        const float side = 0.5f * (dryL - dryR) * (1.0f + width);
        float wetL = mid + side;
        float wetR = mid - side;
        p->delayL[p->delayWritePosition] = wetL;
        p->delayR[p->delayWritePosition] = wetR;
        int readPos = p->delayWritePosition - delaySamples;
        if (readPos < 0)
            readPos += delayBufferSize;
        wetR = p->delayR[readPos];
        left[i] = (dryL + mix * (wetL - dryL)) * outputGain;
        right[i] = (dryR + mix * (wetR - dryR)) * outputGain;
        if (++p->delayWritePosition >= delayBufferSize)
            p->delayWritePosition = 0;
    }
    const Metrics post = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    store_mailboxes(p, &pre, &post);
}

/* JuicyCohereAudioProcessor::processBlock (JuicyCohere/PluginProcessor.cpp:42-131) */
static void cohere_block(Plugin* p, float* const* ch, int n)
{
    const Metrics pre = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    const float matchAmt = rawv(p, "match");
    const int learn = rawv(p, "learn") > 0.5f;
    const float tailAmt = rawv(p, "tail"), decay = rawv(p, "decay"), mix = rawv(p, "mix");
    const float outGain = db_to_gain(rawv(p, "output"));
    const float* second = ch[p->channels - 1 < 1 ? p->channels - 1 : 1];

    float lowEnergy = 0.0f, midEnergy = 0.0f, highEnergy = 0.0f;
    for (int i = 0; i < n; ++i) { /* :62-74 */
        const float mono = 0.5f * (ch[0][i] + second[i]);
        p->lowLp += p->cohLowCoeff * (mono - p->lowLp);
        p->highLp += p->cohHighCoeff * (mono - p->highLp);
        const float low = p->lowLp;
        const float high = mono - p->highLp;
        const float mid = mono - low - high;
        lowEnergy += low * low;
        midEnergy += mid * mid;
        highEnergy += high * high;
    }
    const float inv = 1.0f / (float) (n > 1 ? n : 1);
    lowEnergy *= inv;
    midEnergy *= inv;
    highEnergy *= inv;
    if (learn) { /* :78-84 */
        const float a = 0.02f;
        p->targetLow += (lowEnergy - p->targetLow) * a;
        p->targetMid += (midEnergy - p->targetMid) * a;
        p->targetHigh += (highEnergy - p->targetHigh) * a;
    }
    const float lowErr = fabsf(gain_to_db((lowEnergy + 1.0e-6f) / (p->targetLow + 1.0e-6f)));
    const float midErr = fabsf(gain_to_db((midEnergy + 1.0e-6f) / (p->targetMid + 1.0e-6f)));
    const float highErr = fabsf(gain_to_db((highEnergy + 1.0e-6f) / (p->targetHigh + 1.0e-6f)));
    const float deviation = (lowErr + midErr + highErr) / 3.0f;
    const float contextFit = clampf(0.0f, 100.0f, 100.0f - deviation * 10.0f);
    push_output(p, "contextfit", contextFit);

    const float lowComp = clampf(0.5f, 1.8f, powf((p->targetLow + 1.0e-6f) / (lowEnergy + 1.0e-6f), 0.25f * matchAmt));
    const float midComp = clampf(0.5f, 1.8f, powf((p->targetMid + 1.0e-6f) / (midEnergy + 1.0e-6f), 0.25f * matchAmt));
    const float highComp = clampf(0.5f, 1.8f, powf((p->targetHigh + 1.0e-6f) / (highEnergy + 1.0e-6f), 0.25f * matchAmt));
    const float fb = clampf(0.0f, 0.93f, decay);

    for (int c = 0; c < p->channels; ++c) { /* :99-119 */
        float* x = ch[c];
        float tail = c == 0 ? p->cohTailL : p->cohTailR;
        float lpA = 0.0f, lpB = 0.0f;
        for (int i = 0; i < n; ++i) {
            const float dry = x[i];
            lpA += p->cohLowCoeff * (dry - lpA);
            lpB += p->cohHighCoeff * (dry - lpB);
            const float low = lpA * lowComp;
            const float high = (dry - lpB) * highComp;
            const float mid = (dry - lpA - (dry - lpB)) * midComp;
            const float matched = low + mid + high;
            tail = matched + tail * fb;
            const float wet = matched + tailAmt * 0.35f * tail;
            x[i] = (dry + mix * (wet - dry)) * outGain;
        }
        if (c == 0) p->cohTailL = tail; else p->cohTailR = tail;
    }
    const Metrics post = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    store_mailboxes(p, &pre, &post);
}

/* modeStep lambda (JuicyTexture/PluginProcessor.cpp:77-89) */
static inline float tex_mode_step(TexChannel* st, float srf, int k, float excitation, float freqHz, float t60, float gain)
{
    const float f = clampf(20.0f, 0.45f * srf, freqHz);
    const float t = fmax2(0.02f, t60);
    const float r = expf(logf(0.001f) / (t * srf));
    const float theta = 2.0f * PI_F * f / srf;
    const float a1 = 2.0f * r * cosf(theta);
    const float a2 = -r * r;
    const float y = excitation * gain + a1 * st->modalY1[k] + a2 * st->modalY2[k];
    st->modalY2[k] = st->modalY1[k];
    st->modalY1[k] = y;
    return y;
}

/* waveguideRead lambda (JuicyTexture/PluginProcessor.cpp:91-105) */
static inline float tex_wave_read(const TexChannel* st, float delaySamples)
{
    const int size = st->waveSize;
    if (size <= 1)
        return 0.0f;
    float pos = (float) st->waveIdx - delaySamples;
    while (pos < 0.0f)
        pos += (float) size;
    while (pos >= (float) size)
        pos -= (float) size;
    const int i0 = (int) pos;
    const int i1 = (i0 + 1) % size;
    const float frac = pos - (float) i0;
    return map01(frac, st->waveguide[i0], st->waveguide[i1]);
}

/* JuicyTextureAudioProcessor::processBlock (JuicyTexture/PluginProcessor.cpp:43-290) */
static void texture_block(Plugin* p, float* const* ch, int n)
{
    const Metrics pre = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    const int mode = (int) rawv(p, "material");
    const float tailShape = rawv(p, "tailshape"), damping = rawv(p, "damping"), weight = rawv(p, "weight");
    const float texture = rawv(p, "texture"), mix = rawv(p, "mix");
    const float outGain = db_to_gain(rawv(p, "output"));
    const float srf = (float) p->texSr;

    const float dampingAmt = clampf(0.0f, 1.0f, damping);                         /* :64-75 */
    const float dampingMul = map5(dampingAmt, 0.0f, 1.0f, 1.35f, 0.40f);
    const float decay = map5(tailShape, 0.0f, 1.0f, 0.30f, 0.985f) * map5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.80f);
    const float lowBoost = 1.0f + weight * 1.0f;
    const float splitLowCoeff = 1.0f - expf(-2.0f * PI_F * 140.0f / srf);
    const float splitHighCoeff = 1.0f - expf(-2.0f * PI_F * 2600.0f / srf);
    const float envAtk = expf(-1.0f / (float) (p->texSr * 0.0025));
    const float envRel = expf(-1.0f / (float) (p->texSr * 0.080));
    const float wetEnvAttack = expf(-1.0f / (float) (p->texSr * 0.005));
    const float wetEnvRelease = expf(-1.0f / (float) (p->texSr * 0.090));
    const float dcR = 0.995f;
    const float autoGainBase = map5(texture, 0.0f, 1.0f, 0.78f, 0.54f);

    for (int c = 0; c < p->channels; ++c) {
        float* x = ch[c];
        TexChannel* st = &p->tex[c < 0 ? 0 : (c > 1 ? 1 : c)];
        for (int i = 0; i < n; ++i) {
            const float dry = x[i];
            const float materialInputTrim = (mode == 1 ? 0.58f : (mode == 2 ? 0.62f : (mode == 3 ? 0.60f : 1.0f)));
            const float driven = dry * materialInputTrim;
            const float adry = fabsf(dry);
            const float envCoeff = adry > st->env ? envAtk : envRel;
            st->env = envCoeff * st->env + (1.0f - envCoeff) * adry;
            const float impact = clampf(0.0f, 1.0f, fmax2(0.0f, adry - st->env) * 10.0f);
            const float body = clampf(0.0f, 1.0f, st->env * 3.2f);
            const float trail = clampf(0.0f, 1.0f, 1.0f - impact) * tailShape;

            st->lp += splitLowCoeff * (driven - st->lp);
            st->hp += splitHighCoeff * (driven - st->hp);
            const float low = st->lp * lowBoost;
            const float high = (driven - st->hp);
            const float mid = driven - st->lp - high;
            const float core = low + mid + high * (0.9f + texture * 1.3f);

            float shaped = core;
            float materialTrim = 1.0f;
            switch (mode) {
                case 0: { /* gel :137-151 */
                    const float f0 = 42.0f + texture * 88.0f;
                    const float omega = 2.0f * PI_F * f0 / srf;
                    const float k = omega * omega;
                    const float zeta = map01(trail, 0.62f, 1.45f);
                    const float cc = 2.0f * zeta * omega;
                    const float force = core * (0.52f + 0.62f * body);
                    const float acc = k * (force - st->springPos) - cc * st->springVel;
                    st->springVel += acc;
                    st->springPos += st->springVel;
                    shaped = 0.48f * core + 1.85f * st->springPos;
                    shaped = tanhf(shaped * (0.96f + 0.28f * texture));
                    break;
                }
                case 1: { /* metal :152-169 */
                    const float exc = core * (0.19f + 0.52f * impact);
                    const float f0 = 320.0f + 140.0f * texture;
                    const float bend = 1.0f + 0.09f * impact;
                    const float metalDamp = map5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.55f);
                    const float tScale = map01(tailShape, 0.18f, 0.72f) * dampingMul * metalDamp;
                    const float m0 = tex_mode_step(st, srf, 0, exc, f0 * 1.00f * bend, 0.56f * tScale, 0.34f);
                    const float m1 = tex_mode_step(st, srf, 1, exc, f0 * 2.31f * bend, 0.40f * tScale, 0.20f);
                    const float m2 = tex_mode_step(st, srf, 2, exc, f0 * 4.18f * bend, 0.26f * tScale, 0.13f);
                    const float m3 = tex_mode_step(st, srf, 3, exc, f0 * 6.87f * bend, 0.17f * tScale, 0.09f);
                    const float modes = m0 + m1 + m2 + m3;
                    const float brightExcite = 0.03f * impact * (core - st->hp);
                    shaped = (0.44f * core + 0.42f * modes + brightExcite) * (0.78f + 0.10f * texture);
                    materialTrim = 0.62f;
                    break;
                }
                case 2: { /* wood :170-192 */
                    const float exc = core * (0.10f + 0.34f * impact);
                    const float cavityHz = 92.0f + 95.0f * (0.5f * weight + 0.5f * texture);
                    const float delaySamp = clampf(16.0f, (float) (st->waveSize - 2), srf / cavityHz);
                    const float delayed = tex_wave_read(st, delaySamp);
                    const float damp = map01(tailShape, 0.26f, 0.90f) * map5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.72f);
                    const float newWave = damp * (0.62f * delayed + 0.38f * st->prevWave) + exc * (0.09f + 0.04f * body);
                    st->waveguide[st->waveIdx] = newWave;
                    st->waveIdx = (st->waveIdx + 1) % st->waveSize;
                    st->prevWave = delayed;
                    const float woodDamp = map5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.64f);
                    const float tScale = map01(tailShape, 0.18f, 0.62f) * dampingMul * woodDamp;
                    const float w0 = tex_mode_step(st, srf, 0, exc, 155.0f, 0.40f * tScale, 0.32f);
                    const float w1 = tex_mode_step(st, srf, 1, exc, 355.0f, 0.27f * tScale, 0.18f);
                    const float w2 = tex_mode_step(st, srf, 2, exc, 690.0f, 0.16f * tScale, 0.10f);
                    const float w3 = tex_mode_step(st, srf, 3, exc, 1130.0f, 0.10f * tScale, 0.06f);
                    shaped = (0.56f * core + 0.24f * delayed + 0.30f * (w0 + w1 + w2 + w3)) * (0.74f + 0.08f * texture);
                    materialTrim = 0.54f;
                    break;
                }
                case 3: { /* plastic :193-213 */
                    const float exc = core * (0.20f + 0.60f * impact);
                    const float tubeHz = 210.0f + 340.0f * texture;
                    const float delaySamp = clampf(8.0f, (float) (st->waveSize - 2), srf / tubeHz);
                    const float delayed = tex_wave_read(st, delaySamp);
                    const float damp = map01(tailShape, 0.22f, 0.91f) * map5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.82f);
                    const float newWave = damp * (0.76f * delayed + 0.24f * st->prevWave) + 0.14f * exc;
                    st->waveguide[st->waveIdx] = newWave;
                    st->waveIdx = (st->waveIdx + 1) % st->waveSize;
                    st->prevWave = delayed;
                    const float tScale = map01(tailShape, 0.16f, 0.72f) * dampingMul;
                    const float p0 = tex_mode_step(st, srf, 0, exc, 280.0f, 0.28f * tScale, 0.34f);
                    const float p1 = tex_mode_step(st, srf, 1, exc, 690.0f, 0.18f * tScale, 0.22f);
                    const float p2 = tex_mode_step(st, srf, 2, exc, 1320.0f, 0.11f * tScale, 0.16f);
                    const float p3 = tex_mode_step(st, srf, 3, exc, 2360.0f, 0.07f * tScale, 0.11f);
                    shaped = (0.52f * core + 0.36f * delayed + 0.40f * (p0 + p1 + p2 + p3)) * (0.80f + 0.10f * texture);
                    materialTrim = 0.62f;
                    break;
                }
                default: { /* flesh :214-236 */
                    const float force = core * (0.55f + 0.65f * body);
                    const float wA = 2.0f * PI_F * (38.0f + 52.0f * texture) / srf;
                    const float wB = 2.0f * PI_F * (88.0f + 72.0f * texture) / srf;
                    const float kA = wA * wA;
                    const float kB = wB * wB;
                    const float cA = 2.0f * map01(tailShape, 0.56f, 1.18f) * wA;
                    const float cB = 2.0f * map01(tailShape, 0.70f, 1.34f) * wB;
                    const float kCouple = 0.14f + 0.24f * texture;
                    const float accA = kA * (force - st->fleshPosA) - cA * st->fleshVelA - kCouple * (st->fleshPosA - st->fleshPosB);
                    const float accB = kB * (st->fleshPosA - st->fleshPosB) - cB * st->fleshVelB;
                    st->fleshVelA += accA;
                    st->fleshVelB += accB;
                    st->fleshPosA += st->fleshVelA;
                    st->fleshPosB += st->fleshVelB;
                    const float tissue = 0.92f * st->fleshPosA + 0.58f * st->fleshPosB;
                    const float nl = tissue - 0.19f * tissue * tissue * tissue;
                    shaped = tanhf((0.50f * core + 1.34f * nl) * (0.98f + 0.16f * texture));
                    break;
                }
            }

            p->texRng = 1664525u * p->texRng + 1013904223u;                  /* :239-243 */
            const float white = ((float) ((p->texRng >> 8) & 0xFFFF) / 32768.0f - 1.0f);
            st->noiseHp += 0.08f * (white - st->noiseHp);
            const float rough = white - st->noiseHp;
            shaped += rough * (0.004f + 0.022f * texture) * (0.14f + 0.64f * impact);

            const float dynamics = 1.0f + impact * (0.18f + texture * 0.12f) + body * 0.06f;
            shaped *= dynamics * materialTrim;

            const float tailInput = clampf(-2.0f, 2.0f, shaped) * (0.45f + 0.55f * trail);
            st->tail = tailInput + st->tail * decay;
            float wet = shaped + st->tail * (0.30f + 0.45f * trail);

            const float wetAbs = fabsf(wet);                                 /* :253-257 */
            const float wetCoeff = wetAbs > st->wetEnv ? wetEnvAttack : wetEnvRelease;
            st->wetEnv = wetCoeff * st->wetEnv + (1.0f - wetCoeff) * wetAbs;
            const float autoComp = autoGainBase / (1.0f + 1.8f * st->wetEnv);
            wet *= clampf(0.18f, 1.0f, autoComp);

            const float mixed = dry + mix * (wet - dry);
            float out = mixed * outGain;

            const float dcBlocked = out - st->dcIn + dcR * st->dcOut;        /* :263-265 */
            st->dcIn = out;
            st->dcOut = dcBlocked;

            const float peak = fabsf(dcBlocked);                             /* :268-276 */
            const float ceiling = 0.88f;
            if (peak > ceiling)
                st->protectGain = fmin2(st->protectGain, (ceiling / peak) * 0.98f);
            else
                st->protectGain += (1.0f - st->protectGain) * 0.0028f;
            out = dcBlocked * clampf(0.2f, 1.0f, st->protectGain);
            x[i] = clampf(-0.98f, 0.98f, out);
        }
    }
    const Metrics post = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    store_mailboxes(p, &pre, &post);
}

/* JuicyMotionAudioProcessor::processBlock (JuicyMotion/PluginProcessor.cpp:47-154) */
static void motion_block(Plugin* p, float* const* ch, int n)
{
    const Metrics pre = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    const float microVar = rawv(p, "microvar"), motionDepth = rawv(p, "motiondepth"), repeatCtrl = rawv(p, "repeatctrl");
    const float contrastBudget = rawv(p, "budget"), mix = rawv(p, "mix");
    const float outGain = db_to_gain(rawv(p, "output"));
    const float srf = (float) p->motSr;

    const float envCoeff = expf(-1.0f / (float) (p->motSr * 0.015));                /* :67-73 */
    const float budgetCoeff = expf(-1.0f / (float) (p->motSr * 0.080));
    const float tailFeedback = map5(repeatCtrl, 0.0f, 1.0f, 0.15f, 0.88f);
    const float depth = clampf(0.0f, 2.0f, motionDepth);
    const float motionRateHz = map5(microVar, 0.0f, 1.0f, 0.25f, 2.0f) * map5(depth, 0.0f, 2.0f, 0.75f, 1.6f);
    const float motionInc = (2.0f * PI_F * motionRateHz) / srf;
    const float varSlew = expf(-1.0f / (float) (p->motSr * 0.020));
    const float* second = ch[p->channels - 1 < 1 ? p->channels - 1 : 1];

    for (int i = 0; i < n; ++i) { /* detector pass :75-95 */
        const float mono = 0.5f * (ch[0][i] + second[i]);
        const float absMono = fabsf(mono);
        p->motEnv = envCoeff * p->motEnv + (1.0f - envCoeff) * absMono;
        if (p->motCooldown > 0)
            --p->motCooldown;
        if (absMono > p->motEnv * 1.35f + 0.02f && p->motCooldown <= 0) {
            p->motCooldown = (int) (p->motSr * 0.04);
            p->repetition += 1.0f;
            p->motRng = 1664525u * p->motRng + 1013904223u;
            p->varToneTarget = (((float) ((p->motRng >> 7) & 0x7FFF) / 16384.0f) - 1.0f) * microVar * 0.9f;
            p->motRng = 1664525u * p->motRng + 1013904223u;
            p->varTransientTarget = (((float) ((p->motRng >> 9) & 0x7FFF) / 16384.0f) - 1.0f) * microVar * 0.8f;
            p->motRng = 1664525u * p->motRng + 1013904223u;
            p->varTailTarget = (((float) ((p->motRng >> 11) & 0x7FFF) / 16384.0f) - 1.0f) * microVar * 0.8f;
        }
        p->repetition *= 0.997f;
    }

    const float repNorm = clampf(0.0f, 1.0f, p->repetition * 0.08f);              /* :97-99 */
    const float repetitionScale = 1.0f - repeatCtrl * repNorm * 0.65f;
    const float recovery = 1.0f + repeatCtrl * (1.0f - repNorm) * 0.25f;

    for (int c = 0; c < p->channels; ++c) { /* :101-142 */
        float* x = ch[c];
        float* tail = c == 0 ? &p->motTailL : &p->motTailR;
        float* lp = c == 0 ? &p->lpL : &p->lpR;
        float* prev = c == 0 ? &p->prevL : &p->prevR;
        for (int i = 0; i < n; ++i) {
            p->varTone = varSlew * p->varTone + (1.0f - varSlew) * p->varToneTarget;
            p->varTransient = varSlew * p->varTransient + (1.0f - varSlew) * p->varTransientTarget;
            p->varTail = varSlew * p->varTail + (1.0f - varSlew) * p->varTailTarget;
            p->motionPhase += motionInc;
            if (p->motionPhase > 2.0f * PI_F)
                p->motionPhase -= 2.0f * TWO_PI_F; /* sic: the reference subtracts 4*pi (:114-115) */

            const float dry = x[i];
            const float motionLfo = sinf(p->motionPhase + (c == 0 ? 0.0f : 0.85f));
            const float motionLfoDepth = (250.0f + 550.0f * microVar) * (0.5f + 0.9f * depth);
            const float cutoff = clampf(120.0f, 4200.0f, 900.0f + p->varTone * 1100.0f * (0.6f + 0.6f * depth) + motionLfo * motionLfoDepth);
            const float lpCoeff = 1.0f - expf(-2.0f * PI_F * cutoff / srf);
            *lp += lpCoeff * (dry - *lp);
            const float hp = dry - *lp;
            const float transient = dry - *prev;
            *prev = dry;

            const float transientBoost = 1.0f + p->varTransient * 1.2f * (0.6f + 0.7f * depth) + 0.35f * microVar * motionLfo * (0.6f + 0.8f * depth);
            const float toneShift = *lp * (1.0f + p->varTone * 0.65f * (0.55f + 0.7f * depth))
                + hp * transientBoost
                + transient * (0.12f + 0.30f * microVar) * (0.5f + 0.8f * depth);
            *tail = toneShift + *tail * clampf(0.0f, 0.93f, tailFeedback + p->varTail * 0.06f);

            float wet = toneShift * repetitionScale * recovery + (0.26f + 0.24f * microVar) * (0.6f + 0.7f * depth) * *tail;
            p->budgetEnv = budgetCoeff * p->budgetEnv + (1.0f - budgetCoeff) * fabsf(wet);
            const float budgetTarget = map5(contrastBudget, 0.0f, 1.0f, 0.8f, 0.25f);
            const float limiterGain = p->budgetEnv > budgetTarget ? budgetTarget / (p->budgetEnv + 1.0e-5f) : 1.0f;
            wet *= limiterGain;

            const float wetBoost = 1.0f + 0.9f * microVar * (0.55f + 0.9f * depth);
            x[i] = (dry + mix * (wet * wetBoost - dry)) * outGain;
        }
    }
    const Metrics post = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    store_mailboxes(p, &pre, &post);
}

/* JuicyInferAudioProcessor::processBlock (JuicyInfer/PluginProcessor.cpp:64-102) */
static int approx_equal(float a, float b)
{
    const float diff = fabsf(a - b);
    const float mx = fabsf(a) > fabsf(b) ? fabsf(a) : fabsf(b);
    return diff <= 1.17549435e-38f || diff <= 1.1920929e-7f * mx;
}

static void infer_block(Plugin* p, float* const* ch, int n)
{
    const float trimGain = db_to_gain(rawv(p, "trim"));
    const float sensitivity = rawv(p, "sensitivity");
    const Metrics pre = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    if (!approx_equal(trimGain, 1.0f)) { /* AudioBuffer::applyGain */
        for (int c = 0; c < p->channels; ++c)
            for (int i = 0; i < n; ++i)
                ch[c][i] = trimGain == 0.0f ? 0.0f : ch[c][i] * trimGain;
    }
    Metrics m = analyzer_run(&p->analyzer, ch[0], ch[1], n);
    m.score = clampf(0.0f, 100.0f, m.score * sensitivity);
    p->latestPre = pre.score;                /* :82-89: triangle metrics ride in the five bar mailboxes */
    p->latestPost = m.score;
    p->latestScore = m.score;
    p->latestPunch = m.emphasis;
    p->latestRichness = m.coherence;
    p->latestClarity = m.synesthesia;
    p->latestWidth = m.fatigueRisk;
    p->latestMono = m.repetitionDensity;
    push_output(p, "emphasis", m.emphasis);  /* :91-101 */
    push_output(p, "coherence", m.coherence);
    push_output(p, "synesthesia", m.synesthesia);
    push_output(p, "fatigue", m.fatigueRisk);
    push_output(p, "repetition", m.repetitionDensity);
    push_output(p, "juiciness", m.score);
}

static void plugin_block(Plugin* p, float* const* ch, int n)
{
    const unsigned int saved = _mm_getcsr(); /* juce::ScopedNoDenormals */
    _mm_setcsr(saved | 0x8040u);
    switch (p->kind) {
        case JO_INFER: infer_block(p, ch, n); break;
        case JO_PUNCH: punch_block(p, ch, n); break;
        case JO_SATURATOR: saturator_block(p, ch, n); break;
        case JO_WIDTH: width_block(p, ch, n); break;
        case JO_COHERE: cohere_block(p, ch, n); break;
        case JO_TEXTURE: texture_block(p, ch, n); break;
        case JO_MOTION: motion_block(p, ch, n); break;
        default: break;
    }
    _mm_setcsr(saved);
}

/* getLatestMetrics + output parameters; same 16-float record as ref_harness.cpp::fillRecord */
static void fill_record(const Plugin* p, float* rec)
{
    Metrics m = metrics_default();
    m.preScore = p->latestPre;
    m.postScore = p->latestPost;
    m.score = p->latestScore;
    if (p->kind == JO_INFER) { /* JuicyInfer/PluginProcessor.cpp:164-181 */
        m.emphasis = p->latestPunch;
        m.coherence = p->latestRichness;
        m.synesthesia = p->latestClarity;
        m.fatigueRisk = p->latestWidth;
        m.repetitionDensity = p->latestMono;
        m.punch = m.emphasis;
        m.richness = m.coherence;
        m.clarity = m.synesthesia;
        m.width = m.fatigueRisk;
        m.monoSafety = m.repetitionDensity;
    } else { /* e.g. JuicyPunch/PluginProcessor.cpp:190-202 */
        m.punch = p->latestPunch;
        m.richness = p->latestRichness;
        m.clarity = p->latestClarity;
        m.width = p->latestWidth;
        m.monoSafety = p->latestMono;
    }
    rec[0] = m.score; rec[1] = m.preScore; rec[2] = m.postScore; rec[3] = m.emphasis;
    rec[4] = m.coherence; rec[5] = m.synesthesia; rec[6] = m.fatigueRisk; rec[7] = m.repetitionDensity;
    rec[8] = m.punch; rec[9] = m.richness; rec[10] = m.clarity; rec[11] = m.width; rec[12] = m.monoSafety;
    rec[13] = rawv(p, "juiciness");
    rec[14] = p->kind == JO_COHERE ? rawv(p, "contextfit") : 0.0f;
    rec[15] = 0.0f;
}

/* ------------------------------------------------------------------ C interface */

void* jo_create(int kind, int channels, double sampleRate, int blockSize)
{
    if (kind < 0 || kind >= JO_NUM_KINDS)
        return NULL;
    Plugin* p = (Plugin*) calloc(1, sizeof(Plugin));
    p->kind = kind;
    p->channels = channels;
    p->latestMono = 1.0f;
    p->targetLow = p->targetMid = p->targetHigh = 0.2f; /* JuicyCohere/PluginProcessor.h:55-57 */
    p->motRng = 0x93ab12f0u;                            /* JuicyMotion/PluginProcessor.h:65 */
    p->texRng = 0x12345678u;
    p->analyzer.sr = 44100.0;
    p->analyzer.channels = 2;
    for (int i = 0; i < kSpecCounts[kind]; ++i) { /* APVTS construction: raw = denormalise(getDefaultValue()) */
        const ParamSpec* s = &kSpecs[kind][i];
        p->stored[i] = s->def;
        const int isBool = s->interval > 0.0f && s->hi == 1.0f && s->lo == 0.0f;
        p->raw[i] = param_denorm(s, isBool ? s->def : range_to01(s, s->def));
    }
    apply_program(p, 0); /* constructors of Infer/Punch/Saturator/Width call setCurrentProgram(0) */
    plugin_prepare(p, sampleRate, blockSize);
    return p;
}

void jo_destroy(void* vp)
{
    Plugin* p = (Plugin*) vp;
    if (p == NULL)
        return;
    free(p->delayL);
    free(p->delayR);
    free(p->tex[0].waveguide);
    free(p->tex[1].waveguide);
    free(p);
}

void jo_prepare(void* p, double sampleRate, int blockSize) { plugin_prepare((Plugin*) p, sampleRate, blockSize); }
int jo_num_params(void* p) { return kSpecCounts[((Plugin*) p)->kind]; }
const char* jo_param_id(void* p, int i) { return kSpecs[((Plugin*) p)->kind][i].id; }
void jo_param_range(void* p, int i, float* out3)
{
    const ParamSpec* s = &kSpecs[((Plugin*) p)->kind][i];
    out3[0] = s->lo;
    out3[1] = s->hi;
    out3[2] = s->interval;
}
int jo_get_param(void* p, const char* id, float* out)
{
    const int idx = find_param((Plugin*) p, id);
    if (idx < 0)
        return -1;
    *out = ((Plugin*) p)->raw[idx];
    return 0;
}
int jo_set_param(void* p, const char* id, float plain)
{
    if (find_param((Plugin*) p, id) < 0)
        return -1;
    set_plain((Plugin*) p, id, plain);
    return 0;
}
int jo_set_param_normalised(void* p, const char* id, float n)
{
    const int idx = find_param((Plugin*) p, id);
    if (idx < 0)
        return -1;
    set_normalised((Plugin*) p, idx, n);
    return 0;
}
int jo_num_programs(void* p) { return presets_of(((Plugin*) p)->kind) != NULL ? 5 : 1; }
int jo_get_program(void* p) { return ((Plugin*) p)->program; }
void jo_set_program(void* p, int i) { apply_program((Plugin*) p, i); }
const char* jo_program_name(void* p, int i)
{
    const Preset* ps = presets_of(((Plugin*) p)->kind);
    if (ps == NULL)
        return "";
    return ps[i < 0 ? 0 : (i > 4 ? 4 : i)].name;
}

long jo_process(void* vp, float* audio, long numSamples, int blockSize, float* history)
{
    Plugin* p = (Plugin*) vp;
    long block = 0;
    for (long pos = 0; pos < numSamples; pos += blockSize, ++block) {
        const int n = (int) (numSamples - pos < blockSize ? numSamples - pos : blockSize);
        float* ch[2] = {audio + pos, p->channels > 1 ? audio + numSamples + pos : NULL};
        plugin_block(p, ch, n);
        if (history != NULL)
            fill_record(p, history + block * 16);
    }
    return block;
}

void jo_latest(void* p, float* rec16) { fill_record((Plugin*) p, rec16); }

double jo_render_clips(void* vp, float* audio, long numClips, long numSamples, int blockSize, double sampleRate,
                       float* lastRecords)
{
    Plugin* p = (Plugin*) vp;
    double seconds = 0.0;
    for (long c = 0; c < numClips; ++c) {
        plugin_prepare(p, sampleRate, blockSize);
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        jo_process(p, audio + c * (long) p->channels * numSamples, numSamples, blockSize, NULL);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        seconds += (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
        if (lastRecords != NULL)
            fill_record(p, lastRecords + c * 16);
    }
    return seconds;
}

/* ------------------------------------------------------------------ meter panel */

/* JuicyMeterPanel::smoothValue, src/shared/JuicyMeterPanel.cpp:3-7 */
static float meter_smooth(float current, float target)
{
    const float alpha = target > current ? 0.28f : 0.12f;
    return current + (target - current) * alpha;
}

typedef struct { float min, max, avg; int count; } MeterStats; /* JuicyMeterPanel.h:16-22 */

/* JuicyMeterPanel::updateStats, src/shared/JuicyMeterPanel.cpp:54-71 */
static void meter_update(MeterStats* s, float value)
{
    const float v = clampf(0.0f, 1.0f, value);
    if (s->count == 0) {
        s->min = v;
        s->max = v;
        s->avg = v;
        s->count = 1;
        return;
    }
    s->min = fmin2(s->min, v);
    s->max = fmax2(s->max, v);
    ++s->count;
    const float n = (float) s->count;
    s->avg += (v - s->avg) / n;
}

/* JuicyMeterPanel::setMetrics over a block-ordered history, src/shared/JuicyMeterPanel.cpp:9-34.
 * Record layout: score 0, preScore 1, postScore 2, emphasis 3, coherence 4, synesthesia 5, fatigueRisk 6,
 * repetitionDensity 7, punch 8, richness 9, clarity 10, width 11, monoSafety 12. */
void jo_meter_run(const float* records, int n, float* out)
{
    /* `JuicinessMetrics metrics;` starts from the struct's defaults (JuicinessAnalyzer.h:6-21): monoSafety 1, rest 0 */
    float preScore = 0.0f, postScore = 0.0f, score = 0.0f, punch = 0.0f, richness = 0.0f, clarity = 0.0f, width = 0.0f,
          monoSafety = 1.0f;
    MeterStats st[10];
    for (int k = 0; k < 10; ++k) {
        st[k].min = 1.0f; st[k].max = 0.0f; st[k].avg = 0.0f; st[k].count = 0;
    }
    static const int statField[10] = { 8, 9, 10, 11, 12, 3, 4, 5, 6, 7 };
    for (int i = 0; i < n; ++i) {
        const float* r = records + 16 * i;
        const float newPre = r[1] > 0.0f ? r[1] : r[0];
        const float newPost = r[2] > 0.0f ? r[2] : r[0];
        preScore = meter_smooth(preScore, newPre);
        postScore = meter_smooth(postScore, newPost);
        for (int k = 0; k < 10; ++k)
            meter_update(&st[k], r[statField[k]]);
        score = meter_smooth(score, newPost);
        punch = meter_smooth(punch, r[8]);
        richness = meter_smooth(richness, r[9]);
        clarity = meter_smooth(clarity, r[10]);
        width = meter_smooth(width, r[11]);
        monoSafety = meter_smooth(monoSafety, r[12]);
    }
    out[0] = preScore; out[1] = postScore; out[2] = score;
    out[3] = punch; out[4] = richness; out[5] = clarity; out[6] = width; out[7] = monoSafety;
    for (int k = 0; k < 10; ++k) {
        out[8 + 3 * k] = st[k].min;
        out[9 + 3 * k] = st[k].max;
        out[10 + 3 * k] = st[k].avg;
    }
    out[38] = (float) st[0].count;
    out[39] = 0.0f;
}
