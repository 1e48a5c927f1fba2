// jb_params.cpp -- parameter tables, presets and coefficient derivation (host).
// Compiled with -O2 -ffp-contract=off (no FMA contraction, no fast-math) so the
// fp32 expressions below round exactly like the reference's processBlock
// prologues, which are cited per function (paths relative to /root/reference).
#include "jb_params.h"

#include <cmath>
#include <cstring>

namespace jb {
namespace {

constexpr float kPi = 3.14159265358979323846f;

// juce helpers
inline float jmaxf(float a, float b) { return a < b ? b : a; }
inline float jlimitf(float lo, float hi, float v) { return v < lo ? lo : (hi < v ? hi : v); }
inline float jmap3(float v, float lo, float hi) { return lo + v * (hi - lo); }
inline float jmap5(float v, float s0, float s1, float t0, float t1) { return t0 + ((t1 - t0) * (v - s0)) / (s1 - s0); }
inline float dbToGain(float db) { return db > -100.0f ? std::pow(10.0f, db * 0.05f) : 0.0f; }

ParamSpec F(const char* id, const char* name, float lo, float hi, float def, bool out = false)
{
    return ParamSpec { id, name, lo, hi, 0.0f, def, false, out };
}

// createParameterLayout() of each plugin
const std::vector<ParamSpec> kSpecs[kNumKinds] = {
    // JuicyInfer/PluginProcessor.cpp:183-195
    { F("trim", "Output Trim (dB)", -18.0f, 18.0f, 0.0f), F("sensitivity", "Sensitivity", 0.5f, 2.0f, 1.0f),
      F("juiciness", "Juiciness Score", 0.0f, 100.0f, 0.0f, true), F("emphasis", "Emphasis", 0.0f, 1.0f, 0.0f, true),
      F("coherence", "Coherence", 0.0f, 1.0f, 0.0f, true), F("synesthesia", "Synesthesia", 0.0f, 1.0f, 0.0f, true),
      F("fatigue", "Fatigue Risk", 0.0f, 1.0f, 0.0f, true), F("repetition", "Repetition Density", 0.0f, 1.0f, 0.0f, true) },
    // JuicyPunch/PluginProcessor.cpp:204-215
    { F("punch", "Punch", 0.0f, 1.5f, 0.9f), F("sustain", "Sustain", 0.0f, 1.5f, 0.35f), F("slam", "Slam", 0.0f, 1.0f, 0.65f),
      F("clip", "Clip", 0.0f, 1.0f, 0.25f), F("mix", "Mix", 0.0f, 1.0f, 1.0f), F("output", "Output (dB)", -24.0f, 18.0f, -4.0f),
      F("juiciness", "Juiciness Score", 0.0f, 100.0f, 0.0f, true) },
    // JuicySaturator/PluginProcessor.cpp:189-199
    { F("drive", "Drive (dB)", 0.0f, 24.0f, 6.0f), F("asymmetry", "Asymmetry", -0.5f, 0.5f, 0.1f), F("tone", "Tone", 0.0f, 1.0f, 0.55f),
      F("mix", "Mix", 0.0f, 1.0f, 1.0f), F("output", "Output (dB)", -18.0f, 18.0f, -3.0f),
      F("juiciness", "Juiciness Score", 0.0f, 100.0f, 0.0f, true) },
    // JuicyWidth/PluginProcessor.cpp:229-239
    { F("width", "Stereo Width", 0.0f, 1.0f, 0.45f), F("haasMs", "Haas Delay (ms)", 0.0f, 35.0f, 12.0f), F("monoSafe", "Mono Safety", 0.0f, 1.0f, 0.7f),
      F("mix", "Mix", 0.0f, 1.0f, 1.0f), F("output", "Output (dB)", -18.0f, 18.0f, 0.0f),
      F("juiciness", "Juiciness Score", 0.0f, 100.0f, 0.0f, true) },
    // JuicyCohere/PluginProcessor.cpp:166-178
    { F("match", "Spectral Match", 0.0f, 1.0f, 0.65f), ParamSpec { "learn", "Learn Target", 0.0f, 1.0f, 1.0f, 0.0f, true, false },
      F("tail", "Tail Coherence", 0.0f, 1.0f, 0.45f), F("decay", "Tail Decay", 0.1f, 0.95f, 0.65f), F("mix", "Mix", 0.0f, 1.0f, 1.0f),
      F("output", "Output (dB)", -18.0f, 18.0f, 0.0f), F("contextfit", "Context Fit", 0.0f, 100.0f, 0.0f, true),
      F("juiciness", "Juiciness Score", 0.0f, 100.0f, 0.0f, true) },
    // JuicyTexture/PluginProcessor.cpp:325-337 ("material" is a 5-way choice: Gel, Metal, Wood, Plastic, Flesh-like)
    { ParamSpec { "material", "Material", 0.0f, 4.0f, 1.0f, 0.0f, false, false }, F("tailshape", "Tail Shape", 0.0f, 1.0f, 0.55f),
      F("damping", "Damping", 0.0f, 1.0f, 0.5f), F("weight", "Low-end Weight", 0.0f, 1.0f, 0.45f),
      F("texture", "Texture Layer", 0.0f, 1.0f, 0.5f), F("mix", "Mix", 0.0f, 1.0f, 1.0f), F("output", "Output (dB)", -18.0f, 18.0f, -2.0f),
      F("juiciness", "Juiciness Score", 0.0f, 100.0f, 0.0f, true) },
    // JuicyMotion/PluginProcessor.cpp:189-200
    { F("microvar", "Micro Variation", 0.0f, 1.0f, 0.55f), F("motiondepth", "Motion Depth", 0.0f, 2.0f, 1.0f),
      F("repeatctrl", "Repetition Control", 0.0f, 1.0f, 0.65f), F("budget", "Contrast Budget", 0.0f, 1.0f, 0.5f),
      F("mix", "Mix", 0.0f, 1.0f, 1.0f), F("output", "Output (dB)", -18.0f, 18.0f, -2.0f),
      F("juiciness", "Juiciness Score", 0.0f, 100.0f, 0.0f, true) },
};

// factory programs: JuicyInfer:14-20, JuicyPunch:18-24, JuicySaturator:17-23, JuicyWidth:17-23
const std::vector<Preset> kPresets[kNumKinds] = {
    { { "Reference Lens", 2, { "trim", "sensitivity" }, { 0.0f, 1.0f } },
      { "Detail Hunter", 2, { "trim", "sensitivity" }, { 0.0f, 1.45f } },
      { "Macro Meter", 2, { "trim", "sensitivity" }, { -6.0f, 1.7f } },
      { "Subtle Scout", 2, { "trim", "sensitivity" }, { 0.0f, 0.75f } },
      { "Overdrive Audit", 2, { "trim", "sensitivity" }, { -9.0f, 2.0f } } },
    { { "Solar Snap", 6, { "punch", "sustain", "slam", "clip", "mix", "output" }, { 0.9f, 0.35f, 0.65f, 0.25f, 1.0f, -4.0f } },
      { "Crater Impact", 6, { "punch", "sustain", "slam", "clip", "mix", "output" }, { 1.4f, 0.2f, 0.95f, 0.65f, 1.0f, -8.0f } },
      { "Elastic Slam", 6, { "punch", "sustain", "slam", "clip", "mix", "output" }, { 1.1f, 0.8f, 0.8f, 0.4f, 0.85f, -6.0f } },
      { "Steel Bounce", 6, { "punch", "sustain", "slam", "clip", "mix", "output" }, { 0.7f, 0.55f, 0.45f, 0.1f, 0.75f, -2.0f } },
      { "Apocalypse Tap", 6, { "punch", "sustain", "slam", "clip", "mix", "output" }, { 1.5f, 1.1f, 1.0f, 1.0f, 1.0f, -12.0f } } },
    { { "Amber Heat", 5, { "drive", "asymmetry", "tone", "mix", "output" }, { 6.0f, 0.1f, 0.55f, 1.0f, -3.0f } },
      { "Velvet Burn", 5, { "drive", "asymmetry", "tone", "mix", "output" }, { 11.0f, 0.2f, 0.4f, 0.85f, -6.0f } },
      { "Mirror Glow", 5, { "drive", "asymmetry", "tone", "mix", "output" }, { 8.0f, -0.15f, 0.75f, 0.7f, -4.0f } },
      { "Grain Reactor", 5, { "drive", "asymmetry", "tone", "mix", "output" }, { 18.0f, 0.35f, 0.32f, 1.0f, -10.0f } },
      { "Crystal Edge", 5, { "drive", "asymmetry", "tone", "mix", "output" }, { 4.0f, -0.05f, 0.9f, 0.55f, -1.0f } } },
    { { "Prism Arc", 5, { "width", "haasMs", "monoSafe", "mix", "output" }, { 0.45f, 12.0f, 0.7f, 1.0f, 0.0f } },
      { "Outer Halo", 5, { "width", "haasMs", "monoSafe", "mix", "output" }, { 0.9f, 22.0f, 0.35f, 1.0f, -1.5f } },
      { "Studio Spine", 5, { "width", "haasMs", "monoSafe", "mix", "output" }, { 0.35f, 8.0f, 0.95f, 0.8f, 0.0f } },
      { "Ribbon Drift", 5, { "width", "haasMs", "monoSafe", "mix", "output" }, { 0.7f, 16.0f, 0.55f, 0.65f, -0.5f } },
      { "Monolith Wide", 5, { "width", "haasMs", "monoSafe", "mix", "output" }, { 1.0f, 30.0f, 0.2f, 1.0f, -3.0f } } },
    {}, {}, {},
};

const char* const kKindNames[kNumKinds] = { "JuicyInfer", "JuicyPunch", "JuicySaturator", "JuicyWidth",
                                            "JuicyCohere", "JuicyTexture", "JuicyMotion" };

// juce::NormalisableRange<float> (linear) + RangedAudioParameter::convertFrom0to1
float to01(const ParamSpec& s, float v) { return jlimitf(0.0f, 1.0f, (v - s.lo) / (s.hi - s.lo)); }
float from01(const ParamSpec& s, float p) { return s.lo + (s.hi - s.lo) * jlimitf(0.0f, 1.0f, p); }
float snap(const ParamSpec& s, float v)
{
    if (s.interval > 0.0f)
        v = s.lo + s.interval * std::floor((v - s.lo) / s.interval + 0.5f);
    return jlimitf(s.lo, s.hi, v);
}
float denorm(const ParamSpec& s, float n) { return snap(s, from01(s, jlimitf(0.0f, 1.0f, n))); }

} // namespace

const std::vector<ParamSpec>& paramSpecs(int kind) { return kSpecs[kind]; }
const std::vector<Preset>& presets(int kind) { return kPresets[kind]; }
const char* kindName(int kind) { return kind >= 0 && kind < kNumKinds ? kKindNames[kind] : "?"; }

ParamSet::ParamSet(int kind) : kind_(kind)
{
    const auto& specs = kSpecs[kind];
    stored_.resize(specs.size());
    raw_.resize(specs.size());
    for (size_t i = 0; i < specs.size(); ++i) {
        // APVTS construction: the adapter's raw value is denormalise(getDefaultValue())
        stored_[i] = specs[i].def;
        raw_[i] = denorm(specs[i], specs[i].isBool ? specs[i].def : to01(specs[i], specs[i].def));
    }
    // the constructors of Infer / Punch / Saturator / Width end with setCurrentProgram(0)
    // (e.g. JuicySaturator/PluginProcessor.cpp:26-33)
    if (!kPresets[kind].empty())
        setProgram(0);
}

int ParamSet::find(const char* id) const
{
    const auto& specs = kSpecs[kind_];
    for (size_t i = 0; i < specs.size(); ++i)
        if (std::strcmp(specs[i].id, id) == 0)
            return (int) i;
    return -1;
}

float ParamSet::raw(const char* id) const
{
    const int i = find(id);
    return i >= 0 ? raw_[(size_t) i] : 0.0f;
}

void ParamSet::setNormalised(int index, float n)
{
    const ParamSpec& s = kSpecs[kind_][(size_t) index];
    if (s.isBool) { // AudioParameterBool::setValue
        stored_[(size_t) index] = n >= 0.5f ? 1.0f : 0.0f;
        raw_[(size_t) index] = denorm(s, stored_[(size_t) index]);
    } else {        // RangedAudioParameter value, then the APVTS adapter's denormalise(getValue())
        stored_[(size_t) index] = denorm(s, n);
        raw_[(size_t) index] = denorm(s, to01(s, stored_[(size_t) index]));
    }
}

void ParamSet::setPlain(int index, float plain) { setNormalised(index, to01(kSpecs[kind_][(size_t) index], plain)); }

int ParamSet::numPrograms() const { return kPresets[kind_].empty() ? 1 : (int) kPresets[kind_].size(); }

void ParamSet::setProgram(int index)
{
    const auto& ps = kPresets[kind_];
    if (ps.empty()) { // setCurrentProgram(int) {} for Texture / Motion / Cohere
        program_ = 0;
        return;
    }
    program_ = index < 0 ? 0 : (index >= (int) ps.size() ? (int) ps.size() - 1 : index);
    const Preset& p = ps[(size_t) program_];
    for (int i = 0; i < p.count; ++i) {
        const int idx = find(p.ids[i]);
        if (idx >= 0)
            setPlain(idx, p.values[i]);
    }
}

const char* ParamSet::programName(int index) const
{
    const auto& ps = kPresets[kind_];
    if (ps.empty())
        return "";
    const int safe = index < 0 ? 0 : (index >= (int) ps.size() ? (int) ps.size() - 1 : index);
    return ps[(size_t) safe].name;
}

int widthRingLength(double sr)
{
    const int n = (int) (sr * 0.060);
    return n > 1 ? n : 1;
}

int textureWaveLength(double sr)
{
    const int n = (int) (sr * 0.08);
    return n > 2048 ? n : 2048;
}

// JuicinessAnalyzer::prepare (src/shared/JuicinessAnalyzer.cpp:3-11) and the four
// envelope coefficients analyze() recomputes on every call (:38-41)
AnaCoef makeAnaCoef(double sr)
{
    AnaCoef c {};
    c.aS = std::exp(-1.0f / static_cast<float>(sr * 0.003));
    c.rS = std::exp(-1.0f / static_cast<float>(sr * 0.030));
    c.aL = std::exp(-1.0f / static_cast<float>(sr * 0.050));
    c.rL = std::exp(-1.0f / static_cast<float>(sr * 0.300));
    c.omaS = 1.0f - c.aS;
    c.omrS = 1.0f - c.rS;
    c.omaL = 1.0f - c.aL;
    c.omrL = 1.0f - c.rL;
    c.lowCoeff = 1.0f - std::exp(-2.0f * kPi * 250.0f / static_cast<float>(sr));
    c.highCoeff = 1.0f - std::exp(-2.0f * kPi * 2500.0f / static_cast<float>(sr));
    c.srf = static_cast<float>(sr);
    c.cooldownLen = static_cast<int>(sr * 0.035);
    return c;
}

namespace {

// modeStep's block-constant part (JuicyTexture/PluginProcessor.cpp:79-84)
void modeCoef(TexCoef& t, int k, float freqHz, float t60, float gain, bool constantFreq)
{
    const float srf = t.srf;
    const float tt = jmaxf(0.02f, t60);
    const float r = std::exp(std::log(0.001f) / (tt * srf));
    t.modeGain[k] = gain;
    t.modeF[k] = freqHz;
    t.modeTwoR[k] = 2.0f * r;
    t.modeA2[k] = -r * r;
    if (constantFreq) {
        const float f = jlimitf(20.0f, 0.45f * srf, freqHz);
        const float theta = 2.0f * kPi * f / srf;
        t.modeA1[k] = 2.0f * r * std::cos(theta);
    } else {
        t.modeA1[k] = 0.0f;
    }
}

void makeTexture(const ParamSet& p, double sr, TexCoef& t)
{
    std::memset(&t, 0, sizeof t);
    // JuicyTexture/PluginProcessor.cpp:55-75
    const int mode = static_cast<int>(p.raw("material"));
    const float tailShape = p.raw("tailshape"), damping = p.raw("damping"), weight = p.raw("weight");
    const float texture = p.raw("texture");
    const float srf = static_cast<float>(sr);
    t.material = mode;
    t.srf = srf;
    t.mix = p.raw("mix");
    t.outGain = dbToGain(p.raw("output"));
    t.inTrim = (mode == 1 ? 0.58f : (mode == 2 ? 0.62f : (mode == 3 ? 0.60f : 1.0f))); // :117
    t.matTrim = (mode == 1 ? 0.62f : (mode == 2 ? 0.54f : (mode == 3 ? 0.62f : 1.0f)));
    const float dampingAmt = jlimitf(0.0f, 1.0f, damping);
    const float dampingMul = jmap5(dampingAmt, 0.0f, 1.0f, 1.35f, 0.40f);
    t.tailShape = tailShape;
    t.decay = jmap5(tailShape, 0.0f, 1.0f, 0.30f, 0.985f) * jmap5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.80f);
    t.lowBoost = 1.0f + weight * 1.0f;
    t.splitLow = 1.0f - std::exp(-2.0f * kPi * 140.0f / srf);
    t.splitHigh = 1.0f - std::exp(-2.0f * kPi * 2600.0f / srf);
    t.envAtk = std::exp(-1.0f / static_cast<float>(sr * 0.0025));
    t.envRel = std::exp(-1.0f / static_cast<float>(sr * 0.080));
    t.wetAtk = std::exp(-1.0f / static_cast<float>(sr * 0.005));
    t.wetRel = std::exp(-1.0f / static_cast<float>(sr * 0.090));
    t.omEnvAtk = 1.0f - t.envAtk;
    t.omEnvRel = 1.0f - t.envRel;
    t.omWetAtk = 1.0f - t.wetAtk;
    t.omWetRel = 1.0f - t.wetRel;
    t.autoGainBase = jmap5(texture, 0.0f, 1.0f, 0.78f, 0.54f);
    t.highTilt = 0.9f + texture * 1.3f;      // :131
    t.noiseAmt = 0.004f + 0.022f * texture;  // :243
    t.dynK = 0.18f + texture * 0.12f;        // :245
    t.fMax = 0.45f * srf;
    t.waveSize = textureWaveLength(sr);
    switch (mode) {
        case 0: { // gel :139-149
            const float f0 = 42.0f + texture * 88.0f;
            t.gelOmega = 2.0f * kPi * f0 / srf;
            t.gelK = t.gelOmega * t.gelOmega;
            t.shapeGain = 0.96f + 0.28f * texture;
            break;
        }
        case 1: { // metal :154-166
            const float f0 = 320.0f + 140.0f * texture;
            const float metalDamp = jmap5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.55f);
            const float tScale = jmap3(tailShape, 0.18f, 0.72f) * dampingMul * metalDamp;
            modeCoef(t, 0, f0 * 1.00f, 0.56f * tScale, 0.34f, false);
            modeCoef(t, 1, f0 * 2.31f, 0.40f * tScale, 0.20f, false);
            modeCoef(t, 2, f0 * 4.18f, 0.26f * tScale, 0.13f, false);
            modeCoef(t, 3, f0 * 6.87f, 0.17f * tScale, 0.09f, false);
            t.shapeGain = 0.78f + 0.10f * texture;
            break;
        }
        case 2: { // wood :172-190
            const float cavityHz = 92.0f + 95.0f * (0.5f * weight + 0.5f * texture);
            t.delaySamp = jlimitf(16.0f, static_cast<float>(t.waveSize - 2), srf / cavityHz);
            t.waveDamp = jmap3(tailShape, 0.26f, 0.90f) * jmap5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.72f);
            const float woodDamp = jmap5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.64f);
            const float tScale = jmap3(tailShape, 0.18f, 0.62f) * dampingMul * woodDamp;
            modeCoef(t, 0, 155.0f, 0.40f * tScale, 0.32f, true);
            modeCoef(t, 1, 355.0f, 0.27f * tScale, 0.18f, true);
            modeCoef(t, 2, 690.0f, 0.16f * tScale, 0.10f, true);
            modeCoef(t, 3, 1130.0f, 0.10f * tScale, 0.06f, true);
            t.excA = 0.10f; t.excB = 0.34f;
            t.waveMixA = 0.56f; t.waveMixB = 0.24f; t.waveOut = 0.30f;
            t.shapeGain = 0.74f + 0.08f * texture;
            break;
        }
        case 3: { // plastic :195-211
            const float tubeHz = 210.0f + 340.0f * texture;
            t.delaySamp = jlimitf(8.0f, static_cast<float>(t.waveSize - 2), srf / tubeHz);
            t.waveDamp = jmap3(tailShape, 0.22f, 0.91f) * jmap5(dampingAmt, 0.0f, 1.0f, 1.0f, 0.82f);
            const float tScale = jmap3(tailShape, 0.16f, 0.72f) * dampingMul;
            modeCoef(t, 0, 280.0f, 0.28f * tScale, 0.34f, true);
            modeCoef(t, 1, 690.0f, 0.18f * tScale, 0.22f, true);
            modeCoef(t, 2, 1320.0f, 0.11f * tScale, 0.16f, true);
            modeCoef(t, 3, 2360.0f, 0.07f * tScale, 0.11f, true);
            t.excA = 0.20f; t.excB = 0.60f;
            t.waveMixA = 0.52f; t.waveMixB = 0.36f; t.waveOut = 0.40f;
            t.shapeGain = 0.80f + 0.10f * texture;
            break;
        }
        default: { // flesh :216-234
            const float wA = 2.0f * kPi * (38.0f + 52.0f * texture) / srf;
            const float wB = 2.0f * kPi * (88.0f + 72.0f * texture) / srf;
            t.kA = wA * wA;
            t.kB = wB * wB;
            t.cA = 2.0f * jmap3(tailShape, 0.56f, 1.18f) * wA;
            t.cB = 2.0f * jmap3(tailShape, 0.70f, 1.34f) * wB;
            t.kCouple = 0.14f + 0.24f * texture;
            t.shapeGain = 0.98f + 0.16f * texture;
            break;
        }
    }
}

} // namespace

void makeSlotCoef(const ParamSet& p, double sr, SlotCoef* out)
{
    std::memset(out, 0, sizeof *out);
    const float srf = static_cast<float>(sr);
    switch (p.kind()) {
        case kInfer: { // JuicyInfer/PluginProcessor.cpp:74-79
            InferCoef& c = out->infer;
            c.trimGain = dbToGain(p.raw("trim"));
            c.sensitivity = p.raw("sensitivity");
            // AudioBuffer::applyGain: skipped when the gain is (approximately) 1, clears on 0
            const float diff = std::fabs(c.trimGain - 1.0f);
            const float mx = std::fabs(c.trimGain) > 1.0f ? std::fabs(c.trimGain) : 1.0f;
            const bool isOne = diff <= 1.17549435e-38f || diff <= 1.1920929e-7f * mx;
            c.gainMode = isOne ? 0 : (c.trimGain == 0.0f ? 2 : 1);
            break;
        }
        case kPunch: { // JuicyPunch/PluginProcessor.cpp:74-84, :100-108
            PunchCoef& c = out->punch;
            const float punchAmt = p.raw("punch"), sustainAmt = p.raw("sustain"), slamAmt = p.raw("slam"), clipAmt = p.raw("clip");
            c.mix = p.raw("mix");
            c.outGain = dbToGain(p.raw("output"));
            c.fastCoeff = std::exp(-1.0f / static_cast<float>(sr * 0.0015));
            c.slowCoeff = std::exp(-1.0f / static_cast<float>(sr * 0.110));
            c.omFast = 1.0f - c.fastCoeff;
            c.omSlow = 1.0f - c.slowCoeff;
            c.curveExp = jmap5(slamAmt, 0.0f, 1.0f, 0.95f, 0.55f);
            c.punchK = punchAmt * 12.0f + slamAmt * 22.0f;
            c.sustainK = sustainAmt * 4.0f + slamAmt * 1.5f;
            c.drive = 1.0f + clipAmt * 8.0f + slamAmt * 4.0f;
            c.tanhDrive = std::tanh(c.drive);
            c.hardK = 1.0f + clipAmt * 2.0f;
            c.clipAmt = clipAmt;
            break;
        }
        case kSaturator: { // JuicySaturator/PluginProcessor.cpp:74-81 (rate = getSampleRate())
            SatCoef& c = out->sat;
            c.asym = p.raw("asymmetry");
            c.mix = p.raw("mix");
            c.inGain = dbToGain(p.raw("drive"));
            c.outGain = dbToGain(p.raw("output"));
            const float cutoff = jmap5(p.raw("tone"), 0.0f, 1.0f, 2500.0f, 16000.0f);
            c.toneCoeff = 1.0f - std::exp(-2.0f * kPi * cutoff / srf);
            break;
        }
        case kWidth: { // JuicyWidth/PluginProcessor.cpp:91-97, :110
            WidthCoef& c = out->width;
            c.ringLen = widthRingLength(sr);
            c.delaySamples = static_cast<int>(sr * (p.raw("haasMs") * 0.001f));
            c.width = p.raw("width");
            c.dynamicLimit = jmap5(p.raw("monoSafe"), 0.0f, 1.0f, 1.0f, 0.35f);
            c.mix = p.raw("mix");
            c.outGain = dbToGain(p.raw("output"));
            break;
        }
        case kCohere: { // JuicyCohere/PluginProcessor.cpp:16-17, :54-60, :97, :116
            CohereCoef& c = out->cohere;
            c.lowCoeff = 1.0f - std::exp(-2.0f * kPi * 220.0f / srf);
            c.highCoeff = 1.0f - std::exp(-2.0f * kPi * 2400.0f / srf);
            c.matchQ = 0.25f * p.raw("match");
            c.learn = p.raw("learn") > 0.5f ? 1 : 0;
            c.tailK = p.raw("tail") * 0.35f;
            c.fb = jlimitf(0.0f, 0.93f, p.raw("decay"));
            c.mix = p.raw("mix");
            c.outGain = dbToGain(p.raw("output"));
            break;
        }
        case kTexture:
            makeTexture(p, sr, out->tex);
            break;
        case kMotion: { // JuicyMotion/PluginProcessor.cpp:59-73, :119-141
            MotionCoef& c = out->motion;
            const float microVar = p.raw("microvar"), motionDepth = p.raw("motiondepth"), repeatCtrl = p.raw("repeatctrl");
            c.microVar = microVar;
            c.repeatCtrl = repeatCtrl;
            c.mix = p.raw("mix");
            c.outGain = dbToGain(p.raw("output"));
            c.srf = srf;
            c.envCoeff = std::exp(-1.0f / static_cast<float>(sr * 0.015));
            c.budgetCoeff = std::exp(-1.0f / static_cast<float>(sr * 0.080));
            c.omEnv = 1.0f - c.envCoeff;
            c.omBudget = 1.0f - c.budgetCoeff;
            c.tailFeedback = jmap5(repeatCtrl, 0.0f, 1.0f, 0.15f, 0.88f);
            const float depth = jlimitf(0.0f, 2.0f, motionDepth);
            c.depth = depth;
            const float motionRateHz = jmap5(microVar, 0.0f, 1.0f, 0.25f, 2.0f) * jmap5(depth, 0.0f, 2.0f, 0.75f, 1.6f);
            c.motionInc = (2.0f * kPi * motionRateHz) / srf;
            c.varSlew = std::exp(-1.0f / static_cast<float>(sr * 0.020));
            c.omVarSlew = 1.0f - c.varSlew;
            c.lfoDepth = (250.0f + 550.0f * microVar) * (0.5f + 0.9f * depth);
            c.d06 = 0.6f + 0.6f * depth;
            c.d07 = 0.6f + 0.7f * depth;
            c.d08 = 0.6f + 0.8f * depth;
            c.d0507 = 0.55f + 0.7f * depth;
            c.d0508 = 0.5f + 0.8f * depth;
            c.mv035 = 0.35f * microVar;
            c.mvT = 0.12f + 0.30f * microVar;
            c.tailMix = (0.26f + 0.24f * microVar) * (0.6f + 0.7f * depth);
            c.wetBoost = 1.0f + 0.9f * microVar * (0.55f + 0.9f * depth);
            c.budgetTarget = jmap5(p.raw("budget"), 0.0f, 1.0f, 0.8f, 0.25f);
            c.cooldownLen = static_cast<int>(sr * 0.04);
            break;
        }
        default:
            break;
    }
}

} // namespace jb
