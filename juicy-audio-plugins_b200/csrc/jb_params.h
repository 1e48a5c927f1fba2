// jb_params.h -- host-side mirror of the reference's parameter surface
// (AudioProcessorValueTreeState "PARAMS": ids, ranges, defaults, factory presets;
// SURVEY.md Appendix A) and the derivation of every block-constant coefficient
// from the raw parameter values, with the reference's own fp32 expressions and
// glibc libm calls (this file's .cpp is compiled with -ffp-contract=off).
#pragma once
#include "jb_kernels.h"

#include <string>
#include <vector>

namespace jb {

enum Kind { kInfer = 0, kPunch, kSaturator, kWidth, kCohere, kTexture, kMotion, kNumKinds };

struct ParamSpec {
    const char* id;
    const char* name;
    float lo, hi, interval, def;
    bool isBool;
    bool isOutput;
};

struct Preset {
    const char* name;
    int count;
    const char* ids[6];
    float values[6];
};

const std::vector<ParamSpec>& paramSpecs(int kind);
const std::vector<Preset>& presets(int kind); // empty for plugins with one no-op program
const char* kindName(int kind);

// One plugin's parameter block: the RangedAudioParameter values and what
// getRawParameterValue() reports for them.
class ParamSet {
public:
    explicit ParamSet(int kind);
    int kind() const { return kind_; }
    int count() const { return (int) raw_.size(); }
    int find(const char* id) const;
    float raw(int index) const { return raw_[(size_t) index]; }
    float raw(const char* id) const;
    void setNormalised(int index, float normalised); // setValueNotifyingHost(n)
    void setPlain(int index, float plain);           // setValueNotifyingHost(range.convertTo0to1(v))
    int numPrograms() const;
    int currentProgram() const { return program_; }
    void setProgram(int index);                      // setCurrentProgram(index)
    const char* programName(int index) const;
    bool operator==(const ParamSet& o) const { return kind_ == o.kind_ && program_ == o.program_ && stored_ == o.stored_ && raw_ == o.raw_; }

private:
    int kind_;
    int program_ = 0;
    std::vector<float> stored_, raw_;
};

// prepareToPlay-time sizes
int widthRingLength(double sampleRate);     // jmax(1, int(sr*0.060))   JuicyWidth/PluginProcessor.cpp:38-39
int textureWaveLength(double sampleRate);   // jmax(2048, int(sr*0.08)) JuicyTexture/PluginProcessor.cpp:18

AnaCoef makeAnaCoef(double sampleRate);
void makeSlotCoef(const ParamSet& p, double sampleRate, SlotCoef* out);

// Seeded synthetic clips of SURVEY.md §8(d) on the host (jb_synth.cpp); the device
// generator in jb_kernels.cu follows the same formulas.
void synthFillHost(float* audio, int kind, long long firstClip, int nClips, int nCh, int nSamples,
                   double sampleRate, unsigned int seed);

} // namespace jb
