// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Minimal stand-in for JUCE's juce_audio_processors module (see the note in
// juce_audio_basics.h).  Provides the AudioProcessor / parameter / APVTS surface
// the reference's PluginProcessor.{h,cpp} files use, with the value semantics of
// the real classes: a RangedAudioParameter stores the de-normalised ("raw")
// value; setValueNotifyingHost(n) stores snapToLegalValue(convertFrom0to1(n));
// the APVTS adapter's raw atomic is convertFrom0to1(convertTo0to1(value)).
#pragma once
#include "../juce_audio_basics/juce_audio_basics.h"
#include <map>

namespace juce
{
template <typename T>
class NormalisableRange
{
public:
    NormalisableRange() = default;
    NormalisableRange(T s, T e, T i = T(0)) : start(s), end(e), interval(i) {}
    T convertTo0to1(T v) const noexcept { return jlimit(T(0), T(1), (v - start) / (end - start)); }
    T convertFrom0to1(T p) const noexcept { return start + (end - start) * jlimit(T(0), T(1), p); }
    T snapToLegalValue(T v) const noexcept
    {
        if (interval > T(0))
            v = start + interval * std::floor((v - start) / interval + T(0.5));
        return jlimit(start, end, v);
    }
    T start = T(0), end = T(1), interval = T(0);
};

class AudioProcessorParameter
{
public:
    virtual ~AudioProcessorParameter() = default;
    virtual float getValue() const = 0;
    virtual void setValue(float newValue) = 0;
    virtual float getDefaultValue() const = 0;
    void setValueNotifyingHost(float newValue)
    {
        setValue(newValue);
        if (onValueChanged) onValueChanged();
    }
    std::function<void()> onValueChanged; // stands in for the listener list
};

class AudioProcessorParameterWithID : public AudioProcessorParameter
{
public:
    AudioProcessorParameterWithID(const String& idToUse, const String& nameToUse) : paramID(idToUse), name(nameToUse) {}
    const String paramID, name;
};

class RangedAudioParameter : public AudioProcessorParameterWithID
{
public:
    using AudioProcessorParameterWithID::AudioProcessorParameterWithID;
    virtual const NormalisableRange<float>& getNormalisableRange() const = 0;
    float convertTo0to1(float v) const noexcept { return getNormalisableRange().convertTo0to1(v); }
    float convertFrom0to1(float v) const noexcept
    {
        const auto& r = getNormalisableRange();
        return r.snapToLegalValue(r.convertFrom0to1(jlimit(0.0f, 1.0f, v)));
    }
};

class AudioParameterFloat : public RangedAudioParameter
{
public:
    AudioParameterFloat(const String& id, const String& nm, float minV, float maxV, float def)
        : RangedAudioParameter(id, nm), range(minV, maxV), value(def), defaultValue(def) {}
    const NormalisableRange<float>& getNormalisableRange() const override { return range; }
    float get() const noexcept { return value.load(std::memory_order_relaxed); }
    float getValue() const override { return convertTo0to1(get()); }
    void setValue(float n) override { value.store(convertFrom0to1(n), std::memory_order_relaxed); }
    float getDefaultValue() const override { return convertTo0to1(defaultValue); }
    NormalisableRange<float> range;
private:
    std::atomic<float> value;
    float defaultValue;
};

class AudioParameterBool : public RangedAudioParameter
{
public:
    AudioParameterBool(const String& id, const String& nm, bool def)
        : RangedAudioParameter(id, nm), range(0.0f, 1.0f, 1.0f), value(def ? 1.0f : 0.0f), defaultValue(def ? 1.0f : 0.0f) {}
    const NormalisableRange<float>& getNormalisableRange() const override { return range; }
    float getValue() const override { return value.load(std::memory_order_relaxed); }
    void setValue(float n) override { value.store(n >= 0.5f ? 1.0f : 0.0f, std::memory_order_relaxed); }
    float getDefaultValue() const override { return defaultValue; }
private:
    NormalisableRange<float> range;
    std::atomic<float> value;
    float defaultValue;
};

class AudioParameterChoice : public RangedAudioParameter
{
public:
    AudioParameterChoice(const String& id, const String& nm, const StringArray& c, int defIndex)
        : RangedAudioParameter(id, nm), choices(c), range(0.0f, (float) (c.size() - 1), 1.0f),
          value((float) defIndex), defaultValue(convertTo0to1((float) defIndex)) {}
    const NormalisableRange<float>& getNormalisableRange() const override { return range; }
    float getValue() const override { return convertTo0to1(value.load(std::memory_order_relaxed)); }
    void setValue(float n) override { value.store(convertFrom0to1(n), std::memory_order_relaxed); }
    float getDefaultValue() const override { return defaultValue; }
    const StringArray choices;
private:
    NormalisableRange<float> range;
    std::atomic<float> value;
    float defaultValue;
};

class AudioChannelSet
{
public:
    static AudioChannelSet mono() { return AudioChannelSet(1); }
    static AudioChannelSet stereo() { return AudioChannelSet(2); }
    static AudioChannelSet disabled() { return AudioChannelSet(0); }
    int size() const noexcept { return n; }
    bool operator==(const AudioChannelSet& o) const noexcept { return n == o.n; }
    bool operator!=(const AudioChannelSet& o) const noexcept { return n != o.n; }
private:
    explicit AudioChannelSet(int c) : n(c) {}
    int n;
};

class XmlElement
{
public:
    explicit XmlElement(const String& t) : tag(t) {}
    bool hasTagName(const Identifier& t) const { return tag == t.toString(); }
    bool hasTagName(const String& t) const { return tag == t; }
    String tag;
    std::map<std::string, float> values;
};

class ValueTree
{
public:
    ValueTree() = default;
    explicit ValueTree(const Identifier& t) : type(t) {}
    Identifier getType() const { return type; }
    std::unique_ptr<XmlElement> createXml() const
    {
        auto x = std::make_unique<XmlElement>(type.toString());
        x->values = values;
        return x;
    }
    static ValueTree fromXml(const XmlElement& x)
    {
        ValueTree v { Identifier(x.tag) };
        v.values = x.values;
        return v;
    }
    std::map<std::string, float> values;
private:
    Identifier type;
};

class UndoManager;
class AudioProcessorEditor;

class AudioProcessor
{
public:
    struct BusesLayout
    {
        AudioChannelSet in = AudioChannelSet::stereo(), out = AudioChannelSet::stereo();
        AudioChannelSet getMainInputChannelSet() const { return in; }
        AudioChannelSet getMainOutputChannelSet() const { return out; }
    };
    struct BusesProperties
    {
        BusesProperties withInput(const String&, const AudioChannelSet& s, bool = true) const { auto b = *this; b.layout.in = s; return b; }
        BusesProperties withOutput(const String&, const AudioChannelSet& s, bool = true) const { auto b = *this; b.layout.out = s; return b; }
        BusesLayout layout;
    };

    AudioProcessor() = default;
    explicit AudioProcessor(const BusesProperties& p) : numIn(p.layout.in.size()), numOut(p.layout.out.size()) {}
    virtual ~AudioProcessor() = default;

    virtual void prepareToPlay(double sampleRate, int samplesPerBlock) = 0;
    virtual void releaseResources() = 0;
    virtual bool isBusesLayoutSupported(const BusesLayout&) const { return true; }
    virtual void processBlock(AudioBuffer<float>&, MidiBuffer&) = 0;
    virtual AudioProcessorEditor* createEditor() = 0;
    virtual bool hasEditor() const = 0;
    virtual const String getName() const = 0;
    virtual bool acceptsMidi() const = 0;
    virtual bool producesMidi() const = 0;
    virtual bool isMidiEffect() const { return false; }
    virtual double getTailLengthSeconds() const = 0;
    virtual int getNumPrograms() = 0;
    virtual int getCurrentProgram() = 0;
    virtual void setCurrentProgram(int) = 0;
    virtual const String getProgramName(int) = 0;
    virtual void changeProgramName(int, const String&) = 0;
    virtual void getStateInformation(MemoryBlock&) = 0;
    virtual void setStateInformation(const void*, int) = 0;

    int getTotalNumInputChannels() const noexcept { return numIn; }
    int getTotalNumOutputChannels() const noexcept { return numOut; }
    double getSampleRate() const noexcept { return currentSampleRate; }
    int getBlockSize() const noexcept { return blockSize; }
    void setPlayConfigDetails(int ins, int outs, double sr, int bs) { numIn = ins; numOut = outs; currentSampleRate = sr; blockSize = bs; }
    void setRateAndBufferSizeDetails(double sr, int bs) { currentSampleRate = sr; blockSize = bs; }

    // State blobs: the shim serialises "id=value" pairs; the real XML/binary format
    // is out of scope (SURVEY.md §8(b) "State I/O").
    static void copyXmlToBinary(const XmlElement& xml, MemoryBlock& dest)
    {
        std::string s = xml.tag.toStdString() + "\n";
        for (auto& kv : xml.values) s += kv.first + "=" + std::to_string(kv.second) + "\n";
        dest.setSize(0);
        dest.append(s.data(), s.size());
    }
    static std::unique_ptr<XmlElement> getXmlFromBinary(const void* data, int size)
    {
        if (data == nullptr || size <= 0) return nullptr;
        std::string s(static_cast<const char*>(data), (size_t) size);
        auto nl = s.find('\n');
        if (nl == std::string::npos) return nullptr;
        auto x = std::make_unique<XmlElement>(String(s.substr(0, nl)));
        size_t pos = nl + 1;
        while (pos < s.size())
        {
            auto e = s.find('\n', pos);
            if (e == std::string::npos) e = s.size();
            auto line = s.substr(pos, e - pos);
            auto eq = line.find('=');
            if (eq != std::string::npos) x->values[line.substr(0, eq)] = std::stof(line.substr(eq + 1));
            pos = e + 1;
        }
        return x;
    }

private:
    int numIn = 2, numOut = 2;
    double currentSampleRate = 0.0;
    int blockSize = 0;
};

class AudioProcessorValueTreeState
{
public:
    class ParameterLayout
    {
    public:
        template <typename It>
        ParameterLayout(It b, It e) { for (; b != e; ++b) params.push_back(std::move(*b)); }
        std::vector<std::unique_ptr<RangedAudioParameter>> params;
    };
    class SliderAttachment {};

    AudioProcessorValueTreeState(AudioProcessor& p, UndoManager*, const Identifier& valueTreeType, ParameterLayout layout)
        : processor(p), state(valueTreeType)
    {
        for (auto& up : layout.params)
        {
            auto a = std::make_unique<Adapter>();
            a->param = std::move(up);
            auto* raw = a.get();
            // initial raw value: denormalise(getDefaultValue()), then kept in sync on every change
            raw->unnormalised.store(raw->param->convertFrom0to1(raw->param->getDefaultValue()), std::memory_order_relaxed);
            raw->param->onValueChanged = [raw] {
                raw->unnormalised.store(raw->param->convertFrom0to1(raw->param->getValue()), std::memory_order_relaxed);
            };
            adapters.push_back(std::move(a));
        }
    }

    RangedAudioParameter* getParameter(const String& id) const
    {
        for (auto& a : adapters) if (a->param->paramID == id) return a->param.get();
        return nullptr;
    }
    std::atomic<float>* getRawParameterValue(const String& id) const
    {
        for (auto& a : adapters) if (a->param->paramID == id) return &a->unnormalised;
        return nullptr;
    }
    ValueTree copyState()
    {
        ValueTree v { state.getType() };
        for (auto& a : adapters) v.values[a->param->paramID.toStdString()] = a->unnormalised.load();
        return v;
    }
    void replaceState(const ValueTree& v)
    {
        for (auto& kv : v.values)
            if (auto* p = getParameter(String(kv.first)))
                p->setValueNotifyingHost(p->convertTo0to1(kv.second));
    }
    int getNumParameters() const { return (int) adapters.size(); }
    RangedAudioParameter* getParameterByIndex(int i) const { return adapters[(size_t) i]->param.get(); }

    AudioProcessor& processor;
    ValueTree state;

private:
    struct Adapter
    {
        std::unique_ptr<RangedAudioParameter> param;
        std::atomic<float> unnormalised { 0.0f };
    };
    std::vector<std::unique_ptr<Adapter>> adapters;
};
} // namespace juce
