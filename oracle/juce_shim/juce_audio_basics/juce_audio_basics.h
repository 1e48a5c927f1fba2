// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Minimal stand-in for JUCE's juce_audio_basics module, written from JUCE's
// documented public behaviour (JUCE itself is not vendored by the reference:
// /root/reference/.gitignore:3, /root/reference/CMakeLists.txt:7-11).  It exists
// only so that the reference's own translation units
//   /root/reference/src/shared/JuicinessAnalyzer.cpp
//   /root/reference/src/plugins/Juicy*/PluginProcessor.cpp
// compile UNMODIFIED into oracle/_ref/ (see oracle/Makefile).  Only the members
// those files touch are provided.  Numeric helpers follow SURVEY.md Appendix C.
#pragma once
#include <cstdio>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <memory>
#include <string>
#include <vector>
#include <xmmintrin.h>

#ifndef JUCE_CALLTYPE
#define JUCE_CALLTYPE
#endif
#ifndef JucePlugin_Name
#define JucePlugin_Name "JuicyShimPlugin"
#endif
#define JUCE_DECLARE_NON_COPYABLE(className) \
    className(const className&) = delete;     \
    className& operator=(const className&) = delete;
#define JUCE_DECLARE_NON_COPYABLE_WITH_LEAK_DETECTOR(className) JUCE_DECLARE_NON_COPYABLE(className)

namespace juce
{
using uint32 = std::uint32_t;
using int64 = std::int64_t;

template <typename... Types>
void ignoreUnused(Types&&...) noexcept {}

// jmin/jmax/jlimit/jmap: comparison forms as documented for juce_MathsFunctions.h
template <typename T> constexpr T jmax(T a, T b) { return a < b ? b : a; }
template <typename T> constexpr T jmin(T a, T b) { return b < a ? b : a; }
template <typename T> constexpr T jlimit(T lo, T hi, T v) { return v < lo ? lo : (hi < v ? hi : v); }
template <typename T> constexpr T jmap(T value0To1, T targetMin, T targetMax)
{
    return targetMin + value0To1 * (targetMax - targetMin);
}
template <typename T> T jmap(T v, T srcMin, T srcMax, T tgtMin, T tgtMax)
{
    return tgtMin + ((tgtMax - tgtMin) * (v - srcMin)) / (srcMax - srcMin);
}

template <typename T> struct MathConstants
{
    static constexpr T pi = static_cast<T>(3.141592653589793238L);
    static constexpr T twoPi = static_cast<T>(2 * 3.141592653589793238L);
    static constexpr T halfPi = static_cast<T>(3.141592653589793238L / 2);
};

template <typename T> bool approximatelyEqual(T a, T b)
{
    if (!(std::isfinite(a) && std::isfinite(b)))
        return a == b;
    const T diff = std::abs(a - b);
    return diff <= std::numeric_limits<T>::min()
        || diff <= std::numeric_limits<T>::epsilon() * std::max(std::abs(a), std::abs(b));
}

struct Decibels
{
    template <typename T> static T decibelsToGain(T decibels, T minusInfinityDb = T(-100))
    {
        return decibels > minusInfinityDb ? std::pow(T(10.0), decibels * T(0.05)) : T();
    }
    template <typename T> static T gainToDecibels(T gain, T minusInfinityDb = T(-100))
    {
        return gain > T() ? jmax(minusInfinityDb, static_cast<T>(std::log10(gain)) * T(20.0)) : minusInfinityDb;
    }
};

// Flush-to-zero + denormals-are-zero for the scope (x86 MXCSR bits 15 and 6).
class ScopedNoDenormals
{
public:
    ScopedNoDenormals() noexcept : saved(_mm_getcsr()) { _mm_setcsr(saved | 0x8040u); }
    ~ScopedNoDenormals() noexcept { _mm_setcsr(saved); }
private:
    unsigned int saved;
};

class String
{
public:
    String() = default;
    String(const char* s) : text(s != nullptr ? s : "") {}
    String(const std::string& s) : text(s) {}
    String(float value, int decimals) // juce::String(double, numberOfDecimalPlaces); only the meter panel's labels use it
    {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%.*f", decimals, (double) value);
        text = buf;
    }
    String operator+(const char* s) const { return String(text + (s != nullptr ? s : "")); }
    bool operator==(const String& o) const { return text == o.text; }
    bool operator!=(const String& o) const { return text != o.text; }
    bool operator<(const String& o) const { return text < o.text; }
    int hashCode() const
    {
        int h = 0;
        for (unsigned char c : text) h = 31 * h + (int) c;
        return h;
    }
    const std::string& toStdString() const { return text; }
    bool isEmpty() const { return text.empty(); }
private:
    std::string text;
};

class StringArray
{
public:
    StringArray() = default;
    StringArray(std::initializer_list<const char*> items) { for (auto* s : items) strings.emplace_back(s); }
    int size() const { return (int) strings.size(); }
    const String& operator[](int i) const { return strings[(size_t) i]; }
private:
    std::vector<String> strings;
};

class Identifier
{
public:
    Identifier() = default;
    Identifier(const char* n) : name(n) {}
    Identifier(const String& n) : name(n) {}
    bool operator==(const Identifier& o) const { return name == o.name; }
    bool operator!=(const Identifier& o) const { return name != o.name; }
    const String& toString() const { return name; }
private:
    String name;
};

class MemoryBlock
{
public:
    void setSize(size_t n) { bytes.resize(n); }
    size_t getSize() const { return bytes.size(); }
    void* getData() { return bytes.data(); }
    const void* getData() const { return bytes.data(); }
    void append(const void* d, size_t n)
    {
        auto* p = static_cast<const char*>(d);
        bytes.insert(bytes.end(), p, p + n);
    }
private:
    std::vector<char> bytes;
};

class MidiBuffer {};

// Planar multi-channel buffer.  Semantics the reference relies on: getWritePointer
// clears the "isClear" hint; getRMSLevel accumulates in double; applyGain skips
// a gain that is (approximately) 1 and clears on an exact 0.
template <typename T>
class AudioBuffer
{
public:
    AudioBuffer() = default;
    AudioBuffer(int channelsToAllocate, int samplesToAllocate) { setSize(channelsToAllocate, samplesToAllocate); }

    void setSize(int newNumChannels, int newNumSamples)
    {
        numChannels = newNumChannels;
        numSamples = newNumSamples;
        storage.assign((size_t) numChannels * (size_t) numSamples, T());
        channelPtrs.resize((size_t) numChannels);
        for (int c = 0; c < numChannels; ++c)
            channelPtrs[(size_t) c] = storage.data() + (size_t) c * (size_t) numSamples;
        isClear = false;
    }

    // Refer to caller-owned planar memory (JUCE: setDataToReferTo).
    void setDataToReferTo(T* const* data, int newNumChannels, int newNumSamples)
    {
        numChannels = newNumChannels;
        numSamples = newNumSamples;
        channelPtrs.assign(data, data + newNumChannels);
        isClear = false;
    }

    int getNumChannels() const noexcept { return numChannels; }
    int getNumSamples() const noexcept { return numSamples; }
    const T* getReadPointer(int ch) const noexcept { return channelPtrs[(size_t) ch]; }
    T* getWritePointer(int ch) noexcept { isClear = false; return channelPtrs[(size_t) ch]; }
    T getSample(int ch, int i) const noexcept { return channelPtrs[(size_t) ch][i]; }

    void clear() noexcept
    {
        for (int c = 0; c < numChannels; ++c)
            std::fill(channelPtrs[(size_t) c], channelPtrs[(size_t) c] + numSamples, T());
        isClear = true;
    }
    void clear(int ch, int start, int n) noexcept
    {
        if (!isClear)
            std::fill(channelPtrs[(size_t) ch] + start, channelPtrs[(size_t) ch] + start + n, T());
    }
    bool hasBeenCleared() const noexcept { return isClear; }

    void applyGain(T gain) noexcept
    {
        if (approximatelyEqual(gain, T(1)) || isClear)
            return;
        if (gain == T())
        {
            clear();
            return;
        }
        for (int c = 0; c < numChannels; ++c)
        {
            T* d = channelPtrs[(size_t) c];
            for (int i = 0; i < numSamples; ++i)
                d[i] *= gain;
        }
    }

    T getRMSLevel(int channel, int startSample, int n) const noexcept
    {
        if (n <= 0 || channel < 0 || channel >= numChannels || isClear)
            return T(0);
        const T* data = channelPtrs[(size_t) channel] + startSample;
        double sum = 0.0;
        for (int i = 0; i < n; ++i)
        {
            const double sample = (double) data[i];
            sum += sample * sample;
        }
        return static_cast<T>(std::sqrt(sum / n));
    }

private:
    int numChannels = 0, numSamples = 0;
    std::vector<T> storage;
    std::vector<T*> channelPtrs;
    bool isClear = false;
};
} // namespace juce
