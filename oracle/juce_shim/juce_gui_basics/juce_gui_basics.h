// TEST INFRASTRUCTURE ONLY (oracle).  Declarations only: enough for the
// reference's src/shared/JuicyPluginEditor.h and JuicyMeterPanel.h to parse,
// because every createEditor() does `new JuicyPluginEditor(...)`
// (e.g. /root/reference/src/plugins/JuicyPunch/PluginProcessor.cpp:126-129).
// The GUI is out of scope; nothing here draws.
#pragma once
#include "../juce_audio_processors/juce_audio_processors.h"

namespace juce
{
class Colour
{
public:
    Colour() = default;
    explicit Colour(uint32 c) : argb(c) {}
    uint32 argb = 0;
};
template <typename T> class Rectangle { public: T x {}, y {}, w {}, h {}; };
class Graphics {};
class Component
{
public:
    virtual ~Component() = default;
    virtual void paint(Graphics&) {}
    virtual void resized() {}
};
class Label : public Component {};
class Slider : public Component {};
class Timer
{
public:
    virtual ~Timer() = default;
    virtual void timerCallback() = 0;
};
class AudioProcessorEditor : public Component
{
public:
    explicit AudioProcessorEditor(AudioProcessor& p) : processor(p) {}
    AudioProcessor& processor;
};
} // namespace juce
