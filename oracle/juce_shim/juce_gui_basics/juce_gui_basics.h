// TEST INFRASTRUCTURE ONLY (oracle).  Declarations only: enough for the
// reference's src/shared/JuicyPluginEditor.h and JuicyMeterPanel.h to parse,
// because every createEditor() does `new JuicyPluginEditor(...)`
// (e.g. /root/reference/src/plugins/JuicyPunch/PluginProcessor.cpp:126-129).
// The GUI is out of scope; nothing here draws.  The drawing API below is
// no-op, just wide enough for the reference's src/shared/JuicyMeterPanel.cpp to COMPILE unmodified:
// its setMetrics / updateStats / smoothValue (:3-34, :54-71) are plain arithmetic on the metrics
// history and are what oracle/_ref/libjuicy_ref_MeterPanel.so runs.
#pragma once
#include "../juce_audio_processors/juce_audio_processors.h"

namespace juce
{
class Colour
{
public:
    Colour() = default;
    explicit Colour(uint32 c) : argb(c) {}
    Colour withAlpha(float) const { return *this; }
    Colour interpolatedWith(Colour, float) const { return *this; }
    Colour withMultipliedSaturation(float) const { return *this; }
    uint32 argb = 0;
};
template <typename T> class Rectangle
{
public:
    Rectangle() = default;
    Rectangle(T x_, T y_, T w_, T h_) : x(x_), y(y_), w(w_), h(h_) {}
    T getX() const { return x; }
    T getY() const { return y; }
    T getWidth() const { return w; }
    T getHeight() const { return h; }
    T getBottom() const { return y + h; }
    Rectangle reduced(T dx, T dy) const { return Rectangle(x + dx, y + dy, w - 2 * dx, h - 2 * dy); }
    Rectangle withWidth(T nw) const { return Rectangle(x, y, nw, h); }
    Rectangle removeFromTop(T a) { Rectangle r(x, y, w, a); y += a; h -= a; return r; }
    Rectangle removeFromBottom(T a) { Rectangle r(x, y + h - a, w, a); h -= a; return r; }
    Rectangle removeFromLeft(T a) { Rectangle r(x, y, a, h); x += a; w -= a; return r; }
    T x {}, y {}, w {}, h {};
};
struct Font { enum Style { plain = 0, bold = 1 }; };
struct FontOptions { FontOptions(float, int) {} };
struct Justification { enum Flags { centredLeft = 33, centredRight = 34 }; Justification(Flags) {} };
class Graphics
{
public:
    void setColour(Colour) {}
    void setFont(const FontOptions&) {}
    template <typename R> void fillRect(const R&) {}
    template <typename R> void drawRect(const R&, int) {}
    template <typename S, typename R> void drawText(const S&, const R&, Justification) {}
    void drawVerticalLine(int, float, float) {}
};
class Component
{
public:
    virtual ~Component() = default;
    virtual void paint(Graphics&) {}
    virtual void resized() {}
    void repaint() {}
    Rectangle<int> getLocalBounds() const { return Rectangle<int>(0, 0, 400, 300); }
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
