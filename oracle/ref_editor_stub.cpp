// TEST INFRASTRUCTURE ONLY (oracle).
// Link-time stand-ins for the reference's GUI classes.  Every reference
// createEditor() does `new JuicyPluginEditor(...)`, so the symbols must exist;
// the GUI itself is out of scope and never runs here.
#include "JuicyPluginEditor.h" // found via -I/root/reference/src/shared

JuicyPluginEditor::JuicyPluginEditor(juce::AudioProcessor& audioProcessor,
                                     juce::AudioProcessorValueTreeState& valueTreeState,
                                     MetricsProvider metricsFn,
                                     const juce::String&, bool, bool)
    : AudioProcessorEditor(audioProcessor), state(valueTreeState), metricsProvider(std::move(metricsFn))
{
}
void JuicyPluginEditor::resized() {}
void JuicyPluginEditor::paint(juce::Graphics&) {}
void JuicyPluginEditor::timerCallback() {}
void JuicyPluginEditor::createControls() {}
void JuicyMeterPanel::paint(juce::Graphics&) {}
