// TEST INFRASTRUCTURE ONLY (oracle).
//
// Headless host for ONE reference plugin class, compiled together with the
// reference's unmodified PluginProcessor.cpp + JuicinessAnalyzer.cpp into
// oracle/_ref/libjuicy_ref_<Plugin>.so (see oracle/Makefile).  It plays the role
// of the DAW: createPluginFilter() -> setPlayConfigDetails -> prepareToPlay ->
// processBlock over fixed-size blocks (ragged last block), and after every
// block records getLatestMetrics() plus the output parameters -- the call
// sequence described in SURVEY.md §3.1-3.2.
//
// JUICY_PLUGIN_CLASS / JUICY_PLUGIN_HEADER are supplied by the Makefile.
// Everything the reference header pulls in is included first, so that the
// access-specifier override below touches nothing but the reference's own
// two class definitions (the plugin class and JuicinessAnalyzer).  The
// reference keeps its AudioProcessorValueTreeState member `parameters`
// private and offers no accessor; a host reaches parameters through the
// plugin-format wrapper, which this harness replaces.  Access specifiers do
// not change g++'s object layout or name mangling, and the reference .cpp
// files themselves are compiled without the override.
#include <juce_audio_processors/juce_audio_processors.h>
#include <juce_gui_basics/juce_gui_basics.h>
#include <juce_dsp/juce_dsp.h>
#include <array>
#include <atomic>
#include <chrono>
#include <vector>
#define private public
#include JUICY_PLUGIN_HEADER
#undef private

juce::AudioProcessor* JUCE_CALLTYPE createPluginFilter();

namespace
{
struct Host
{
    std::unique_ptr<JUICY_PLUGIN_CLASS> plugin;
    juce::AudioProcessorValueTreeState* apvts = nullptr;
    int channels = 2;
};
} // namespace

extern "C"
{
void* ref_create(int channels, double sampleRate, int blockSize)
{
    auto* h = new Host();
    h->plugin.reset(static_cast<JUICY_PLUGIN_CLASS*>(createPluginFilter()));
    h->apvts = &h->plugin->parameters;
    h->channels = channels;
    h->plugin->setPlayConfigDetails(channels, channels, sampleRate, blockSize);
    h->plugin->prepareToPlay(sampleRate, blockSize);
    return h;
}

void ref_destroy(void* p) { delete static_cast<Host*>(p); }

void ref_prepare(void* p, double sampleRate, int blockSize)
{
    auto* h = static_cast<Host*>(p);
    h->plugin->setPlayConfigDetails(h->channels, h->channels, sampleRate, blockSize);
    h->plugin->prepareToPlay(sampleRate, blockSize);
}

int ref_num_params(void* p) { return static_cast<Host*>(p)->apvts->getNumParameters(); }

const char* ref_param_id(void* p, int i)
{
    return static_cast<Host*>(p)->apvts->getParameterByIndex(i)->paramID.toStdString().c_str();
}

// display name as createParameterLayout() gives it (what a host shows next to the control)
const char* ref_param_name(void* p, int i)
{
    return static_cast<Host*>(p)->apvts->getParameterByIndex(i)->name.toStdString().c_str();
}

void ref_param_range(void* p, int i, float* out3)
{
    const auto& r = static_cast<Host*>(p)->apvts->getParameterByIndex(i)->getNormalisableRange();
    out3[0] = r.start; out3[1] = r.end; out3[2] = r.interval;
}

// raw (de-normalised) value as processBlock reads it: *getRawParameterValue(id)
int ref_get_param(void* p, const char* id, float* out)
{
    auto* a = static_cast<Host*>(p)->apvts->getRawParameterValue(id);
    if (a == nullptr) return -1;
    *out = a->load();
    return 0;
}

// what a host does when it automates a parameter to a plain value
int ref_set_param(void* p, const char* id, float plainValue)
{
    auto* prm = static_cast<Host*>(p)->apvts->getParameter(id);
    if (prm == nullptr) return -1;
    prm->setValueNotifyingHost(prm->getNormalisableRange().convertTo0to1(plainValue));
    return 0;
}

int ref_set_param_normalised(void* p, const char* id, float n)
{
    auto* prm = static_cast<Host*>(p)->apvts->getParameter(id);
    if (prm == nullptr) return -1;
    prm->setValueNotifyingHost(n);
    return 0;
}

int ref_num_programs(void* p) { return static_cast<Host*>(p)->plugin->getNumPrograms(); }
int ref_get_program(void* p) { return static_cast<Host*>(p)->plugin->getCurrentProgram(); }
void ref_set_program(void* p, int i) { static_cast<Host*>(p)->plugin->setCurrentProgram(i); }
const char* ref_program_name(void* p, int i)
{
    static thread_local std::string s;
    s = static_cast<Host*>(p)->plugin->getProgramName(i).toStdString();
    return s.c_str();
}

static void fillRecord(Host* h, float* rec)
{
    const JuicinessMetrics m = h->plugin->getLatestMetrics();
    rec[0] = m.score; rec[1] = m.preScore; rec[2] = m.postScore; rec[3] = m.emphasis;
    rec[4] = m.coherence; rec[5] = m.synesthesia; rec[6] = m.fatigueRisk; rec[7] = m.repetitionDensity;
    rec[8] = m.punch; rec[9] = m.richness; rec[10] = m.clarity; rec[11] = m.width; rec[12] = m.monoSafety;
    auto* j = h->apvts->getRawParameterValue("juiciness");
    rec[13] = j != nullptr ? j->load() : 0.0f;
    auto* c = h->apvts->getRawParameterValue("contextfit");
    rec[14] = c != nullptr ? c->load() : 0.0f;
    rec[15] = 0.0f;
}

// In-place render of one clip held as planar [channels][numSamples] floats.
// history (may be null) receives 16 floats per block; returns the block count.
long ref_process(void* p, float* audio, long numSamples, int blockSize, float* history)
{
    auto* h = static_cast<Host*>(p);
    juce::MidiBuffer midi;
    juce::AudioBuffer<float> buffer;
    float* chans[2] = { nullptr, nullptr };
    long block = 0;
    for (long pos = 0; pos < numSamples; pos += blockSize, ++block)
    {
        const int n = (int) std::min<long>(blockSize, numSamples - pos);
        for (int c = 0; c < h->channels; ++c)
            chans[c] = audio + (long) c * numSamples + pos;
        buffer.setDataToReferTo(chans, h->channels, n);
        h->plugin->processBlock(buffer, midi);
        if (history != nullptr)
            fillRecord(h, history + block * 16);
    }
    return block;
}

void ref_latest(void* p, float* rec16) { fillRecord(static_cast<Host*>(p), rec16); }

// Render `numClips` independent clips ([clip][channel][sample]) one after the
// other on the calling thread, re-preparing the instance for each; returns the
// seconds spent inside the processBlock loops (CPU baseline).
double ref_render_clips(void* p, float* audio, long numClips, long numSamples, int blockSize, double sampleRate, float* lastRecords)
{
    auto* h = static_cast<Host*>(p);
    double seconds = 0.0;
    for (long c = 0; c < numClips; ++c)
    {
        ref_prepare(p, sampleRate, blockSize);
        const auto t0 = std::chrono::steady_clock::now();
        ref_process(p, audio + c * (long) h->channels * numSamples, numSamples, blockSize, nullptr);
        seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (lastRecords != nullptr)
            fillRecord(h, lastRecords + c * 16);
    }
    return seconds;
}
}
