// JuicyBatchProcessor.h -- header-only C++17 mirror of the reference's processor API
// (juce::AudioProcessor + AudioProcessorValueTreeState as the Juicy* plugins use them,
// e.g. /root/reference/src/plugins/JuicyPunch/PluginProcessor.h:9-59) over the C ABI of
// include/juicy_batch.h, for N identical instances ("clips") of a plugin or a chain.
//
// Same method names and argument meaning as the reference; errors surface as
// juicy::Error (thrown on THIS side of the ABI only).  No CPU fallback: without a CUDA
// device every compute call throws with JB_ERR_CUDA.
#pragma once
#include "juicy_batch.h"

#include <cstddef>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <vector>

namespace juicy {

struct Error : std::runtime_error {
    int code;
    Error(int c, const char* msg) : std::runtime_error(msg ? msg : "juicy_batch error"), code(c) {}
};

inline void check(int rc)
{
    if (rc != JB_OK)
        throw Error(rc, jb_last_error());
}

// Page-locked planar audio [clip][channel][sample] (N AudioBuffer<float> images end to end).
class PinnedAudio {
public:
    PinnedAudio(int nClips, int nChannels, int nSamples) : clips_(nClips), ch_(nChannels), n_(nSamples)
    {
        void* p = nullptr;
        check(jb_host_alloc(sizeof(float) * (size_t) nClips * (size_t) nChannels * (size_t) nSamples, &p));
        data_ = static_cast<float*>(p);
    }
    ~PinnedAudio() { jb_host_free(data_); }
    PinnedAudio(const PinnedAudio&) = delete;
    PinnedAudio& operator=(const PinnedAudio&) = delete;
    float* data() { return data_; }
    float* channel(int clip, int ch) { return data_ + ((size_t) clip * (size_t) ch_ + (size_t) ch) * (size_t) n_; } // getWritePointer(ch) of clip
    int getNumSamples() const { return n_; }
    int getNumChannels() const { return ch_; }
    int getNumClips() const { return clips_; }

private:
    float* data_ = nullptr;
    int clips_, ch_, n_;
};

class BatchProcessor {
public:
    // createPluginFilter() x nClips for every plugin of `chain` (jb_plugin_kind values, applied in order)
    BatchProcessor(std::initializer_list<int> chain, int nClips, int device = 0, int nChannels = 2)
    {
        std::vector<int> kinds(chain);
        check(jb_create(kinds.data(), (int) kinds.size(), nClips, nChannels, device, &e_));
        nClips_ = nClips;
    }
    ~BatchProcessor() { jb_destroy(e_); }
    BatchProcessor(const BatchProcessor&) = delete;
    BatchProcessor& operator=(const BatchProcessor&) = delete;

    // ---- juce::AudioProcessor
    void prepareToPlay(double sampleRate, int samplesPerBlock) { check(jb_prepare(e_, sampleRate, samplesPerBlock)); }
    void releaseResources() {}
    void reset() { check(jb_reset(e_)); }
    // processBlock over host memory: every clip, every samplesPerBlock-sized block of nSamples
    void processBlock(const float* in, float* out, int nSamples) { check(jb_process_host(e_, in, out, nSamples)); }
    // the same over device memory (asynchronous on the engine's stream)
    void processBlockDevice(const float* dIn, float* dOut, int nSamples) { check(jb_process(e_, dIn, dOut, nSamples)); }
    void synchronize() { check(jb_synchronize(e_)); }

    int getNumPrograms(int slot = 0) const { return jb_num_programs(e_, slot); }
    int getCurrentProgram(int slot = 0) const { return jb_get_program(e_, slot); }
    void setCurrentProgram(int index, int slot = 0) { check(jb_set_program(e_, slot, index)); }
    std::string getProgramName(int index, int slot = 0) const
    {
        const char* s = jb_program_name(e_, slot, index);
        return s ? s : "";
    }

    // ---- AudioProcessorValueTreeState "PARAMS"
    float getRawParameterValue(const char* id, int slot = 0) const
    {
        float v = 0.0f;
        check(jb_get_param(e_, slot, id, &v));
        return v;
    }
    void setParameter(const char* id, float plainValue, int slot = 0) { check(jb_set_param(e_, slot, id, plainValue)); }
    void setValueNotifyingHost(const char* id, float normalised, int slot = 0) { check(jb_set_param_normalised(e_, slot, id, normalised)); }
    std::vector<jb_param_info> getParameters(int slot = 0) const
    {
        std::vector<jb_param_info> out((size_t) jb_num_params(e_, slot));
        for (size_t i = 0; i < out.size(); ++i)
            check(jb_param_info_at(e_, slot, (int) i, &out[i]));
        return out;
    }

    // ---- outputs: getLatestMetrics() of every clip after the most recent block
    std::vector<jb_metrics> getLatestMetrics(int slot = 0)
    {
        std::vector<jb_metrics> out((size_t) nClips_);
        check(jb_get_metrics(e_, slot, out.data()));
        return out;
    }
    void enableHistory(int maxBlocks) { check(jb_enable_history(e_, maxBlocks)); }
    int historyBlocks() const { return jb_history_blocks(e_); }
    std::vector<jb_metrics> getHistory(int slot, int firstBlock, int nBlocks) // [block][clip]
    {
        std::vector<jb_metrics> out((size_t) nBlocks * (size_t) nClips_);
        if (nBlocks > 0)
            check(jb_get_history(e_, slot, firstBlock, nBlocks, out.data()));
        return out;
    }
    // per-instance settings and host automation: the reference re-reads its parameters every block
    void setParameterForClips(int slot, const char* id, float plain, int firstClip, int nClips)
    {
        check(jb_set_param_clips(e_, slot, id, plain, firstClip, nClips));
    }
    void setCurrentProgramForClips(int slot, int index, int firstClip, int nClips)
    {
        check(jb_set_program_clips(e_, slot, index, firstClip, nClips));
    }
    float getRawParameterValueOfClip(int slot, const char* id, int clip) const
    {
        float v = 0.0f;
        check(jb_get_param_clip(e_, slot, id, clip, &v));
        return v;
    }
    // takes effect at the top of absolute host block `atBlock` (counted since prepareToPlay), like
    // setValueNotifyingHost between two processBlock callbacks
    void scheduleParameter(int slot, const char* id, long long atBlock, float plain, int firstClip = JB_ALL_CLIPS, int nClips = 0)
    {
        check(jb_schedule_param(e_, slot, id, atBlock, plain, firstClip, nClips));
    }
    void setMathMode(int mode) { check(jb_set_math_mode(e_, mode)); }
    // get/setStateInformation: the copyXmlToBinary blob of the APVTS "PARAMS" tree
    std::vector<unsigned char> getStateInformation(int slot, int clip = JB_ALL_CLIPS) const
    {
        size_t size = 0;
        check(jb_get_state(e_, slot, clip, nullptr, 0, &size));
        std::vector<unsigned char> blob(size);
        check(jb_get_state(e_, slot, clip, blob.data(), blob.size(), &size));
        return blob;
    }
    void setStateInformation(int slot, const void* data, size_t size, int firstClip = JB_ALL_CLIPS, int nClips = 0)
    {
        check(jb_set_state(e_, slot, data, size, firstClip, nClips));
    }

    // JuicyMeterPanel state of every clip after the render (src/shared/JuicyMeterPanel.cpp:9-34,54-71)
    std::vector<jb_meter_stats> getMeterStatistics(int slot, int firstBlock, int nBlocks, int blockStride = 1)
    {
        std::vector<jb_meter_stats> out((size_t) nClips_);
        check(jb_meter_statistics(e_, slot, firstBlock, nBlocks, blockStride, out.data()));
        return out;
    }

    int getNumClips() const { return nClips_; }
    jb_engine* handle() { return e_; }

private:
    jb_engine* e_ = nullptr;
    int nClips_ = 0;
};

} // namespace juicy
