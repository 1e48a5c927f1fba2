// demo_render.cpp -- the C++ host flow of INTEGRATION.md §2 without JUCE: render a batch of
// seeded synthetic drum-hit clips through Punch -> Width and print per-clip records.
//   demo_render [n_clips] [n_samples] [--params-only]
// --params-only exercises the parameter/program surface on a device-less engine (CPU test).
// Output: one line "clip <i> juiciness <v> score <v> checksum <sum of |samples|>" per clip.
#include "JuicyBatchProcessor.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char** argv)
{
    int nClips = 8, nSamples = 2048;
    bool paramsOnly = false;
    int pos = 0;
    for (int i = 1; i < argc; ++i) {
        if (std::strcmp(argv[i], "--params-only") == 0)
            paramsOnly = true;
        else if (pos++ == 0)
            nClips = std::atoi(argv[i]);
        else
            nSamples = std::atoi(argv[i]);
    }
    try {
        juicy::BatchProcessor batch({ JB_PUNCH, JB_WIDTH }, nClips, paramsOnly ? -1 : 0);
        batch.setCurrentProgram(2, 0); // "Elastic Slam" (JuicyPunch/PluginProcessor.cpp:18-24)
        batch.setParameter("haasMs", 16.0f, 1);
        std::printf("program %s punch %.9g sustain %.9g haasMs %.9g\n", batch.getProgramName(2, 0).c_str(),
                    batch.getRawParameterValue("punch", 0), batch.getRawParameterValue("sustain", 0),
                    batch.getRawParameterValue("haasMs", 1));
        for (const auto& p : batch.getParameters(1))
            std::printf("param %s [%g, %g] default %g%s\n", p.id, p.min_value, p.max_value, p.default_value, p.is_output ? " (output)" : "");
        if (paramsOnly)
            return 0;
        batch.prepareToPlay(48000.0, 512);
        juicy::PinnedAudio audio(nClips, 2, nSamples);
        juicy::check(jb_synth_fill_host(audio.data(), /*drum*/ 3, 0, nClips, 2, nSamples, 48000.0, 0x4A554943u));
        batch.processBlock(audio.data(), audio.data(), nSamples);
        const auto rec = batch.getLatestMetrics(1);
        for (int c = 0; c < nClips; ++c) {
            double sum = 0.0;
            for (int ch = 0; ch < 2; ++ch)
                for (int i = 0; i < nSamples; ++i)
                    sum += std::fabs((double) audio.channel(c, ch)[i]);
            std::printf("clip %d juiciness %.6f score %.6f checksum %.9g\n", c, rec[(size_t) c].juiciness, rec[(size_t) c].score, sum);
        }
    } catch (const juicy::Error& e) {
        std::fprintf(stderr, "juicy_batch error %d: %s\n", e.code, e.what());
        return 2;
    }
    return 0;
}
