// TEST INFRASTRUCTURE ONLY (oracle).  The reference includes juce_dsp
// (/root/reference/src/shared/JuicinessAnalyzer.h:4) but uses nothing from it.
#pragma once
#include "../juce_audio_basics/juce_audio_basics.h"
