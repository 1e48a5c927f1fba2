// TEST INFRASTRUCTURE ONLY (oracle).
//
// Headless driver for the reference's JuicyMeterPanel (src/shared/JuicyMeterPanel.{h,cpp}, compiled
// unmodified into oracle/_ref/libjuicy_ref_MeterPanel.so): feeds it one JuicinessMetrics record per
// host block -- what the editor's timer hands to setMetrics() -- and reads back the smoothed values
// and the min / max / average "ghost" statistics.  The panel keeps them private
// and has no accessor (they only reach the screen); the access override below touches nothing but
// the reference's own class definition and changes neither layout nor mangling.
#include <juce_gui_basics/juce_gui_basics.h>
#define private public
#include "JuicyMeterPanel.h"
#undef private

extern "C" {

// records: [n][16] floats in JuicinessMetrics field order (score, preScore, postScore, emphasis, coherence,
// synesthesia, fatigueRisk, repetitionDensity, punch, richness, clarity, width, monoSafety, ...).
// out: 40 floats: smoothed preScore, postScore, score, punch, richness, clarity, width, monoSafety; then
// (min, max, avg) of punch, richness, clarity, width, monoSafety, emphasis, coherence, synesthesia, fatigue,
// repetition; then the sample count; one pad.
void ref_meter_run(const float* records, int n, float* out)
{
    JuicyMeterPanel panel;
    for (int i = 0; i < n; ++i) {
        const float* r = records + 16 * i;
        JuicinessMetrics m;
        m.score = r[0]; m.preScore = r[1]; m.postScore = r[2];
        m.emphasis = r[3]; m.coherence = r[4]; m.synesthesia = r[5]; m.fatigueRisk = r[6]; m.repetitionDensity = r[7];
        m.punch = r[8]; m.richness = r[9]; m.clarity = r[10]; m.width = r[11]; m.monoSafety = r[12];
        panel.setMetrics(m);
    }
    const JuicinessMetrics& s = panel.metrics;
    out[0] = s.preScore; out[1] = s.postScore; out[2] = s.score;
    out[3] = s.punch; out[4] = s.richness; out[5] = s.clarity; out[6] = s.width; out[7] = s.monoSafety;
    const JuicyMeterPanel::MetricStats* stats[10] = { &panel.punchStats, &panel.richnessStats, &panel.clarityStats, &panel.widthStats,
                                                      &panel.monoSafetyStats, &panel.emphasisStats, &panel.coherenceStats,
                                                      &panel.synesthesiaStats, &panel.fatigueStats, &panel.repetitionStats };
    for (int k = 0; k < 10; ++k) {
        out[8 + 3 * k] = stats[k]->min;
        out[9 + 3 * k] = stats[k]->max;
        out[10 + 3 * k] = stats[k]->avg;
    }
    out[38] = (float) panel.punchStats.count;
    out[39] = 0.0f;
}

} // extern "C"
