/* TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") for the JuicySuite hot path.
 *
 * A plain-C restatement of the reference's per-sample DSP and analyzer, one
 * function per reference routine, each citing the reference file:line it
 * follows.  It is pinned against the reference's own C++ (oracle/_ref, built by
 * oracle/Makefile from /root/reference) bit-for-bit by tests/test_oracle_port.py
 * and against the committed golden vectors in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use this library.  The product never links it.
 *
 * The C interface deliberately has the same shape as oracle/ref_harness.cpp's
 * (prefix jo_ instead of ref_) so one Python driver serves both.
 */
#ifndef JUICY_ORACLE_H
#define JUICY_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
    JO_INFER = 0, JO_PUNCH = 1, JO_SATURATOR = 2, JO_WIDTH = 3,
    JO_COHERE = 4, JO_TEXTURE = 5, JO_MOTION = 6, JO_NUM_KINDS = 7
};

void* jo_create(int kind, int channels, double sampleRate, int blockSize);
void jo_destroy(void* p);
void jo_prepare(void* p, double sampleRate, int blockSize);

int jo_num_params(void* p);
const char* jo_param_id(void* p, int index);
void jo_param_range(void* p, int index, float* out3);
int jo_get_param(void* p, const char* id, float* out);
int jo_set_param(void* p, const char* id, float plainValue);
int jo_set_param_normalised(void* p, const char* id, float normalised);

int jo_num_programs(void* p);
int jo_get_program(void* p);
void jo_set_program(void* p, int index);
const char* jo_program_name(void* p, int index);

/* In-place render of one clip, planar [channels][numSamples]; 16 floats per block into history. */
long jo_process(void* p, float* audio, long numSamples, int blockSize, float* history);
void jo_latest(void* p, float* rec16);
double jo_render_clips(void* p, float* audio, long numClips, long numSamples, int blockSize,
                       double sampleRate, float* lastRecords);

/* Meter-panel statistics over a render: feed n records ([n][16], JuicinessMetrics field order)
 * to JuicyMeterPanel::setMetrics in block order; out = 40 floats: smoothed preScore, postScore, score, punch,
 * richness, clarity, width, monoSafety; (min, max, avg) of punch, richness, clarity, width, monoSafety, emphasis,
 * coherence, synesthesia, fatigue, repetition; sample count; pad.  Same layout as ref_meter_run. */
void jo_meter_run(const float* records, int n, float* out);

#ifdef __cplusplus
}
#endif
#endif
