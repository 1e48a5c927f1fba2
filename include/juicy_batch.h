/* juicy_batch.h -- C ABI of the B200-native JuicySuite batch engine.
 *
 * One engine = N identical instances ("clips") of one plugin, or of a chain of
 * plugins, rendered together on one GPU.  The entry points are what a binding
 * of the reference's processor/parameter API would call; each cites the
 * reference interface it replaces (paths relative to /root/reference).
 * INTEGRATION.md shows the C++ and ctypes stubs a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns JB_OK (0) or a
 *     negative jb_status and records a message readable via jb_last_error();
 *   - audio is planar fp32, [clip][channel][sample], exactly N concatenated
 *     juce::AudioBuffer<float> images; "device" pointers are CUDA device
 *     pointers on the engine's GPU, "host" pointers are ordinary memory;
 *   - `slot` is the position of a plugin inside the engine's chain (0-based);
 *   - parameter values are the plain (de-normalised) floats that
 *     `*parameters.getRawParameterValue(id)` yields in the reference.
 *   - there is no CPU fallback: without a CUDA device every call that needs
 *     one fails with JB_ERR_CUDA.
 */
#ifndef JUICY_BATCH_H
#define JUICY_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_ABI_VERSION 1
#define JB_MAX_CHAIN 8
#define JB_ALL_CLIPS (-1)

typedef enum jb_status {
    JB_OK = 0,
    JB_ERR_ARG = -1,       /* bad argument (null pointer, unknown id, size mismatch) */
    JB_ERR_STATE = -2,     /* call order violated (e.g. process before prepare)      */
    JB_ERR_CUDA = -3,      /* CUDA runtime / driver error, or no device               */
    JB_ERR_UNSUPPORTED = -4
} jb_status;

/* Plugin kinds, in the order of CMakeLists.txt:63-69 (add_juicy_plugin calls). */
typedef enum jb_plugin_kind {
    JB_INFER = 0,     /* src/plugins/JuicyInfer     */
    JB_PUNCH = 1,     /* src/plugins/JuicyPunch     */
    JB_SATURATOR = 2, /* src/plugins/JuicySaturator */
    JB_WIDTH = 3,     /* src/plugins/JuicyWidth     */
    JB_COHERE = 4,    /* src/plugins/JuicyCohere    */
    JB_TEXTURE = 5,   /* src/plugins/JuicyTexture   */
    JB_MOTION = 6,    /* src/plugins/JuicyMotion    */
    JB_NUM_KINDS = 7
} jb_plugin_kind;

/* Mirrors `struct JuicinessMetrics` (src/shared/JuicinessAnalyzer.h:6-21), field
 * for field, followed by the plugin's output parameters as the host sees them:
 * `juiciness` (every plugin, e.g. JuicyPunch/PluginProcessor.cpp:56-62,123) and
 * `aux` = Cohere's `contextfit` (JuicyCohere/PluginProcessor.cpp:91-92), else 0.
 * The values are what `getLatestMetrics()` returns after a processBlock, e.g.
 * JuicyPunch/PluginProcessor.cpp:190-202; for Infer the five feature slots carry
 * the triangle metrics (JuicyInfer/PluginProcessor.cpp:164-181). */
typedef struct jb_metrics {
    float score, preScore, postScore;
    float emphasis, coherence, synesthesia, fatigueRisk, repetitionDensity;
    float punch, richness, clarity, width, monoSafety;
    float juiciness; /* de-normalised `juiciness` output parameter */
    float aux;       /* `contextfit` for JB_COHERE, otherwise 0 */
    float reserved;
} jb_metrics;

typedef struct jb_param_info {
    const char* id;   /* APVTS parameter id */
    const char* name;
    float min_value, max_value, interval, default_value;
    int is_output;    /* 1 for juiciness / contextfit / Infer's triangle outputs */
} jb_param_info;

typedef struct jb_engine jb_engine;

const char* jb_last_error(void);
int jb_abi_version(void);
/* Number of CUDA devices visible (0 without a GPU; never an error). */
int jb_device_count(void);

/* createPluginFilter() x N (e.g. JuicySaturator/PluginProcessor.cpp:201-204) and the
 * constructors' bus layout + setCurrentProgram(0) (:26-33).  `chain` holds
 * chain_len jb_plugin_kind values applied in order to every clip; n_channels is
 * the bus layout (isBusesLayoutSupported: mono or stereo in == out, e.g.
 * JuicyPunch/PluginProcessor.cpp:48-54): 1 or 2.  Mono audio is [clip][1][sample]. */
int jb_create(const int* chain, int chain_len, int n_clips, int n_channels, int device, jb_engine** out);
int jb_destroy(jb_engine* e);

int jb_chain_length(const jb_engine* e);
int jb_chain_kind(const jb_engine* e, int slot);
int jb_num_clips(const jb_engine* e);

/* AudioProcessor::prepareToPlay(sampleRate, samplesPerBlock) for every plugin of the
 * chain (e.g. JuicyMotion/PluginProcessor.cpp:12-29): sets the rate, (re)allocates
 * and clears all DSP + analyzer state.  jb_reset clears state again at the same
 * rate (a second prepareToPlay). */
int jb_prepare(jb_engine* e, double sample_rate, int samples_per_block);
int jb_reset(jb_engine* e);

/* Parameters: AudioProcessorValueTreeState "PARAMS" (ids, ranges, defaults of
 * createParameterLayout(), e.g. JuicySaturator/PluginProcessor.cpp:189-199). */
int jb_num_params(const jb_engine* e, int slot);
int jb_param_info_at(const jb_engine* e, int slot, int index, jb_param_info* out);
/* `*parameters.getRawParameterValue(id)` */
int jb_get_param(const jb_engine* e, int slot, const char* id, float* out);
/* host automation to a plain value: param->setValueNotifyingHost(range.convertTo0to1(v)) */
int jb_set_param(jb_engine* e, int slot, const char* id, float plain_value);
/* param->setValueNotifyingHost(normalised) */
int jb_set_param_normalised(jb_engine* e, int slot, const char* id, float normalised);

/* Per-clip parameters and per-block automation.  The reference re-reads its parameters at the top of
 * every processBlock (37 getRawParameterValue sites, e.g. JuicyPunch/PluginProcessor.cpp:71-78), so a host can give every
 * plugin instance its own settings and automate them between callbacks.  Here:
 *  - jb_set_param_clips / jb_set_program_clips give clips [first_clip, first_clip + n_clips) their own values (first_clip =
 *    JB_ALL_CLIPS: every clip; jb_set_param & co. always address every clip).  Clips with identical settings share a
 *    parameter set and render in one launch; jb_num_param_sets reports how many distinct sets are in use.
 *  - jb_get_param_clip reads one clip's value (jb_get_param reads parameter set 0, the engine-wide default).
 *  - jb_schedule_param makes a change take effect at the top of absolute block `at_block` (counted in host blocks since
 *    jb_prepare / jb_reset, the index jb_get_history uses): the equivalent of the host calling setValueNotifyingHost between
 *    two processBlock callbacks inside one jb_process / jb_process_host render.  Applied changes persist like any other
 *    parameter change; jb_prepare / jb_reset / jb_clear_schedule drop what is still pending. */
int jb_set_param_clips(jb_engine* e, int slot, const char* id, float plain_value, int first_clip, int n_clips);
int jb_set_program_clips(jb_engine* e, int slot, int index, int first_clip, int n_clips);
int jb_get_param_clip(const jb_engine* e, int slot, const char* id, int clip, float* out);
int jb_num_param_sets(const jb_engine* e);
int jb_schedule_param(jb_engine* e, int slot, const char* id, long long at_block, float plain_value, int first_clip,
                      int n_clips);
int jb_clear_schedule(jb_engine* e);

/* State blobs: get/setStateInformation (e.g. JuicySaturator/PluginProcessor.cpp:117-131) -- the APVTS
 * "PARAMS" tree (<PARAM id=".." value=".."/> per parameter) wrapped the way AudioProcessor::copyXmlToBinary does (magic
 * 0x21324356, text length, single-line XML, NUL), so a state saved by the DAW plugin configures the batch engine and back.
 * jb_get_state: clip = JB_ALL_CLIPS reads parameter set 0; buffer == NULL only reports the size.  jb_set_state: like
 * replaceState -- every parameter takes the blob's value (through the same range handling as host automation), one the
 * blob does not mention returns to its default; first_clip = JB_ALL_CLIPS applies to every clip.  JUCE itself is not
 * available offline: the format follows its documentation (parity unpinned). */
int jb_get_state(const jb_engine* e, int slot, int clip, void* buffer, size_t capacity, size_t* size_out);
int jb_set_state(jb_engine* e, int slot, const void* data, size_t size, int first_clip, int n_clips);

/* Programs: getNumPrograms / getCurrentProgram / setCurrentProgram / getProgramName
 * (e.g. JuicyWidth/PluginProcessor.cpp:171-210). */
int jb_num_programs(const jb_engine* e, int slot);
int jb_get_program(const jb_engine* e, int slot);
int jb_set_program(jb_engine* e, int slot, int index);
const char* jb_program_name(const jb_engine* e, int slot, int index);

/* processBlock(AudioBuffer<float>&, MidiBuffer&) for every block of every clip and
 * every plugin of the chain (e.g. JuicyPunch/PluginProcessor.cpp:64-124): walks
 * n_samples in blocks of samples_per_block (ragged last block), carrying state
 * across calls like consecutive host callbacks.  d_in/d_out: device pointers,
 * [n_clips][n_channels][n_samples] fp32; d_out may equal d_in (in place).
 * Asynchronous on the engine's stream. */
int jb_process(jb_engine* e, const float* d_in, float* d_out, int n_samples);
/* Same through HOST buffers: uploads, renders and downloads clip ranges in a
 * pipelined fashion (copies overlapped with kernels); returns when h_out is
 * complete.  h_out may equal h_in. */
int jb_process_host(jb_engine* e, const float* h_in, float* h_out, int n_samples);
/* The same with 16-bit PCM host buffers, planar [clip][channel][sample] (either side of the render is
 * host<->device streaming, and 16-bit sources need not cross PCIe as 32-bit floats).  The device converts with the rule
 * of jb_wav_read / jb_wav_write -- s / 32768 on the way in, round-half-even(v * 32768) limited to +-32767 on the way out --
 * i.e. what the host's file reader / writer does around processBlock; the render in between is the fp32 one, bit for
 * bit.  Half the bytes per direction of jb_process_host. */
int jb_process_host_pcm16(jb_engine* e, const int16_t* h_in, int16_t* h_out, int n_samples);
int jb_synchronize(jb_engine* e);
/* Run on a caller-owned cudaStream_t (e.g. the framework's current stream). */
int jb_set_stream(jb_engine* e, void* cuda_stream);

/* getLatestMetrics() of every clip after the most recent block (host buffer,
 * n_clips records).  Synchronises the engine's stream. */
int jb_get_metrics(jb_engine* e, int slot, jb_metrics* out);
/* Device-resident copy, structure-of-arrays [16][n_clips] floats in jb_metrics field
 * order (for gathering scores across GPUs without a host round trip). */
int jb_metrics_device(jb_engine* e, int slot, const float** d_out);

/* Optional per-block history (what host automation of `juiciness` sees over a
 * render): keep the record of every block for up to max_blocks blocks since the
 * last prepare/reset.  out: [n_blocks][n_clips] records. */
int jb_enable_history(jb_engine* e, int max_blocks);
int jb_history_blocks(const jb_engine* e);
int jb_get_history(jb_engine* e, int slot, int first_block, int n_blocks, jb_metrics* out);

/* Meter-panel statistics of a render: what the plugin editor's JuicyMeterPanel of every clip would
 * hold after being handed the records of blocks first_block, first_block + block_stride, ... of the
 * history in order (JuicyMeterPanel::setMetrics / smoothValue / updateStats,
 * src/shared/JuicyMeterPanel.cpp:3-34,54-71; the editor calls it from a 20 Hz timer with
 * getLatestMetrics(), src/shared/JuicyPluginEditor.cpp:36,85-89 -- block_stride 1 feeds every block).
 * Reduced on the device from the history (jb_enable_history first).  out: n_clips records. */
typedef struct jb_meter_stat { float min, max, avg; } jb_meter_stat;
typedef struct jb_meter_stats {
    float preScore, postScore, score, punch, richness, clarity, width, monoSafety; /* smoothed bar values */
    jb_meter_stat punchStats, richnessStats, clarityStats, widthStats, monoSafetyStats;
    jb_meter_stat emphasisStats, coherenceStats, synesthesiaStats, fatigueStats, repetitionStats;
    float count;    /* records fed (MetricStats::count) */
    float reserved;
} jb_meter_stats;
int jb_meter_statistics(jb_engine* e, int slot, int first_block, int n_blocks, int block_stride, jb_meter_stats* out);

/* Audio files.  The reference never touches files -- its DAW host decodes them into the fp32
 * AudioBuffers processBlock receives -- so an offline batch renderer does that step itself: RIFF/WAVE, PCM 16 / 24 / 32-bit
 * integer or 32-bit float, plain or WAVE_FORMAT_EXTENSIBLE header, to / from planar [channel][sample] fp32 host buffers
 * (one clip of jb_process_host's layout).  Integer samples convert like JUCE's WavAudioFormat: value / 2^(bits-1) on
 * reading, round(value * 2^(bits-1)) limited to +-(2^(bits-1) - 1) on writing.  Host code only (works without a GPU).
 * Errors: JB_ERR_ARG with a message in jb_wav_last_error(). */
typedef struct jb_wav_info {
    int n_channels, n_samples;
    double sample_rate;
    int bits_per_sample, is_float;
} jb_wav_info;
const char* jb_wav_last_error(void);
int jb_wav_info_read(const char* path, jb_wav_info* out);
int jb_wav_read(const char* path, float* h_planar, int n_channels, int n_samples);
int jb_wav_write(const char* path, const float* h_planar, int n_channels, int n_samples, double sample_rate,
                 int bits_per_sample, int is_float);

/* Seeded synthetic clips written straight into device memory:
 * kind 0 sweep, 1 noise, 2 impulse train, 3 drum hit, 4 mixed (clip mod 4).
 * first_clip offsets the per-clip seeds so shards of one job stay distinct. */
int jb_synth_fill(float* d_audio, int kind, long long first_clip, int n_clips, int n_channels,
                  int n_samples, double sample_rate, unsigned int seed, int device, void* cuda_stream);

/* The same clips generated on the host (plain C++; needs no GPU): the parity tests feed
 * these to both the engine and the CPU oracle.  The device generator uses the GPU's
 * own sinf/expf, so its samples agree with these to rounding, not bit for bit. */
int jb_synth_fill_host(float* h_audio, int kind, long long first_clip, int n_clips, int n_channels,
                       int n_samples, double sample_rate, unsigned int seed);

/* Render-kernel choice.  The library has two sm_100a kernels for processBlock: one lane per clip
 * (any chain; fills the GPU from ~75k clips up) and a block-cooperative, time-parallel one (chains of
 * Punch-first / Width / Infer, host block <= 512, 16-byte aligned audio) for smaller batches.
 * mode JB_PATH_AUTO picks per call; JB_PATH_LANE / JB_PATH_COOP force one (COOP fails with
 * JB_ERR_UNSUPPORTED when the chain or call shape is outside its scope).  Both produce the same
 * samples; their records agree to rounding (the cooperative one tree-reduces analyze()'s plain sums). */
#define JB_PATH_AUTO 0
#define JB_PATH_LANE 1
#define JB_PATH_COOP 2
int jb_set_path(jb_engine* e, int mode);
/* Per-sample std::tanh / std::pow of Saturator and Punch (JuicySaturator/PluginProcessor.cpp:92, JuicyPunch/
 * PluginProcessor.cpp:100,106).  JB_MATH_FAST: special-function-unit based, within 3e-6 relative of the C library (the
 * stated sample tolerance for the plugin itself).  JB_MATH_EXACT: the C library's own algorithms restated (fdlibm tanhf,
 * glibc powf), bit-identical to the reference built against glibc 2.28 - 2.39, about 1.5x the arithmetic.  JB_MATH_AUTO
 * (default): see DESIGN.md §4.4 for the rule -- exact wherever a later plugin of the chain would amplify or threshold
 * the shaper's last-bit differences, fast otherwise.  Mono and stereo buses honour the mode alike.  The cooperative
 * kernel only runs in fast mode. */
#define JB_MATH_AUTO 0
#define JB_MATH_EXACT 1
#define JB_MATH_FAST 2
int jb_set_math_mode(jb_engine* e, int mode);
/* Launches of each render kernel by this engine so far (either pointer may be null). */
int jb_path_launches(const jb_engine* e, long long* cooperative, long long* lane_per_clip);

/* ---- Sharding across the GPUs of one box and the score gather ------------------------------------
 * The reference has no counterpart: a DAW runs one plugin instance per track and instances share no state
 * (all state is instance members, e.g. JuicyTexture/PluginProcessor.h:79-81).  That independence is the sharding
 * axis: rank r of N owns the contiguous clip range jb_shard_range gives it, renders it with its own engine, and
 * nothing is exchanged while rendering.  The ONE collective is an ncclAllGather of the per-clip records
 * (`getLatestMetrics()` of every instance, 16 floats per clip) after the render, over NVLink / NVSwitch.
 * NCCL is loaded at run time (dlopen "libnccl.so.2"; JB_NCCL_LIB overrides): hosts that never shard never need it.
 *
 *   multi-process (one process per GPU):  rank 0 calls jb_comm_unique_id and hands the 128 bytes to the others
 *       (any channel the launcher has); every rank calls jb_comm_init_rank(engine, id, N, rank).
 *   single process driving N engines on N distinct GPUs:  jb_comm_init_all(engines, N).
 */
#define JB_COMM_ID_BYTES 128
/* [first, first + count) of the clips rank `rank` of `world` renders: contiguous, balanced to within one clip. */
int jb_shard_range(long long n_clips, int rank, int world, long long* first, long long* count);
/* Clips per field in the engine's structure-of-arrays record block (n_clips rounded up to 32). */
long long jb_record_pitch(const jb_engine* e);
int jb_comm_version(int* nccl_version);
int jb_comm_unique_id(void* id_out /* JB_COMM_ID_BYTES */);
int jb_comm_init_rank(jb_engine* e, const void* id, int n_ranks, int rank);
int jb_comm_init_all(jb_engine* const* engines, int n_engines);
int jb_comm_destroy(jb_engine* e);
int jb_comm_size(const jb_engine* e);
int jb_comm_rank(const jb_engine* e);
/* ncclAllGather of plugin `slot`'s latest records on the engine's stream (asynchronous): every rank's block
 * [16][jb_record_pitch] (what jb_metrics_device points at) -> d_out[rank][16][pitch] on every rank.  All engines of a
 * communicator must hold the same number of clips. */
int jb_gather_records(jb_engine* e, int slot, float* d_out);
/* The same for a single process that drives all engines of the communicator (grouped NCCL calls). */
int jb_gather_records_all(jb_engine* const* engines, int n_engines, int slot, float* const* d_out);
/* Gather + download + unpack: out[rank * clips_per_rank + clip], jb_metrics records in global clip order. */
int jb_gather_records_host(jb_engine* e, int slot, jb_metrics* out, int clips_per_rank);

/* Number of kernel launches issued by this library since load (bench evidence). */
long long jb_launch_count(void);
/* Device time spent in the render kernel since the previous call of this function:
 * the sum over launches of (stop - start) CUDA events recorded on the engine's stream
 * around every launch, and the number of launches summed.  Synchronises the stream. */
int jb_kernel_time_ms(jb_engine* e, double* ms, long long* launches);
/* Per-plugin device time of a chain that renders as one launch per plugin (the lane-per-clip kernels on batches too
 * large for the plugin pipeline): while enabled, an event is recorded on the render's stream between the plugins'
 * launches.  jb_slot_time_ms returns, for chain slot `slot`, the time and the number of launches accumulated since
 * its previous call for that slot (0 launches where the render took another path: cooperative kernel, pipelined
 * chain, a single plugin -- jb_kernel_time_ms covers those).  Both synchronise the stream.  No counterpart in the
 * reference; measurement only (bench.py's per-kernel roofline). */
/* The time slices jb_process_host cuts a render of `total_blocks` host blocks into when a slice holds `slice_blocks` of them
 * (the last slices halve down to one block when `taper`): writes the first block of every slice, then total_blocks, into
 * first[0 .. capacity) and returns the number of slices (or a negative jb_status).  Needs no device; for tests and tools. */
int jb_plan_slices(int total_blocks, int slice_blocks, int taper, int* first, int capacity);
int jb_enable_slot_timing(jb_engine* e, int on);
int jb_slot_time_ms(jb_engine* e, int slot, double* ms, long long* launches);

/* Plumbing for callers without a CUDA runtime of their own (the ctypes tests, the C++
 * demo): page-locked host memory and raw device memory on `device`. */
int jb_host_alloc(size_t bytes, void** out);
int jb_host_free(void* p);
int jb_device_alloc(int device, size_t bytes, void** out);
int jb_device_free(int device, void* p);
int jb_copy_to_device(int device, void* d_dst, const void* h_src, size_t bytes);
int jb_copy_to_host(int device, void* h_dst, const void* d_src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* JUICY_BATCH_H */
