"""ctypes binding of libjuicy_batch.so (include/juicy_batch.h) plus a thin host-side
mirror of the reference's processor API for batches of clips.

`BatchProcessor` keeps the names of juce::AudioProcessor / APVTS as the reference
plugins use them (prepareToPlay, processBlock, getRawParameterValue,
setCurrentProgram, getLatestMetrics; e.g. /root/reference/src/plugins/JuicyPunch/
PluginProcessor.h:9-59) over N identical instances.  There is NO CPU fallback: if
the shared library is missing the import fails, and without a CUDA device every
compute call raises JuicyBatchError.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JUICY_BATCH_LIB") or os.path.join(HERE, "libjuicy_batch.so")  # override: kernel-variant experiments

JB_OK = 0
KINDS = ("JuicyInfer", "JuicyPunch", "JuicySaturator", "JuicyWidth", "JuicyCohere", "JuicyTexture", "JuicyMotion")
KIND_INDEX = {name: i for i, name in enumerate(KINDS)}
RECORD_FIELDS = ("score", "preScore", "postScore", "emphasis", "coherence", "synesthesia", "fatigueRisk",
                 "repetitionDensity", "punch", "richness", "clarity", "width", "monoSafety",
                 "juiciness", "aux", "reserved")
SYNTH_KINDS = {"sweep": 0, "noise": 1, "impulse": 2, "drum": 3, "mixed": 4}


class JuicyBatchError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("juicy_batch error %d: %s" % (code, message))
        self.code = code


class ParamInfo(ctypes.Structure):
    _fields_ = [("id", ctypes.c_char_p), ("name", ctypes.c_char_p), ("min_value", ctypes.c_float),
                ("max_value", ctypes.c_float), ("interval", ctypes.c_float), ("default_value", ctypes.c_float),
                ("is_output", ctypes.c_int)]


_lib = None


def lib():
    """Load the C-ABI library.  Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make -C %s` (python __graft_entry__.py build); "
                          "there is no CPU fallback" % (LIB_PATH, HERE))
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cf, cd, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_longlong
    cs = ctypes.c_char_p
    sig = {
        "jb_last_error": (cs, []),
        "jb_abi_version": (ci, []),
        "jb_device_count": (ci, []),
        "jb_create": (ci, [ctypes.POINTER(ci), ci, ci, ci, ci, ctypes.POINTER(vp)]),
        "jb_destroy": (ci, [vp]),
        "jb_chain_length": (ci, [vp]),
        "jb_chain_kind": (ci, [vp, ci]),
        "jb_num_clips": (ci, [vp]),
        "jb_prepare": (ci, [vp, cd, ci]),
        "jb_reset": (ci, [vp]),
        "jb_num_params": (ci, [vp, ci]),
        "jb_param_info_at": (ci, [vp, ci, ci, ctypes.POINTER(ParamInfo)]),
        "jb_get_param": (ci, [vp, ci, cs, ctypes.POINTER(cf)]),
        "jb_set_param": (ci, [vp, ci, cs, cf]),
        "jb_set_param_normalised": (ci, [vp, ci, cs, cf]),
        "jb_set_param_clips": (ci, [vp, ci, ctypes.c_char_p, ctypes.c_float, ci, ci]),
        "jb_set_program_clips": (ci, [vp, ci, ci, ci, ci]),
        "jb_get_param_clip": (ci, [vp, ci, ctypes.c_char_p, ci, ctypes.POINTER(ctypes.c_float)]),
        "jb_num_param_sets": (ci, [vp]),
        "jb_schedule_param": (ci, [vp, ci, ctypes.c_char_p, cll, ctypes.c_float, ci, ci]),
        "jb_clear_schedule": (ci, [vp]),
        "jb_get_state": (ci, [vp, ci, ci, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
        "jb_set_state": (ci, [vp, ci, vp, ctypes.c_size_t, ci, ci]),
        "jb_wav_last_error": (ctypes.c_char_p, []),
        "jb_wav_info_read": (ci, [ctypes.c_char_p, vp]),
        "jb_wav_read": (ci, [ctypes.c_char_p, vp, ci, ci]),
        "jb_wav_write": (ci, [ctypes.c_char_p, vp, ci, ci, cd, ci, ci]),
        "jb_num_programs": (ci, [vp, ci]),
        "jb_get_program": (ci, [vp, ci]),
        "jb_set_program": (ci, [vp, ci, ci]),
        "jb_program_name": (cs, [vp, ci, ci]),
        "jb_process": (ci, [vp, vp, vp, ci]),
        "jb_process_host": (ci, [vp, vp, vp, ci]),
        "jb_process_host_pcm16": (ci, [vp, vp, vp, ci]),
        "jb_synchronize": (ci, [vp]),
        "jb_set_stream": (ci, [vp, vp]),
        "jb_get_metrics": (ci, [vp, ci, vp]),
        "jb_metrics_device": (ci, [vp, ci, ctypes.POINTER(vp)]),
        "jb_enable_history": (ci, [vp, ci]),
        "jb_history_blocks": (ci, [vp]),
        "jb_get_history": (ci, [vp, ci, ci, ci, vp]),
        "jb_meter_statistics": (ci, [vp, ci, ci, ci, ci, vp]),
        "jb_synth_fill": (ci, [vp, ci, cll, ci, ci, ci, cd, ctypes.c_uint, ci, vp]),
        "jb_synth_fill_host": (ci, [vp, ci, cll, ci, ci, ci, cd, ctypes.c_uint]),
        "jb_launch_count": (cll, []),
        "jb_host_alloc": (ci, [ctypes.c_size_t, ctypes.POINTER(vp)]),
        "jb_host_free": (ci, [vp]),
        "jb_device_alloc": (ci, [ci, ctypes.c_size_t, ctypes.POINTER(vp)]),
        "jb_device_free": (ci, [ci, vp]),
        "jb_copy_to_device": (ci, [ci, vp, vp, ctypes.c_size_t]),
        "jb_copy_to_host": (ci, [ci, vp, vp, ctypes.c_size_t]),
        "jb_kernel_time_ms": (ci, [vp, ctypes.POINTER(cd), ctypes.POINTER(cll)]),
        "jb_plan_slices": (ci, [ci, ci, ci, ctypes.POINTER(ci), ci]),
        "jb_enable_slot_timing": (ci, [vp, ci]),
        "jb_slot_time_ms": (ci, [vp, ci, ctypes.POINTER(cd), ctypes.POINTER(cll)]),
        "jb_set_path": (ci, [vp, ci]),
        "jb_set_math_mode": (ci, [vp, ci]),
        "jb_path_launches": (ci, [vp, ctypes.POINTER(cll), ctypes.POINTER(cll)]),
        "jb_shard_range": (ci, [cll, ci, ci, ctypes.POINTER(cll), ctypes.POINTER(cll)]),
        "jb_record_pitch": (cll, [vp]),
        "jb_comm_version": (ci, [ctypes.POINTER(ci)]),
        "jb_comm_unique_id": (ci, [vp]),
        "jb_comm_init_rank": (ci, [vp, vp, ci, ci]),
        "jb_comm_init_all": (ci, [ctypes.POINTER(vp), ci]),
        "jb_comm_destroy": (ci, [vp]),
        "jb_comm_size": (ci, [vp]),
        "jb_comm_rank": (ci, [vp]),
        "jb_gather_records": (ci, [vp, ci, vp]),
        "jb_gather_records_all": (ci, [ctypes.POINTER(vp), ci, ci, ctypes.POINTER(vp)]),
        "jb_gather_records_host": (ci, [vp, ci, vp, ci]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def exported_symbols():
    """Names declared in include/juicy_batch.h (used by the CPU-side load test)."""
    import re
    header = os.path.join(HERE, "..", "include", "juicy_batch.h")
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jb_[a-z_0-9]+)\s*\(", text)))


def _check(rc):
    if rc != JB_OK:
        raise JuicyBatchError(rc, lib().jb_last_error().decode("utf-8", "replace"))
    return rc


def device_count():
    return lib().jb_device_count()


def launch_count():
    return lib().jb_launch_count()


def _kind(k):
    if isinstance(k, str):
        if k not in KIND_INDEX:
            raise ValueError("unknown plugin %r" % (k,))
        return KIND_INDEX[k]
    return int(k)


class PinnedBuffer:
    """Page-locked host memory (cudaHostAlloc) exposed as a numpy float32 array."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.shape))
        self.ptr = ctypes.c_void_p()
        _check(lib().jb_host_alloc(n * 4, ctypes.byref(self.ptr)))
        buf = (ctypes.c_float * n).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float32).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().jb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceBuffer:
    """Raw device allocation on one GPU (cudaMalloc); the engine itself never needs torch."""

    def __init__(self, nbytes, device=0):
        self.device = device
        self.nbytes = int(nbytes)
        self.ptr = ctypes.c_void_p()
        _check(lib().jb_device_alloc(device, self.nbytes, ctypes.byref(self.ptr)))

    def upload(self, array):
        a = np.ascontiguousarray(array)
        assert a.nbytes <= self.nbytes
        _check(lib().jb_copy_to_device(self.device, self.ptr, a.ctypes.data, a.nbytes))

    def download(self, shape, dtype=np.float32):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        _check(lib().jb_copy_to_host(self.device, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().jb_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BatchProcessor:
    """N identical instances of one reference plugin -- or of a chain of them -- on one GPU.

    chain: a plugin name ("JuicyPunch") or a list of names applied in order.
    device: CUDA ordinal, or -1 for a parameter-only engine (no compute possible).
    """

    def __init__(self, chain, n_clips, n_channels=2, device=0):
        if isinstance(chain, (str, int)):
            chain = [chain]
        kinds = (ctypes.c_int * len(chain))(*[_kind(k) for k in chain])
        self._h = ctypes.c_void_p()
        _check(lib().jb_create(kinds, len(chain), int(n_clips), int(n_channels), int(device), ctypes.byref(self._h)))
        self.chain = [KINDS[k] for k in kinds]
        self.n_clips = int(n_clips)
        self.n_channels = int(n_channels)
        self.device = int(device)
        self.sample_rate = None
        self.block_size = None

    # ---- lifecycle (AudioProcessor)
    def close(self):
        if getattr(self, "_h", None):
            lib().jb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def prepareToPlay(self, sample_rate, samples_per_block):
        _check(lib().jb_prepare(self._h, float(sample_rate), int(samples_per_block)))
        self.sample_rate, self.block_size = float(sample_rate), int(samples_per_block)

    def reset(self):
        _check(lib().jb_reset(self._h))

    def set_stream(self, cuda_stream):
        _check(lib().jb_set_stream(self._h, ctypes.c_void_p(int(cuda_stream))))

    def synchronize(self):
        _check(lib().jb_synchronize(self._h))

    # ---- parameters (AudioProcessorValueTreeState)
    def slot(self, plugin):
        if isinstance(plugin, int):
            return plugin
        return self.chain.index(plugin)

    def parameterInfo(self, slot=0):
        slot = self.slot(slot)
        n = lib().jb_num_params(self._h, slot)
        out = []
        for i in range(n):
            info = ParamInfo()
            _check(lib().jb_param_info_at(self._h, slot, i, ctypes.byref(info)))
            out.append({"id": info.id.decode(), "name": info.name.decode(), "min": info.min_value, "max": info.max_value,
                        "interval": info.interval, "default": info.default_value, "is_output": bool(info.is_output)})
        return out

    def getRawParameterValue(self, pid, slot=0):
        v = ctypes.c_float()
        _check(lib().jb_get_param(self._h, self.slot(slot), pid.encode(), ctypes.byref(v)))
        return v.value

    def setParameter(self, pid, plain_value, slot=0):
        _check(lib().jb_set_param(self._h, self.slot(slot), pid.encode(), float(plain_value)))

    def setValueNotifyingHost(self, pid, normalised, slot=0):
        _check(lib().jb_set_param_normalised(self._h, self.slot(slot), pid.encode(), float(normalised)))

    # ---- per-clip parameters and per-block automation
    def setParameterClips(self, pid, plain_value, first_clip, n_clips, slot=0):
        """Give clips [first_clip, first_clip + n_clips) their own value of one parameter."""
        _check(lib().jb_set_param_clips(self._h, self.slot(slot), pid.encode(), float(plain_value), int(first_clip), int(n_clips)))

    def setCurrentProgramClips(self, index, first_clip, n_clips, slot=0):
        _check(lib().jb_set_program_clips(self._h, self.slot(slot), int(index), int(first_clip), int(n_clips)))

    def getParameterClip(self, pid, clip, slot=0):
        v = ctypes.c_float()
        _check(lib().jb_get_param_clip(self._h, self.slot(slot), pid.encode(), int(clip), ctypes.byref(v)))
        return v.value

    def numParameterSets(self):
        return lib().jb_num_param_sets(self._h)

    def scheduleParameter(self, pid, at_block, plain_value, slot=0, first_clip=-1, n_clips=0):
        """The change takes effect at the top of absolute block `at_block` (host blocks since prepare/reset)."""
        _check(lib().jb_schedule_param(self._h, self.slot(slot), pid.encode(), int(at_block), float(plain_value),
                                       int(first_clip), int(n_clips)))

    def clearSchedule(self):
        _check(lib().jb_clear_schedule(self._h))

    # ---- state blobs (get/setStateInformation)
    def getStateInformation(self, slot=0, clip=-1):
        size = ctypes.c_size_t()
        _check(lib().jb_get_state(self._h, self.slot(slot), int(clip), None, 0, ctypes.byref(size)))
        buf = ctypes.create_string_buffer(size.value)
        _check(lib().jb_get_state(self._h, self.slot(slot), int(clip), buf, size.value, ctypes.byref(size)))
        return buf.raw[:size.value]

    def setStateInformation(self, blob, slot=0, first_clip=-1, n_clips=0):
        data = bytes(blob)
        _check(lib().jb_set_state(self._h, self.slot(slot), data, len(data), int(first_clip), int(n_clips)))

    def getNumPrograms(self, slot=0):
        return lib().jb_num_programs(self._h, self.slot(slot))

    def getCurrentProgram(self, slot=0):
        return lib().jb_get_program(self._h, self.slot(slot))

    def setCurrentProgram(self, index, slot=0):
        _check(lib().jb_set_program(self._h, self.slot(slot), int(index)))

    def getProgramName(self, index, slot=0):
        return lib().jb_program_name(self._h, self.slot(slot), int(index)).decode()

    # ---- audio
    def processBlock(self, audio):
        """Render host audio [n_clips][n_channels][n] (float32) through the chain, walking
        samples_per_block blocks like consecutive host callbacks.  Returns a new array."""
        a = np.ascontiguousarray(audio, dtype=np.float32)
        assert a.shape[:2] == (self.n_clips, self.n_channels), a.shape
        out = np.empty_like(a)
        _check(lib().jb_process_host(self._h, a.ctypes.data, out.ctypes.data, a.shape[2]))
        return out

    def processBlockPcm16(self, audio):
        """Host audio as 16-bit PCM [n_clips][n_channels][n] (int16): converted on the device on the way in and out."""
        a = np.ascontiguousarray(audio, dtype=np.int16)
        assert a.shape[:2] == (self.n_clips, self.n_channels), a.shape
        out = np.empty_like(a)
        _check(lib().jb_process_host_pcm16(self._h, a.ctypes.data, out.ctypes.data, a.shape[2]))
        return out

    def process_host_pcm16_ptr(self, in_ptr, out_ptr, n_samples):
        _check(lib().jb_process_host_pcm16(self._h, ctypes.c_void_p(int(in_ptr)), ctypes.c_void_p(int(out_ptr)), int(n_samples)))

    def process_host_ptr(self, in_ptr, out_ptr, n_samples):
        _check(lib().jb_process_host(self._h, ctypes.c_void_p(int(in_ptr)), ctypes.c_void_p(int(out_ptr)), int(n_samples)))

    def process_device(self, d_in, d_out, n_samples):
        """Device-resident render (asynchronous on the engine's stream)."""
        _check(lib().jb_process(self._h, ctypes.c_void_p(int(d_in)), ctypes.c_void_p(int(d_out)), int(n_samples)))

    # ---- outputs
    def getLatestMetrics(self, slot=0):
        """[n_clips][16] float32 records in RECORD_FIELDS order (after the most recent block)."""
        out = np.zeros((self.n_clips, 16), dtype=np.float32)
        _check(lib().jb_get_metrics(self._h, self.slot(slot), out.ctypes.data))
        return out

    def metrics_device_ptr(self, slot=0):
        p = ctypes.c_void_p()
        _check(lib().jb_metrics_device(self._h, self.slot(slot), ctypes.byref(p)))
        return p.value

    def enableHistory(self, max_blocks):
        _check(lib().jb_enable_history(self._h, int(max_blocks)))

    def historyBlocks(self):
        return lib().jb_history_blocks(self._h)

    def getHistory(self, slot=0, first_block=0, n_blocks=None):
        """[n_blocks][n_clips][16] per-block records since the last prepare/reset."""
        if n_blocks is None:
            n_blocks = self.historyBlocks() - first_block
        out = np.zeros((n_blocks, self.n_clips, 16), dtype=np.float32)
        if n_blocks > 0:
            _check(lib().jb_get_history(self._h, self.slot(slot), int(first_block), int(n_blocks), out.ctypes.data))
        return out

    def meterStatistics(self, slot=0, first_block=0, n_blocks=None, block_stride=1):
        """[n_clips][40] JuicyMeterPanel state after the render (jb_meter_stats field order): 8 smoothed bar
        values, (min, max, avg) of the 10 tracked metrics, records fed."""
        if n_blocks is None:
            n_blocks = self.historyBlocks() - first_block
        out = np.zeros((self.n_clips, 40), dtype=np.float32)
        _check(lib().jb_meter_statistics(self._h, self.slot(slot), int(first_block), int(n_blocks), int(block_stride),
                                         out.ctypes.data))
        return out

    def set_path(self, mode):
        """Render-kernel choice: "auto", "lane" (one lane per clip) or "coop" (block-cooperative)."""
        _check(lib().jb_set_path(self._h, {"auto": 0, "lane": 1, "coop": 2}[mode]))

    def set_math_mode(self, mode):
        """Saturator / Punch tanh and pow: "auto", "exact" (glibc's algorithms, bit-identical) or "fast" (MUFU-based)."""
        _check(lib().jb_set_math_mode(self._h, {"auto": 0, "exact": 1, "fast": 2}[mode]))

    def path_launches(self):
        """(cooperative, lane_per_clip) render-kernel launches so far."""
        c, l = ctypes.c_longlong(), ctypes.c_longlong()
        _check(lib().jb_path_launches(self._h, ctypes.byref(c), ctypes.byref(l)))
        return c.value, l.value

    # ---- score gather across GPUs (jb_comm_* / jb_gather_records)
    def comm_init_rank(self, unique_id, n_ranks, rank):
        _check(lib().jb_comm_init_rank(self._h, ctypes.c_char_p(bytes(unique_id)), int(n_ranks), int(rank)))

    def record_pitch(self):
        return lib().jb_record_pitch(self._h)

    def gather_records_host(self, slot=0, clips_per_rank=None):
        """[n_ranks * clips_per_rank][16]: every rank's latest records in global clip order (ncclAllGather + download)."""
        n = lib().jb_comm_size(self._h)
        cpr = self.n_clips if clips_per_rank is None else int(clips_per_rank)
        out = np.zeros((n * cpr, 16), dtype=np.float32)
        _check(lib().jb_gather_records_host(self._h, self.slot(slot), out.ctypes.data, cpr))
        return out

    def kernel_time_ms(self):
        """(milliseconds, launches) spent in the render kernel since the last call (CUDA events
        recorded on the engine's stream around every launch)."""
        ms = ctypes.c_double()
        n = ctypes.c_longlong()
        _check(lib().jb_kernel_time_ms(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def enable_slot_timing(self, on=True):
        """Per-plugin timing of a chain rendered as one launch per plugin (jb_enable_slot_timing)."""
        _check(lib().jb_enable_slot_timing(self._h, 1 if on else 0))

    def slot_times_ms(self):
        """[(milliseconds, launches)] per chain slot since the last call (jb_slot_time_ms)."""
        out = []
        for s in range(len(self.chain)):
            ms = ctypes.c_double()
            n = ctypes.c_longlong()
            _check(lib().jb_slot_time_ms(self._h, s, ctypes.byref(ms), ctypes.byref(n)))
            out.append((ms.value, n.value))
        return out


def plan_slices(total_blocks, slice_blocks, taper=True):
    """First host block of every time slice of a jb_process_host render, then total_blocks (jb_plan_slices; no device needed)."""
    cap = int(total_blocks) + 2
    buf = (ctypes.c_int * cap)()
    n = lib().jb_plan_slices(int(total_blocks), int(slice_blocks), 1 if taper else 0, buf, cap)
    if n < 0:
        _check(n)
    return list(buf[:n + 1])


def shard_range(n_clips, rank, world):
    """[lo, hi) of the clips rank `rank` of `world` renders (jb_shard_range): contiguous, balanced to within one clip."""
    first, count = ctypes.c_longlong(), ctypes.c_longlong()
    _check(lib().jb_shard_range(int(n_clips), int(rank), int(world), ctypes.byref(first), ctypes.byref(count)))
    return first.value, first.value + count.value


def comm_unique_id():
    """128 opaque bytes (ncclGetUniqueId) that rank 0 hands to the other ranks."""
    buf = ctypes.create_string_buffer(128)
    _check(lib().jb_comm_unique_id(buf))
    return buf.raw


def comm_init_all(engines):
    """One process driving one engine per GPU: a communicator over all of them (ncclCommInitAll)."""
    hs = (ctypes.c_void_p * len(engines))(*[e._h for e in engines])
    _check(lib().jb_comm_init_all(hs, len(engines)))


def synth_fill_device(d_ptr, kind, first_clip, n_clips, n_channels, n_samples, sample_rate=48000.0,
                      seed=0x4A554943, device=0, stream=0):
    k = SYNTH_KINDS[kind] if isinstance(kind, str) else int(kind)
    _check(lib().jb_synth_fill(ctypes.c_void_p(int(d_ptr)), k, int(first_clip), int(n_clips), int(n_channels), int(n_samples),
                               float(sample_rate), ctypes.c_uint(seed), int(device), ctypes.c_void_p(int(stream))))


class WavInfo(ctypes.Structure):
    _fields_ = [("n_channels", ctypes.c_int), ("n_samples", ctypes.c_int), ("sample_rate", ctypes.c_double),
                ("bits_per_sample", ctypes.c_int), ("is_float", ctypes.c_int)]


def _check_wav(rc):
    if rc != JB_OK:
        raise JuicyBatchError(rc, lib().jb_wav_last_error().decode("utf-8", "replace"))


def wav_info(path):
    info = WavInfo()
    _check_wav(lib().jb_wav_info_read(os.fsencode(path), ctypes.byref(info)))
    return {"n_channels": info.n_channels, "n_samples": info.n_samples, "sample_rate": info.sample_rate,
            "bits_per_sample": info.bits_per_sample, "is_float": bool(info.is_float)}


def wav_read(path):
    """(float32 [n_channels][n_samples], sample_rate) of a RIFF/WAVE file (PCM 16/24/32 or float32)."""
    info = wav_info(path)
    out = np.zeros((info["n_channels"], info["n_samples"]), dtype=np.float32)
    _check_wav(lib().jb_wav_read(os.fsencode(path), out.ctypes.data, info["n_channels"], info["n_samples"]))
    return out, info["sample_rate"]


def wav_write(path, audio, sample_rate=48000.0, bits_per_sample=24, is_float=False):
    a = np.ascontiguousarray(audio, dtype=np.float32)
    assert a.ndim == 2
    _check_wav(lib().jb_wav_write(os.fsencode(path), a.ctypes.data, a.shape[0], a.shape[1], float(sample_rate),
                                  int(bits_per_sample), int(bool(is_float))))


def synth_clips(kind, first_clip, n_clips, n_samples, n_channels=2, sample_rate=48000.0, seed=0x4A554943):
    """The same seeded synthetic clips generated on the host (plain C++ in the library, no GPU needed):
    float32 [n_clips][n_channels][n_samples]."""
    k = SYNTH_KINDS[kind] if isinstance(kind, str) else int(kind)
    out = np.zeros((n_clips, n_channels, n_samples), dtype=np.float32)
    _check(lib().jb_synth_fill_host(out.ctypes.data, k, int(first_clip), int(n_clips), int(n_channels), int(n_samples),
                                    float(sample_rate), ctypes.c_uint(seed)))
    return out
