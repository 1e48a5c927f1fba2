"""TEST INFRASTRUCTURE ONLY -- ctypes driver for oracle/_ref/libjuicy_ref_<Plugin>.so.

Each library is the reference's unmodified PluginProcessor.cpp +
JuicinessAnalyzer.cpp (/root/reference/src/...) plus oracle/ref_harness.cpp.
`RefPlugin` plays the DAW host: prepareToPlay, processBlock over 512-sample
blocks (ragged tail), getLatestMetrics after every block.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

PLUGINS = ("JuicyInfer", "JuicyPunch", "JuicySaturator", "JuicyWidth", "JuicyCohere", "JuicyTexture", "JuicyMotion")
# record layout written by ref_harness.cpp::fillRecord (16 floats per block)
RECORD_FIELDS = ("score", "preScore", "postScore", "emphasis", "coherence", "synesthesia", "fatigueRisk",
                 "repetitionDensity", "punch", "richness", "clarity", "width", "monoSafety",
                 "juiciness", "aux", "reserved")

_libs = {}


def available():
    return all(os.path.exists(os.path.join(REF_DIR, "libjuicy_ref_%s.so" % p)) for p in PLUGINS)


class _Api:
    """Resolves <prefix>_<name> symbols of one oracle library."""

    def __init__(self, lib, prefix):
        self._lib, self._prefix = lib, prefix

    def __getattr__(self, name):
        return getattr(self._lib, "%s_%s" % (self._prefix, name))


def bind(path, prefix, create_argtypes):
    """Load an oracle library and declare the harness signatures (shared by oracle.port)."""
    raw = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    lib = _Api(raw, prefix)
    lib.create.restype = ctypes.c_void_p
    lib.create.argtypes = create_argtypes
    lib.destroy.argtypes = [ctypes.c_void_p]
    lib.prepare.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
    lib.num_params.argtypes = [ctypes.c_void_p]
    lib.param_id.restype = ctypes.c_char_p
    lib.param_id.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.param_range.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    lib.get_param.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_float)]
    lib.set_param.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_float]
    lib.set_param_normalised.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_float]
    lib.num_programs.argtypes = [ctypes.c_void_p]
    lib.get_program.argtypes = [ctypes.c_void_p]
    lib.set_program.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.program_name.restype = ctypes.c_char_p
    lib.program_name.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.process.restype = ctypes.c_long
    lib.process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_void_p]
    lib.latest.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.render_clips.restype = ctypes.c_double
    lib.render_clips.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_long,
                                         ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    return lib


def _lib(plugin):
    if plugin not in _libs:
        path = os.path.join(REF_DIR, "libjuicy_ref_%s.so" % plugin)
        _libs[plugin] = bind(path, "ref", [ctypes.c_int, ctypes.c_double, ctypes.c_int])
    return _libs[plugin]


class OraclePlugin:
    """Common driver over the ref_/jo_ harness interface (one plugin instance)."""

    plugin = None
    lib = None

    def _init(self, lib, create_args, channels, sample_rate, block_size):
        self.lib = lib
        self.channels = channels
        self.sample_rate = float(sample_rate)
        self.block_size = int(block_size)
        self.h = ctypes.c_void_p(self.lib.create(*create_args))
        if not self.h:
            raise RuntimeError("oracle create failed")

    def close(self):
        if self.h:
            self.lib.destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def prepare(self, sample_rate=None, block_size=None):
        if sample_rate is not None:
            self.sample_rate = float(sample_rate)
        if block_size is not None:
            self.block_size = int(block_size)
        self.lib.prepare(self.h, self.sample_rate, self.block_size)

    def param_ids(self):
        return [self.lib.param_id(self.h, i).decode() for i in range(self.lib.num_params(self.h))]

    def param_range(self, index):
        out = (ctypes.c_float * 3)()
        self.lib.param_range(self.h, index, out)
        return tuple(out)

    def get_param(self, pid):
        v = ctypes.c_float()
        if self.lib.get_param(self.h, pid.encode(), ctypes.byref(v)) != 0:
            raise KeyError(pid)
        return v.value

    def set_param(self, pid, plain):
        if self.lib.set_param(self.h, pid.encode(), float(plain)) != 0:
            raise KeyError(pid)

    def set_param_normalised(self, pid, n):
        if self.lib.set_param_normalised(self.h, pid.encode(), float(n)) != 0:
            raise KeyError(pid)

    def params(self):
        return {p: self.get_param(p) for p in self.param_ids()}

    def num_programs(self):
        return self.lib.num_programs(self.h)

    def set_program(self, i):
        self.lib.set_program(self.h, int(i))

    def program_name(self, i):
        return self.lib.program_name(self.h, int(i)).decode()

    def process(self, audio):
        """audio: float32 [channels][n] (C-contiguous). Returns (out, history[nblocks][16])."""
        a = np.ascontiguousarray(audio, dtype=np.float32).copy()
        assert a.ndim == 2 and a.shape[0] == self.channels
        n = a.shape[1]
        nblocks = (n + self.block_size - 1) // self.block_size
        hist = np.zeros((max(nblocks, 1), 16), dtype=np.float32)
        got = self.lib.process(self.h, a.ctypes.data, n, self.block_size, hist.ctypes.data)
        return a, hist[:got]

    def render_clips(self, clips):
        """clips: float32 [n_clips][channels][n]; each clip gets a freshly prepared instance state.
        Returns (out, last_records[n_clips][16], seconds inside processBlock loops)."""
        a = np.ascontiguousarray(clips, dtype=np.float32).copy()
        assert a.ndim == 3 and a.shape[1] == self.channels
        rec = np.zeros((a.shape[0], 16), dtype=np.float32)
        secs = self.lib.render_clips(self.h, a.ctypes.data, a.shape[0], a.shape[2], self.block_size,
                                         self.sample_rate, rec.ctypes.data)
        return a, rec, secs


class RefPlugin(OraclePlugin):
    """One instance of a reference plugin class behind the headless harness."""

    def __init__(self, plugin, channels=2, sample_rate=48000.0, block_size=512):
        if plugin not in PLUGINS:
            raise ValueError("unknown plugin %r" % (plugin,))
        self.plugin = plugin
        self._init(_lib(plugin), (channels, float(sample_rate), int(block_size)), channels, sample_rate, block_size)


def run_chain(chain, audio, channels=2, sample_rate=48000.0, block_size=512, programs=None, params=None, cls=None):
    """Push one clip [channels][n] through `chain` (list of plugin names), plugin by plugin.
    Equivalent to the per-block chain because every plugin is causal and sees identical blocking.
    Returns (out, [history per plugin]).  programs: {slot: idx}; params: {slot: {id: plain}}."""
    x = np.ascontiguousarray(audio, dtype=np.float32)
    hists = []
    for slot, name in enumerate(chain):
        p = (cls or RefPlugin)(name, channels, sample_rate, block_size)
        if programs and slot in programs:
            p.set_program(programs[slot])
        if params and slot in params:
            for k, v in params[slot].items():
                p.set_param(k, v)
        p.prepare()
        x, h = p.process(x)
        hists.append(h)
        p.close()
    return x, hists


def meter_run(records):
    """The reference's own JuicyMeterPanel (oracle/_ref/libjuicy_ref_MeterPanel.so) fed [n][16] records."""
    rec = np.ascontiguousarray(records, dtype=np.float32)
    out = np.zeros(40, dtype=np.float32)
    fn = ctypes.CDLL(os.path.join(REF_DIR, "libjuicy_ref_MeterPanel.so")).ref_meter_run
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    fn.restype = None
    fn(rec.ctypes.data, int(rec.shape[0]), out.ctypes.data)
    return out


def meter_available():
    return os.path.exists(os.path.join(REF_DIR, "libjuicy_ref_MeterPanel.so"))
