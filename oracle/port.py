"""TEST INFRASTRUCTURE ONLY -- ctypes driver for oracle/liboracle_port.so (juicy_oracle.c),
the repo's own C restatement of the reference algorithms.  Same interface as oracle.refhost."""
import ctypes
import os
import subprocess

from . import refhost

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_port.so")
KIND = {name: i for i, name in enumerate(refhost.PLUGINS)}  # Infer=0 ... Motion=6 (juicy_oracle.h)

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = refhost.bind(LIB_PATH, "jo", [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int])
    return _lib


class PortPlugin(refhost.OraclePlugin):
    def __init__(self, plugin, channels=2, sample_rate=48000.0, block_size=512):
        if plugin not in KIND:
            raise ValueError("unknown plugin %r" % (plugin,))
        self.plugin = plugin
        self._init(lib(), (KIND[plugin], channels, float(sample_rate), int(block_size)), channels, sample_rate, block_size)


def run_chain(chain, audio, **kw):
    return refhost.run_chain(chain, audio, cls=PortPlugin, **kw)


def meter_run(records):
    """JuicyMeterPanel statistics (40 floats) after feeding [n][16] records in block order (jo_meter_run)."""
    import numpy as np
    rec = np.ascontiguousarray(records, dtype=np.float32)
    out = np.zeros(40, dtype=np.float32)
    fn = ctypes.CDLL(LIB_PATH).jo_meter_run
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    fn.restype = None
    lib()
    fn(rec.ctypes.data, int(rec.shape[0]), out.ctypes.data)
    return out
