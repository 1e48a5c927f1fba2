"""Writes tests/golden/param_names_v1.json: parameter ids and display names of every plugin as the compiled reference
(oracle/_ref, built from /root/reference by oracle/Makefile) reports them through createParameterLayout().
Run here, where /root/reference exists:  python tests/golden/make_param_names.py"""
import ctypes
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refhost  # noqa: E402

out = {}
for plugin in refhost.PLUGINS:
    p = refhost.RefPlugin(plugin)
    raw = p.lib._lib
    raw.ref_param_name.restype = ctypes.c_char_p
    raw.ref_param_name.argtypes = [ctypes.c_void_p, ctypes.c_int]
    out[plugin] = [[pid, raw.ref_param_name(p.h, i).decode()] for i, pid in enumerate(p.param_ids())]
    p.close()
json.dump(out, open(os.path.join(HERE, "param_names_v1.json"), "w"), indent=1)
print("wrote", len(out), "plugins")
