"""The opt-in data paths of the lane kernels (read once per process from the environment, hence a subprocess each):
JB_TMA=1 -- TMA tile streaming (cp.async.bulk.tensor.3d, swizzled tiles) -- and JB_LDG256=1 -- 32-byte register loads -- must
render the bits of the default cp.async rings (same arithmetic, another way of fetching the samples)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.join(%(root)r, "tests"))
from conftest import load_juicy_batch
import numpy as np
jb = load_juicy_batch()
out = {}
for chain, n_clips in ((["JuicySaturator"], 32768), (["JuicyCohere"], 32768), (["JuicyWidth"], 32768), (["JuicyInfer"], 32768),
                       (["JuicySaturator", "JuicyWidth", "JuicyCohere", "JuicyInfer"], 32768)):
    n = 2 * 512 + 72          # ragged last block, not a multiple of the 16-sample tile
    d = jb.DeviceBuffer(n_clips * 2 * n * 4)
    jb.synth_fill_device(d.ptr.value, "mixed", 0, n_clips, 2, n)
    eng = jb.BatchProcessor(chain, n_clips)
    eng.set_math_mode("fast")
    eng.prepareToPlay(48000.0, 512)
    eng.process_device(d.ptr.value, d.ptr.value, n)
    eng.synchronize()
    audio = d.download((n_clips, 2, n))
    rec = eng.getLatestMetrics(len(chain) - 1)
    out["+".join(chain)] = [hashlib.sha256(audio.tobytes()).hexdigest(), hashlib.sha256(rec.tobytes()).hexdigest()]
    eng.close()
    d.free()
print(json.dumps(out))
"""


def _run(env):
    e = dict(os.environ)
    e.pop("JB_TMA", None)
    e.pop("JB_LDG256", None)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_tma_and_register_load_paths_render_the_default_bits():
    base = _run({})
    assert _run({"JB_TMA": "1"}) == base
    assert _run({"JB_LDG256": "1"}) == base
