"""The C++ host mirror (host/JuicyBatchProcessor.h) through host/demo_render, the flow INTEGRATION.md §2 shows."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG

DEMO = os.path.join(PKG, "host", "demo_render")


def _run(*args):
    assert os.path.exists(DEMO), "build the demo with `make -C juicy-audio-plugins_b200` (python __graft_entry__.py build)"
    return subprocess.run([DEMO, *args], capture_output=True, text=True, timeout=300)


def test_cpp_parameter_surface_matches_python_mirror(jb):
    r = _run("--params-only")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    eng = jb.BatchProcessor(["JuicyPunch", "JuicyWidth"], 8, device=-1)
    eng.setCurrentProgram(2, 0)
    eng.setParameter("haasMs", 16.0, 1)
    head = lines[0].split()
    assert " ".join(head[1:3]) == eng.getProgramName(2, 0)
    assert float(head[4]) == np.float32(eng.getRawParameterValue("punch", 0))
    assert float(head[6]) == np.float32(eng.getRawParameterValue("sustain", 0))
    assert float(head[8]) == np.float32(eng.getRawParameterValue("haasMs", 1))
    ids = [l.split()[1] for l in lines[1:]]
    assert ids == [p["id"] for p in eng.parameterInfo(1)]
    eng.close()


@pytest.mark.gpu
def test_cpp_render_matches_python_mirror_and_oracle(jb, port):
    n_clips, n = 6, 1500
    r = _run(str(n_clips), str(n))
    assert r.returncode == 0, r.stderr
    rows = [l.split() for l in r.stdout.splitlines() if l.startswith("clip ")]
    assert len(rows) == n_clips
    clips = jb.synth_clips("drum", 0, n_clips, n)
    eng = jb.BatchProcessor(["JuicyPunch", "JuicyWidth"], n_clips)
    eng.setCurrentProgram(2, 0)
    eng.setParameter("haasMs", 16.0, 1)
    eng.prepareToPlay(48000.0, 512)
    out = eng.processBlock(clips)
    rec = eng.getLatestMetrics(1)
    eng.close()
    for c, row in enumerate(rows):
        assert abs(float(row[3]) - rec[c, 13]) < 1e-4
        assert abs(float(row[5]) - rec[c, 0]) < 1e-4
        assert float(row[7]) == pytest.approx(float(np.abs(out[c].astype(np.float64)).sum()), rel=1e-6)
        ref, hists = port.run_chain(["JuicyPunch", "JuicyWidth"], clips[c], programs={0: 2}, params={1: {"haasMs": 16.0}})
        assert float(row[7]) == pytest.approx(float(np.abs(ref.astype(np.float64)).sum()), rel=1e-4)
        assert abs(float(row[3]) - hists[-1][-1][13]) <= 0.01


def test_cpp_demo_fails_loudly_without_gpu(jb):
    if jb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run("4", "512")
    assert r.returncode == 2 and "juicy_batch error" in r.stderr
