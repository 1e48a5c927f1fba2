"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, mirrors the
reference's parameter surface (ids, ranges, defaults, programs) and refuses to compute without
a GPU (no CPU fallback).  No compute calls here."""
import ctypes

import numpy as np
import pytest

from cases import PLUGINS


def test_library_exports_every_declared_symbol(jb):
    L = jb.lib()
    names = jb.exported_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), "include/juicy_batch.h declares %s but the library does not export it" % name
    assert L.jb_abi_version() == 1


@pytest.mark.parametrize("plugin", PLUGINS)
def test_parameter_surface_matches_oracle(plugin, jb, port):
    eng = jb.BatchProcessor(plugin, 4, device=-1)
    ora = port.PortPlugin(plugin)
    info = eng.parameterInfo()
    assert [p["id"] for p in info] == ora.param_ids()
    for i, p in enumerate(info):
        lo, hi, interval = ora.param_range(i)
        assert (p["min"], p["max"]) == (lo, hi)
        assert eng.getRawParameterValue(p["id"]) == ora.get_param(p["id"]), p["id"]
    assert eng.getNumPrograms() == ora.num_programs()
    for g in range(eng.getNumPrograms()):
        eng.setCurrentProgram(g)
        ora.set_program(g)
        assert eng.getCurrentProgram() == ora.lib.get_program(ora.h)
        assert eng.getProgramName(g) == ora.program_name(g)
        for p in info:
            assert eng.getRawParameterValue(p["id"]) == ora.get_param(p["id"]), (g, p["id"])
    # host automation: plain values, out-of-range values (clamped), normalised values
    for p in info:
        if p["is_output"]:
            continue
        for v in (p["min"] - 1.0, p["min"], 0.3 * p["min"] + 0.7 * p["max"], p["max"], p["max"] + 5.0):
            eng.setParameter(p["id"], v)
            ora.set_param(p["id"], v)
            assert eng.getRawParameterValue(p["id"]) == ora.get_param(p["id"]), (p["id"], v)
        for n in (0.0, 0.123, 0.5, 0.77, 1.0):
            eng.setValueNotifyingHost(p["id"], n)
            ora.set_param_normalised(p["id"], n)
            assert eng.getRawParameterValue(p["id"]) == ora.get_param(p["id"]), (p["id"], n)
    eng.close()


def test_unknown_ids_and_slots_are_errors(jb):
    eng = jb.BatchProcessor(["JuicyPunch", "JuicyWidth"], 2, device=-1)
    with pytest.raises(jb.JuicyBatchError):
        eng.getRawParameterValue("drive", slot=0)
    with pytest.raises(jb.JuicyBatchError):
        eng.setParameter("punch", 1.0, slot=5)
    assert eng.getRawParameterValue("haasMs", slot="JuicyWidth") == 12.0
    eng.close()
    with pytest.raises(jb.JuicyBatchError):
        jb.BatchProcessor([99], 2, device=-1)
    with pytest.raises(jb.JuicyBatchError):
        jb.BatchProcessor("JuicyPunch", 0, device=-1)
    with pytest.raises(jb.JuicyBatchError):
        jb.BatchProcessor(["JuicyPunch"] * 9, 1, device=-1)


def test_no_cpu_fallback(jb):
    """Without a device the engine must refuse to render -- never fall back to host code."""
    eng = jb.BatchProcessor("JuicySaturator", 2, device=-1)
    with pytest.raises(jb.JuicyBatchError) as e:
        eng.prepareToPlay(48000.0, 512)
    assert e.value.code == -3
    with pytest.raises(jb.JuicyBatchError):
        eng.processBlock(np.zeros((2, 2, 64), dtype=np.float32))
    eng.close()
    if jb.device_count() == 0:
        with pytest.raises(jb.JuicyBatchError) as e:
            jb.BatchProcessor("JuicySaturator", 2, device=0)
        assert e.value.code == -3


def test_host_synth_shapes(jb):
    for kind in ("sweep", "noise", "impulse", "drum", "mixed"):
        x = jb.synth_clips(kind, 3, 5, 2048)
        assert x.shape == (5, 2, 2048) and np.isfinite(x).all()
        assert np.abs(x).max() <= 1.3
    a = jb.synth_clips("noise", 0, 4, 256)
    b = jb.synth_clips("noise", 2, 2, 256)
    assert np.array_equal(a[2:], b), "first_clip must offset the per-clip seeds"
    m = jb.synth_clips("mixed", 0, 4, 256)
    for k, kind in enumerate(("sweep", "noise", "impulse", "drum")):
        assert np.array_equal(m[k], jb.synth_clips(kind, k, 1, 256)[0])


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/juicy_batch.h must stay a C header (the reference-side binding may be C, cgo, ...): compile a C99 program
    against it with -pedantic, link the library, run the host-only part (device -1: parameter logic without a GPU)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "juicy-audio-plugins_b200")
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c_abi_smoke.c"), "-o", exe, "-L", pkg, "-ljuicy_batch",
                           "-Wl,-rpath," + pkg])
    out = subprocess.check_output([exe]).decode()
    assert out.startswith("sets 4 state bytes"), out


def test_synth_numpy_port_matches_the_library_generator(jb):
    """oracle/synth_np.py (inputs of bench.py's CPU reference arm, which must not load the product library)."""
    from oracle import synth_np
    for kind, tol in (("noise", 0.0), ("impulse", 0.0), ("drum", 1e-6), ("sweep", 1e-3), ("mixed", 1e-3)):
        a = jb.synth_clips(kind, 4090, 6, 5000)
        b = synth_np.synth_clips(kind, 4090, 6, 5000)
        assert float(np.abs(a - b).max()) <= tol, kind


def test_shard_range_partitions(jb):
    for n_clips, world in ((4096, 1), (262144, 8), (10, 4), (3, 8)):
        ranges = [jb.shard_range(n_clips, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n_clips
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1
    with pytest.raises(jb.JuicyBatchError):
        jb.shard_range(10, 4, 4)


def test_host_render_slice_plan(jb):
    """jb_plan_slices (the time slices of jb_process_host; no device needed): every plan covers [0, total) exactly once in
    order, uniform slices of the requested size, and -- tapered -- a tail that halves down to one block."""
    for total in (1, 2, 3, 5, 12, 13, 94, 938):
        for per in (1, 2, 3, 5, 6, 8, 94, 200):
            for taper in (False, True):
                plan = jb.plan_slices(total, per, taper)
                assert plan[0] == 0 and plan[-1] == total
                sizes = [b - a for a, b in zip(plan, plan[1:])]
                assert all(s > 0 for s in sizes) and max(sizes) <= min(per, total)
                if not taper or total < 4 * min(per, total):
                    assert sizes[:-1] == [min(per, total)] * (len(sizes) - 1)      # only the last slice may be ragged
    assert [b - a for a, b in zip(*(lambda p: (p, p[1:]))(jb.plan_slices(94, 6, True)))][-4:] == [3, 4, 2, 1]   # 66..84 uniform, ragged 3, then 4, 2, 1
    assert jb.plan_slices(94, 1, True) == list(range(95))
    with pytest.raises(jb.JuicyBatchError):
        jb.plan_slices(0, 4)
