"""Regression tests for the round-1 review findings: staging-buffer capacities, the math mode on mono buses,
parameter display names, the latest-metrics mailboxes across prepareToPlay."""
import json
import os

import numpy as np
import pytest

from cases import SAMPLE_RATE, BLOCK, PLUGINS
from conftest import assert_samples_close, assert_records_close

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("plugin", PLUGINS)
def test_parameter_display_names_match_the_reference_layout(plugin, jb):
    """ids AND display names of createParameterLayout() (e.g. JuicyWidth/PluginProcessor.cpp:232-234), from the compiled
    reference: tests/golden/param_names_v1.json (tests/golden/make_param_names.py)."""
    golden = json.load(open(os.path.join(HERE, "golden", "param_names_v1.json")))
    eng = jb.BatchProcessor(plugin, 1, device=-1)
    got = [[p["id"], p["name"]] for p in eng.parameterInfo()]
    eng.close()
    assert got == golden[plugin]


def test_parameter_names_fixture_is_what_the_reference_reports(refhost):
    if not refhost.available() or not os.path.exists("/root/reference"):
        pytest.skip("compiled reference not present on this box")
    import ctypes
    golden = json.load(open(os.path.join(HERE, "golden", "param_names_v1.json")))
    for plugin in refhost.PLUGINS:
        p = refhost.RefPlugin(plugin)
        raw = p.lib._lib
        if not hasattr(raw, "ref_param_name"):
            pytest.skip("oracle/_ref predates ref_param_name")
        raw.ref_param_name.restype = ctypes.c_char_p
        raw.ref_param_name.argtypes = [ctypes.c_void_p, ctypes.c_int]
        assert [[pid, raw.ref_param_name(p.h, i).decode()] for i, pid in enumerate(p.param_ids())] == golden[plugin]
        p.close()


@pytest.mark.gpu
def test_host_staging_buffers_grow_independently(jb, monkeypatch):
    """multi-pass call (two staging buffers of X) -> single-pass call that grows buffer 0 only -> multi-pass call needing
    X < Z <= Y: buffer 1 must be reallocated too (it used to be written past its end)."""
    chain = ["JuicySaturator"]
    n_clips = 96

    def render(eng, n, env):
        for k in ("JB_HOST_PASS_MIB", "JB_HOST_SLICE_MIB"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        clips = jb.synth_clips("mixed", 5, n_clips, n)
        eng.reset()
        return clips, eng.processBlock(clips)

    eng = jb.BatchProcessor(chain, n_clips)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    ref = jb.BatchProcessor(chain, n_clips)
    ref.prepareToPlay(SAMPLE_RATE, BLOCK)
    # 1 MiB pass budget: 32 clips per pass (the minimum), X = 32 clips x 2 x n x 4 bytes
    for n, env in ((4 * BLOCK, {"JB_HOST_PASS_MIB": "1"}),        # multi-pass, small X
                   (40 * BLOCK, {}),                               # single pass, grows buffer 0 to Y
                   (12 * BLOCK, {"JB_HOST_PASS_MIB": "1"}),        # multi-pass, X < Z <= Y
                   (30 * BLOCK, {"JB_HOST_PASS_MIB": "1"})):
        clips, out = render(eng, n, env)
        monkeypatch.delenv("JB_HOST_PASS_MIB", raising=False)
        ref.reset()
        want = ref.processBlock(clips)
        assert np.array_equal(out, want), "n = %d" % n
    eng.close()
    ref.close()


@pytest.mark.gpu
@pytest.mark.parametrize("material", [1, 2], ids=["metal", "wood"])
def test_mono_shaper_into_resonant_texture_honours_the_math_mode(material, jb, port):
    """A mono Saturator -> Texture chain with a resonant material: JB_MATH_AUTO must pick the exact routines on a
    one-channel bus as it does on a stereo one (the mono kernel used to run the MUFU routines whatever the mode)."""
    chain = ["JuicySaturator", "JuicyTexture"]
    n_clips, n = 33, 6 * BLOCK + 40
    clips = jb.synth_clips("mixed", 11, n_clips, n, 1)
    eng = jb.BatchProcessor(chain, n_clips, n_channels=1)
    eng.setParameter("material", float(material), 1)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(clips)
    recs = [eng.getLatestMetrics(s) for s in range(2)]
    eng.close()
    for c in range(n_clips):
        plugs = [port.PortPlugin(p, 1, SAMPLE_RATE, BLOCK) for p in chain]
        plugs[1].set_param("material", float(material))
        cur, last = clips[c], []
        for p in plugs:
            p.prepare()
            cur, h = p.process(cur)
            last.append(h[-1])
        assert_samples_close(out[c], cur, "mono Saturator->Texture clip %d" % c)
        for s in range(2):
            assert_records_close(recs[s][c], last[s], "clip %d slot %d" % (c, s))


@pytest.mark.gpu
def test_latest_metrics_survive_prepare_to_play(jb):
    """The latest* mailboxes are set by the constructor and by processBlock only (e.g. JuicyPunch/PluginProcessor.h:44-51,
    .cpp:115-123): before any block monoSafety reads 1 and the rest 0; after a re-prepare the previous block's values."""
    eng = jb.BatchProcessor(["JuicyPunch"], 8)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    fresh = eng.getLatestMetrics(0)
    assert (fresh[:, 12] == 1.0).all() and (np.delete(fresh, 12, axis=1) == 0.0).all()
    eng.processBlock(jb.synth_clips("drum", 0, 8, 2 * BLOCK))
    after = eng.getLatestMetrics(0)
    assert (after[:, 0] > 0).all()
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    assert np.array_equal(eng.getLatestMetrics(0), after)
    eng.reset()
    assert np.array_equal(eng.getLatestMetrics(0), after)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chain", [["JuicyPunch", "JuicyWidth"], ["JuicySaturator", "JuicyMotion", "JuicyCohere"]],
                         ids=["punch-width", "sat-motion-cohere"])
def test_host_slice_geometry_does_not_change_the_render(chain, jb, monkeypatch):
    """jb_process_host cuts the call into time slices of whole host blocks (JB_HOST_SLICE_MIB, the last ones halving down to
    one block: JB_HOST_TAPER) and, beyond the pass budget, into clip ranges -- state carries across the slices like across
    host callbacks, so every geometry renders the same bits and the same records as one slice over the whole call."""
    n_clips, n = 40, 31 * BLOCK + 100
    clips = jb.synth_clips("mixed", 3, n_clips, n)
    outs = []
    # one slice; 6-block slices with the tapered tail (6 ... 6, 4, 2, 1); the same untapered; one block per slice
    for env in ({"JB_HOST_SLICE_MIB": "4096"}, {"JB_HOST_SLICE_MIB": "1"}, {"JB_HOST_SLICE_MIB": "1", "JB_HOST_TAPER": "0"},
                {"JB_HOST_SLICE_MIB": "1", "JB_HOST_MIN_SLICE_BLOCKS": "1", "JB_HOST_PASS_MIB": "1"}):
        for k in ("JB_HOST_PASS_MIB", "JB_HOST_SLICE_MIB", "JB_HOST_TAPER", "JB_HOST_MIN_SLICE_BLOCKS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = jb.BatchProcessor(chain, n_clips)
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        out = eng.processBlock(clips)
        rec = eng.getLatestMetrics(len(chain) - 1)
        eng.close()
        outs.append((out, rec))
    for out, rec in outs[1:]:
        assert np.array_equal(out, outs[0][0])
        assert np.array_equal(rec, outs[0][1])


@pytest.mark.gpu
def test_slot_timing_reports_every_plugin_of_a_chain(jb):
    """jb_enable_slot_timing / jb_slot_time_ms: one (time, launches) pair per plugin of a chain that renders as one launch
    per plugin; nothing for a render that takes the cooperative kernel (jb_kernel_time_ms covers that one)."""
    chain = ["JuicySaturator", "JuicyCohere", "JuicyInfer"]
    n_clips, n = 64, 3 * BLOCK          # fewer than four host blocks: no plugin pipeline, one launch per plugin
    d = jb.DeviceBuffer(n_clips * 2 * n * 4)
    jb.synth_fill_device(d.ptr.value, "mixed", 0, n_clips, 2, n, SAMPLE_RATE, device=0, stream=0)
    eng = jb.BatchProcessor(chain, n_clips)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.enable_slot_timing(True)
    for _ in range(2):
        eng.process_device(d.ptr.value, d.ptr.value, n)
    times = eng.slot_times_ms()
    assert [k for _, k in times] == [2, 2, 2]
    assert all(ms > 0.0 for ms, _ in times)
    assert [k for _, k in eng.slot_times_ms()] == [0, 0, 0]      # read once, then reset
    eng.enable_slot_timing(False)
    eng.process_device(d.ptr.value, d.ptr.value, n)
    assert [k for _, k in eng.slot_times_ms()] == [0, 0, 0]
    eng.close()
    coop = jb.BatchProcessor(["JuicyPunch", "JuicyWidth"], n_clips)
    coop.prepareToPlay(SAMPLE_RATE, BLOCK)
    coop.enable_slot_timing(True)
    coop.process_device(d.ptr.value, d.ptr.value, n)
    if coop.path_launches()[0] > 0:
        assert [k for _, k in coop.slot_times_ms()] == [0, 0]
    coop.close()
    d.free()


@pytest.mark.gpu
def test_scoring_in_place_on_the_host_leaves_the_buffer_and_returns_the_same_records(jb):
    """JuicyInfer with trim = 0 dB does not change the audio (applyGain(1), JuicyInfer/PluginProcessor.cpp:79): rendered in
    place on the host, jb_process_host brings back the records only.  Same records as the out-of-place render, buffer
    untouched; with a trim the download happens and the buffer changes."""
    n_clips, n = 48, 9 * BLOCK + 17
    clips = jb.synth_clips("mixed", 21, n_clips, n)

    def render(in_place, trim):
        eng = jb.BatchProcessor(["JuicyInfer"], n_clips)
        eng.setParameter("trim", trim)
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        buf = clips.copy()
        out = buf if in_place else np.empty_like(buf)
        eng.process_host_ptr(buf.ctypes.data, out.ctypes.data, n)
        rec = eng.getLatestMetrics(0)
        eng.close()
        return out, rec

    out_a, rec_a = render(False, 0.0)
    out_b, rec_b = render(True, 0.0)
    assert np.array_equal(out_a, clips) and np.array_equal(out_b, clips)
    assert np.array_equal(rec_a, rec_b)
    out_c, _ = render(True, -6.0)
    out_d, _ = render(False, -6.0)
    assert np.array_equal(out_c, out_d) and not np.array_equal(out_c, clips)
