"""The clip-per-CTA kernel for few live streams (csrc/jb_solo.cu): samples bit-identical to the lane kernels (same
operations per sample), records within tolerance of them and of the oracle, state carried across calls, ragged blocks,
in place and out of place, per-clip parameter sets -- and BASELINE.json configs[0] (one 10 s clip) against the oracle."""
import numpy as np
import pytest

from cases import SAMPLE_RATE, BLOCK
from conftest import assert_samples_close, assert_records_close

pytestmark = pytest.mark.gpu


def _engine(jb, plugin, n_clips, path, math="auto", settings=None, program=None, block=BLOCK):
    eng = jb.BatchProcessor([plugin], n_clips)
    if program is not None:
        eng.setCurrentProgram(program)
    for k, v in (settings or {}).items():
        eng.setParameter(k, v)
    eng.set_path(path)
    eng.set_math_mode(math)
    eng.prepareToPlay(SAMPLE_RATE, block)
    eng.enableHistory(32)
    return eng


@pytest.mark.parametrize("plugin,math,program", [("JuicySaturator", "fast", 0), ("JuicySaturator", "exact", 3), ("JuicyInfer", "auto", 0),
                                                 ("JuicyInfer", "auto", 2), ("JuicyInfer", "auto", 4),
                                                 ("JuicyCohere", "auto", 0)])
def test_solo_kernel_equals_lane_kernels_and_oracle(plugin, math, program, jb, port):
    n_clips, n = 37, 6 * BLOCK + 132          # ragged last block (132 = 4 * 33)
    clips = jb.synth_clips("mixed", 5, n_clips, n)
    clips *= np.linspace(0.2, 2.0, n_clips, dtype=np.float32)[:, None, None]
    solo = _engine(jb, plugin, n_clips, "auto", math, program=program)
    lane = _engine(jb, plugin, n_clips, "lane", math, program=program)
    # two calls (state carries across them), the second one ragged
    cut = 2 * BLOCK
    outs = {}
    for name, eng in (("solo", solo), ("lane", lane)):
        a = eng.processBlock(clips[:, :, :cut])
        b = eng.processBlock(clips[:, :, cut:])
        outs[name] = np.concatenate([a, b], axis=2)
    assert np.array_equal(outs["solo"].view(np.uint32), outs["lane"].view(np.uint32))
    hs, hl = solo.getHistory(0), lane.getHistory(0)
    assert hs.shape == hl.shape and hs.shape[0] == 2 + 5
    assert_records_close(hs, hl, "solo vs lane records")
    for c in range(0, n_clips, 6):
        p = port.PortPlugin(plugin)
        p.set_program(program)
        p.prepare()
        ref_a, h_a = p.process(clips[c][:, :cut])
        ref_b, h_b = p.process(clips[c][:, cut:])
        assert_samples_close(outs["solo"][c], np.concatenate([ref_a, ref_b], axis=1), "clip %d" % c)
        assert_records_close(hs[:, c], np.concatenate([h_a, h_b]), "clip %d records" % c)
    solo.close()
    lane.close()


def test_solo_kernel_cohere_learning_targets_across_calls(jb, port):
    """JuicyCohere on the few-streams kernel with `learn` on: the block pre-pass moves the three band targets every block
    (JuicyCohere/PluginProcessor.cpp:78-84) and they persist across calls; contextfit (the record's aux field) follows."""
    n_clips, n = 11, 7 * BLOCK
    clips = jb.synth_clips("mixed", 9, n_clips, n)
    settings = {"learn": 1.0, "match": 0.8, "tail": 0.5, "decay": 0.6}
    solo = _engine(jb, "JuicyCohere", n_clips, "auto", settings=settings)
    lane = _engine(jb, "JuicyCohere", n_clips, "lane", settings=settings)
    cut = 3 * BLOCK
    outs = {}
    for name, eng in (("solo", solo), ("lane", lane)):
        outs[name] = np.concatenate([eng.processBlock(clips[:, :, :cut]), eng.processBlock(clips[:, :, cut:])], axis=2)
    assert np.array_equal(outs["solo"].view(np.uint32), outs["lane"].view(np.uint32))
    hs, hl = solo.getHistory(0), lane.getHistory(0)
    assert_records_close(hs, hl, "solo vs lane records")
    assert np.array_equal(hs[:, :, 14], hl[:, :, 14])          # contextfit: the same operations on the same sums
    for c in range(0, n_clips, 3):
        p = port.PortPlugin("JuicyCohere")
        for k, v in settings.items():
            p.set_param(k, v)
        p.prepare()
        ref_a, h_a = p.process(clips[c][:, :cut])
        ref_b, h_b = p.process(clips[c][:, cut:])
        assert_samples_close(outs["solo"][c], np.concatenate([ref_a, ref_b], axis=1), "clip %d" % c)
        assert_records_close(hs[:, c], np.concatenate([h_a, h_b]), "clip %d records" % c)
    solo.close()
    lane.close()


def test_solo_kernel_small_blocks_in_place_and_per_clip_parameters(jb, port):
    n_clips, n, block = 9, 5 * 128 + 64, 128
    clips = jb.synth_clips("drum", 0, n_clips, n)
    eng = jb.BatchProcessor(["JuicySaturator"], n_clips)
    for c in range(n_clips):
        eng.setParameterClips("drive", 3.0 + 2.0 * (c % 3), c, 1)     # three scattered parameter sets (clip maps)
    eng.prepareToPlay(SAMPLE_RATE, block)
    d = jb.DeviceBuffer(clips.nbytes)
    d.upload(clips)
    eng.process_device(d.ptr.value, d.ptr.value, n)                     # in place
    eng.synchronize()
    out = d.download(clips.shape)
    rec = eng.getLatestMetrics(0)
    eng.close()
    d.free()
    for c in range(n_clips):
        p = port.PortPlugin("JuicySaturator", 2, SAMPLE_RATE, block)
        p.set_param("drive", 3.0 + 2.0 * (c % 3))
        p.prepare()
        ref, h = p.process(clips[c])
        assert_samples_close(out[c], ref, "clip %d" % c)
        assert_records_close(rec[c], h[-1], "clip %d" % c)


def test_config1_one_ten_second_clip_every_sample_and_every_block(jb, port):
    """BASELINE.json configs[0] through the few-streams kernel: 938 blocks, every sample and every block's record."""
    n = 480000
    clip = jb.synth_clips("sweep", 0, 1, n)
    eng = jb.BatchProcessor(["JuicySaturator"], 1)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.enableHistory(1000)
    out = eng.processBlock(clip)
    hist = eng.getHistory(0)
    eng.close()
    ref, h = port.run_chain(["JuicySaturator"], clip[0])
    assert_samples_close(out[0], ref, "configs[0]")
    assert hist.shape[0] == 938
    assert_records_close(hist[:, 0], h[0], "configs[0] records")
