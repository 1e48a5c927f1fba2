"""Per-clip parameters and per-block automation.

The reference re-reads its parameters at the top of every processBlock, so every plugin instance can
carry its own settings and a host can change them between callbacks.  The oracle does literally that
(one instance per clip, set_param between process() calls); the engine must match it within the
tolerances of BASELINE.json (samples 1e-5 of clip peak, records 0.01 absolute).

CPU part: the host-side bookkeeping (parameter sets, ranges, errors) through the C ABI -- no GPU."""
import numpy as np
import pytest

from cases import SAMPLE_RATE, BLOCK, FULL_CHAIN
from conftest import assert_samples_close, assert_records_close


# ------------------------------------------------------------------------------------------ host logic (no GPU)

def test_parameter_sets_bookkeeping(jb):
    eng = jb.BatchProcessor(["JuicySaturator", "JuicyWidth"], 100, device=-1)
    assert eng.numParameterSets() == 1
    eng.setParameterClips("drive", 12.0, 10, 20, slot=0)
    assert eng.numParameterSets() == 2
    assert eng.getParameterClip("drive", 9, 0) == eng.getRawParameterValue("drive", 0)
    assert eng.getParameterClip("drive", 10, 0) == pytest.approx(12.0)
    assert eng.getParameterClip("drive", 29, 0) == pytest.approx(12.0)
    assert eng.getParameterClip("drive", 30, 0) == eng.getRawParameterValue("drive", 0)
    # the value goes through the same normalise -> denormalise round trip as setParameter
    ref = jb.BatchProcessor(["JuicySaturator"], 1, device=-1)
    ref.setParameter("asymmetry", 0.33)
    eng.setParameterClips("asymmetry", 0.33, 0, 5, slot=0)
    assert eng.getParameterClip("asymmetry", 3, 0) == ref.getRawParameterValue("asymmetry")
    ref.close()
    assert eng.numParameterSets() == 3        # [0,5) asym, [10,30) drive, the rest
    # overlapping range in another slot splits sets further; identical settings merge again
    eng.setParameterClips("width", 1.5, 20, 30, slot=1)
    assert eng.numParameterSets() == 5        # [0,5) [5,10)+[50,100) [10,20) [20,30) [30,50)
    eng.setParameterClips("width", eng.getRawParameterValue("width", 1), 20, 30, slot=1)
    assert eng.numParameterSets() == 3
    # an engine-wide change reaches every set and keeps them distinct
    eng.setParameter("mix", 0.5, 0)
    assert eng.numParameterSets() == 3
    assert eng.getParameterClip("mix", 15, 0) == pytest.approx(0.5)
    assert eng.getParameterClip("drive", 15, 0) == pytest.approx(12.0)
    # undoing the differences collapses back to one set
    eng.setParameterClips("drive", eng.getRawParameterValue("drive", 0), 10, 20, slot=0)
    eng.setParameterClips("asymmetry", eng.getRawParameterValue("asymmetry", 0), 0, 5, slot=0)
    assert eng.numParameterSets() == 1
    # programs per clip
    eng.setCurrentProgramClips(3, 40, 10, slot=0)
    assert eng.numParameterSets() == 2
    one = jb.BatchProcessor(["JuicySaturator"], 1, device=-1)
    one.setCurrentProgram(3)
    for pid in ("drive", "asymmetry", "tone", "mix", "output"):
        assert eng.getParameterClip(pid, 45, 0) == one.getRawParameterValue(pid)
    one.close()
    eng.close()


def test_per_clip_errors(jb):
    eng = jb.BatchProcessor(["JuicySaturator"], 8, device=-1)
    with pytest.raises(jb.JuicyBatchError):
        eng.setParameterClips("drive", 3.0, 4, 5)          # runs past the last clip
    with pytest.raises(jb.JuicyBatchError):
        eng.setParameterClips("drive", 3.0, -2, 1)
    with pytest.raises(jb.JuicyBatchError):
        eng.setParameterClips("nosuch", 3.0, 0, 1)
    with pytest.raises(jb.JuicyBatchError):
        eng.getParameterClip("drive", 8)
    with pytest.raises(jb.JuicyBatchError):
        eng.scheduleParameter("nosuch", 2, 1.0)
    eng.scheduleParameter("drive", 2, 9.0)                 # accepted without a GPU; applied by the next render
    eng.clearSchedule()
    eng.close()


# ------------------------------------------------------------------------------------------ GPU parity

def oracle_clip(port, chain, x, settings, block=BLOCK, events=()):
    """One clip through fresh oracle instances.  settings: [(slot, id, value)] applied before prepare;
    events: [(at_block, slot, id, value)] applied between process() calls, like host automation."""
    plugs = [port.PortPlugin(p, 2, SAMPLE_RATE, block) for p in chain]
    for slot, pid, v in settings:
        if pid == "__program__":
            plugs[slot].set_program(int(v))
        else:
            plugs[slot].set_param(pid, v)
    for p in plugs:
        p.prepare()
    n = x.shape[1]
    cuts = sorted({0, n} | {b * block for b, _, _, _ in events if 0 < b * block < n})
    outs, hists = [], [[] for _ in chain]
    for t0, t1 in zip(cuts[:-1], cuts[1:]):
        for b, slot, pid, v in events:
            if b * block == t0:
                plugs[slot].set_param(pid, v)
        cur = x[:, t0:t1]
        for s, p in enumerate(plugs):
            cur, h = p.process(cur)
            hists[s].append(h)
        outs.append(cur)
    for p in plugs:
        p.close()
    return np.concatenate(outs, axis=1), [np.concatenate(h) for h in hists]


@pytest.mark.gpu
def test_texture_material_per_clip_matches_oracle(jb, port):
    """BASELINE config 3's layout: material = clip mod 5 (scattered sets -> clip maps, concurrent launches)."""
    n_clips, n = 67, 3 * BLOCK + 40
    clips = jb.synth_clips("impulse", 5, n_clips, n)
    eng = jb.BatchProcessor("JuicyTexture", n_clips)
    for c in range(n_clips):
        eng.setParameterClips("material", float(c % 5), c, 1)
    assert eng.numParameterSets() == 5
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.enableHistory(8)
    out = eng.processBlock(clips)
    rec = eng.getLatestMetrics(0)
    hist = eng.getHistory(0)
    for c in range(n_clips):
        ref, h = oracle_clip(port, ["JuicyTexture"], clips[c], [(0, "material", float(c % 5))])
        assert_samples_close(out[c], ref, "clip %d material %d" % (c, c % 5))
        assert_records_close(rec[c], h[0][-1], "clip %d" % c)
        assert_records_close(hist[:, c, :], h[0], "clip %d history" % c)
    eng.close()


@pytest.mark.gpu
def test_full_chain_with_ranges_programs_and_scattered_clips(jb, port):
    n_clips, n = 45, 2 * BLOCK + 128
    chain = FULL_CHAIN
    clips = jb.synth_clips("mixed", 11, n_clips, n)
    eng = jb.BatchProcessor(chain, n_clips)
    settings = {c: [] for c in range(n_clips)}

    def give(first, count, slot, pid, v):
        if pid == "__program__":
            eng.setCurrentProgramClips(int(v), first, count, slot)
        else:
            eng.setParameterClips(pid, v, first, count, slot)
        for c in range(first, first + count):
            settings[c].append((slot, pid, v))

    give(0, 20, 0, "__program__", 2)            # Punch preset on a contiguous range
    give(10, 25, 1, "drive", 14.0)              # Saturator, overlapping range
    give(30, 15, 3, "haasMs", 21.0)             # Width: another delay length
    for c in range(1, n_clips, 3):              # scattered: Texture wood on every third clip
        give(c, 1, 2, "material", 2.0)
    give(40, 5, 5, "learn", 1.0)                # Cohere learning on the tail
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(clips)
    recs = [eng.getLatestMetrics(s) for s in range(len(chain))]
    for c in range(n_clips):
        ref, h = oracle_clip(port, chain, clips[c], settings[c])
        assert_samples_close(out[c], ref, "clip %d" % c)
        for s in range(len(chain)):
            assert_records_close(recs[s][c], h[s][-1], "clip %d slot %d" % (c, s))
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["host", "device"])
def test_block_automation_matches_host_automation(path, jb, port):
    """Scheduled changes take effect at block boundaries exactly like set_param between processBlock calls."""
    n_clips, n = 12, 7 * BLOCK + 200
    chain = ["JuicyPunch", "JuicySaturator", "JuicyWidth"]
    clips = jb.synth_clips("drum", 3, n_clips, n)
    eng = jb.BatchProcessor(chain, n_clips)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.enableHistory(16)
    common = [(2, 1, "drive", 3.0), (4, 1, "drive", 15.0), (4, 2, "width", 0.2), (6, 0, "punch", 1.6)]
    for b, slot, pid, v in common:
        eng.scheduleParameter(pid, b, v, slot)
    eng.scheduleParameter("mix", 3, 0.4, 1, first_clip=5, n_clips=4)   # only clips 5..8, from block 3
    if path == "host":
        out = eng.processBlock(clips)
    else:
        d = jb.DeviceBuffer(clips.nbytes)
        d.upload(clips)
        eng.process_device(d.ptr.value, d.ptr.value, n)
        eng.synchronize()
        out = d.download(clips.shape)
    hist = [eng.getHistory(s) for s in range(len(chain))]
    assert eng.getRawParameterValue("drive", 1) == pytest.approx(15.0)        # applied changes persist, like host automation
    assert eng.getParameterClip("mix", 6, 1) == pytest.approx(0.4)
    for c in range(n_clips):
        events = list(common) + ([(3, 1, "mix", 0.4)] if 5 <= c < 9 else [])
        ref, h = oracle_clip(port, chain, clips[c], [], events=events)
        assert_samples_close(out[c], ref, "clip %d" % c)
        for s in range(len(chain)):
            assert_records_close(hist[s][:, c, :], h[s], "clip %d slot %d" % (c, s))
    # a second render continues the block count: schedule relative to blocks already done
    eng.scheduleParameter("drive", eng.historyBlocks() + 1, 2.0, 1)
    with pytest.raises(jb.JuicyBatchError):
        eng.scheduleParameter("drive", 1, 2.0, 1)                      # in the past
    eng.close()


@pytest.mark.gpu
def test_parameter_sets_do_not_change_untouched_clips(jb):
    """Clips whose settings were not touched render bit-identically whether or not other clips have their own."""
    n_clips, n = 64, 2 * BLOCK
    clips = jb.synth_clips("mixed", 0, n_clips, n)
    plain = jb.BatchProcessor(["JuicySaturator", "JuicyCohere"], n_clips)
    plain.prepareToPlay(SAMPLE_RATE, BLOCK)
    want = plain.processBlock(clips)
    plain.close()
    eng = jb.BatchProcessor(["JuicySaturator", "JuicyCohere"], n_clips)
    for c in range(0, n_clips, 2):
        eng.setParameterClips("drive", 18.0, c, 1, 0)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    got = eng.processBlock(clips)
    eng.close()
    assert np.array_equal(got[1::2], want[1::2])
    assert not np.array_equal(got[0::2], want[0::2])


@pytest.mark.gpu
def test_mono_bus_with_per_clip_parameters_and_automation(jb, port):
    """The one-channel kernel under clip maps and a parameter schedule."""
    n_clips, n = 21, 6 * BLOCK + 36
    chain = ["JuicySaturator", "JuicyTexture", "JuicyMotion"]
    clips = jb.synth_clips("mixed", 8, n_clips, n, 1)
    eng = jb.BatchProcessor(chain, n_clips, n_channels=1)
    eng.set_math_mode("fast")
    for c in range(0, n_clips, 2):
        eng.setParameterClips("material", 4.0, c, 1, 1)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.scheduleParameter("drive", 3, 11.0, 0)
    out = eng.processBlock(clips)
    rec = eng.getLatestMetrics(2)
    eng.close()
    for c in range(n_clips):
        plugs = [port.PortPlugin(p, 1, SAMPLE_RATE, BLOCK) for p in chain]
        if c % 2 == 0:
            plugs[1].set_param("material", 4.0)
        for p in plugs:
            p.prepare()
        outs, last = [], None
        for t0, t1, drive in ((0, 3 * BLOCK, None), (3 * BLOCK, n, 11.0)):
            if drive is not None:
                plugs[0].set_param("drive", drive)
            cur = clips[c][:, t0:t1]
            for p in plugs:
                cur, h = p.process(cur)
            outs.append(cur)
            last = h[-1]
        assert_samples_close(out[c], np.concatenate(outs, axis=1), "mono clip %d" % c)
        assert_records_close(rec[c], last, "mono clip %d" % c)
