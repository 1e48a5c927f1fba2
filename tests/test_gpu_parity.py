"""GPU parity: the CUDA engine, called through the C ABI (libjuicy_batch.so), against
  * the committed golden vectors (made from the reference's own C++), and
  * the CPU oracle on seeded inputs,
within the tolerances BASELINE.json states: samples |gpu-ref| <= 1e-5 x clip peak, metric records
<= 0.01 absolute.  Every test here needs a B200 and fails (not skips) if the library cannot render."""
import numpy as np
import pytest

from cases import (GOLDEN_CASES, N_SAMPLES, SAMPLE_RATE, BLOCK, SEED, FULL_CHAIN, PLUGINS, case_input,
                   apply_case_settings, load_golden)
from conftest import assert_samples_close, assert_records_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return load_golden()


def oracle_render(port, chain, clips, block=BLOCK, programs=None, params=None):
    """Fresh oracle instances per clip (what N independent plugin instances do)."""
    outs, hists = [], []
    for x in clips:
        o, h = port.run_chain(chain, x, sample_rate=SAMPLE_RATE, block_size=block, programs=programs, params=params)
        outs.append(o)
        hists.append(h)
    return np.stack(outs), hists


@pytest.mark.parametrize("case", GOLDEN_CASES, ids=[c["name"] for c in GOLDEN_CASES])
def test_engine_matches_golden(case, golden, jb):
    z, _ = golden
    x = z["in/%s/%d" % (case["input"], case["clip"])]
    # the golden clip sits in lane 1 of a 3-clip batch whose neighbours carry other signals
    other = jb.synth_clips("mixed", 100, 2, N_SAMPLES)
    batch = np.stack([other[0], x, other[1]])
    eng = jb.BatchProcessor(case["chain"], 3)
    apply_case_settings(eng, case)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.enableHistory(16)
    out = eng.processBlock(batch)
    assert_samples_close(out[1], z["out/" + case["name"]], case["name"])
    for slot in range(len(case["chain"])):
        ref = z["hist/%s/%d" % (case["name"], slot)]
        got = eng.getHistory(slot)[:, 1, :]
        assert got.shape == ref.shape
        assert_records_close(got, ref, "%s slot %d" % (case["name"], slot))
        assert_records_close(eng.getLatestMetrics(slot)[1], ref[-1], "%s latest" % case["name"])
    eng.close()


@pytest.mark.parametrize("chain", [[p] for p in PLUGINS] + [["JuicyPunch", "JuicyWidth"], FULL_CHAIN],
                         ids=list(PLUGINS) + ["punch-width", "full-chain"])
def test_engine_matches_oracle_on_mixed_batch(chain, jb, port):
    """70 clips (not a multiple of the warp size) of all four signal kinds, ragged length."""
    n_clips, n = 70, 2 * BLOCK + 300
    clips = jb.synth_clips("mixed", 7, n_clips, n)
    clips[:, :, :] *= np.linspace(0.3, 1.6, n_clips, dtype=np.float32)[:, None, None]  # drive some into clipping
    eng = jb.BatchProcessor(chain, n_clips)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(clips)
    ref, hists = oracle_render(port, chain, clips)
    assert_samples_close(out, ref, "+".join(chain))
    for slot in range(len(chain)):
        got = eng.getLatestMetrics(slot)
        want = np.stack([h[slot][-1] for h in hists])
        assert_records_close(got, want, "%s slot %d" % ("+".join(chain), slot))
    eng.close()


@pytest.mark.parametrize("material", range(5))
def test_texture_materials_match_oracle(material, jb, port):
    n_clips, n = 33, 3 * BLOCK + 11
    clips = jb.synth_clips("mixed", 40, n_clips, n)
    params = {0: {"material": float(material), "texture": 0.8, "tailshape": 0.7}}
    eng = jb.BatchProcessor("JuicyTexture", n_clips)
    for k, v in params[0].items():
        eng.setParameter(k, v)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(clips)
    ref, hists = oracle_render(port, ["JuicyTexture"], clips, params=params)
    assert_samples_close(out, ref, "texture material %d" % material)
    assert_records_close(eng.getLatestMetrics(0), np.stack([h[0][-1] for h in hists]), "texture material %d" % material)
    eng.close()


@pytest.mark.parametrize("block", [64, 500, 512, 1024])
def test_block_sizes(block, jb, port):
    """Results depend on the host's block size; the engine must follow it, odd sizes included."""
    chain = ["JuicyPunch", "JuicyWidth", "JuicyCohere"]
    clips = jb.synth_clips("drum", 9, 5, 2600)
    eng = jb.BatchProcessor(chain, 5)
    eng.prepareToPlay(SAMPLE_RATE, block)
    out = eng.processBlock(clips)
    ref, hists = oracle_render(port, chain, clips, block=block)
    assert_samples_close(out, ref, "block %d" % block)
    for slot in range(len(chain)):
        assert_records_close(eng.getLatestMetrics(slot), np.stack([h[slot][-1] for h in hists]), "block %d" % block)
    eng.close()


def test_state_carries_across_calls_and_reset(jb, port):
    """Two consecutive host callbacks equal one long one; prepareToPlay/reset restores the initial state."""
    chain = FULL_CHAIN
    clips = jb.synth_clips("mixed", 0, 8, 4 * BLOCK)
    eng = jb.BatchProcessor(chain, 8)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    whole = eng.processBlock(clips)
    split = jb.BatchProcessor(chain, 8)  # a fresh set of instances (Motion's LCG survives prepareToPlay)
    split.prepareToPlay(SAMPLE_RATE, BLOCK)
    first = split.processBlock(clips[:, :, :BLOCK])
    second = split.processBlock(clips[:, :, BLOCK:])
    assert np.array_equal(np.concatenate([first, second], axis=2), whole)
    assert np.array_equal(split.getLatestMetrics(6), eng.getLatestMetrics(6))
    split.close()
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    again = eng.processBlock(clips)
    # Motion's LCG is seeded at construction only (JuicyMotion/PluginProcessor.h:65), so a second
    # render after prepareToPlay continues its sequence -- compare against the oracle doing the same
    outs = []
    for x in clips:
        cur = x
        plugs = [port.PortPlugin(p) for p in chain]
        for p in plugs:
            p.prepare()
        for p in plugs:
            cur, _ = p.process(cur)
        cur = x
        for p in plugs:
            p.prepare()
        for p in plugs:
            cur, _ = p.process(cur)
        outs.append(cur)
    assert_samples_close(again, np.stack(outs), "second render after prepareToPlay")
    eng.close()


def test_device_resident_path_equals_host_path(jb):
    chain = ["JuicyPunch", "JuicyWidth"]
    n_clips, n = 96, 3 * BLOCK + 128
    clips = jb.synth_clips("drum", 0, n_clips, n)
    eng = jb.BatchProcessor(chain, n_clips)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    via_host = eng.processBlock(clips)
    rec_host = eng.getLatestMetrics(1)
    eng.reset()
    d_in = jb.DeviceBuffer(clips.nbytes)
    d_out = jb.DeviceBuffer(clips.nbytes)
    d_in.upload(clips)
    eng.process_device(d_in.ptr.value, d_out.ptr.value, n)
    eng.synchronize()
    assert np.array_equal(d_out.download(clips.shape), via_host)
    assert np.array_equal(d_in.download(clips.shape), clips), "out-of-place render must leave the input intact"
    assert np.array_equal(eng.getLatestMetrics(1), rec_host)
    # in place
    eng.reset()
    eng.process_device(d_in.ptr.value, d_in.ptr.value, n)
    eng.synchronize()
    assert np.array_equal(d_in.download(clips.shape), via_host)
    ms, launches = eng.kernel_time_ms()
    assert launches >= 2 and ms > 0.0
    # device generator agrees with the host generator to rounding
    jb.synth_fill_device(d_out.ptr.value, "drum", 0, n_clips, 2, n)
    dev = d_out.download(clips.shape)
    assert np.abs(dev - clips).max() < 2.0e-4
    eng.close()


def test_silence_scores_forty_and_lanes_are_independent(jb):
    """Known answer + a clip's result must not depend on its neighbours or lane."""
    n_clips, n = 257, 2 * BLOCK
    eng = jb.BatchProcessor("JuicyInfer", n_clips)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.processBlock(np.zeros((n_clips, 2, n), dtype=np.float32))
    rec = eng.getLatestMetrics(0)
    assert np.allclose(rec[:, 0], 40.0, atol=1e-6)
    eng.close()

    chain = FULL_CHAIN
    one = jb.synth_clips("drum", 9, 1, n)
    batch = jb.synth_clips("mixed", 0, n_clips, n)
    batch[5] = one[0]
    batch[200] = one[0]
    eng = jb.BatchProcessor(chain, n_clips)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(batch)
    solo = jb.BatchProcessor(chain, 1)
    solo.prepareToPlay(SAMPLE_RATE, BLOCK)
    ref = solo.processBlock(one)
    assert np.array_equal(out[5], ref[0]) and np.array_equal(out[200], ref[0])
    assert np.array_equal(eng.getLatestMetrics(6)[5], solo.getLatestMetrics(6)[0])
    eng.close()
    solo.close()


# ---------------------------------------------------------------- the block-cooperative kernel (jb_coop.cu)

COOP_CHAINS = [["JuicyPunch", "JuicyWidth"], ["JuicyPunch"], ["JuicyWidth"], ["JuicyInfer"],
               ["JuicyPunch", "JuicyWidth", "JuicyInfer"], ["JuicyWidth", "JuicyInfer"]]


@pytest.mark.parametrize("chain", COOP_CHAINS, ids=["+".join(c) for c in COOP_CHAINS])
def test_coop_kernel_matches_oracle_and_lane_kernel(chain, jb, port):
    """Forced cooperative path vs the oracle (tolerances of BASELINE.json) and vs the lane-per-clip kernel."""
    n_clips, n = 70, 2 * BLOCK + 300
    clips = jb.synth_clips("mixed", 7, n_clips, n)
    clips[:, :, :] *= np.linspace(0.3, 1.6, n_clips, dtype=np.float32)[:, None, None]
    outs, recs = {}, {}
    for path in ("coop", "lane"):
        eng = jb.BatchProcessor(chain, n_clips)
        eng.set_path(path)
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        eng.enableHistory(4)
        outs[path] = eng.processBlock(clips)
        recs[path] = [eng.getHistory(slot) for slot in range(len(chain))]
        coop_launches, lane_launches = eng.path_launches()
        assert (coop_launches > 0 and lane_launches == 0) if path == "coop" else (lane_launches > 0 and coop_launches == 0)
        eng.close()
    ref, hists = oracle_render(port, chain, clips)
    assert_samples_close(outs["coop"], ref, "coop " + "+".join(chain))
    assert_samples_close(outs["coop"], outs["lane"], "coop vs lane " + "+".join(chain))
    for slot in range(len(chain)):
        want = np.stack([h[slot] for h in hists], axis=1)  # [block][clip][16]
        assert_records_close(recs["coop"][slot], want, "coop %s slot %d" % ("+".join(chain), slot))
        assert_records_close(recs["coop"][slot], recs["lane"][slot], "coop vs lane slot %d" % slot)


@pytest.mark.parametrize("block", [64, 256, 500, 512])
def test_coop_kernel_block_sizes_and_split_calls(block, jb, port):
    chain = ["JuicyPunch", "JuicyWidth"]
    n_clips, n = 5, 2600
    clips = jb.synth_clips("drum", 9, n_clips, n)
    eng = jb.BatchProcessor(chain, n_clips)
    eng.set_path("coop")
    eng.setParameter("haasMs", 3.0, 1)  # 144-sample delay: the ring read lands inside the same step
    eng.prepareToPlay(SAMPLE_RATE, block)
    out = eng.processBlock(clips)
    ref, hists = oracle_render(port, chain, clips, block=block, params={1: {"haasMs": 3.0}})
    assert_samples_close(out, ref, "coop block %d" % block)
    for slot in range(len(chain)):
        assert_records_close(eng.getLatestMetrics(slot), np.stack([h[slot][-1] for h in hists]), "coop block %d" % block)
    # two host callbacks == one (state carried through the SoA arrays between launches)
    eng.reset()
    cut = 3 * block
    first = eng.processBlock(clips[:, :, :cut])
    second = eng.processBlock(clips[:, :, cut:])
    assert np.array_equal(np.concatenate([first, second], axis=2), out)
    eng.close()


def test_coop_kernel_many_groups(jb, port):
    """More clips than one wave of CTA groups (148 x 32): persistent CTAs loop over groups."""
    chain = ["JuicyPunch", "JuicyWidth"]
    n_clips, n = 148 * 32 + 37, BLOCK + 256
    clips = jb.synth_clips("mixed", 3, n_clips, n)
    outs = {}
    for path in ("coop", "lane"):
        eng = jb.BatchProcessor(chain, n_clips)
        eng.set_path(path)
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        outs[path] = (eng.processBlock(clips), eng.getLatestMetrics(0), eng.getLatestMetrics(1))
        eng.close()
    assert_samples_close(outs["coop"][0], outs["lane"][0], "coop vs lane, many groups")
    assert_records_close(outs["coop"][1], outs["lane"][1], "many groups slot 0")
    assert_records_close(outs["coop"][2], outs["lane"][2], "many groups slot 1")
    pick = list(range(0, n_clips, 151))
    ref, hists = oracle_render(port, chain, clips[pick])
    assert_samples_close(outs["coop"][0][pick], ref, "coop vs oracle, many groups")
    assert_records_close(outs["coop"][2][pick], np.stack([h[1][-1] for h in hists]), "many groups records")


def test_host_streaming_slices_and_passes_are_exact(jb, monkeypatch):
    """jb_process_host cuts the render into time slices (whole host blocks) and, for big batches, clip
    passes; any cut must give the bits of the uncut render (state carries like across host callbacks)."""
    chain = ["JuicyPunch", "JuicyWidth", "JuicyInfer"]
    n_clips, n = 96, 9 * BLOCK + 100
    clips = jb.synth_clips("mixed", 3, n_clips, n)

    def render(env):
        for k in ("JB_HOST_PASS_MIB", "JB_HOST_SLICE_MIB"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = jb.BatchProcessor(chain, n_clips)
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        eng.enableHistory(16)
        out = eng.processBlock(clips)
        hist = [eng.getHistory(s) for s in range(len(chain))]
        launches = sum(eng.path_launches())
        eng.close()
        return out, hist, launches

    whole, hist_whole, one = render({"JB_HOST_SLICE_MIB": "4096"})
    assert one == 1
    sliced, hist_sliced, many = render({"JB_HOST_SLICE_MIB": "1", "JB_HOST_PASS_MIB": "1"})  # 1 MiB: 2 blocks x 32 clips
    assert many > 4
    assert np.array_equal(whole, sliced)
    for a, b in zip(hist_whole, hist_sliced):
        assert a.shape == b.shape and np.array_equal(a, b)


def test_lane_kernel_split_chain_equals_fused_chain(jb, monkeypatch):
    """Big lane-kernel batches render a chain as one launch per plugin (plugin-by-plugin == block-by-block for
    causal plugins with identical blocking); both modes must give the same bits."""
    n_clips, n = 64, 5 * BLOCK + 100
    clips = jb.synth_clips("mixed", 11, n_clips, n)

    def render(mode):
        monkeypatch.setenv("JB_LANE_SPLIT", mode)
        eng = jb.BatchProcessor(FULL_CHAIN, n_clips)
        eng.set_path("lane")
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        eng.enableHistory(8)
        out = eng.processBlock(clips)
        hist = [eng.getHistory(s) for s in range(len(FULL_CHAIN))]
        eng.close()
        return out, hist

    fused, hist_fused = render("0")
    split, hist_split = render("1")
    assert np.array_equal(fused, split)
    for a, b in zip(hist_fused, hist_split):
        assert np.array_equal(a, b)


# ------------------------------------------------------------------ two lanes per clip (csrc/jb_pair.cu)

PAIR_CASES = [("JuicySaturator", (), "fast"), ("JuicySaturator", (), "exact"), ("JuicyPunch", (), "fast"), ("JuicyPunch", (), "exact")] + \
             [("JuicyTexture", ((0, "material", float(m)),), "auto") for m in range(5)]


def _render_in_fresh_process(env, chain, settings, math, n_clips, n, tmp, tag):
    """The library reads its kernel-choice environment once, at load: render in a fresh interpreter."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, os, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import load_juicy_batch
jb = load_juicy_batch()
clips = jb.synth_clips("mixed", 17, %d, %d)
clips *= np.linspace(0.3, 1.8, %d, dtype=np.float32)[:, None, None]
chain = %r
eng = jb.BatchProcessor(chain, %d)
for slot, pid, v in %r:
    eng.setParameter(pid, v, slot)
eng.set_path("lane"); eng.set_math_mode(%r)
eng.prepareToPlay(48000.0, 512); eng.enableHistory(64)
a = eng.processBlock(clips[:, :, :1024]); b = eng.processBlock(clips[:, :, 1024:])
np.savez(sys.argv[1], out=np.concatenate([a, b], axis=2), clips=clips, **{"hist%%d" %% s: eng.getHistory(s) for s in range(len(chain))})
''' % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
       n_clips, n, n_clips, list(chain), n_clips, list(settings), math)
    path = os.path.join(tmp, "%s.npz" % tag)
    subprocess.check_call([sys.executable, "-c", code, path], env=dict(os.environ, **env))
    return dict(np.load(path))


@pytest.mark.parametrize("chain", [["JuicySaturator"], ["JuicyCohere"], ["JuicyWidth"], ["JuicyInfer"],
                                   ["JuicySaturator", "JuicyWidth", "JuicyCohere", "JuicyInfer"]],
                         ids=["saturator", "cohere", "width", "infer", "light-chain"])
def test_tile_streaming_is_bit_identical_to_lane_streaming(chain, jb, port):
    """Warp-transposed tile streaming (JB_TILE=1; default from 32768 clips) == per-lane ring streaming (JB_TILE=0):
    same arithmetic, only the way samples travel differs.  64 clips, 5 blocks + a 12-sample tail in two calls."""
    import tempfile
    n_clips, n = 64, 5 * BLOCK + 12
    settings = [(chain.index("JuicyInfer"), "trim", 3.0)] if chain == ["JuicyInfer"] else []
    with tempfile.TemporaryDirectory() as tmp:
        res = {m: _render_in_fresh_process({"JB_TILE": m}, chain, settings, "fast", n_clips, n, tmp, "t" + m) for m in ("0", "1")}
    assert np.array_equal(res["0"]["out"].view(np.uint32), res["1"]["out"].view(np.uint32)), \
        "max diff %g" % float(np.abs(res["0"]["out"] - res["1"]["out"]).max())
    for s in range(len(chain)):
        assert np.array_equal(res["0"]["hist%d" % s].view(np.uint32), res["1"]["hist%d" % s].view(np.uint32)), s
    clips = res["1"]["clips"]
    params = {s: {pid: v} for s, pid, v in settings} if settings else None
    for c in (0, 31, 63):
        ref, h = port.run_chain(chain, clips[c], sample_rate=SAMPLE_RATE, block_size=BLOCK, params=params)
        assert_samples_close(res["1"]["out"][c], ref, "clip %d" % c)
        for s in range(len(chain)):
            assert_records_close(res["1"]["hist%d" % s][:, c, :], h[s], "clip %d slot %d" % (c, s))


@pytest.mark.parametrize("plugin,settings,math", PAIR_CASES,
                         ids=["sat-fast", "sat-exact", "punch-fast", "punch-exact"] + ["texture-%s" % m for m in ("gel", "metal", "wood", "plastic", "flesh")])
def test_pair_kernel_is_bit_identical_to_the_one_lane_kernel(plugin, settings, math, jb, port, monkeypatch):
    """Channel-per-lane kernel == clip-per-lane kernel, samples and every record of every block, bit for bit;
    49 clips (a partial warp, odd count), ragged split into two calls; and both match the oracle."""
    import juicy_batch  # noqa: F401  (the library reads JB_PAIR once, at load: spawn fresh interpreters)
    import subprocess, sys, os, tempfile, json
    n_clips, n = 49, 5 * BLOCK
    code = r'''
import sys, os, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import load_juicy_batch
jb = load_juicy_batch()
clips = jb.synth_clips("mixed", 17, %d, %d)
clips *= np.linspace(0.3, 1.8, %d, dtype=np.float32)[:, None, None]
eng = jb.BatchProcessor([%r], %d)
for slot, pid, v in %r:
    eng.setParameter(pid, v, slot)
eng.set_path("lane"); eng.set_math_mode(%r)
eng.prepareToPlay(48000.0, 512); eng.enableHistory(64)
a = eng.processBlock(clips[:, :, :1024]); b = eng.processBlock(clips[:, :, 1024:])
np.savez(sys.argv[1], out=np.concatenate([a, b], axis=2), hist=eng.getHistory(0), clips=clips)
''' % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
       n_clips, n, n_clips, plugin, n_clips, list(settings), math)
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        for mode in ("0", "1"):
            path = os.path.join(tmp, "m%s.npz" % mode)
            env = dict(os.environ, JB_PAIR=mode)
            subprocess.check_call([sys.executable, "-c", code, path], env=env)
            res[mode] = dict(np.load(path))
    assert np.array_equal(res["0"]["out"].view(np.uint32), res["1"]["out"].view(np.uint32)), \
        "max diff %g" % float(np.abs(res["0"]["out"] - res["1"]["out"]).max())
    assert np.array_equal(res["0"]["hist"].view(np.uint32), res["1"]["hist"].view(np.uint32))
    clips = res["1"]["clips"]
    params = {0: {pid: v for _, pid, v in settings}} if settings else None
    for c in (0, 7, 48):
        ref, h = port.run_chain([plugin], clips[c], sample_rate=SAMPLE_RATE, block_size=BLOCK, params=params)
        assert_samples_close(res["1"]["out"][c], ref, "%s clip %d" % (plugin, c))
        assert_records_close(res["1"]["hist"][:, c, :], h[0], "%s clip %d" % (plugin, c))


@pytest.mark.parametrize("sr,block", [(44100.0, 512), (96000.0, 1024), (22050.0, 256), (192000.0, 2048)])
def test_other_sample_rates_and_block_sizes(sr, block, jb, port):
    """prepareToPlay(sampleRate, samplesPerBlock) at rates other than 48 kHz: every rate-dependent length (Width's 60 ms
    ring, Texture's 80 ms waveguide, the analyzer's and Motion's cooldowns) and coefficient follows the reference."""
    n_clips, n = 40, 3 * block + 64
    for chain, settings in ((FULL_CHAIN, {}), (["JuicyPunch", "JuicyWidth"], {}), (["JuicyTexture"], {0: {"material": 2.0}}),
                            (["JuicyTexture"], {0: {"material": 1.0}}), (["JuicyWidth", "JuicyInfer"], {0: {"haasMs": 17.0}})):
        clips = jb.synth_clips("mixed", 5, n_clips, n, 2, sr)
        eng = jb.BatchProcessor(chain, n_clips)
        for slot, kv in settings.items():
            for k, v in kv.items():
                eng.setParameter(k, v, slot)
        eng.prepareToPlay(sr, block)
        out = eng.processBlock(clips)
        recs = [eng.getLatestMetrics(s) for s in range(len(chain))]
        eng.close()
        for c in range(0, n_clips, 3):
            ref, h = port.run_chain(chain, clips[c], sample_rate=sr, block_size=block, params=settings or None)
            assert_samples_close(out[c], ref, "%s at %g Hz clip %d" % ("+".join(chain), sr, c))
            for s in range(len(chain)):
                assert_records_close(recs[s][c], h[s][-1], "%s at %g Hz clip %d slot %d" % ("+".join(chain), sr, c, s))


@pytest.mark.parametrize("chain", [[p] for p in PLUGINS] + [FULL_CHAIN], ids=list(PLUGINS) + ["full-chain"])
def test_mono_bus_matches_oracle(chain, jb, port):
    """One-channel buses (the reference's other supported layout): [clip][1][sample] audio, analyzer with right = left,
    Width without DSP, Motion / Texture advancing their shared state for one channel only."""
    n_clips, n = 37, 3 * BLOCK + 50
    clips = jb.synth_clips("mixed", 3, n_clips, n, 1)
    clips *= np.linspace(0.3, 1.7, n_clips, dtype=np.float32)[:, None, None]
    settings = {chain.index("JuicyTexture"): {"material": 3.0}} if chain == ["JuicyTexture"] else {}
    eng = jb.BatchProcessor(chain, n_clips, n_channels=1)
    for slot, kv in settings.items():
        for k, v in kv.items():
            eng.setParameter(k, v, slot)
    eng.set_math_mode("fast")
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.enableHistory(8)
    first = eng.processBlock(clips[:, :, :BLOCK + 10])          # ragged split: state carries across calls
    second = eng.processBlock(clips[:, :, BLOCK + 10:])
    recs = [eng.getLatestMetrics(s) for s in range(len(chain))]
    eng.close()
    for c in range(n_clips):
        plugs = [port.PortPlugin(p, 1, SAMPLE_RATE, BLOCK) for p in chain]
        for slot, kv in settings.items():
            for k, v in kv.items():
                plugs[slot].set_param(k, v)
        for p in plugs:
            p.prepare()
        outs, last = [], None
        for seg in (clips[c][:, :BLOCK + 10], clips[c][:, BLOCK + 10:]):
            cur, last = seg, []
            for p in plugs:
                cur, h = p.process(cur)
                last.append(h[-1])
            outs.append(cur)
        ref = np.concatenate(outs, axis=1)
        got = np.concatenate([first[c], second[c]], axis=1)
        assert_samples_close(got, ref, "%s mono clip %d" % ("+".join(chain), c))
        for s in range(len(chain)):
            assert_records_close(recs[s][c], last[s], "%s mono clip %d slot %d" % ("+".join(chain), c, s))


@pytest.mark.parametrize("chain", [FULL_CHAIN, ["JuicySaturator", "JuicyWidth", "JuicyCohere"]], ids=["full-chain", "light-chain"])
def test_pipelined_chain_is_bit_identical(chain, jb, port):
    """Plugins of a chain on separate streams, pipelined over segments of whole blocks (JB_CHAIN_PIPELINE=1), against one
    launch per plugin over the whole call (=0): same samples, same per-block records, in place and with history."""
    import tempfile
    n_clips, n = 40, 23 * BLOCK + 100
    with tempfile.TemporaryDirectory() as tmp:
        res = {m: _render_in_fresh_process({"JB_CHAIN_PIPELINE": m}, chain, [], "auto", n_clips, n, tmp, "p" + m) for m in ("0", "1")}
    assert np.array_equal(res["0"]["out"].view(np.uint32), res["1"]["out"].view(np.uint32)), \
        "max diff %g" % float(np.abs(res["0"]["out"] - res["1"]["out"]).max())
    for s in range(len(chain)):
        assert np.array_equal(res["0"]["hist%d" % s].view(np.uint32), res["1"]["hist%d" % s].view(np.uint32)), s
    ref, h = port.run_chain(chain, res["1"]["clips"][5], sample_rate=SAMPLE_RATE, block_size=BLOCK)
    assert_samples_close(res["1"]["out"][5], ref, "clip 5")
