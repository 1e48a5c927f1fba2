"""Saturator / Punch with the C library's own tanh / pow (JB_MATH_EXACT, csrc/jb_libm.h): bit-identical samples, and
chains whose Texture resonators amplify any upstream difference stay inside the stated tolerance."""
import numpy as np
import pytest

from cases import SAMPLE_RATE, BLOCK, FULL_CHAIN, load_golden, GOLDEN_CASES, apply_case_settings
from conftest import assert_samples_close, assert_records_close

pytestmark = pytest.mark.gpu


def oracle_render(port, chain, clips, programs=None, params=None):
    outs, hists = [], []
    for x in clips:
        o, h = port.run_chain(chain, x, sample_rate=SAMPLE_RATE, block_size=BLOCK, programs=programs, params=params)
        outs.append(o)
        hists.append(h)
    return np.stack(outs), hists


@pytest.mark.parametrize("plugin,program", [("JuicySaturator", 0), ("JuicySaturator", 3), ("JuicyPunch", 0), ("JuicyPunch", 2),
                                            ("JuicyPunch", 4)])
def test_exact_mode_is_bit_identical(plugin, program, jb, port):
    n_clips, n = 40, 3 * BLOCK + 64
    clips = jb.synth_clips("mixed", 21, n_clips, n)
    clips *= np.linspace(0.2, 2.5, n_clips, dtype=np.float32)[:, None, None]   # well into the shapers' curved range
    eng = jb.BatchProcessor(plugin, n_clips)
    eng.setCurrentProgram(program)
    eng.set_math_mode("exact")
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(clips)
    ref, hists = oracle_render(port, [plugin], clips, programs={0: program})
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), \
        "max |gpu-ref| = %g" % float(np.abs(out - ref).max())
    assert_records_close(eng.getLatestMetrics(0), np.stack([h[0][-1] for h in hists]), plugin)
    if plugin == "JuicyPunch":
        assert eng.path_launches()[0] > 0       # small Punch batches: the cooperative kernel's exact-math instantiation
    eng.close()


def test_exact_mode_matches_golden_bit_for_bit(jb):
    """The committed vectors come from the reference's own C++: Saturator / Punch cases must match to the bit."""
    z, _ = load_golden()
    n = 0
    for case in GOLDEN_CASES:
        if case["chain"] not in (["JuicySaturator"], ["JuicyPunch"]):
            continue
        x = z["in/%s/%d" % (case["input"], case["clip"])]
        eng = jb.BatchProcessor(case["chain"], 1)
        apply_case_settings(eng, case)
        eng.set_math_mode("exact")
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        out = eng.processBlock(x[None])[0]
        eng.close()
        assert np.array_equal(out.view(np.uint32), z["out/" + case["name"]].view(np.uint32)), case["name"]
        n += 1
    assert n >= 12


@pytest.mark.parametrize("material", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("chain", [["JuicyPunch", "JuicyTexture"], ["JuicySaturator", "JuicyTexture"], FULL_CHAIN],
                         ids=["punch-texture", "saturator-texture", "full-chain"])
def test_resonant_materials_downstream_stay_in_tolerance(chain, material, jb, port):
    """Auto mode: a shaper that feeds another plugin runs the exact routines.  Texture's metal / wood / plastic resonators
    amplify a 1e-6 input difference ~200x: with fast math this test fails at 1e-4 .. 4e-3 of clip peak."""
    n_clips, n = 24, 2 * BLOCK + 128
    slot = chain.index("JuicyTexture")
    clips = jb.synth_clips("mixed", 11, n_clips, n)
    eng = jb.BatchProcessor(chain, n_clips)
    eng.setParameter("material", float(material), slot)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(clips)
    ref, hists = oracle_render(port, chain, clips, params={slot: {"material": float(material)}})
    assert_samples_close(out, ref, "+".join(chain))
    for s in range(len(chain)):
        assert_records_close(eng.getLatestMetrics(s), np.stack([h[s][-1] for h in hists]), "slot %d" % s)
    eng.close()


def test_auto_mode_is_exact_where_a_shaper_feeds_another_plugin(jb, port):
    """JB_MATH_AUTO: Punch -> Width runs the cooperative kernel's exact instantiation (C library pow / tanh, the reference's
    unfused operand order) and the samples are the reference's bit for bit, so Width's `corrProxy < -0.1f` decisions
    (JuicyWidth/PluginProcessor.cpp:109-112) are the reference's; JB_MATH_FAST stays inside the plugin tolerance on
    these clips.  A shaper that ends the chain keeps the fast routines."""
    clips = jb.synth_clips("mixed", 0, 64, 4 * BLOCK)
    ref, _ = oracle_render(port, ["JuicyPunch", "JuicyWidth"], clips)
    eng = jb.BatchProcessor(["JuicyPunch", "JuicyWidth"], 64)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(clips)
    assert eng.path_launches()[0] > 0
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), "max |gpu-ref| = %g" % float(np.abs(out - ref).max())
    eng.set_math_mode("fast")
    eng.reset()
    fast = eng.processBlock(clips)
    assert not np.array_equal(fast, out)
    assert_samples_close(fast, ref, "punch-width fast")
    eng.close()
    for chain in (["JuicyWidth", "JuicySaturator"], ["JuicySaturator"]):
        eng = jb.BatchProcessor(chain, 64)
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        auto = eng.processBlock(clips)
        eng.set_math_mode("fast")
        eng.reset()
        assert np.array_equal(eng.processBlock(clips), auto), chain
        eng.close()


@pytest.mark.parametrize("material", [0, 4], ids=["gel", "flesh"])
def test_texture_gel_and_flesh_are_bit_identical(material, jb, port):
    """Their std::tanh is the C library's own algorithm on the device (jb_libm.h), like the other three materials' cos."""
    clips = jb.synth_clips("mixed", 3, 48, 4 * BLOCK + 32)
    eng = jb.BatchProcessor(["JuicyTexture"], 48)
    eng.setParameter("material", float(material))
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(clips)
    eng.close()
    ref, _ = oracle_render(port, ["JuicyTexture"], clips, params={0: {"material": float(material)}})
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), "max |gpu-ref| = %g" % float(np.abs(out - ref).max())
