"""Pins the CPU oracle (oracle/juicy_oracle.c, "port") to the reference:
  * bit for bit against the committed golden vectors, which tests/golden/make_golden.py produced
    from the reference's own unmodified C++ (oracle/_ref);
  * bit for bit against oracle/_ref itself whenever those libraries are present (always in the
    build container; on the GPU box they travel with the snapshot)."""
import numpy as np
import pytest

from cases import GOLDEN_CASES, N_SAMPLES, SAMPLE_RATE, BLOCK, PLUGINS, case_input, load_golden


@pytest.fixture(scope="module")
def golden():
    return load_golden()


@pytest.mark.parametrize("case", GOLDEN_CASES, ids=[c["name"] for c in GOLDEN_CASES])
def test_port_matches_golden_bit_for_bit(case, golden, port):
    z, meta = golden
    assert meta["n_samples"] == N_SAMPLES and meta["block"] == BLOCK
    x = z["in/%s/%d" % (case["input"], case["clip"])]
    out, hists = port.run_chain(case["chain"], x, sample_rate=SAMPLE_RATE, block_size=BLOCK,
                                programs=case.get("programs"), params=case.get("params"))
    ref = z["out/" + case["name"]]
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), "samples differ from the reference's"
    for slot, h in enumerate(hists):
        href = z["hist/%s/%d" % (case["name"], slot)]
        assert np.array_equal(h.view(np.uint32), href.view(np.uint32)), "metric records differ (slot %d)" % slot


def test_host_generator_reproduces_golden_inputs(golden):
    """The inputs the GPU parity tests regenerate are the ones the golden outputs were made from."""
    z, _ = golden
    for case in GOLDEN_CASES:
        x = z["in/%s/%d" % (case["input"], case["clip"])]
        assert np.array_equal(case_input(case).view(np.uint32), x.view(np.uint32)), case["name"]


def test_silence_scores_exactly_forty(port):
    """a silent block with rested filter state scores 40.0 (known answer)."""
    p = port.PortPlugin("JuicyInfer")
    p.prepare()
    _, hist = p.process(np.zeros((2, 1024), dtype=np.float32))
    assert hist[-1, 0] == pytest.approx(40.0, abs=1e-6)


@pytest.mark.parametrize("plugin", PLUGINS)
def test_port_matches_compiled_reference(plugin, port, refhost):
    if not refhost.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this box and no prebuilt libraries)")
    rng = np.random.default_rng(1234 + PLUGINS.index(plugin))
    n = 3 * BLOCK + 77
    x = (0.4 * rng.standard_normal((2, n))).astype(np.float32)
    x[:, 700:760] *= 4.0  # an onset and some clipping
    for block in (BLOCK, 64):
        a = refhost.RefPlugin(plugin, 2, SAMPLE_RATE, block)
        b = port.PortPlugin(plugin, 2, SAMPLE_RATE, block)
        assert a.param_ids() == b.param_ids()
        assert a.params() == b.params()
        if plugin == "JuicyTexture":
            settings = [{"material": float(m)} for m in range(5)]
        else:
            settings = [{}]
        for s in settings:
            for k, v in s.items():
                a.set_param(k, v)
                b.set_param(k, v)
            a.prepare()
            b.prepare()
            oa, ha = a.process(x)
            ob, hb = b.process(x)
            assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32)), (plugin, s, block)
            assert np.array_equal(ha.view(np.uint32), hb.view(np.uint32)), (plugin, s, block)


@pytest.mark.parametrize("plugin", ("JuicyInfer", "JuicyPunch", "JuicySaturator", "JuicyWidth"))
def test_port_programs_match_compiled_reference(plugin, port, refhost):
    if not refhost.available():
        pytest.skip("oracle/_ref not built")
    a = refhost.RefPlugin(plugin)
    b = port.PortPlugin(plugin)
    assert a.num_programs() == b.num_programs() == 5
    for i in range(5):
        a.set_program(i)
        b.set_program(i)
        assert a.program_name(i) == b.program_name(i)
        assert a.params() == b.params()


@pytest.mark.parametrize("plugin", PLUGINS)
def test_port_matches_compiled_reference_on_a_mono_bus(plugin, port, refhost):
    """isBusesLayoutSupported allows mono == mono (e.g. JuicyPunch/PluginProcessor.cpp:48-54): one channel loop, the analyzer
    reads right = left, Width returns before its DSP.  The port must follow the reference there too, bit for bit."""
    if not refhost.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this box and no prebuilt libraries)")
    rng = np.random.default_rng(4321 + PLUGINS.index(plugin))
    n = 3 * BLOCK + 77
    x = (0.4 * rng.standard_normal((1, n))).astype(np.float32)
    x[:, 700:760] *= 4.0
    mats = range(5) if plugin == "JuicyTexture" else (None,)
    for m in mats:
        a = refhost.RefPlugin(plugin, 1, SAMPLE_RATE, BLOCK)
        b = port.PortPlugin(plugin, 1, SAMPLE_RATE, BLOCK)
        if m is not None:
            a.set_param("material", float(m))
            b.set_param("material", float(m))
        a.prepare()
        b.prepare()
        oa, ha = a.process(x)
        ob, hb = b.process(x)
        assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32)), (plugin, m)
        assert np.array_equal(ha.view(np.uint32), hb.view(np.uint32)), (plugin, m)
        a.close()
        b.close()
