"""Meter-panel statistics: JuicyMeterPanel::setMetrics / smoothValue / updateStats
(src/shared/JuicyMeterPanel.cpp:3-34,54-71) over the per-block record history of a render.

CPU: the oracle's restatement (oracle/juicy_oracle.c: jo_meter_run) bit for bit against the golden
vectors made from the reference's own panel, and against the compiled panel itself where present.
GPU: jb_meter_statistics (device reduction over the engine's history) bit for bit against the oracle
fed the same records, and against the golden vectors within the record tolerance."""
import os

import numpy as np
import pytest

from cases import GOLDEN_CASES, N_SAMPLES, SAMPLE_RATE, BLOCK, apply_case_settings, load_golden
from conftest import METRIC_TOL


def load_meter_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_meter_v1.npz"))


def meter_sequences():
    z, _ = load_golden()
    zm = load_meter_golden()
    for key in zm.files:
        if not key.startswith("meter/"):
            continue
        name = key[len("meter/"):]
        seq = zm["in/" + name] if name.startswith("synthetic/") else z[name]
        yield name, seq, zm[key]


def test_port_meter_matches_golden_bit_for_bit(port):
    n = 0
    for name, seq, want in meter_sequences():
        got = port.meter_run(seq)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
        n += 1
    assert n >= len(GOLDEN_CASES) + 6


def test_port_meter_matches_compiled_panel(port, refhost):
    if not refhost.meter_available():
        pytest.skip("oracle/_ref/libjuicy_ref_MeterPanel.so not built")
    rng = np.random.default_rng(77)
    for n in (0, 1, 3, 50, 2000):
        r = rng.uniform(-0.5, 1.5, (n, 16)).astype(np.float32)
        r[:, 0:3] = rng.uniform(-5.0, 100.0, (n, 3)).astype(np.float32)
        assert np.array_equal(port.meter_run(r).view(np.uint32), refhost.meter_run(r).view(np.uint32)), n


def test_meter_known_answers(port):
    """One record: stats collapse onto the clamped value, bars move 28 % of the way up (12 % down)."""
    r = np.zeros((1, 16), dtype=np.float32)
    r[0, 0] = 50.0          # score; pre/post 0 -> both fall back to score
    r[0, 8] = 1.5           # punch above 1 -> stat clamps, bar does not
    r[0, 12] = 0.5          # monoSafety starts at 1 -> falls with alpha 0.12
    out = port.meter_run(r)
    assert out[0] == pytest.approx(14.0) and out[1] == pytest.approx(14.0) and out[2] == pytest.approx(14.0)
    assert out[3] == pytest.approx(0.42)
    assert out[7] == pytest.approx(0.94)
    assert tuple(out[8:11]) == (1.0, 1.0, 1.0)
    assert out[38] == 1.0


# ------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
@pytest.mark.parametrize("chain,stride", [(["JuicyPunch", "JuicyWidth"], 1), (["JuicyInfer"], 1), (["JuicySaturator"], 5),
                                          (["JuicyPunch", "JuicySaturator", "JuicyTexture", "JuicyWidth", "JuicyMotion",
                                            "JuicyCohere", "JuicyInfer"], 1)],
                         ids=["punch-width", "infer", "saturator-stride5", "full-chain"])
def test_engine_meter_statistics_match_oracle(chain, stride, jb, port):
    n_clips, n = 70, 20 * BLOCK + 100
    clips = jb.synth_clips("mixed", 3, n_clips, n)
    eng = jb.BatchProcessor(chain, n_clips)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.enableHistory(32)
    eng.processBlock(clips)
    assert eng.historyBlocks() == 21
    for slot in range(len(chain)):
        hist = eng.getHistory(slot)                 # [block][clip][16], the engine's own records
        got = eng.meterStatistics(slot, block_stride=stride)
        for c in range(n_clips):
            want = port.meter_run(hist[::stride, c, :])
            assert np.array_equal(got[c].view(np.uint32), want.view(np.uint32)), "slot %d clip %d" % (slot, c)
    # a sub-range, as a host that opens the editor mid-render would see it
    got = eng.meterStatistics(0, first_block=4, n_blocks=9)
    hist = eng.getHistory(0)
    for c in (0, 33, 69):
        assert np.array_equal(got[c].view(np.uint32), port.meter_run(hist[4:13, c, :]).view(np.uint32))
    # empty range: the panel's initial state
    got = eng.meterStatistics(0, first_block=0, n_blocks=0)
    assert np.array_equal(got[5].view(np.uint32), port.meter_run(np.zeros((0, 16), np.float32)).view(np.uint32))
    with pytest.raises(jb.JuicyBatchError):
        eng.meterStatistics(0, first_block=10, n_blocks=40)
    eng.close()


@pytest.mark.gpu
def test_engine_meter_statistics_match_golden(jb):
    """End to end against the reference: render the golden clips, reduce on the device, compare with the
    panel fed the reference's own records (tolerance of the records themselves)."""
    z, _ = load_golden()
    zm = load_meter_golden()
    for case in GOLDEN_CASES[::4]:
        x = z["in/%s/%d" % (case["input"], case["clip"])]
        eng = jb.BatchProcessor(case["chain"], 1)
        apply_case_settings(eng, case)
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        eng.enableHistory(16)
        eng.processBlock(x[None])
        for slot in range(len(case["chain"])):
            want = zm["meter/hist/%s/%d" % (case["name"], slot)]
            got = eng.meterStatistics(slot)[0]
            assert np.abs(got[0:3] - want[0:3]).max() <= METRIC_TOL, case["name"]       # 0..100 scores
            assert np.abs(got[3:38] - want[3:38]).max() <= METRIC_TOL, case["name"]     # 0..1 bars and stats
            assert got[38] == want[38]
        eng.close()


@pytest.mark.gpu
def test_meter_statistics_need_history(jb):
    eng = jb.BatchProcessor(["JuicyWidth"], 4)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    with pytest.raises(jb.JuicyBatchError):
        eng.meterStatistics(0)
    eng.close()
