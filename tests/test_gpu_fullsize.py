"""Full-size runs of the BASELINE.json configurations on one B200, checked through size-independent
properties (the CPU oracle is far too slow for 4096 x 48000) plus oracle parity on a strided subset."""
import numpy as np
import pytest

from cases import SAMPLE_RATE, BLOCK
from conftest import assert_samples_close, assert_records_close

pytestmark = pytest.mark.gpu

N_CLIPS, N_SAMPLES = 4096, 48000


def test_config2_punch_width_full_size(jb, port):
    """configs[1]: Punch -> Width on 4096 stereo drum-hit clips of 1 s."""
    chain = ["JuicyPunch", "JuicyWidth"]
    nbytes = N_CLIPS * 2 * N_SAMPLES * 4
    d = jb.DeviceBuffer(nbytes)
    jb.synth_fill_device(d.ptr.value, "drum", 0, N_CLIPS, 2, N_SAMPLES)
    # duplicate clip 17 into lane 4001: identical inputs must give identical outputs anywhere in the batch
    clip_bytes = 2 * N_SAMPLES * 4
    tmp = np.empty((2, N_SAMPLES), dtype=np.float32)
    jb._check(jb.lib().jb_copy_to_host(0, tmp.ctypes.data, d.ptr.value + 17 * clip_bytes, clip_bytes))
    jb._check(jb.lib().jb_copy_to_device(0, d.ptr.value + 4001 * clip_bytes, tmp.ctypes.data, clip_bytes))
    subset = list(range(0, N_CLIPS, 512)) + [17, 4001, N_CLIPS - 1]
    inputs = {}
    for c in subset:
        buf = np.empty((2, N_SAMPLES), dtype=np.float32)
        jb._check(jb.lib().jb_copy_to_host(0, buf.ctypes.data, d.ptr.value + c * clip_bytes, clip_bytes))
        inputs[c] = buf
    eng = jb.BatchProcessor(chain, N_CLIPS)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    eng.process_device(d.ptr.value, d.ptr.value, N_SAMPLES)
    eng.synchronize()
    rec = [eng.getLatestMetrics(s) for s in range(2)]
    outs = {}
    for c in subset:
        buf = np.empty((2, N_SAMPLES), dtype=np.float32)
        jb._check(jb.lib().jb_copy_to_host(0, buf.ctypes.data, d.ptr.value + c * clip_bytes, clip_bytes))
        outs[c] = buf
    assert np.array_equal(outs[17], outs[4001])
    assert np.array_equal(rec[1][17], rec[1][4001])
    for r in rec:
        assert np.isfinite(r).all()
        assert (r[:, 0] >= 0).all() and (r[:, 0] <= 100).all()
    for c in subset:
        ref, hists = port.run_chain(chain, inputs[c], sample_rate=SAMPLE_RATE, block_size=BLOCK)
        assert_samples_close(outs[c], ref, "clip %d" % c)
        for s in range(2):
            assert_records_close(rec[s][c], hists[s][-1], "clip %d slot %d" % (c, s))
    eng.close()
    d.free()
