"""jb_process_host_pcm16: 16-bit PCM host buffers, converted on the device with jb_wav.cpp's rule (s / 32768 in,
round-half-even(v * 32768) limited to +-32767 out) around the unchanged fp32 render."""
import numpy as np
import pytest

from cases import SAMPLE_RATE, BLOCK, FULL_CHAIN

pytestmark = pytest.mark.gpu


def _quantize(x):
    q = np.rint(x.astype(np.float64) * 32768.0)          # numpy rounds half to even, like nearbyint
    return np.clip(q, -32767, 32767).astype(np.int16)


@pytest.mark.parametrize("chain,n", [(["JuicyPunch", "JuicyWidth"], 5 * BLOCK + 96), (FULL_CHAIN, 3 * BLOCK + 37), (["JuicyInfer"], 2 * BLOCK)])
def test_pcm16_path_equals_float_path_with_the_wav_conversion_rule(chain, n, jb, monkeypatch):
    n_clips = 70
    clips = jb.synth_clips("mixed", 2, n_clips, n)
    clips *= np.linspace(0.5, 3.0, n_clips, dtype=np.float32)[:, None, None]     # loud clips clip at +-32767
    pcm_in = _quantize(clips)
    as_float = pcm_in.astype(np.float32) * np.float32(1.0 / 32768.0)
    ref = jb.BatchProcessor(chain, n_clips)
    ref.prepareToPlay(SAMPLE_RATE, BLOCK)
    want = _quantize(ref.processBlock(as_float))
    rec_want = ref.getLatestMetrics(len(chain) - 1)
    ref.close()
    for env in ({}, {"JB_HOST_SLICE_MIB": "1", "JB_HOST_PASS_MIB": "1"}):       # one slice, and many slices / passes
        for k in ("JB_HOST_PASS_MIB", "JB_HOST_SLICE_MIB"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = jb.BatchProcessor(chain, n_clips)
        eng.prepareToPlay(SAMPLE_RATE, BLOCK)
        got = eng.processBlockPcm16(pcm_in)
        assert got.dtype == np.int16 and np.array_equal(got, want)
        assert np.array_equal(eng.getLatestMetrics(len(chain) - 1), rec_want)
        eng.close()


def test_pcm16_conversion_matches_the_wav_writer(jb, tmp_path):
    """The device's float -> int16 is what jb_wav_write produces for the same floats (Infer with trim = 0 leaves the audio alone)."""
    n = 4 * BLOCK
    x = (np.random.default_rng(3).standard_normal((1, 2, n)) * 0.4).astype(np.float32)
    x[0, 0, :8] = [1.0, -1.0, 0.99998474, -0.99998474, 1.5, -1.5, 0.5 / 32768.0, 1.5 / 32768.0]   # limits and ties
    pcm = _quantize(x)
    eng = jb.BatchProcessor(["JuicyInfer"], 1)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlockPcm16(pcm)
    eng.close()
    assert np.array_equal(out, pcm)        # s / 32768 -> round(v * 32768) is the identity on legal samples
    path = str(tmp_path / "q.wav")
    jb.wav_write(path, x[0], SAMPLE_RATE, 16)
    back, _ = jb.wav_read(path)
    assert np.array_equal(_quantize(x[0]).astype(np.float32) / np.float32(32768.0), back)
