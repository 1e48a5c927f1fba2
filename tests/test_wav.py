"""Audio file I/O: RIFF/WAVE <-> planar fp32 buffers, host only.  Checked against Python's own `wave`
module (an independent decoder / encoder) and against hand-built headers; conversion rule: integer / 2^(bits-1) in,
round(x * 2^(bits-1)) limited to +-(2^(bits-1) - 1) out."""
import struct
import wave

import numpy as np
import pytest


def python_wave_write(path, ints, width, rate):
    """ints: [channels][n] python-int array -> interleaved little-endian PCM through the stdlib encoder."""
    ch, n = ints.shape
    w = wave.open(str(path), "wb")
    w.setnchannels(ch)
    w.setsampwidth(width)
    w.setframerate(rate)
    frames = bytearray()
    for i in range(n):
        for c in range(ch):
            frames += int(ints[c, i]).to_bytes(width, "little", signed=True)
    w.writeframes(bytes(frames))
    w.close()


@pytest.mark.parametrize("width", [2, 3, 4])
def test_reads_what_pythons_wave_module_writes(width, jb, tmp_path):
    rng = np.random.default_rng(width)
    bits = 8 * width
    lim = 2 ** (bits - 1)
    ints = rng.integers(-lim, lim, size=(2, 1237), dtype=np.int64)
    ints[0, :3] = (-lim, lim - 1, 0)
    path = tmp_path / ("pcm%d.wav" % bits)
    python_wave_write(path, ints, width, 44100)
    info = jb.wav_info(path)
    assert info == {"n_channels": 2, "n_samples": 1237, "sample_rate": 44100.0, "bits_per_sample": bits, "is_float": False}
    audio, rate = jb.wav_read(path)
    assert rate == 44100.0
    want = (ints.astype(np.float64) / lim).astype(np.float32)
    assert np.array_equal(audio, want)


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_python_reads_back_what_the_library_writes(bits, jb, tmp_path):
    rng = np.random.default_rng(bits)
    x = rng.uniform(-1.2, 1.2, size=(2, 999)).astype(np.float32)   # beyond full scale: must be limited, not wrapped
    x[1, :4] = (1.0, -1.0, 0.0, 0.5)
    path = tmp_path / ("out%d.wav" % bits)
    jb.wav_write(path, x, 96000.0, bits)
    w = wave.open(str(path), "rb")
    assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, bits // 8, 96000, 999)
    raw = w.readframes(999)
    w.close()
    width = bits // 8
    got = np.array([int.from_bytes(raw[i:i + width], "little", signed=True) for i in range(0, len(raw), width)]).reshape(999, 2).T
    lim = 2 ** (bits - 1)
    want = np.clip(np.rint(x.astype(np.float64) * lim), -(lim - 1), lim - 1).astype(np.int64)
    assert np.array_equal(got, want)
    back, _ = jb.wav_read(path)                                     # and the library's own round trip
    assert np.array_equal(back, (want.astype(np.float64) / lim).astype(np.float32))


def test_float32_files_round_trip_bit_for_bit(jb, tmp_path):
    x = np.random.default_rng(5).standard_normal((1, 2048)).astype(np.float32) * 3.0
    x[0, :2] = (np.float32(1e-30), np.float32(-0.0))
    path = tmp_path / "f32.wav"
    jb.wav_write(path, x, 48000.0, 32, is_float=True)
    back, rate = jb.wav_read(path)
    assert rate == 48000.0 and np.array_equal(back.view(np.uint32), x.view(np.uint32))
    assert jb.wav_info(path)["is_float"]


def test_extensible_header_extra_chunks_and_errors(jb, tmp_path):
    # WAVE_FORMAT_EXTENSIBLE (what DAWs write for 24-bit), a LIST chunk with odd size before the data, a trailing chunk
    n, ch, bits = 10, 2, 24
    pcm = b"".join(int(v).to_bytes(3, "little", signed=True) for v in range(-10, 10))
    fmt = struct.pack("<HHIIHHHHI", 0xFFFE, ch, 48000, 48000 * ch * 3, ch * 3, bits, 22, bits, 3) + \
        struct.pack("<H", 1) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 5) + b"INFOx" + b"\0" + \
        b"data" + struct.pack("<I", len(pcm)) + pcm + b"cue " + struct.pack("<I", 4) + b"\0\0\0\0"
    path = tmp_path / "ext.wav"
    path.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    audio, rate = jb.wav_read(path)
    assert audio.shape == (2, 10) and rate == 48000.0
    assert np.array_equal(audio[0], (np.arange(-10, 10, 2) / 8388608.0).astype(np.float32))
    assert np.array_equal(audio[1], (np.arange(-9, 10, 2) / 8388608.0).astype(np.float32))
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFF\0\0\0\0WAVX")
    with pytest.raises(jb.JuicyBatchError):
        jb.wav_info(bad)
    with pytest.raises(jb.JuicyBatchError):
        jb.wav_info(tmp_path / "missing.wav")
    with pytest.raises(jb.JuicyBatchError):
        jb.wav_write(tmp_path / "x.wav", np.zeros((2, 4), np.float32), 48000.0, 12)


@pytest.mark.gpu
def test_render_files_end_to_end(jb, port, tmp_path):
    """WAV in -> engine -> WAV out, against the oracle fed the same decoded samples."""
    from conftest import assert_samples_close
    clips = jb.synth_clips("drum", 2, 3, 4096)
    for c in range(3):
        jb.wav_write(tmp_path / ("in%d.wav" % c), clips[c], 48000.0, 24)
    decoded = np.stack([jb.wav_read(tmp_path / ("in%d.wav" % c))[0] for c in range(3)])
    eng = jb.BatchProcessor(["JuicyPunch", "JuicyWidth"], 3)
    eng.prepareToPlay(48000.0, 512)
    out = eng.processBlock(decoded)
    eng.close()
    for c in range(3):
        jb.wav_write(tmp_path / ("out%d.wav" % c), out[c], 48000.0, 32, is_float=True)
        back, _ = jb.wav_read(tmp_path / ("out%d.wav" % c))
        ref, _ = port.run_chain(["JuicyPunch", "JuicyWidth"], decoded[c])
        assert_samples_close(back, ref, "file %d" % c)
