"""TEST / BENCH INFRASTRUCTURE -- the seeded synthetic clips of SURVEY.md §8(d) in numpy.

Same formulas as the engine's generators (juicy-audio-plugins_b200/csrc/jb_synth.cpp, jb_synth_kernel), so the CPU
reference arm of bench.py can make its inputs WITHOUT loading the product library.  Noise and impulse clips are
bit-identical to the library's host generator; sweep and drum clips agree to rounding (numpy's float32 sin / exp).
"""
import numpy as np

SEED = 0x4A554943
KINDS = {"sweep": 0, "noise": 1, "impulse": 2, "drum": 3, "mixed": 4}


def _hash32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7feb352d)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846ca68b)
    x ^= x >> np.uint32(16)
    return x


def _unit_noise(seed, ch, n):
    with np.errstate(over="ignore"):
        h = _hash32(np.uint32(seed) ^ _hash32(n * np.uint32(2) + np.uint32(ch) + np.uint32(0x9E3779B9)))
    return (h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)


def synth_clips(kind, first_clip, n_clips, n_samples, n_channels=2, sample_rate=48000.0, seed=SEED):
    """float32 [n_clips][n_channels][n_samples]"""
    kind = KINDS[kind] if isinstance(kind, str) else int(kind)
    out = np.zeros((n_clips, n_channels, n_samples), dtype=np.float32)
    sr = np.float32(sample_rate)
    n = np.arange(n_samples, dtype=np.uint32)
    nf = n.astype(np.float32)
    f32 = np.float32
    with np.errstate(over="ignore"):
        for c in range(n_clips):
            clip_id = first_clip + c
            cseed = np.uint32(seed) ^ np.uint32((clip_id * 0x9E3779B9) & 0xFFFFFFFF)
            k = (clip_id & 3) if kind == 4 else kind
            if k == 0:
                T = f32(n_samples) / sr
                K = np.log(f32(1000.0))
                t = nf / sr
                ph = f32(2.0) * f32(np.pi) * f32(20.0) * T / K * (np.exp(t / T * K) - f32(1.0))
                left = f32(0.5) * np.sin(ph)
                right = f32(0.5) * np.sin(ph + f32(0.3))
            elif k == 1:
                left = f32(0.25) * _unit_noise(cseed, 0, n)
                right = f32(0.5) * (left + f32(0.25) * _unit_noise(cseed, 1, n))
            elif k == 2:
                period = 2400 + 37 * (clip_id & 63)
                left = np.where(n % np.uint32(period) == 0, f32(0.9), f32(0.0)).astype(np.float32)
                right = np.zeros(n_samples, dtype=np.float32)
                right[7::period] = f32(0.9)
            else:
                h = int(_hash32(np.array([cseed], dtype=np.uint32))[0])
                onset = 480 + h % 4800
                f0 = f32(45.0) + f32(45.0) * f32((h >> 13) & 1023) / f32(1023.0)
                m = np.maximum(nf - f32(onset), f32(0.0))
                body = f32(0.8) * np.exp(-m / f32(2400.0)) * np.sin(f32(2.0) * f32(np.pi) * f0 * m / sr)
                burst = f32(0.4) * np.exp(-m / f32(600.0))
                left = body + burst * _unit_noise(cseed, 0, n)
                right = f32(0.8) * left + f32(0.2) * burst * _unit_noise(cseed, 1, n)
                left = np.where(n >= onset, left, f32(0.0)).astype(np.float32)
                right = np.where(n >= onset, right, f32(0.0)).astype(np.float32)
            out[c, 0] = left
            if n_channels > 1:
                out[c, 1] = right
    return out
