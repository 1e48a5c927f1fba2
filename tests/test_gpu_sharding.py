"""Sharding invariance on the GPU: the same clips rendered by ONE engine or split over several
engines (one per shard, as one process per GPU would hold them) give bitwise identical records and samples, and the
library's own NCCL gather (jb_comm_* / jb_gather_records, no torch anywhere) returns them in global clip order.
With one visible GPU the shards run one after the other on device 0 and the gather is exercised with a 1-rank
communicator; with two or more (gpurun --gpus 2) every shard gets its own device and the gather crosses NVLink."""
import numpy as np
import pytest

from cases import SAMPLE_RATE, BLOCK, FULL_CHAIN

pytestmark = pytest.mark.gpu


def _render(jb, chain, clips, device):
    eng = jb.BatchProcessor(chain, clips.shape[0], device=device)
    eng.prepareToPlay(SAMPLE_RATE, BLOCK)
    out = eng.processBlock(clips)
    return eng, out


@pytest.mark.parametrize("chain", [["JuicyPunch", "JuicyWidth"], FULL_CHAIN], ids=["punch-width", "full-chain"])
@pytest.mark.parametrize("world", [2, 4])
def test_records_and_samples_are_bitwise_identical_for_any_shard_count(chain, world, jb):
    n_clips, n = 256, 5 * BLOCK + 64
    clips = jb.synth_clips("mixed", 0, n_clips, n)
    whole, out_whole = _render(jb, chain, clips, 0)
    rec_whole = [whole.getLatestMetrics(s) for s in range(len(chain))]
    whole.close()
    n_dev = jb.device_count()
    outs, recs = [], [[] for _ in chain]
    for r in range(world):
        lo, hi = jb.shard_range(n_clips, r, world)
        eng, out = _render(jb, chain, clips[lo:hi], r % n_dev)
        outs.append(out)
        for s in range(len(chain)):
            recs[s].append(eng.getLatestMetrics(s))
        eng.close()
    assert np.array_equal(np.concatenate(outs), out_whole)
    for s in range(len(chain)):
        assert np.array_equal(np.concatenate(recs[s]), rec_whole[s]), "slot %d" % s


def test_native_gather_single_rank(jb):
    """jb_comm_init_rank with one rank + jb_gather_records_host: NCCL is loaded by the library (dlopen), not by torch."""
    chain = ["JuicyPunch", "JuicyWidth"]
    clips = jb.synth_clips("drum", 0, 70, 3 * BLOCK)
    eng, _ = _render(jb, chain, clips, 0)
    eng.comm_init_rank(jb.comm_unique_id(), 1, 0)
    got = eng.gather_records_host(1)
    assert np.array_equal(got, eng.getLatestMetrics(1))
    eng.close()


def test_native_gather_across_devices(jb):
    """One process, one engine per GPU (ncclCommInitAll): every engine receives every shard's records, in clip order."""
    n_dev = jb.device_count()
    if n_dev < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = min(n_dev, 4)
    chain = ["JuicyPunch", "JuicyWidth"]
    per, n = 96, 4 * BLOCK
    clips = jb.synth_clips("mixed", 0, world * per, n)
    whole, _ = _render(jb, chain, clips, 0)
    want = whole.getLatestMetrics(1)
    whole.close()
    engines = [_render(jb, chain, clips[r * per:(r + 1) * per], r)[0] for r in range(world)]
    jb.comm_init_all(engines)
    import ctypes
    L = jb.lib()
    pitch = engines[0].record_pitch()
    bufs = [jb.DeviceBuffer(world * 16 * pitch * 4, r) for r in range(world)]
    hs = (ctypes.c_void_p * world)(*[e._h for e in engines])
    outs = (ctypes.c_void_p * world)(*[b.ptr.value for b in bufs])
    jb._check(L.jb_gather_records_all(hs, world, 1, outs))
    for r, e in enumerate(engines):
        e.synchronize()
        flat = bufs[r].download((world, 16, pitch))
        got = np.concatenate([flat[k, :, :per].T for k in range(world)])
        assert np.array_equal(got, want), "engine %d" % r
    for e in engines:
        e.close()
    for b in bufs:
        b.free()
