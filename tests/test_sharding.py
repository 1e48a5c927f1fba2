"""N > 1 host logic on CPU: clip sharding (jb_shard_range, the library's own) + the record gather, 2 ranks over gloo.

Each rank renders ITS clip range (here with the CPU oracle standing in for the GPU, which this
container does not have), packs the records in the engine's device layout and all_gathers them;
the result must equal one process rendering every clip.  The GPU version of the same flow is
jb_comm_init_rank + jb_gather_records (NCCL inside the library: tests/test_gpu_sharding.py, bench.py --gpus N).
torch appears here only as the CPU stand-in for that collective."""
import importlib.util
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT, load_juicy_batch


REC = 16  # floats per record (jb_metrics)


class _Sharding:
    """Shard arithmetic from the library (jb_shard_range); the record block layout [16][pitch] of jb_metrics_device."""

    @staticmethod
    def clip_pitch(n_clips):
        return (int(n_clips) + 31) // 32 * 32

    @staticmethod
    def shard_range(n_clips, rank, world):
        jb = load_juicy_batch()
        try:
            return jb.shard_range(n_clips, rank, world)
        except jb.JuicyBatchError as exc:
            raise ValueError(str(exc))

    @staticmethod
    def pack_records_soa(records, pitch):
        records = np.asarray(records, dtype=np.float32)
        out = np.zeros((REC, pitch), dtype=np.float32)
        out[:, :records.shape[0]] = records.T
        return out.reshape(-1)

    @staticmethod
    def unpack_gathered(flat, counts, pitch):
        flat = np.asarray(flat, dtype=np.float32).reshape(len(counts), REC, pitch)
        return np.concatenate([flat[r, :, :counts[r]].T for r in range(len(counts))], axis=0)

    @staticmethod
    def gather_records(local_soa, world, dist):
        import torch
        if world == 1:
            return local_soa.clone()
        out = torch.empty(world * local_soa.numel(), dtype=local_soa.dtype, device=local_soa.device)
        dist.all_gather_into_tensor(out, local_soa)
        return out


def load_sharding():
    return _Sharding


@pytest.mark.parametrize("n_clips,world", [(4096, 1), (4096, 8), (10, 4), (3, 8), (262144, 8), (7, 2)])
def test_shard_ranges_partition_the_clips(n_clips, world):
    sh = load_sharding()
    ranges = [sh.shard_range(n_clips, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_clips
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [hi - lo for lo, hi in ranges]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(n_clips, world, world)


def test_record_soa_round_trip():
    sh = load_sharding()
    rng = np.random.default_rng(1)
    counts = [5, 4]
    pitch = sh.clip_pitch(max(counts))
    recs = [rng.random((c, 16), dtype=np.float32) for c in counts]
    flat = np.concatenate([sh.pack_records_soa(r, pitch) for r in recs])
    assert np.array_equal(sh.unpack_gathered(flat, counts, pitch), np.concatenate(recs))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port_no, n_clips, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import port
    jb = load_juicy_batch()
    sh = load_sharding()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sh.shard_range(n_clips, rank, world)
    clips = jb.synth_clips("drum", lo, hi - lo, n)           # shard r's clips are clips [lo, hi) of the job
    chain = ["JuicyPunch", "JuicyWidth"]
    recs = np.stack([port.run_chain(chain, x)[1][-1][-1] for x in clips])
    counts = [b - a for a, b in (sh.shard_range(n_clips, r, world) for r in range(world))]
    pitch = sh.clip_pitch(max(counts))
    local = torch.from_numpy(sh.pack_records_soa(recs, pitch))
    gathered = sh.gather_records(local, world, dist)
    everything = sh.unpack_gathered(gathered.numpy(), counts, pitch)
    np.save(os.path.join(out_dir, "gathered_%d.npy" % rank), everything)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_render_and_gather_equals_single_process(tmp_path, port):
    import torch.multiprocessing as mp
    jb = load_juicy_batch()
    n_clips, n, world = 7, 1100, 2
    mp.start_processes(_rank_main, args=(world, _free_port(), n_clips, n, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    clips = jb.synth_clips("drum", 0, n_clips, n)
    want = np.stack([port.run_chain(["JuicyPunch", "JuicyWidth"], x)[1][-1][-1] for x in clips])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "gathered_%d.npy" % r))
        assert got.shape == want.shape
        assert np.array_equal(got, want), "rank %d gathered records differ from the single-process render" % r
