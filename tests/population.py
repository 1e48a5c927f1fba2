"""Population-level parity: EVERY clip of a full-size BASELINE.json configuration against the CPU oracle.

TEST INFRASTRUCTURE.  The GPU renders the whole batch in one call (so the engine takes exactly the code path it takes
at that size: kernel choice, samples per trip, pipelining); the batch then comes back to the host in chunks and a pool
of oracle processes -- one per host core, each driving the compiled reference (oracle/_ref) or, where that build is
absent, the C port -- re-renders every clip block by block and compares inside the worker:

  * every sample:  |gpu - ref| <= 1e-5 * max|ref| over the clip   (BASELINE.json north_star)
  * every record of every block of every plugin: <= 0.01 absolute (needs the engine's record history)

Only small per-clip error arrays travel back.  Used by tests/test_gpu_population.py and tools/parity_population.py.
Worker mode: `python tests/population.py --worker` (commands as JSON lines on stdin; audio in POSIX shared memory).
"""
import json
import os
import subprocess
import sys
import time
from multiprocessing import shared_memory

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SAMPLE_TOL = 1.0e-5
METRIC_TOL = 0.01


def host_procs():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------ worker side

def _worker_main():
    from oracle import refhost, port
    plugins = None
    spec = None
    shms = {}

    def attach(name):
        if name not in shms:
            shms[name] = shared_memory.SharedMemory(name=name)
            try:  # the driver owns (and unlinks) the segment; keep this process's resource tracker out of it
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shms[name]._name, "shared_memory")
            except Exception:
                pass
        return shms[name]

    for line in sys.stdin:
        cmd = json.loads(line)
        if cmd["op"] == "quit":
            break
        if cmd["op"] == "config":
            spec = cmd
            cls = refhost.RefPlugin if cmd["kind"] == "reference" else port.PortPlugin

            def make_plugins(cfg=cmd, cls=cls):
                # a FRESH instance per clip, like the engine's clips: some members are set by the constructor only and
                # survive prepareToPlay (Motion's LCG, JuicyMotion/PluginProcessor.h:65; Cohere's learnt targets, .h:55-57)
                ps = [cls(p, 2, cfg["sample_rate"], cfg["block"]) for p in cfg["chain"]]
                for slot, kv in (cfg.get("params") or {}).items():
                    for k, v in kv.items():
                        ps[int(slot)].set_param(k, v)
                return ps

            make_plugins()  # fail early if the oracle library is missing
            sys.stdout.write("ok\n")
            sys.stdout.flush()
            continue
        # op == "check": clips [lo, hi) of the chunk
        k, n, nb, L = cmd["chunk"], spec["samples"], cmd["blocks"], len(spec["chain"])
        lo, hi = cmd["lo"], cmd["hi"]
        a_in = np.ndarray((k, 2, n), dtype=np.float32, buffer=attach(cmd["in"]).buf)
        a_gpu = np.ndarray((k, 2, n), dtype=np.float32, buffer=attach(cmd["gpu"]).buf)
        a_hist = np.ndarray((L, nb, k, 16), dtype=np.float32, buffer=attach(cmd["hist"]).buf) if cmd.get("hist") else None
        a_last = np.ndarray((L, k, 16), dtype=np.float32, buffer=attach(cmd["last"]).buf) if cmd.get("last") else None
        res = np.ndarray((k, 6), dtype=np.float64, buffer=attach(cmd["res"]).buf)
        per_clip = spec.get("per_clip")  # {"slot": s, "id": pid, "mod": m}: value = absolute clip index mod m
        for c in range(lo, hi):
            x = a_in[c]
            rec_err, rec_block, rec_slot, rec_field, rec_ref = 0.0, -1, -1, -1, 0.0
            plugins = make_plugins()
            for s, p in enumerate(plugins):
                if per_clip and per_clip["slot"] == s:
                    p.set_param(per_clip["id"], float((cmd["first_clip"] + c) % per_clip["mod"]))
                p.prepare()
                x, h = p.process(x)
                if a_hist is not None:
                    d = np.abs(a_hist[s, :h.shape[0], c, :].astype(np.float64) - h.astype(np.float64))
                    # NaN-safe: a non-finite GPU record is an error of +inf
                    d = np.where(np.isfinite(d), d, np.inf)
                    m = float(d.max())
                    if m > rec_err:
                        bi, fi = np.unravel_index(int(d.argmax()), d.shape)
                        rec_err, rec_block, rec_slot, rec_field, rec_ref = m, int(bi), s, int(fi), float(h[bi, fi])
                elif a_last is not None:
                    d = np.abs(a_last[s, c, :].astype(np.float64) - h[-1].astype(np.float64))
                    d = np.where(np.isfinite(d), d, np.inf)
                    m = float(d.max())
                    if m > rec_err:
                        rec_err, rec_block, rec_slot, rec_field, rec_ref = m, h.shape[0] - 1, s, int(d.argmax()), float(h[-1, int(d.argmax())])
            peak = max(float(np.abs(x).max()), 1.0e-30)
            g = a_gpu[c]
            err = np.abs(g.astype(np.float64) - x.astype(np.float64))
            worst = float(err.max()) if np.isfinite(g).all() else float("inf")
            res[c, 0] = worst / peak
            res[c, 1] = rec_err
            res[c, 2] = rec_block
            res[c, 3] = rec_slot
            res[c, 4] = rec_field
            res[c, 5] = rec_ref
            for p in plugins:
                p.close()
        sys.stdout.write("done\n")
        sys.stdout.flush()
    for s in shms.values():
        s.close()
    return 0


# ------------------------------------------------------------------------------------------------ driver side

class OraclePool:
    """A pool of oracle worker processes configured for one chain / parameter setting."""

    def __init__(self, chain, n_samples, sample_rate=48000.0, block=512, params=None, per_clip=None, procs=None, kind=None):
        from oracle import refhost, port
        if kind is None:
            kind = "reference" if refhost.available() else "port"
        if kind == "port":
            port.lib()  # builds it if needed, before the workers race for it
        self.kind = kind
        self.chain = list(chain)
        self.n = int(n_samples)
        self.block = int(block)
        self.n_blocks = (self.n + self.block - 1) // self.block
        self.procs = procs or host_procs()
        env = dict(os.environ)
        env["CUDA_VISIBLE_DEVICES"] = ""
        self.workers = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker"], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, text=True, env=env) for _ in range(self.procs)]
        cfg = {"op": "config", "kind": kind, "chain": self.chain, "samples": self.n, "sample_rate": float(sample_rate),
               "block": self.block, "params": {str(k): v for k, v in (params or {}).items()}, "per_clip": per_clip}
        for w in self.workers:
            w.stdin.write(json.dumps(cfg) + "\n")
            w.stdin.flush()
        for w in self.workers:
            if w.stdout.readline().strip() != "ok":
                raise RuntimeError("oracle worker failed to configure")
        self._shm = {}

    def _buf(self, key, nbytes):
        s = self._shm.get(key)
        if s is None or s.size < nbytes:
            if s is not None:
                s.close()
                s.unlink()
            s = shared_memory.SharedMemory(create=True, size=max(nbytes, 16))
            self._shm[key] = s
        return s

    def arrays(self, k, with_hist):
        """Shared chunk buffers for k clips: (in, gpu, hist-or-last) numpy views to be filled by the caller."""
        L, nb = len(self.chain), self.n_blocks
        a_in = np.ndarray((k, 2, self.n), dtype=np.float32, buffer=self._buf("in", k * 2 * self.n * 4).buf)
        a_gpu = np.ndarray((k, 2, self.n), dtype=np.float32, buffer=self._buf("gpu", k * 2 * self.n * 4).buf)
        if with_hist:
            a_rec = np.ndarray((L, nb, k, 16), dtype=np.float32, buffer=self._buf("hist", L * nb * k * 64).buf)
        else:
            a_rec = np.ndarray((L, k, 16), dtype=np.float32, buffer=self._buf("last", L * k * 64).buf)
        return a_in, a_gpu, a_rec

    def check(self, k, first_clip, with_hist):
        """Run the oracle over the k clips currently in the shared buffers.  Returns [k][6]: sample error / peak,
        worst record error, its block, its slot."""
        res_shm = self._buf("res", k * 6 * 8)
        res = np.ndarray((k, 6), dtype=np.float64, buffer=res_shm.buf)
        res[:] = -1.0
        per = (k + self.procs - 1) // self.procs
        active = []
        for i, w in enumerate(self.workers):
            lo, hi = i * per, min(k, (i + 1) * per)
            if lo >= hi:
                continue
            cmd = {"op": "check", "chunk": k, "blocks": self.n_blocks, "lo": lo, "hi": hi, "first_clip": int(first_clip),
                   "in": self._shm["in"].name, "gpu": self._shm["gpu"].name, "res": res_shm.name,
                   "hist": self._shm["hist"].name if with_hist else None,
                   "last": None if with_hist else self._shm["last"].name}
            w.stdin.write(json.dumps(cmd) + "\n")
            w.stdin.flush()
            active.append(w)
        for w in active:
            if w.stdout.readline().strip() != "done":
                raise RuntimeError("oracle worker died")
        return res.copy()

    def close(self):
        for w in self.workers:
            try:
                w.stdin.write(json.dumps({"op": "quit"}) + "\n")
                w.stdin.flush()
                w.stdin.close()
            except Exception:
                pass
        for w in self.workers:
            try:
                w.wait(timeout=10)
            except Exception:
                w.kill()
        for s in self._shm.values():
            s.close()
            s.unlink()
        self._shm = {}


def run_population(jb, chain, n_clips, synth, n_samples=48000, sample_rate=48000.0, block=512, params=None, per_clip=None,
                   math="auto", history=True, chunk=2048, procs=None, device=0, first_clip=0, limit_clips=None, name=None):
    """Render `n_clips` synthetic clips through `chain` on the GPU in ONE call (out of place), then check every clip
    (or the first `limit_clips`) against the oracle.  Returns a summary dict (counts of clips out of tolerance)."""
    clip_bytes = 2 * n_samples * 4
    n_blocks = (n_samples + block - 1) // block
    L = len(chain)
    t0 = time.time()
    d_in = jb.DeviceBuffer(n_clips * clip_bytes, device)
    d_out = jb.DeviceBuffer(n_clips * clip_bytes, device)
    jb.synth_fill_device(d_in.ptr.value, synth, first_clip, n_clips, 2, n_samples, sample_rate, device=device)
    eng = jb.BatchProcessor(chain, n_clips, device=device)
    for slot, kv in (params or {}).items():
        for k, v in kv.items():
            eng.setParameter(k, v, slot)
    if per_clip:
        for c in range(n_clips):
            eng.setParameterClips(per_clip["id"], float((first_clip + c) % per_clip["mod"]), c, 1, per_clip["slot"])
    eng.set_math_mode(math)
    if history:
        eng.enableHistory(n_blocks)
    eng.prepareToPlay(sample_rate, block)
    eng.process_device(d_in.ptr.value, d_out.ptr.value, n_samples)
    eng.synchronize()
    kernel_ms, launches = eng.kernel_time_ms()
    coop, lane = eng.path_launches()
    t_gpu = time.time() - t0

    n_check = n_clips if limit_clips is None else min(n_clips, limit_clips)
    pool = OraclePool(chain, n_samples, sample_rate, block, params, per_clip, procs)
    res_all = np.zeros((n_check, 6), dtype=np.float64)
    t1 = time.time()
    try:
        # records of the whole batch, once per plugin: [n_blocks][n_clips][16] (history) or [n_clips][16] (last block)
        recs = [eng.getHistory(s, 0, n_blocks) if history else eng.getLatestMetrics(s) for s in range(L)]
        for c0 in range(0, n_check, chunk):
            k = min(chunk, n_check - c0)
            a_in, a_gpu, a_rec = pool.arrays(k, history)
            jb._check(jb.lib().jb_copy_to_host(device, a_in.ctypes.data, d_in.ptr.value + c0 * clip_bytes, k * clip_bytes))
            jb._check(jb.lib().jb_copy_to_host(device, a_gpu.ctypes.data, d_out.ptr.value + c0 * clip_bytes, k * clip_bytes))
            for s in range(L):
                a_rec[s] = recs[s][:, c0:c0 + k, :] if history else recs[s][c0:c0 + k]
            res_all[c0:c0 + k] = pool.check(k, first_clip + c0, history)
    finally:
        pool.close()
        eng.close()
        d_in.free()
        d_out.free()
    t_cpu = time.time() - t1
    bad_s = np.nonzero(~(res_all[:, 0] <= SAMPLE_TOL))[0]
    bad_r = np.nonzero(~(res_all[:, 1] <= METRIC_TOL))[0]
    worst_s = int(np.argmax(res_all[:, 0])) if n_check else -1
    worst_r = int(np.argmax(res_all[:, 1])) if n_check else -1
    return {
        "config": name or "+".join(chain), "chain": list(chain), "clips_rendered": n_clips, "clips_checked": int(n_check),
        "samples_per_clip": n_samples, "block": block, "synth": synth, "params": params, "per_clip": per_clip, "math": math,
        "records": "every block of every plugin" if history else "last block of every plugin",
        "oracle": pool.kind, "oracle_procs": pool.procs,
        "kernel_ms": kernel_ms, "kernel_launches": int(launches), "coop_launches": int(coop), "lane_launches": int(lane),
        "clips_over_sample_tol": int(bad_s.size), "clips_over_record_tol": int(bad_r.size),
        "worst_sample_err_of_peak": float(res_all[worst_s, 0]) if n_check else 0.0, "worst_sample_clip": worst_s,
        "worst_record_err": float(res_all[worst_r, 1]) if n_check else 0.0, "worst_record_clip": worst_r,
        "worst_record_block": int(res_all[worst_r, 2]) if n_check else -1, "worst_record_slot": int(res_all[worst_r, 3]) if n_check else -1,
        "worst_record_field": int(res_all[worst_r, 4]) if n_check else -1, "worst_record_ref_value": float(res_all[worst_r, 5]) if n_check else 0.0,
        "bad_record_detail": [{"clip": int(c), "err": float(res_all[c, 1]), "block": int(res_all[c, 2]), "slot": int(res_all[c, 3]),
                               "field": int(res_all[c, 4]), "ref": float(res_all[c, 5])} for c in bad_r[:16]],
        "bad_sample_clips": [int(c) for c in bad_s[:32]], "bad_record_clips": [int(c) for c in bad_r[:32]],
        "median_sample_err_of_peak": float(np.median(res_all[:, 0])) if n_check else 0.0,
        "sample_tol": SAMPLE_TOL, "record_tol": METRIC_TOL, "gpu_seconds": t_gpu, "oracle_seconds": t_cpu,
    }


if __name__ == "__main__":
    if "--worker" in sys.argv:
        sys.exit(_worker_main())
    sys.exit("tests/population.py is a helper: see tools/parity_population.py")
