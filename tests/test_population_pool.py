"""The population checker's own plumbing (tests/population.py), on the CPU: a pool of oracle workers fed through shared
memory must report zero error for the oracle's own output and must see a single perturbed sample / record."""
import numpy as np

import population
from cases import SAMPLE_RATE, BLOCK


def test_pool_reports_zero_for_oracle_output_and_sees_perturbations(jb, port):
    chain = ["JuicyPunch", "JuicyWidth"]
    k, n = 6, 3 * BLOCK + 128
    clips = jb.synth_clips("drum", 0, k, n)
    pool = population.OraclePool(chain, n, SAMPLE_RATE, BLOCK, procs=2, kind="port")
    try:
        for with_hist in (True, False):
            a_in, a_gpu, a_rec = pool.arrays(k, with_hist)
            a_in[:] = clips
            for c in range(k):
                out, hists = port.run_chain(chain, clips[c], sample_rate=SAMPLE_RATE, block_size=BLOCK)
                a_gpu[c] = out
                for s in range(len(chain)):
                    if with_hist:
                        a_rec[s, :, c, :] = hists[s]
                    else:
                        a_rec[s, c, :] = hists[s][-1]
            res = pool.check(k, 0, with_hist)
            assert (res[:, 0] == 0.0).all() and (res[:, 1] == 0.0).all()
            a_gpu[3, 1, 700] += 1.0e-3
            if with_hist:
                a_rec[1, 2, 4, 0] += 0.5
            else:
                a_rec[1, 4, 0] += 0.5
            res = pool.check(k, 0, with_hist)
            assert res[3, 0] > population.SAMPLE_TOL and (np.delete(res[:, 0], 3) == 0.0).all()
            assert abs(res[4, 1] - 0.5) < 1e-3 and res[4, 3] == 1 and (np.delete(res[:, 1], 4) == 0.0).all()
            if with_hist:
                assert res[4, 2] == 2
    finally:
        pool.close()


def test_pool_per_clip_parameter(jb, port):
    """material = absolute clip index mod 5, as BASELINE.json configs[2] asks."""
    chain = ["JuicyTexture"]
    k, n = 5, 2 * BLOCK
    clips = jb.synth_clips("impulse", 0, k, n)
    pool = population.OraclePool(chain, n, SAMPLE_RATE, BLOCK, per_clip={"slot": 0, "id": "material", "mod": 5}, procs=2, kind="port")
    try:
        a_in, a_gpu, a_rec = pool.arrays(k, True)
        a_in[:] = clips
        for c in range(k):
            out, hists = port.run_chain(chain, clips[c], sample_rate=SAMPLE_RATE, block_size=BLOCK, params={0: {"material": float((10 + c) % 5)}})
            a_gpu[c] = out
            a_rec[0, :, c, :] = hists[0]
        res = pool.check(k, 10, True)
        assert (res[:, 0] == 0.0).all() and (res[:, 1] == 0.0).all()
    finally:
        pool.close()
