"""Population-level parity on the BASELINE.json shapes: EVERY clip of the batch against the CPU oracle
(tests/population.py: the GPU renders the whole batch in one call, a pool of oracle processes re-renders every clip).
The count of clips outside the north_star tolerances (1e-5 of clip peak per sample, 0.01 per record field, every block
of every plugin) must be ZERO.  The heavier sweeps (all five Texture materials alone, the 32768-clip shard with a
resonant material) are run by tools/parity_population.py and committed as profiles/r02_parity_population.json;
JB_POPULATION_FULL=1 runs them here as well."""
import json
import os

import pytest

import population
from cases import SAMPLE_RATE, BLOCK, FULL_CHAIN

pytestmark = pytest.mark.gpu

FULL = os.environ.get("JB_POPULATION_FULL", "0") not in ("", "0")


def _assert_clean(r):
    print(json.dumps(r))
    assert r["clips_checked"] > 0
    assert r["clips_over_sample_tol"] == 0, "%(config)s: %(clips_over_sample_tol)d clips over 1e-5 of peak (worst %(worst_sample_err_of_peak).3e, clip %(worst_sample_clip)d)" % r
    assert r["clips_over_record_tol"] == 0, "%(config)s: %(clips_over_record_tol)d clips with a record off by more than 0.01 (worst %(worst_record_err).4f, clip %(worst_record_clip)d block %(worst_record_block)d slot %(worst_record_slot)d)" % r


def test_population_c2_punch_width_4096(jb):
    """configs[1]: all 4096 drum-hit clips through the cooperative kernel."""
    r = population.run_population(jb, ["JuicyPunch", "JuicyWidth"], 4096, "drum", name="C2")
    assert r["coop_launches"] >= 1
    _assert_clean(r)


def test_population_c3_texture_clip_mod_5(jb):
    """configs[2]: 8192 stereo impulse-train clips, material = clip mod 5 (five concurrent parameter sets)."""
    r = population.run_population(jb, ["JuicyTexture"], 8192, "impulse", per_clip={"slot": 0, "id": "material", "mod": 5}, name="C3 mod 5")
    _assert_clean(r)


def test_population_c4_infer_65536(jb):
    """configs[3]: all 65536 noise / sweep / impulse / drum clips."""
    r = population.run_population(jb, ["JuicyInfer"], 65536, "mixed", name="C4")
    _assert_clean(r)


def test_population_c5_shard_32768(jb):
    """configs[4]: one GPU's shard of the 7-plugin chain at its REAL size -- 32768 clips take other code than small
    batches (no plugin pipeline, no channel-pair kernels, eight samples per trip for the light plugins)."""
    limit = None if FULL else 8192  # every clip is RENDERED at the real shape; the default run checks a quarter of them
    r = population.run_population(jb, FULL_CHAIN, 32768, "mixed", limit_clips=limit, name="C5 shard")
    _assert_clean(r)


@pytest.mark.skipif(not FULL, reason="JB_POPULATION_FULL=1 (tools/parity_population.py commits the result)")
@pytest.mark.parametrize("material", [2], ids=["wood"])
def test_population_c5_shard_resonant_material(material, jb):
    r = population.run_population(jb, FULL_CHAIN, 32768, "mixed", params={2: {"material": float(material)}}, name="C5 shard, wood")
    _assert_clean(r)
