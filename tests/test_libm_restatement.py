"""csrc/jb_libm.h restates the C library's tanhf and powf (the per-sample transcendentals of JuicySaturator /
JuicyPunch, JuicySaturator/PluginProcessor.cpp:92, JuicyPunch/PluginProcessor.cpp:100,106) so that the CUDA
engine can reproduce them bit for bit.  This pins the restatement itself, on the host, against the libm the
oracle and the compiled reference link (glibc 2.39 in this image): no argument may differ."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_restated_tanhf_and_powf_are_bit_identical_to_libm(tmp_path):
    exe = str(tmp_path / "libm_harness")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off",
                           "-I", os.path.join(ROOT, "juicy-audio-plugins_b200", "csrc"),
                           os.path.join(ROOT, "tests", "libm_harness.cpp"), "-o", exe, "-lm"])
    out = json.loads(subprocess.check_output([exe, "10000000"]).decode())
    assert out["tanhf_checked"] > 7_000_000 and out["powf_checked"] > 7_000_000
    assert out["tanhf_mismatches"] == 0, out
    assert out["powf_mismatches"] == 0, out
    assert out["pow_zero_ok"] == 1
