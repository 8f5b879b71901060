"""Determinism of the captured training step (runs right after test_model_gpu.py, i.e. "in-suite").

Forward kernels and activation-gradient kernels (dgrad, LayerNorm backward, resize backward ...) use no atomics:
on identical weights and inputs their outputs must be BIT-identical between repetitions and between the
single-stream step and the step with the filter-gradient kernels on a second stream -- a differing bit is a
race or an uninitialised read, not summation-order noise.  Parameter gradients are summed with fp32 atomics by
default, so they and the loss trajectories are compared against the band measured between repetitions of the
SAME configuration.  The full report goes to gpurun_out/ when that directory exists."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_step_determinism_single_vs_side_stream():
    import trajectory_probe as TP
    out = os.path.join(ROOT, "gpurun_out", "trajectory_probe_insuite.json") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
    rep = TP.main(["--reps", "2", "--steps", "4", "--tag", "in-suite"] + (["--out", out] if out else []))
    forced = [r for r in rep["runs"] if r["state_forced"]]
    assert forced
    for r in forced:
        # same weights, same inputs: every activation, LayerNorm statistic and activation gradient of step 1 bit-identical
        assert r["acts_first_bit_difference"] is None, r
        assert r["extra_first_bit_difference"] is None, r
        assert r["grads_first_bit_difference"] is None, r
        assert r["loss_absdiff"][1] <= 2e-6, r          # the loss sum itself is an atomic reduction
        assert all(t["rel_l2"] < 1e-5 for t in r["G1_worst_tensors"]), r
