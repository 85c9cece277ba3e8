"""CPU: the pieces of bench.py that do not need a GPU - flop accounting, the committed ncu
summary the roofline object quotes, the workload table."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_algorithmic_flops_match_the_survey():
    f_fwd, f_train = bench.flops_per_sample(64, 256, 4, poses=False)
    assert f_fwd == 426_496 and f_train == 1_246_720          # 4x256, map-only
    assert bench.flops_per_sample(64, 256, 4, poses=True)[1] == 1_279_488


def test_ncu_summary_is_readable_and_plausible():
    for kern, lo, hi in (("mlp_fwd", 9e9, 11e9), ("mlp_dgrad", 8e9, 10e9), ("mlp_wgrad", 15e9, 16.5e9)):
        t = bench.ncu_traffic_bytes(kern)
        assert t is not None and lo < t < hi, (kern, t)          # C2: stash / dZ / stash+dZ bytes


def test_workloads_name_the_baseline_configs():
    assert bench.WORKLOADS["c2"]["rays_per_gpu"] == 8192 and bench.WORKLOADS["c2"]["S"] == 512
    assert bench.WORKLOADS["c5"]["poses"] and bench.WORKLOADS["c5"]["K"] == 16
    assert bench.WORKLOADS["c2hash"]["encoding"] == "HashGrid"
