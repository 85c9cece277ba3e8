"""CPU: the pieces of bench.py that do not need a GPU - flop accounting, the committed ncu
summary the roofline object quotes, the workload table."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_algorithmic_flops_match_the_survey():
    e_pad, f_fwd, f_train = bench.flops_per_sample(bench.WORKLOADS["c2"])
    assert e_pad == 64 and f_fwd == 426_496 and f_train == 1_246_720          # 4x256, map-only
    assert bench.flops_per_sample(bench.WORKLOADS["c5"])[2] == 1_279_488       # joint pose + map
    assert bench.flops_per_sample(bench.WORKLOADS["c1"])[1] == 16_512          # 2x64
    assert bench.flops_per_sample(bench.WORKLOADS["c2hash"])[0] == 32


def test_ncu_summary_is_readable_and_plausible():
    # C2-sized launches: stash / dZ / stash+dZ bytes; the upper bounds are round 1's design, round 2 moves less
    for kern, lo, hi in (("mlp_fwd", 6.5e9, 11e9), ("mlp_dgrad", 5e9, 10e9), ("mlp_wgrad", 9e9, 16.5e9)):   # fwd: 7.5 GB since A_L left the stash
        t, src = bench.ncu_traffic_bytes(kern)
        assert t is not None and src.startswith("profiles/") and lo < t < hi, (kern, t)


def test_workloads_name_the_baseline_configs():
    assert bench.WORKLOADS["c2"]["rays_per_gpu"] == 8192 and bench.WORKLOADS["c2"]["S"] == 512
    assert bench.WORKLOADS["c5"]["poses"] and bench.WORKLOADS["c5"]["K"] == 16
    assert bench.WORKLOADS["c2hash"]["encoding"] == "HashGrid"
    assert bench.rays_per_gpu(bench.WORKLOADS["c4"], 4) == 16384 and bench.WORKLOADS["c4"]["S"] == 256
    assert bench.describe(bench.WORKLOADS["c2hash"], 1)["encoding"].startswith("HashGrid")
    assert bench.describe(bench.WORKLOADS["c2"], 1)["encoding"].startswith("Frequency")
