"""The drop-in `models.*` package exposes the reference's public names with the reference's signatures.
(Container only: compares against the reference's own modules imported through oracle/ref_harness.)"""
import importlib
import inspect
import sys

import pytest

PUBLIC = {
    "models.model_tcnn": ["Model", "OccupancyGridModel"],
    "models.nerf_tcnn": ["DecoupledNeRF"],
    "models.ray_sampling": ["UniformRaySampler", "OccGridRaySampler"],
    "models.rendering_tcnn": ["render_rays", "inference", "sample_pdf"],
    "models.losses": ["get_weights_gt", "get_logits_grad", "img_to_mse", "mse_to_psnr"],
}
METHODS = {
    "Model": ["forward", "get_sigma_parameters", "get_rgb_parameters", "freeze_sigma_head", "freeze_rgb_head",
              "inference_points", "get_rgb_mlp_parameters", "get_rgb_feature_parameters"],
    "OccupancyGridModel": ["forward", "interpolate"],
    "DecoupledNeRF": ["forward"],
    "UniformRaySampler": ["get_samples"],
    "OccGridRaySampler": ["get_samples", "update_occ_grid"],
}


def _params(fn):
    return [(p.name, p.default if p.default is not inspect._empty else "<req>")
            for p in inspect.signature(fn).parameters.values()]


@pytest.mark.refonly
def test_names_and_signatures_match_the_reference():
    from oracle import ref_harness as rh
    rh.import_reference()
    ref = {m: importlib.import_module(m) for m in PUBLIC}
    for m in list(sys.modules):
        if m == "models" or m.startswith("models."):
            del sys.modules[m]
    from loner_b200 import dropin
    path = dropin.install()
    try:
        ours = {m: importlib.import_module(m) for m in PUBLIC}
        for mod, names in PUBLIC.items():
            assert ours[mod].__file__.startswith(path)
            for n in names:
                a, b = getattr(ref[mod], n), getattr(ours[mod], n)
                if inspect.isclass(a):
                    assert _params(a.__init__) == _params(b.__init__), f"{mod}.{n}.__init__"
                    for meth in METHODS[n]:
                        assert _params(getattr(a, meth)) == _params(getattr(b, meth)), f"{mod}.{n}.{meth}"
                else:
                    assert _params(a) == _params(b), f"{mod}.{n}"
    finally:
        sys.path.remove(path)
        for m in list(sys.modules):
            if m == "models" or m.startswith("models."):
                del sys.modules[m]


def test_dropin_imports_without_a_gpu():
    from loner_b200 import dropin
    path = dropin.install()
    try:
        mt = importlib.import_module("models.model_tcnn")
        ls = importlib.import_module("models.losses")
        assert hasattr(mt, "Model") and hasattr(ls, "get_weights_gt")
    finally:
        sys.path.remove(path)
        for m in list(sys.modules):
            if m == "models" or m.startswith("models."):
                del sys.modules[m]


@pytest.mark.refonly
def test_fused_optimizer_has_the_reference_constructor_and_entry_points():
    from oracle import ref_harness as rh
    ns = rh.import_reference()
    from loner_b200.dropin.mapping_optimizer import FusedOptimizer, OptimizationSettings
    ref = ns.optimizer.Optimizer
    assert _params(ref.__init__) == _params(FusedOptimizer.__init__)
    assert _params(ref.iterate_optimizer) == _params(FusedOptimizer.iterate_optimizer)
    assert _params(ref._do_iterate_optimizer) == _params(FusedOptimizer._do_iterate_optimizer)
    import dataclasses
    assert [f.name for f in dataclasses.fields(ns.optimizer.OptimizationSettings)] == \
        [f.name for f in dataclasses.fields(OptimizationSettings)]
