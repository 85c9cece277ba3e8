"""TEST INFRASTRUCTURE — runs the reference's OWN Python (read-only at /root/reference) on CPU.

Only usable where /root/reference exists (this container; never the GPU box).  It installs
`sys.modules` stand-ins for the packages the reference imports but this image lacks
(tinycudann, pytorch3d, attrdict, open3d, kornia, torchviz, matplotlib), then imports the
reference's own `common/*`, `mapping/*`, `models/*` unmodified.  Used by oracle/make_golden.py
to mint tests/golden/*.npz and by the container-only tests that pin oracle/loner_oracle.py.

Randomness injection: the four random draws on the path (optimizer.py:288 randint;
ray_sampling.py:71-72 rand; rendering_tcnn.py:48 rand; rendering_tcnn.py:104 randn) are
replaced by a replay queue so that the oracle / CUDA path can be fed identical numbers.
"""
import contextlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("LONER_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "models"))


class _AttrDict(dict):
    """attrdict.AttrDict stand-in: attribute reads wrap nested mappings (as a shallow copy, like
    the real package) and turn sequences into tuples; item access is plain dict access."""

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return tuple(cls._wrap(x) for x in v)
        return v

    def __getattr__(self, k):
        if k.startswith("__") or k not in self:
            raise AttributeError(k)
        return self._wrap(dict.__getitem__(self, k))

    def __setattr__(self, k, v):
        self[k] = v


class _Permissive(types.ModuleType):
    """Module whose every attribute is another permissive object (open3d, kornia, ...)."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Permissive(self.__name__ + "." + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **kw):
        return _Permissive(self.__name__ + "()")


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from oracle import p3d_standin, tcnn_standin

    sys.modules["tinycudann"] = tcnn_standin

    p3d = types.ModuleType("pytorch3d")
    p3d_t = types.ModuleType("pytorch3d.transforms")
    for n in ("axis_angle_to_matrix", "matrix_to_axis_angle", "matrix_to_quaternion",
              "quaternion_to_axis_angle", "quaternion_to_matrix", "axis_angle_to_quaternion"):
        setattr(p3d_t, n, getattr(p3d_standin, n))
    p3d.transforms = p3d_t
    sys.modules["pytorch3d"] = p3d
    sys.modules["pytorch3d.transforms"] = p3d_t

    ad = types.ModuleType("attrdict")
    ad.AttrDict = _AttrDict
    sys.modules["attrdict"] = ad

    for name in ("open3d", "kornia", "kornia.geometry", "kornia.geometry.calibration",
                 "kornia.morphology", "torchviz", "matplotlib", "matplotlib.pyplot",
                 "cv2", "rosbag", "rospy", "ros_numpy", "trimesh", "skimage"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Permissive(name)
    # `from matplotlib import pyplot as plt` needs the attribute
    if isinstance(sys.modules["matplotlib"], _Permissive):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["kornia"].geometry = sys.modules["kornia.geometry"]
        sys.modules["kornia.geometry"].calibration = sys.modules["kornia.geometry.calibration"]

    for p in (REF_ROOT, os.path.join(REF_ROOT, "src")):
        if p not in sys.path:
            sys.path.append(p)
    _installed = True


def import_reference():
    """Returns a namespace of the reference modules on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    install_stubs()
    import importlib
    ns = types.SimpleNamespace()
    ns.rendering = importlib.import_module("models.rendering_tcnn")
    ns.losses = importlib.import_module("models.losses")
    ns.nerf = importlib.import_module("models.nerf_tcnn")
    ns.model = importlib.import_module("models.model_tcnn")
    ns.ray_sampling = importlib.import_module("models.ray_sampling")
    ns.settings = importlib.import_module("common.settings")
    ns.pose_utils = importlib.import_module("common.pose_utils")
    ns.pose = importlib.import_module("common.pose")
    ns.sensors = importlib.import_module("common.sensors")
    ns.frame = importlib.import_module("common.frame")
    ns.ray_utils = importlib.import_module("common.ray_utils")
    ns.keyframe = importlib.import_module("mapping.keyframe")
    ns.optimizer = importlib.import_module("mapping.optimizer")
    return ns


class Replay:
    """Replay queue for torch.rand / torch.randn / torch.randint inside the reference."""

    def __init__(self):
        self.rand, self.randn, self.randint = [], [], []
        self.log = []

    def _pop(self, q, kind, shape):
        if not q:
            raise RuntimeError(f"replay queue for {kind} exhausted (asked {tuple(shape)})")
        t = q.pop(0)
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"replay {kind}: queued {tuple(t.shape)} but reference asked {tuple(shape)}")
        self.log.append((kind, tuple(shape)))
        return t.clone()


@contextlib.contextmanager
def injected_randomness(replay: Replay):
    """Patch torch.rand/randn/randint (as seen by the reference modules) with the replay."""
    def _shape(args):
        if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)):
            return tuple(args[0])
        return tuple(args)

    o_rand, o_randn, o_randint = torch.rand, torch.randn, torch.randint

    def rand(*a, **kw):
        return replay._pop(replay.rand, "rand", _shape(a))

    def randn(*a, **kw):
        return replay._pop(replay.randn, "randn", _shape(a))

    def randint(*a, **kw):
        shape = a[-1]
        return replay._pop(replay.randint, "randint", tuple(shape))

    torch.rand, torch.randn, torch.randint = rand, randn, randint
    try:
        yield
    finally:
        torch.rand, torch.randn, torch.randint = o_rand, o_randn, o_randint


def load_settings(sequence_yaml: str = None):
    """The reference's own Settings for cfg/defaults.yaml + an optional sequence file's changes."""
    ns = import_reference()
    S = ns.settings.Settings
    s = S.load_from_file(os.path.join(REF_ROOT, "cfg", "defaults.yaml"))
    if sequence_yaml is not None:
        import yaml
        with open(os.path.join(REF_ROOT, "cfg", sequence_yaml)) as f:
            seq = yaml.full_load(f)
        s.augment(seq.get("changes"))
    return s
