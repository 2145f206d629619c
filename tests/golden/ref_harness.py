"""Import the UNMODIFIED reference (/root/reference/offpolicy) in the build container.

Only used to (a) generate the committed golden fixtures (make_goldens.py) and (b) pin
the oracle port against the live reference in `-m "not gpu"` tests when /root/reference
is present.  Never imported by the product, bench.py's default arm or any `-m gpu` test
(the GPU box has no /root/reference).

Recipe = SURVEY.md §8(c): a bare namespace module instead of offpolicy/__init__.py
(which drags in wandb/pysc2) and a 20-line `gym` shim for offpolicy/utils/util.py:2-4.
"""
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("OFFPOLICY_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "offpolicy", "algorithms"))


def _install_gym_shim():
    if "gym" in sys.modules and hasattr(sys.modules["gym"], "_b200_shim"):
        return
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Space(object):
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape) if shape is not None else np.asarray(low).shape
            self.low = np.full(self.shape, low, dtype=dtype)
            self.high = np.full(self.shape, high, dtype=dtype)
            self.dtype = dtype

    class Discrete(Space):
        def __init__(self, n):
            self.n = int(n)

    class Tuple(Space):
        def __init__(self, spaces_):
            self.spaces = tuple(spaces_)

        def __iter__(self):
            return iter(self.spaces)

    gym.Space = Space
    gym.spaces = spaces
    gym._b200_shim = True
    spaces.Box, spaces.Discrete, spaces.Tuple, spaces.Space = Box, Discrete, Tuple, Space
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces


def import_reference():
    """Returns the `offpolicy` namespace module rooted at the reference tree."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_gym_shim()
    mod = sys.modules.get("offpolicy")
    ref_pkg = os.path.join(REF_ROOT, "offpolicy")
    if mod is None or ref_pkg not in list(getattr(mod, "__path__", [])):
        for k in [k for k in sys.modules if k == "offpolicy" or k.startswith("offpolicy.")]:
            del sys.modules[k]
        mod = types.ModuleType("offpolicy")
        mod.__path__ = [ref_pkg]
        sys.modules["offpolicy"] = mod
    return mod


def make_args(extra=()):
    import_reference()
    from offpolicy.config import get_config
    args = get_config().parse_known_args(list(extra))[0]
    args.use_same_share_obs = True  # defined by train_*.py, not config.py
    return args


def gym_spaces():
    _install_gym_shim()
    return sys.modules["gym.spaces"]
