"""Import the UNMODIFIED reference (/root/reference/nerf.py) in the build container.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (the
build container).  Nothing that runs on the GPU box imports this module: the
goldens it produces are committed under tests/golden/ by oracle/make_golden.py.

The reference imports `imageio` and `matplotlib.pyplot` at module scope
(nerf.py:7, 12); neither is installed here, so empty stub modules are injected
first.  No reference source is copied into this repo.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REF_DIR = os.environ.get("NERF_TINY_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "nerf.py"))


def import_reference():
    if not available():
        raise RuntimeError(f"reference not present at {REF_DIR}")
    for name in ("imageio", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    # our drop-in package also has a module called `nerf`; the reference is imported under its own name
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_nerf", os.path.join(REF_DIR, "nerf.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_nerf"] = mod
    spec.loader.exec_module(mod)          # runs seed_everything(624) (nerf.py:50)
    mod.device = torch.device("cpu")      # normally set by NeRFRunner.__init__ (nerf.py:387)
    return mod


def make_model(ref, n_rays: int, sd=None, n_coarse: int = 64, n_fine: int = 128):
    m = ref.NeRFModel(num_coarse=n_coarse, num_fine=n_fine, batch_ray=n_rays)
    if sd is not None:
        m.load_state_dict(sd)
    return m
