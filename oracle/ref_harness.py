"""Import the UNMODIFIED reference (nerf.py of D-Hank/NeRF-tiny).

TEST / BASELINE INFRASTRUCTURE ONLY.  Where it is found:
  * /root/reference (the build container): used by oracle/make_golden.py & co to produce tests/golden/, and by the CPU
    tests that cross-check the oracle;
  * oracle/_ref/ : a byte-for-byte staging copy of nerf.py + loader.py made by `stage()` at build time (git-ignored, so
    it never enters the history, but it travels to the GPU box with the snapshot like the built .so).  Its only consumer
    is `bench.py --impl reference` / `cpu_baseline`, which time the reference's own CPU path on the box's host cores.
The parity tests that run on the GPU box use the committed goldens and the oracle port, never this module.

The reference imports `imageio` and `matplotlib.pyplot` at module scope
(nerf.py:7, 12); neither is installed here, so empty stub modules are injected
first.  No reference source is copied into this repo.
"""
from __future__ import annotations

import os
import sys
import types

import torch

SRC_DIR = "/root/reference"
STAGE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _resolve() -> str:
    env = os.environ.get("NERF_TINY_REFERENCE")
    if env:
        return env
    return SRC_DIR if os.path.isfile(os.path.join(SRC_DIR, "nerf.py")) else STAGE_DIR


REF_DIR = _resolve()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "nerf.py"))


def stage() -> bool:
    """Copy the reference's two modules, unmodified, into oracle/_ref/ (git-ignored).  No-op where /root/reference is
    absent (the GPU box uses the staged copy that travelled with the snapshot)."""
    import filecmp
    import shutil
    if not os.path.isfile(os.path.join(SRC_DIR, "nerf.py")):
        return os.path.isfile(os.path.join(STAGE_DIR, "nerf.py"))
    os.makedirs(STAGE_DIR, exist_ok=True)
    for name in ("nerf.py", "loader.py"):
        src, dst = os.path.join(SRC_DIR, name), os.path.join(STAGE_DIR, name)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    return True


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
