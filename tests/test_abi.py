"""The C-ABI library loads and exports every symbol include/nerftiny.h declares (no compute, CPU only)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from nerf_tiny_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "nerftiny.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in nerftiny.h but not exported"


def test_binding_matches_header(lib_path):
    from nerf_tiny_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    src = open(os.path.join(ROOT, "include", "nerftiny.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, binding has {len(args)}"


def test_layer_table_and_param_count(lib_path):
    from nerf_tiny_b200 import _lib
    from nerf_tiny_b200.nerf_keys import LAYER_SHAPES
    lib = _lib.load()
    assert lib.nt_param_count() == 593924 == _lib.N_PARAMS
    off = 0
    for (o, i, wo, bo), (so, si) in zip(_lib.layer_table(), LAYER_SHAPES):
        assert (o, i) == (so, si)
        assert wo == off
        off += o * i
        assert bo == off
        off += o
    assert off == 593924


def test_create_fails_loudly_without_gpu(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nerf_tiny_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.nt_create(ctypes.byref(h), 0, 64, 128)
    assert rc != 0 and lib.nt_last_error()
    with pytest.raises(_lib.NerfTinyError):
        _lib.check(rc)


def test_no_oracle_import_in_product():
    """The product package never imports the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "nerf_tiny_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# ", ""), f
