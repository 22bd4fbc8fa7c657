"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/soc_b200.h declares.  No compute calls (there is no GPU on the build box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from soc_b200 import build
    return build.build()


def _declared():
    txt = open(os.path.join(ROOT, "include", "soc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(soc_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_documented_surface():
    names = _declared()
    for must in ("soc_create", "soc_set_grid", "soc_set_params", "soc_upload", "soc_download", "soc_zero_amc",
                 "soc_sim_pb", "soc_sim_hp", "soc_sim_cl", "soc_mapping", "soc_mapping_levels", "soc_healpix_mapping", "soc_sca_ps",
                 "soc_sca_pb", "soc_eq_temperature", "soc_emission", "soc_get_counters"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    L = ctypes.CDLL(lib_path)
    for name in _declared():
        assert hasattr(L, name), name


def test_python_binding_covers_the_header(lib_path):
    from soc_b200 import backend
    assert sorted(backend.exported_symbols()) == _declared()
    backend.load_library()


def test_no_device_is_a_loud_error(lib_path):
    """On a box without a GPU the product path must fail, not fall back."""
    from soc_b200 import backend
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(backend.SocError):
        backend.Device(0)


def test_product_does_not_import_the_oracle():
    """soc_b200/ must never import, dlopen or execute anything under oracle/ (the oracle is the checker)."""
    pkg = os.path.join(ROOT, "soc_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|libsoc_oracle|libsocref|oracle[/\\.](orc|ref|_ref|build_ref|soc_oracle)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not pat.search(txt), os.path.join(dirpath, f)


def test_missing_library_is_a_loud_error(lib_path, monkeypatch, tmp_path):
    """SOC_B200_LIB points the binding at a development build; a path that does not exist raises (no fallback to anything)."""
    from soc_b200 import backend
    monkeypatch.setattr(backend, "_lib", None)
    monkeypatch.setenv("SOC_B200_LIB", str(tmp_path / "no_such_library.so"))
    with pytest.raises(backend.SocError):
        backend.load_library()
    monkeypatch.setenv("SOC_B200_LIB", lib_path)
    assert backend.load_library() is not None
