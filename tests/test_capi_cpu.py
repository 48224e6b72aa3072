"""CPU: the C-ABI library loads and exports every symbol include/pdeb200.h declares;
argument validation that needs no GPU; no compute calls."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    txt = (ROOT / "include" / "pdeb200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pdeb200_[a-z0-9_]+)\s*\(", txt)))


def test_exports_match_header(pkg):
    lib = pkg.lib.load()
    names = header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libpdeb200.so does not export %s" % n
    assert sorted(pkg.lib.exported_symbols()) == names, "python prototypes out of sync with the header"


def test_config_struct_abi(pkg):
    L = pkg.lib
    cfg = L.Config()
    assert L.load().pdeb200_default_config(L.KS, C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(L.Config)
    assert (cfg.nx, cfg.oversampling, cfg.window_size) == (240, 30, 1)
    assert cfg.reward_pow == pytest.approx(1.3) and cfg.reward_div == pytest.approx(90.0)
    assert L.load().pdeb200_default_config(L.KSEG1D, C.byref(cfg)) == 0
    assert (cfg.nx, cfg.temporal_steps, cfg.window_size) == (100, 2, 3)
    assert L.load().pdeb200_default_config(L.NS2D, C.byref(cfg)) == 0
    assert cfg.sensors_per_axis == 16 and cfg.reward_pow == pytest.approx(1.1)
    assert L.load().pdeb200_default_config(99, C.byref(cfg)) != 0


def test_create_rejects_bad_config_without_gpu(pkg):
    L = pkg.lib
    lib = L.load()
    cfg = L.Config()
    lib.pdeb200_default_config(L.KS, C.byref(cfg))
    ctx = C.c_void_p()
    cfg.struct_size = 8
    assert lib.pdeb200_create(C.byref(cfg), 0, C.byref(ctx)) == -1
    assert b"struct_size" in lib.pdeb200_last_error(None)
    lib.pdeb200_default_config(L.KS, C.byref(cfg))
    cfg.n_sensors = 0
    assert lib.pdeb200_create(C.byref(cfg), 0, C.byref(ctx)) == -1


def test_no_oracle_import_in_product():
    """The product path must never route through oracle/ (it is test infrastructure)."""
    pkgdir = ROOT / "distributedconvrl-pde-control_b200"
    for p in pkgdir.rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p
        assert "oracle." not in src and "/oracle" not in src, p
