"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header declares,
the ctypes table covers the header, struct layouts match, and -- without a GPU -- the product fails loudly
instead of falling back to a CPU path."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

import spinoza_b200 as sb

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "spinoza_b200.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"SPZ_API[^;(]*?\b(spz_\w+)\s*\(", HEADER)))


def test_header_declares_a_real_surface():
    syms = declared_symbols()
    assert len(syms) >= 30
    for must in ("spz_create", "spz_apply", "spz_c_apply", "spz_mc_apply", "spz_execute", "spz_measure_qubit",
                 "spz_sample", "spz_download"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(sb.library_path())
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/spinoza_b200.h but not exported"


def test_ctypes_table_matches_header():
    assert sorted(sb._SIGNATURES) == declared_symbols()


def test_exported_symbols_are_only_the_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", sb.library_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    ours = {s for s in exported if s.startswith("spz_")}
    assert ours == set(declared_symbols())


def test_struct_layouts():
    assert C.sizeof(sb._Gate) == 40 and C.sizeof(sb._Op) == 64
    assert sb._Op.ctrl_mask.offset == 48 and sb._Op.p.offset == 16


def test_abi_version_and_status_strings():
    assert sb._lib.spz_abi_version() == int(re.search(r"#define SPZ_ABI_VERSION (\d+)", HEADER).group(1))
    assert sb._lib.spz_status_string(2) == b"unsupported gate/control combination"


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", sb.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu():
    if sb.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(sb.SpinozaError) as e:
        sb.State(3)
    assert e.value.status == sb.ERR_NO_DEVICE


def test_product_never_imports_oracle():
    # the oracle is the checker, never the product: no import, include, link or dlopen of oracle/ from the package
    pat = re.compile(r"^\s*(import|from)\s+oracle\b|#\s*include\s*[\"<][^\">]*oracle|dlopen\([^)]*oracle|"
                     r"CDLL\([^)]*oracle|libspinoza_oracle", re.M)
    for p in list((ROOT / "spinoza_b200").rglob("*")) + [ROOT / "include" / "spinoza_b200.h"]:
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"):
            assert not pat.search(p.read_text()), p
