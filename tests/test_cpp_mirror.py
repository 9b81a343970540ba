"""The C++ host mirror over the C ABI: compiles everywhere; runs the reference-style tests on a GPU box and fails
loudly (exit 3, SPZ_ERR_NO_DEVICE) without one."""
import subprocess
from pathlib import Path

import pytest

import spinoza_b200 as sb

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "cpp" / "test_mirror.cpp"
EXE = ROOT / "tests" / "cpp" / "_build" / "test_mirror"


def build():
    EXE.parent.mkdir(exist_ok=True)
    lib = Path(sb.library_path())
    cmd = ["/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++", "-std=c++17", "-O1", "-pthread", str(SRC), "-o", str(EXE),
           f"-L{lib.parent}", "-lspinoza_b200", f"-Wl,-rpath,{lib.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


C_SRC = ROOT / "tests" / "c" / "abi_smoke.c"
C_EXE = ROOT / "tests" / "cpp" / "_build" / "abi_smoke"


def build_c():
    C_EXE.parent.mkdir(exist_ok=True)
    lib = Path(sb.library_path())
    cmd = ["/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else "gcc", "-std=c11", "-Wall", "-Werror", "-O1", str(C_SRC), "-o", str(C_EXE),
           f"-L{lib.parent}", "-lspinoza_b200", f"-Wl,-rpath,{lib.parent}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C_EXE


def test_c_client_compiles_as_c11_and_fails_loudly_without_gpu():
    exe = build_c()
    if sb.device_count() > 0:
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 3, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_client_on_gpu():
    exe = build_c()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "C_ABI_OK" in r.stdout, r.stdout + r.stderr


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu():
    exe = build()
    if sb.device_count() > 0:
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 3, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_reference_tests_on_gpu():
    exe = build()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CPP_MIRROR_OK" in r.stdout, r.stdout + r.stderr
