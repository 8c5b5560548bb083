"""The C ABI from plain C (tests/abi/abi_check.c): the header compiles as C11, the program links against the shared library,
and -- without a GPU -- piql_create refuses to run (no CPU fallback); with one (-m gpu) it solves a small gridworld."""
import os
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "abi", "abi_check.c")
LIBDIR = os.path.join(ROOT, "piql.jl_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "abi_check")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lpiql_b200", "-lm", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="the no-device contract is checked where there is no device")
def test_c_client_without_a_device_gets_no_device(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "PIQL_ERR_NO_DEVICE" in out.stdout


@pytest.mark.gpu
def test_c_client_solves_a_gridworld(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "device path" in out.stdout
