"""Host-side check of the register-resident pivoted QR used by the loop's plane fits: it must agree bit for bit with the
generic local-memory version (the one the seams and the legacy kernel use) on random, ill-conditioned and rank-deficient
5x3 systems.  Compiles tools/test_qr_reg.cu as HOST code with nvcc (no GPU involved)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_register_qr_is_bit_identical_to_the_generic_qr(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "test_qr_reg"
    subprocess.run([nvcc, "-O2", "-o", str(exe), os.path.join(ROOT, "tools", "test_qr_reg.cu")], check=True,
                   capture_output=True, text=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "400000 systems, 0 mismatches" in res.stdout


def test_fast_solve_step_pieces_on_the_host(tmp_path):
    """k2_fast.cuh (warm-started 3x3 Jacobi, pivoted L D L^T inverse with the FullPivLU invertibility decision) compiled
    as host code: residuals, orthogonality, warm == cold to rounding, decisions against a plain full-pivot LU."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "test_k2_fast"
    subprocess.run([nvcc, "-O2", "-o", str(exe), os.path.join(ROOT, "tools", "test_k2_fast.cu")], check=True,
                   capture_output=True, text=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "K2_FAST_OK" in res.stdout


def test_loop_tile_plan(tmp_path):
    """loop_plan.hpp (how a run's source slots are cut into blocks of the iteration kernel) as plain host C++: coverage,
    tile bounds, resident-block cap, the small-cloud rule and the values used for the shipped cloud / C2 / C4 / C5."""
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = tmp_path / "test_loop_plan"
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tools", "test_loop_plan.cpp")], check=True,
                   capture_output=True, text=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "LOOP_PLAN_OK" in res.stdout
