"""The C++ host shim (include/upc_b200_shim.hpp): compiles against the C ABI on CPU; on the GPU box the compiled
driver (tests/cpp/shim_smoke.cpp) runs the reference-style call sequence and checks it against the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "shim_smoke")


def _build(oracle):
    src = os.path.join(ROOT, "tests", "cpp", "shim_smoke.cpp")
    hdr = os.path.join(ROOT, "include", "upc_b200_shim.hpp")
    if not os.path.exists(EXE) or any(os.path.getmtime(f) > os.path.getmtime(EXE) for f in (src, hdr)):
        pkg = os.path.join(ROOT, "uncertainty_planning_core_b200")
        orc = os.path.join(ROOT, "oracle")
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-fopenmp", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L" + pkg, "-L" + orc, "-lupc_b200", "-loracle", "-Wl,-rpath," + pkg, "-Wl,-rpath," + orc])


def test_shim_compiles_and_links_against_the_c_abi(oracle):
    _build(oracle)
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_shim_round_trip_matches_oracle(oracle):
    _build(oracle)
    out = subprocess.run([EXE], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim ok" in out.stdout


@pytest.mark.gpu
def test_sharded_simulator_matches_single_gpu(oracle):
    """ShardedSimulator (edges split over the GPUs of the box, NCCL all-gather issued by the library) == one GPU, bit for bit"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    _build(oracle)
    out = subprocess.run([EXE, "sharded", str(min(torch.cuda.device_count(), 8))], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "sharded ok" in out.stdout
