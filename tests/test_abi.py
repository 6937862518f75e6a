"""The C-ABI library: loads without a GPU, exports every symbol include/upc_b200.h declares, keeps its POD
layouts in sync with the ctypes mirror, and reports errors the way the header promises."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from uncertainty_planning_core_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "upc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(upc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(abi.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in lib.upc_version()


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "upc_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(upc_env_desc),sizeof(upc_axis_params),sizeof(upc_joint_desc),sizeof(upc_robot_desc),sizeof(upc_sim_params),"
                   "sizeof(upc_cluster_params),sizeof(upc_stats),offsetof(upc_robot_desc,points),offsetof(upc_robot_desc,joints));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(abi.EnvDesc), C.sizeof(abi.AxisParams), C.sizeof(abi.JointDesc), C.sizeof(abi.RobotDesc), C.sizeof(abi.SimParams),
            C.sizeof(abi.ClusterParams), C.sizeof(abi.Stats), abi.RobotDesc.points.offset, abi.RobotDesc.joints.offset]
    assert got == want


def test_size_helpers_and_tape_generation_need_no_gpu():
    lib = abi.load_library()
    assert lib.upc_config_width(abi.ROBOT_SE2, 3) == 3
    assert lib.upc_config_width(abi.ROBOT_SE3, 6) == 12
    assert lib.upc_config_width(abi.ROBOT_LINKED, 7) == 7
    p = abi.make_sim_params(64, 0.01, 0.04, max_microsteps=2)
    assert lib.upc_tape_doubles_per_particle(C.byref(p), 6) == 64 * 3 * 6
    from uncertainty_planning_core_b200 import scenarios
    robot = scenarios.se3_robot(max_sensor_noise=0.0)
    tape = abi.fill_noise_tape(9, robot, p, 5)
    assert tape.shape == (64, 3, 6, 5)
    assert np.all(tape[:, 0] == 0.0)                 # zero sensor noise -> exact zeros
    assert np.all(np.abs(tape[:, 1:]) <= 1.0) and np.any(tape[:, 1:] != 0.0)
    assert np.array_equal(tape, abi.fill_noise_tape(9, robot, p, 5))
    assert not np.array_equal(tape, abi.fill_noise_tape(10, robot, p, 5))


def test_invalid_arguments_are_reported_not_crashed():
    lib = abi.load_library()
    assert lib.upc_create(None, 0, None) == abi.UPC_ERR_INVALID_ARGUMENT
    assert b"null" in lib.upc_last_error()
    assert lib.upc_set_robot(None, None) == abi.UPC_ERR_INVALID_ARGUMENT
    assert lib.upc_forward_simulate(None, None, 1, None, 1, None, None, None, None) == abi.UPC_ERR_INVALID_ARGUMENT
    assert lib.upc_cluster_particles(None, None, 1, 1, None, None, None) == abi.UPC_ERR_INVALID_ARGUMENT
    with pytest.raises(ValueError):
        abi.check(lib.upc_set_lanes_per_particle(None, 3))


def test_no_silent_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from uncertainty_planning_core_b200 import B200Simulator, scenarios
    sc = scenarios.c1_se2(n=4, shape=(32, 32, 1), resolution=0.32)
    with pytest.raises((abi.UpcError, ValueError)):
        B200Simulator(sc["env"], sc["robot"])
