"""CPU-only parity: the product's per-particle device program (uncertainty_planning_core_b200/csrc/upc_sim.cuh,
compiled for the host with one lane per particle by tests/emulation) against the independent oracle.
Bit-exact: configurations 0 ulp, collision flags and solver counters equal."""
import numpy as np
import pytest

import cases
from common import emulate_forward_simulate, max_ulp_diff
from uncertainty_planning_core_b200 import abi


@pytest.mark.parametrize("corner_cells", [True, False])
@pytest.mark.parametrize("name", sorted(cases.GOLDEN_SIM_CASES))
def test_device_program_matches_oracle(oracle, name, corner_cells):
    sc = cases.GOLDEN_SIM_CASES[name]()
    n = sc["starts"].shape[0]
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
    ocfg, ocoll, ostats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    ecfg, ecoll, estats = emulate_forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape, corner_cells)
    assert np.array_equal(ocoll, ecoll)
    assert max_ulp_diff(ocfg, ecfg) == 0
    for k in cases.SIM_STAT_KEYS:
        assert ostats[k] == estats[k], k
