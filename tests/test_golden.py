"""The oracle against the committed fixtures under tests/golden/ (written by tests/golden/make_golden.py):
guards against silent drift of the oracle between the container that generated them and any other box."""
import os

import numpy as np
import pytest

import cases
from uncertainty_planning_core_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_SIM_CASES))
def test_oracle_simulation_matches_fixture(oracle, name):
    g = np.load(os.path.join(HERE, "golden", "sim_%s.npz" % name))
    sc = cases.GOLDEN_SIM_CASES[name]()
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], sc["starts"].shape[0])
    assert float(np.sum(tape)) == float(g["tape_checksum"][0])
    cfg, coll, stats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    assert np.array_equal(cfg, g["configs"]) and np.array_equal(coll, g["collided"])
    assert [stats[k] for k in cases.SIM_STAT_KEYS] == g["stats"].tolist()


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_SIM_CASES))
def test_oracle_seeded_simulation_matches_fixture(oracle, name):
    g = np.load(os.path.join(HERE, "golden", "sim_seeded_%s.npz" % name))
    sc = cases.GOLDEN_SIM_CASES[name]()
    tape = oracle.fill_noise_tape_counter(cases.SEEDED_GOLDEN_SEED, sc["robot"], sc["params"], sc["starts"].shape[0], first_particle=cases.SEEDED_GOLDEN_FIRST)
    assert np.array_equal(tape[:2, :, :, :4], g["tape_head"])
    cfg, coll, stats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    assert np.array_equal(cfg, g["configs"]) and np.array_equal(coll, g["collided"])
    assert [stats[k] for k in cases.SIM_STAT_KEYS] == g["stats"].tolist()


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_CLUSTER_CASES))
def test_oracle_clustering_matches_fixture(oracle, name):
    g = np.load(os.path.join(HERE, "golden", "cluster_%s.npz" % name))
    c = cases.GOLDEN_CLUSTER_CASES[name]()
    labels, ncl, stats = oracle.cluster_particles(c["env"], c["robot"], c["cparams"], c["configs"], c.get("n_sets", 1))
    assert np.array_equal(labels, g["labels"]) and np.array_equal(ncl, g["n_clusters"])
