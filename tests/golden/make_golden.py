"""Regenerates the committed fixtures under tests/golden/.

  pid_golden.json      produced by the REFERENCE's own, unmodified simple_pid_controller.hpp
                       (compiled into oracle/_ref/libref_pid.so by oracle/Makefile; needs /root/reference)
  sim_*.npz, sim_seeded_*.npz, cluster_*.npz   small seeded scenarios (tape noise / counter-based noise) with the ORACLE's outputs, so that the GPU parity tests
                       can also be checked against files that travel with the repository

Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import binding as ob  # noqa: E402
from uncertainty_planning_core_b200 import abi  # noqa: E402
import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pid_golden():
    ob.build()
    assert ob.load_ref_pid() is not None, "oracle/_ref/libref_pid.so missing (needs /root/reference)"
    rng = np.random.default_rng(20261019)
    out = []
    gains = [(1.0, 0.1, 0.01, 5.0), (5.0, 0.0, 0.0, 0.0), (-2.0, 0.5, -0.05, 0.2), (0.7, 3.0, 0.0, 0.05)]
    for kp, ki, kd, clamp in gains:
        errors = rng.normal(scale=0.5, size=16)
        dts = rng.choice([0.01, 0.02, 0.005], size=16)
        terms = ob.ref_pid_sequence(kp, ki, kd, clamp, errors, dts)
        out.append(dict(kp=kp, ki=ki, kd=kd, clamp=clamp, errors=[float.hex(float(e)) for e in errors],
                        dts=[float.hex(float(d)) for d in dts], terms=[float.hex(float(t)) for t in terms]))
    # the known-answer pair quoted in SURVEY.md section 4
    terms = ob.ref_pid_sequence(1.0, 0.1, 0.01, 5.0, np.array([0.5, 0.4]), np.array([0.01, 0.01]))
    out.append(dict(kp=1.0, ki=0.1, kd=0.01, clamp=5.0, errors=[float.hex(0.5), float.hex(0.4)],
                    dts=[float.hex(0.01), float.hex(0.01)], terms=[float.hex(float(t)) for t in terms]))
    with open(os.path.join(HERE, "pid_golden.json"), "w") as f:
        json.dump(dict(source="reference simple_pid_controller.hpp:122-135 compiled unmodified", cases=out), f, indent=1)


def sim_golden():
    for name, make in cases.GOLDEN_SIM_CASES.items():
        sc = make()
        tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], sc["starts"].shape[0])
        cfg, coll, stats = ob.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
        np.savez_compressed(os.path.join(HERE, "sim_%s.npz" % name), configs=cfg, collided=coll,
                            stats=np.array([stats[k] for k in cases.SIM_STAT_KEYS], dtype=np.int64),
                            tape_checksum=np.array([np.sum(tape)]))


def seeded_sim_golden():
    """the same scenarios with the counter-based noise of DESIGN.md 4.2 (the oracle's own generator)"""
    for name, make in cases.GOLDEN_SIM_CASES.items():
        sc = make()
        tape = ob.fill_noise_tape_counter(cases.SEEDED_GOLDEN_SEED, sc["robot"], sc["params"], sc["starts"].shape[0], first_particle=cases.SEEDED_GOLDEN_FIRST)
        cfg, coll, stats = ob.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
        np.savez_compressed(os.path.join(HERE, "sim_seeded_%s.npz" % name), configs=cfg, collided=coll,
                            stats=np.array([stats[k] for k in cases.SIM_STAT_KEYS], dtype=np.int64),
                            tape_head=tape[:2, :, :, :4].copy())


def cluster_golden():
    for name, make in cases.GOLDEN_CLUSTER_CASES.items():
        c = make()
        labels, ncl, stats = ob.cluster_particles(c["env"], c["robot"], c["cparams"], c["configs"], c.get("n_sets", 1))
        np.savez_compressed(os.path.join(HERE, "cluster_%s.npz" % name), labels=labels, n_clusters=ncl,
                            fallback=np.array([stats["cluster_fallback_calls"]]))


if __name__ == "__main__":
    pid_golden()
    sim_golden()
    seeded_sim_golden()
    cluster_golden()
    print("golden fixtures written to", HERE)
