"""Counter-based noise: the oracle's restatement is pinned to published known answers (Random123 Philox
vectors, scipy's inverse normal CDF), the product's generator must equal the oracle's bit for bit, and the seeded device
program (CPU emulation here, the kernels in test_gpu_parity.py) must equal the oracle replaying the oracle's tape."""
import numpy as np
import pytest

import cases
from common import emulate_forward_simulate_seeded, max_ulp_diff
from uncertainty_planning_core_b200 import abi

# Random123 kat_vectors, philox4x32 with 10 rounds: (counter, key, expected)
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("counter,key,expected", PHILOX_KAT)
def test_philox_known_answers(oracle, counter, key, expected):
    assert oracle.philox4x32_10(key, counter).tolist() == expected


def test_truncated_quantile_matches_scipy(oracle):
    """the degree-8 / degree-8 rational approximation of DESIGN.md 4.2 against scipy's inverse normal CDF on the whole truncated range"""
    from scipy.special import ndtri
    W = 0.9544997361036415856
    t = np.concatenate([np.linspace(-0.5, 0.5, 400001)[1:-1], [2.0 ** -53, -2.0 ** -53, 0.5 - 2.0 ** -53, -0.5 + 2.0 ** -53]])
    z = oracle.truncated_normal_quantile(t)
    assert np.max(np.abs(z - np.clip(ndtri(0.5 + t * W), -2.0, 2.0))) < 1e-13
    assert np.all(np.diff(z[:-4]) >= 0) and z.min() >= -2.0 and z.max() <= 2.0       # monotone, bounded
    assert np.array_equal(z, -oracle.truncated_normal_quantile(-t))                        # odd
    assert abs(oracle.truncated_normal_quantile([0.5 - 2.0 ** -53])[0] - 2.0) < 1e-12


def test_centered_uniform_is_exact(oracle):
    rng = np.random.default_rng(5)
    hi = rng.integers(0, 2 ** 32, 100000, dtype=np.uint64).astype(np.uint32)
    lo = rng.integers(0, 2 ** 32, 100000, dtype=np.uint64).astype(np.uint32)
    hi[:4] = [0, 0xffffffff, 0x80000000, 0x7fffffff]
    lo[:4] = [0, 0xffffffff, 0, 0xffffffff]
    t = oracle.centered_uniform52(hi, lo)
    k = (hi.astype(np.uint64) << np.uint64(20)) | (lo.astype(np.uint64) >> np.uint64(12))
    import fractions
    for i in list(range(8)) + [17, 999]:
        want = fractions.Fraction(2 * int(k[i]) + 1, 2 ** 53) - fractions.Fraction(1, 2)
        assert fractions.Fraction(float(t[i])) == want
    assert t.min() > -0.5 and t.max() < 0.5 and not np.any(t == 0.0)
    assert t[0] == -0.5 + 2.0 ** -53 and t[1] == 0.5 - 2.0 ** -53


def test_log_is_within_one_ulp_of_libm(oracle):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(0.02, 0.08, 200000), np.exp(rng.uniform(-700, 700, 200000)), [1.0, 2.0, 0.5, np.sqrt(2.0), 1.0 / np.sqrt(2.0)]])
    assert max_ulp_diff(oracle.log(x), np.log(x)) <= 1
    assert oracle.log([1.0])[0] == 0.0


def test_draws_follow_the_truncated_normal(oracle):
    from scipy import stats
    z = np.concatenate([oracle.standard_draws(1234, 0, 200000, step, 1, axis) for step, axis in ((0, 0), (0, 1), (5, 2))])
    assert z.min() >= -2.0 and z.max() <= 2.0
    assert stats.kstest(z, stats.truncnorm(-2.0, 2.0).cdf).pvalue > 1e-3
    assert abs(z.mean()) < 0.005 and abs(z.var() - stats.truncnorm(-2.0, 2.0).var()) < 0.005
    # streams are keyed: another seed, particle offset, step, slot or axis gives another stream; the same key the same values
    a = oracle.standard_draws(1234, 0, 1000, 0, 1, 0)
    assert np.array_equal(a, z[:1000])
    assert np.array_equal(oracle.standard_draws(1234, 100, 900, 0, 1, 0), a[100:])
    for other in (oracle.standard_draws(1235, 0, 1000, 0, 1, 0), oracle.standard_draws(1234, 0, 1000, 1, 1, 0),
                  oracle.standard_draws(1234, 0, 1000, 0, 2, 0), oracle.standard_draws(1234, 0, 1000, 0, 1, 1)):
        assert abs(np.corrcoef(a, other)[0, 1]) < 0.15
    # axes 2j and 2j+1 share one Philox block but use disjoint halves of it
    assert abs(np.corrcoef(oracle.standard_draws(9, 0, 100000, 3, 1, 2), oracle.standard_draws(9, 0, 100000, 3, 1, 3))[0, 1]) < 0.02


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_SIM_CASES))
def test_product_tape_equals_oracle_tape(oracle, name):
    sc = cases.GOLDEN_SIM_CASES[name]()
    n = sc["starts"].shape[0]
    for seed, first in ((77, 0), (2 ** 63 + 12345, 5), (1, 2 ** 33 + 7)):
        ot = oracle.fill_noise_tape_counter(seed, sc["robot"], sc["params"], n, first_particle=first)
        pt = abi.fill_noise_tape_seeded(seed, sc["robot"], sc["params"], n, first_particle=first)
        assert np.array_equal(ot, pt), (name, seed, first)
    # noise-free axes get exact zeros; sensor slot scaled by max_sensor_noise / 2, actuator slots by 1/2
    d = sc["robot"].desc
    for a in range(sc["robot"].dof):
        sb, ab = abs(d.axis[a].max_sensor_noise), abs(d.axis[a].max_actuator_noise)
        assert np.max(np.abs(ot[:, 0, a, :])) <= sb * (1.0 + 1e-15) and (sb > 0 or not ot[:, 0, a, :].any())
        assert np.max(np.abs(ot[:, 1:, a, :])) <= 1.0 and (ab > 0 or not ot[:, 1:, a, :].any())


@pytest.mark.parametrize("name", sorted(cases.GOLDEN_SIM_CASES))
def test_seeded_device_program_matches_oracle(oracle, name):
    """simulate_particle<M, 1, SEEDED = true> (the code the seeded kernels run) against the oracle replaying the oracle's tape"""
    sc = cases.GOLDEN_SIM_CASES[name]()
    n = sc["starts"].shape[0]
    seed, first = 4242, 1000
    tape = oracle.fill_noise_tape_counter(seed, sc["robot"], sc["params"], n, first_particle=first)
    ocfg, ocoll, ostats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    ecfg, ecoll, estats = emulate_forward_simulate_seeded(sc["env"], sc["robot"], sc["params"], seed, sc["starts"], sc["targets"], first_particle=first)
    assert np.array_equal(ecoll, ocoll), name
    assert max_ulp_diff(ecfg, ocfg) == 0, name
    for k in cases.SIM_STAT_KEYS:
        assert estats[k] == ostats[k], (name, k)


def test_seeded_edge_batch_form_and_trace(oracle):
    """one start / target per block of particles (edge batches) equals the per-particle form, and the per-interval trace equals
    the oracle's (ForwardSimulateRobot with enable_tracing, simple_simulator_interface.hpp:40-86, 621)"""
    sc = cases.se2_contact()
    P, E = 12, 8
    rng = np.random.default_rng(8)
    es = sc["starts"][:E] + rng.uniform(-0.02, 0.02, size=(E, 3))
    et = np.repeat(sc["targets"], E, axis=0) + rng.uniform(-0.05, 0.05, size=(E, 3))
    full_s, full_t = np.repeat(es, P, axis=0), np.repeat(et, P, axis=0)
    a = emulate_forward_simulate_seeded(sc["env"], sc["robot"], sc["params"], 5, full_s, full_t)
    b = emulate_forward_simulate_seeded(sc["env"], sc["robot"], sc["params"], 5, es, et, n=E * P)
    c = emulate_forward_simulate_seeded(sc["env"], sc["robot"], sc["params"], 5, full_s, et, n=E * P, trace=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[0], c[0])
    tape = oracle.fill_noise_tape_counter(5, sc["robot"], sc["params"], E * P)
    ocfg, ocoll, otrace = oracle.forward_simulate_traced(sc["env"], sc["robot"], sc["params"], full_s, full_t, tape)
    assert np.array_equal(ocfg, a[0]) and np.array_equal(ocoll, a[1])
    etrace = c[3]
    assert np.array_equal(etrace["steps"], otrace["steps"])
    assert np.array_equal(etrace["configs"], otrace["configs"]) and np.array_equal(etrace["controls"], otrace["controls"])
    # the last traced configuration of every particle is its final configuration
    last = otrace["steps"][0] - 1
    assert np.array_equal(otrace["configs"][last, np.arange(E * P)], ocfg)
