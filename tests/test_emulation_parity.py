"""CPU-only parity: the product's per-particle device program (uncertainty_planning_core_b200/csrc/upc_sim.cuh,
compiled for the host with one lane per particle by tests/emulation) against the independent oracle.
Bit-exact: configurations 0 ulp, collision flags and solver counters equal."""
import numpy as np
import pytest

import cases
from common import emulate_cull_check, emulate_forward_simulate, emulate_group_cull_check, max_ulp_diff
from uncertainty_planning_core_b200 import scenarios
from uncertainty_planning_core_b200 import abi


@pytest.mark.parametrize("culls", ["all", "no_fp32", "no_link", "no_group", "group_only", "no_shared_masks", "none"])
@pytest.mark.parametrize("corner_cells", [True, False])
@pytest.mark.parametrize("name", sorted(cases.GOLDEN_SIM_CASES))
def test_device_program_matches_oracle(oracle, name, corner_cells, culls):
    """the shortcuts (corner-cell table, single-precision pre-cull, link-level cull, group-level cull) in their combinations give the oracle's bits"""
    sc = cases.GOLDEN_SIM_CASES[name]()
    n = sc["starts"].shape[0]
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
    ocfg, ocoll, ostats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape)
    ecfg, ecoll, estats = emulate_forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], tape, corner_cells,
                                                   fp32_cull=culls in ("all", "no_link", "no_group", "no_shared_masks"),
                                                   link_cull=culls in ("all", "no_fp32", "no_group", "no_shared_masks"),
                                                   group_cull=culls in ("all", "no_fp32", "no_link", "group_only", "no_shared_masks"),
                                                   shared_masks=culls != "no_shared_masks")
    assert np.array_equal(ocoll, ecoll)
    assert max_ulp_diff(ocfg, ecfg) == 0
    for k in cases.SIM_STAT_KEYS:
        assert ostats[k] == estats[k], k


def _random_frames(rng, n, lo, hi, snap=None):
    frames = np.zeros((n, 12))
    for i in range(n):
        R = scenarios.rotation_from_axis_angle(rng.normal(size=3), rng.uniform(0.0, np.pi))
        t = rng.uniform(lo, hi, size=3)
        if snap is not None and i % 2 == 0:
            # adversarial: identity rotation, body origin within a few float ulps of a cell border (voxel centre planes)
            R = np.eye(3)
            t = (np.round(t / snap) + 0.5) * snap + rng.uniform(-3e-7, 3e-7, size=3)
        frames[i] = np.concatenate([R, np.asarray(t, dtype=float)[:, None]], axis=1).ravel()   # row-major [R | t]
    return frames


def test_single_precision_pre_cull_is_conservative():
    """every point the float pre-cull calls clear is clear for the exact FP64 test - random poses,
    poses whose points sit on cell borders, poses reaching outside the grid, several thresholds"""
    rng = np.random.default_rng(77)
    sc = scenarios.c2_se3(seed=21, n=1, shape=(64, 64, 64), resolution=0.0625)
    extent = 64 * 0.0625
    total_culled = total_hits = total_links = 0
    for threshold in (0.0, 0.01, 0.1):
        for frames in (_random_frames(rng, 300, 0.3, extent - 0.3), _random_frames(rng, 300, -0.2, extent + 0.2),
                       _random_frames(rng, 400, 0.3, extent - 0.3, snap=0.0625)):
            tested, culled, violations, hits, links_cleared, link_violations = emulate_cull_check(sc["env"], sc["robot"], frames, threshold)
            assert violations == 0 and link_violations == 0
            assert tested == frames.shape[0] * sc["robot"].desc.num_points
            total_culled += culled
            total_hits += hits
            total_links += links_cleared
    assert total_culled > 0 and total_hits > 0 and total_links > 0      # every outcome was exercised
    # planar grid (single layer in z): SE(2) body points sit exactly on the layer's centre plane
    s1 = scenarios.c1_se2(seed=11, n=1, shape=(128, 128, 1), resolution=0.08)
    frames = np.zeros((500, 12))
    for i in range(500):
        th = rng.uniform(-np.pi, np.pi)
        R = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
        frames[i] = np.concatenate([R, np.array([[rng.uniform(0.2, 10.0)], [rng.uniform(0.2, 10.0)], [0.0]])], axis=1).ravel()
    tested, culled, violations, hits, links_cleared, link_violations = emulate_cull_check(s1["env"], s1["robot"], frames, 0.0)
    assert violations == 0 and culled > 0 and hits > 0
    assert link_violations == 0 and 0 < links_cleared < 500


def test_link_level_cull_is_conservative_for_the_arm():
    """a link the dilated-minimum lookup calls clear has no body point the exact FP64 test would flag - every link of
    the 7-DoF arm as its own single-link robot, poses all over (and beyond) the grid, spheres with radius"""
    rng = np.random.default_rng(78)
    sc = cases.linked_contact()
    extent = 64 * 0.0625
    pts = sc["robot"].points[:16].copy()                 # one link: 16 spheres along z, radius 0.02
    link = abi.Robot(abi.ROBOT_SE3, [pts], [abi.make_axis() for _ in range(6)])
    cleared = hits = 0
    for threshold in (0.0, 0.00625, 0.05):
        for frames in (_random_frames(rng, 600, 0.2, extent - 0.2), _random_frames(rng, 300, -0.3, extent + 0.3),
                       _random_frames(rng, 300, 0.2, extent - 0.2, snap=0.0625)):
            tested, culled, violations, point_hits, links_cleared, link_violations = emulate_cull_check(sc["env"], link, frames, threshold)
            assert link_violations == 0 and violations == 0
            cleared += links_cleared
            hits += point_hits
    assert cleared > 100 and hits > 0


def test_group_level_cull_is_conservative():
    """a point group the dilated-minimum lookup calls clear holds no body point the exact FP64 test would flag - the
    SE(3) lattice (128 points: four chunks of four groups), one arm link (16 spheres with radius: four groups of four) and the planar
    SE(2) lattice, poses all over and beyond the grid and on cell borders, several thresholds"""
    rng = np.random.default_rng(79)
    sc = scenarios.c2_se3(seed=21, n=1, shape=(64, 64, 64), resolution=0.0625)
    arm = cases.linked_contact()
    link = abi.Robot(abi.ROBOT_SE3, [arm["robot"].points[:16].copy()], [abi.make_axis() for _ in range(6)])
    extent = 64 * 0.0625
    cleared_total = tested_total = 0
    for env, robot in ((sc["env"], sc["robot"]), (arm["env"], link)):
        for threshold in (0.0, 0.00625, 0.05):
            for frames in (_random_frames(rng, 400, 0.2, extent - 0.2), _random_frames(rng, 300, -0.3, extent + 0.3),
                           _random_frames(rng, 300, 0.2, extent - 0.2, snap=0.0625)):
                tested, cleared, violations, points = emulate_group_cull_check(env, robot, frames, threshold)
                assert violations == 0
                assert tested == frames.shape[0] * 4 * (robot.desc.num_points // 32 if robot.desc.num_points >= 32 else 1)
                cleared_total += cleared
                tested_total += tested
    assert 0 < cleared_total < tested_total                                 # both outcomes were exercised
    s1 = scenarios.c1_se2(seed=11, n=1, shape=(128, 128, 1), resolution=0.08)
    frames = np.zeros((600, 12))
    for i in range(600):
        th = rng.uniform(-np.pi, np.pi)
        R = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
        frames[i] = np.concatenate([R, np.array([[rng.uniform(0.2, 10.0)], [rng.uniform(0.2, 10.0)], [0.0]])], axis=1).ravel()
    tested, cleared, violations, points = emulate_group_cull_check(s1["env"], s1["robot"], frames, 0.0)
    assert violations == 0 and 0 < cleared < tested and tested == 600 * 8
