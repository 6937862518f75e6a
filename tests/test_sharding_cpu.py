"""world_size-2 gloo test of the N>1 path's host logic: contiguous sharding, per-particle tape keys, the single
gather of outcome records.  The per-rank compute is the CPU oracle here (the GPU test suite covers the kernels);
the sharded, gathered result must equal the unsharded run bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    from oracle import binding as oracle
    from uncertainty_planning_core_b200 import abi, sharding
    oracle.set_num_threads(1)
    sc = cases.se2_shortcut_per_target()          # per-particle targets: shards need their slice of both arrays
    n = sc["starts"].shape[0] - sc["starts"].shape[0] % world
    lo, hi = sharding.shard_bounds(n, rank, world)
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], hi - lo, first_particle=lo)
    cfg, coll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"][lo:hi], sc["targets"][lo:hi], tape)
    cp = cases._cparams(abi.FEATURE_NONE, 0.125, 0.1)
    labels, ncl, _ = oracle.cluster_particles(sc["env"], sc["robot"], cp, cfg)
    packed, m = sharding.pack_outcomes(torch.from_numpy(labels), torch.from_numpy(coll.astype(np.uint8)),
                                       torch.from_numpy(np.ascontiguousarray(cfg.T)))
    gathered = sharding.gather_outcomes(packed, world)
    parts = sharding.merge_cluster_counts(gathered, m, 3, world)
    all_cfg = np.concatenate([p[2] for p in parts], axis=0)
    all_coll = np.concatenate([p[1] for p in parts], axis=0)
    np.savez(os.path.join(result_dir, "rank%d.npz" % rank), cfg=all_cfg, coll=all_coll, labels=np.concatenate([p[0] for p in parts]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_single_run(tmp_path, oracle):
    import cases
    from uncertainty_planning_core_b200 import abi, sharding
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sc = cases.se2_shortcut_per_target()
    n = sc["starts"].shape[0] - sc["starts"].shape[0] % world
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], n)
    cfg, coll, _ = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], sc["starts"][:n], sc["targets"][:n], tape)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    for r in (r0, r1):
        assert np.array_equal(r["cfg"], cfg) and np.array_equal(r["coll"], coll)
    assert np.array_equal(r0["labels"], r1["labels"])
    # each shard clustered on its own: labels restart at 0 inside every shard
    lo, hi = sharding.shard_bounds(n, 1, world)
    assert r0["labels"][lo] == 0


def test_shard_bounds_cover_everything_once():
    from uncertainty_planning_core_b200 import sharding
    for total in (0, 1, 7, 4096, 10001):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
