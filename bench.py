#!/usr/bin/env python
"""Benchmark of the particle hot path (BASELINE.json: particle-steps/sec for sim + SDF collision, and particles
clustered/sec).

A "step" is one pass of the hot path over one planner edge: ForwardSimulateRobots on N particles for S controller
intervals (noisy actuation, sphere-vs-SDF collision, contact resolution) followed by ClusterParticles of the N
outcomes.  Default workload = BASELINE config 2: SE(3) robot, 10,000 particles per edge, 128^3 SDF, 128 body
points, 64 controller intervals.  With --gpus N every rank owns its own edge (weak scaling), the SDF is
replicated, and the step ends with ONE NCCL all-gather of the outcome records.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # the CUDA path
  python bench.py --impl reference ...                           # the CPU path (oracle) on the box's host cores

Prints ONE JSON line (see DESIGN.md section 7 for every key).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3"])
    ap.add_argument("--particles", type=int, default=0, help="override the workload's particle count")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per particle (0 = built-in choice)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-big-batch", action="store_true", help="skip the 10^6-particle throughput-regime measurement of the simulation kernel")
    return ap.parse_args()


def make_workload(args, edge):
    from uncertainty_planning_core_b200 import scenarios
    if args.workload == "c1":
        sc = scenarios.c1_se2(n=args.particles or 1000)
        name = "C1 SE(2), %d particles/edge, 256x256 SDF, P=64, S=64"
    elif args.workload == "c3":
        sc = scenarios.c3_linked(n=args.particles or 100000)
        name = "C3 7-DoF arm, %d particles/edge, 256^3 SDF, P=128, S=64"
    else:
        sc = scenarios.c2_se3(n=args.particles or 10000, edge=edge)
        name = "C2 SE(3), %d particles/edge, 128^3 SDF, P=128, S=64, contact resolution"
    sc["label"] = name % sc["starts"].shape[0]
    return sc


def cluster_params():
    from uncertainty_planning_core_b200 import abi
    cp = abi.ClusterParams()
    cp.feature_type = abi.FEATURE_CRS
    cp.signature_matching_threshold = 0.125
    cp.distance_clustering_threshold = 0.1
    return cp


def sdf_bytes_per_point(sc):
    """one trilinear cell fetch = 8 corners x 4 B (4 corners for the 2-D grid)."""
    return 16 if sc["env"].sdf.shape[2] == 1 else 32


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe).
    The sampler is started ahead of the warm-up (nvidia-smi needs ~0.2 s to come up); samples are kept when their
    timestamp falls inside [mark_begin, mark_end], the timed region."""

    def __init__(self, index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.t0 = self.t1 = None
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(f[1]), float(f[2]), [nm for k, nm in enumerate(names) if f[5 + k].lower().startswith("active")]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except OSError:
            pass
        window = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        scope = "timed region"
        if len(window) < 3:
            window, scope = rows, "warm-up + timed region (the timed region is shorter than three sampling periods)"
        if window:
            reasons = sorted({nm for r in window for nm in r[3]})
            out.update(sm_mhz=float(np.median([r[1] for r in window])), sm_max_mhz=float(max(r[2] for r in window)), reasons=reasons,
                       samples=len(window), scope=scope)
        return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- CPU arm

def cpu_path(sc, cp, sample, tape, reps=1):
    """The CPU path (oracle, all host threads) on the first `sample` particles of the workload: returns
    (seconds for sim, seconds for clustering, particle-steps executed, clusters)."""
    from oracle import binding as oracle
    starts = sc["starts"][:sample]
    targets = sc["targets"] if sc["targets"].shape[0] == 1 else sc["targets"][:sample]
    t_sim = t_cl = 0.0
    steps = 0
    ncl = 0
    for _ in range(reps):
        t0 = time.perf_counter()
        cfg, coll, stats = oracle.forward_simulate(sc["env"], sc["robot"], sc["params"], starts, targets, tape)
        t1 = time.perf_counter()
        labels, n_clusters, _ = oracle.cluster_particles(sc["env"], sc["robot"], cp, cfg)
        t2 = time.perf_counter()
        t_sim += t1 - t0
        t_cl += t2 - t1
        steps += stats["particle_steps"]
        ncl = int(n_clusters[0])
    return t_sim, t_cl, steps, ncl


def auto_cpu_sample(sc, cp, want_seconds=12.0):
    """Size the CPU sample to roughly `want_seconds` of host work.  Simulation time is linear in the particle count;
    the clustering half is between quadratic (one blob) and cubic (hierarchical path), so its growth is fitted from
    two probes."""
    from uncertainty_planning_core_b200 import abi
    total = sc["starts"].shape[0]
    p1, p2 = min(96, total), min(192, total)
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], p2)
    s1, c1, _, _ = cpu_path(sc, cp, p1, np.ascontiguousarray(tape[..., :p1]))
    s2, c2, _, _ = cpu_path(sc, cp, p2, tape)
    sim_per = max(s2 / p2, 1e-9)
    power = 2.0
    if p2 > p1 and c1 > 1e-5 and c2 > c1:
        power = min(3.2, max(1.0, np.log(c2 / c1) / np.log(p2 / p1)))
    coef = max(c2, 1e-7) / (p2 ** power)
    cap = min(total, 8192)
    n = p2
    while n * 2 <= cap and sim_per * (n * 2) + coef * (n * 2) ** power <= want_seconds:
        n *= 2
    lo, hi = n, min(cap, n * 2)
    while hi - lo > 16:
        mid = (lo + hi) // 2
        if sim_per * mid + coef * mid ** power <= want_seconds:
            lo = mid
        else:
            hi = mid
    return int(lo)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as oracle
    from uncertainty_planning_core_b200 import abi
    oracle.build()
    oracle.set_num_threads(os.cpu_count() or 1)
    sc = make_workload(args, 0)
    cp = cluster_params()
    cores = oracle.num_threads()
    sample = args.cpu_sample or auto_cpu_sample(sc, cp, want_seconds=8.0)
    tape = abi.fill_noise_tape(sc["seed"], sc["robot"], sc["params"], sample)
    S = sc["params"].num_steps
    for _ in range(min(args.warmup, 1)):
        cpu_path(sc, cp, sample, tape)
    t0 = time.perf_counter()
    t_sim = t_cl = 0.0
    for _ in range(args.steps):
        a, b, _, ncl = cpu_path(sc, cp, sample, tape)
        t_sim += a
        t_cl += b
    elapsed = time.perf_counter() - t0
    value = sample * S * args.steps / elapsed
    desc = "first %d of %d particles of the edge, %d controller intervals, sim + CRS/distance clustering, per step" % (sample, sc["starts"].shape[0], S)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (seeded boxes SDF, replicated start, truncated-normal noise tape)",
        "config": {"workload": sc["label"], "sample": desc, "threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "sim_value": sample * S * args.steps / t_sim, "clustered_per_sec": sample * args.steps / max(t_cl, 1e-12)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm

def run_b200(args):
    import torch
    import torch.distributed as dist
    from uncertainty_planning_core_b200 import B200Simulator, abi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # every rank runs its own attempt of the same planner edge (the planner repeats an edge `edge_attempt_count` times,
    # uncertainty_planning_core.hpp:61): same endpoints, independent noise, equal work per GPU
    sc = make_workload(args, edge=0)
    cp = cluster_params()
    params = sc["params"]
    robot, env = sc["robot"], sc["env"]
    n = sc["starts"].shape[0]
    S = params.num_steps
    width, dof = robot.width, robot.dof
    lib = abi.load_library()
    sim = B200Simulator(env, robot, device=local_rank)
    if args.lanes:
        sim.SetLanesPerParticle(args.lanes)
    h = sim.handle

    # ---- inputs: a ring of distinct noise tapes whose total footprint exceeds L2, so no step finds its tape cached
    tape_bytes = int(np.prod(abi.tape_shape(params, dof, n))) * 8
    ring = max(2, int(np.ceil(1.5 * L2_BYTES / tape_bytes)))
    ring = min(ring, 8)
    pinned_tapes = []
    for r in range(ring):
        t = torch.empty(abi.tape_shape(params, dof, n), dtype=torch.float64).pin_memory()
        # particle keys are global: rank r, ring slot k uses particles [(r*ring+k)*n, ...)
        abi.check(lib.upc_fill_noise_tape(sc["seed"], C.byref(robot.desc), C.byref(params), (rank * ring + r) * n, n, n, t.data_ptr()))
        pinned_tapes.append(t)
    d_tapes = [t.to(dev, non_blocking=True) for t in pinned_tapes]
    h_starts = torch.from_numpy(np.ascontiguousarray(sc["starts"].T)).pin_memory()
    h_targets = torch.from_numpy(np.ascontiguousarray(sc["targets"].T)).pin_memory()
    n_targets = sc["targets"].shape[0]
    d_starts, d_targets = h_starts.to(dev), h_targets.to(dev)
    d_out = torch.zeros((width, n), dtype=torch.float64, device=dev)
    d_coll = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_labels = torch.zeros(n, dtype=torch.int32, device=dev)
    d_ncl = torch.zeros(1, dtype=torch.int32, device=dev)
    h_out = torch.empty((width, n), dtype=torch.float64).pin_memory()
    h_coll = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_labels = torch.empty(n, dtype=torch.int32).pin_memory()
    h_ncl = torch.empty(1, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()
    # a side stream: its handle is non-NULL (NULL would select the library's own stream, which torch events cannot see)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    assert stream != 0

    # two sets of outcome buffers: while batch k simulates, the host packs and ships batch k-1 (multi-GPU runs)
    outs = [(d_out, d_coll, d_labels, d_ncl)]
    if world > 1:
        outs.append((torch.zeros_like(d_out), torch.zeros_like(d_coll), torch.zeros_like(d_labels), torch.zeros_like(d_ncl)))
    in_flight = []
    pending = []

    def ship(buf):
        """pack + ONE all-gather of a finished batch; NCCL's own stream carries it, at most two are in flight"""
        o, c, l, _ = outs[buf]
        packed, _ = sharding.pack_outcomes(l, c, o)
        gathered, work = sharding.gather_outcomes(packed, world, async_op=True)
        in_flight.append((gathered, packed, work))
        while len(in_flight) > 2:
            _, _, w = in_flight.pop(0)
            if w is not None:
                w.wait()

    def device_step(k, ev=None):
        tape = d_tapes[k % ring]
        buf = k % len(outs)
        o, c, l, ncl = outs[buf]
        if ev:
            ev[0].record()
        abi.check(lib.upc_forward_simulate_device(h, C.byref(params), n, d_starts.data_ptr(), n_targets, d_targets.data_ptr(),
                                                  tape.data_ptr(), o.data_ptr(), c.data_ptr(), stream))
        if ev:
            ev[1].record()
        if world > 1 and pending:
            ship(pending.pop(0))   # host-side packing of batch k-1 overlaps the simulation of batch k just enqueued
        abi.check(lib.upc_cluster_particles_device(h, C.byref(cp), 1, n, o.data_ptr(), l.data_ptr(), ncl.data_ptr(), stream))
        if ev:
            ev[2].record()
        if world > 1:
            pending.append(buf)

    def drain():
        while pending:
            ship(pending.pop(0))
        while in_flight:
            _, _, w = in_flight.pop(0)
            if w is not None:
                w.wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for k in range(args.warmup):
        device_step(k)
    drain()
    barrier()
    sim.ResetStatistics()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_begin()
    t_begin.record()
    for k in range(args.steps):
        device_step(args.warmup + k, evs[k])
    drain()
    t_end.record()
    barrier()
    if sampler:
        sampler.mark_end()
    total_ms = t_begin.elapsed_time(t_end)
    clocks = sampler.stop() if sampler else None
    sim_ms = sum(e[0].elapsed_time(e[1]) for e in evs)
    cl_ms = sum(e[1].elapsed_time(e[2]) for e in evs)
    stats = sim.GetStatistics()
    t = torch.tensor([total_ms, sim_ms, cl_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, sim_ms, cl_ms = [float(v) for v in t.tolist()]
    particle_steps = n * S * args.steps * world
    value = particle_steps / (total_ms * 1e-3)

    # ---- end to end through the host-pointer C ABI: pinned host buffers in, results back on the host, every step
    def e2e_step(k):
        tape = pinned_tapes[k % ring]
        abi.check(lib.upc_forward_simulate(h, C.byref(params), n, h_starts.data_ptr(), n_targets, h_targets.data_ptr(), tape.data_ptr(),
                                           h_out.data_ptr(), h_coll.data_ptr()))
        abi.check(lib.upc_cluster_particles(h, C.byref(cp), 1, n, h_out.data_ptr(), h_labels.data_ptr(), h_ncl.data_ptr()))
        return int(h_ncl[0])

    for k in range(min(args.warmup, 3)):
        e2e_step(k)
    barrier()
    sim.ResetStatistics()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(args.warmup + k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_stats = sim.GetStatistics()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = particle_steps / e2e_s

    if rank == 0:
        # ---- roofline of the dominant kernel (forward_simulate_kernel): algorithmic SDF gather bytes / measured launch time
        hbm_peak, peak_src = measured_peaks()
        P = robot.num_points
        bpp = sdf_bytes_per_point(sc)
        ps_launch = stats["particle_steps"] / args.steps
        ri_launch = stats["resolver_iterations"] / args.steps
        alg_bytes = (ps_launch + ri_launch) * P * bpp            # M*P*corners*4 per particle-step + R_used*P*32
        sim_ms_launch = sim_ms / args.steps
        achieved = alg_bytes / (sim_ms_launch * 1e-3) / 1e9
        l2_gbs = C.c_double(0.0)
        fp64 = C.c_double(0.0)
        abi.check(lib.upc_measure_l2_gather_gbs(h, env.sdf.nbytes, C.byref(l2_gbs)))
        abi.check(lib.upc_measure_fp64_tflops(h, C.byref(fp64)))
        compulsory = (tape_bytes + 2 * width * n * 8 + n) / (sim_ms_launch * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_sim_kernel_traffic.json")
        if args.workload == "c2" and not args.particles and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("traffic")   # dram bytes of one launch of this kernel, from the committed ncu capture
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "kernel": "forward_simulate_kernel", "kernel_ms": sim_ms_launch, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "l2_gather_peak_gbs": l2_gbs.value, "frac_of_l2_gather": achieved / max(l2_gbs.value, 1e-9),
                    "compulsory_hbm_gbs": compulsory, "fp64_fma_peak_tflops": fp64.value,
                    "note": "SDF (%.0f MiB) is L2/L1-resident by design, so the gather rate may exceed the HBM copy peak" % (env.sdf.nbytes / 2 ** 20)}
        # ---- the same kernel with the machine full (rank 0, N = 1, default workload): 10^6 particles of the same edge, one lane each,
        # first 16 controller intervals - the throughput regime, where the SDF gathers are the bound (north_star: >= 60 % of the L2 roofline)
        if world == 1 and args.workload == "c2" and not args.particles and not args.no_big_batch:
            nb, sb = 1000000, 16
            pb = abi.make_sim_params(sb, params.controller_interval, env.resolution, allow_contacts=True, max_microsteps=params.max_microsteps)
            tb = torch.from_numpy(abi.fill_noise_tape(sc["seed"] + 99, robot, pb, nb)).to(dev)
            sbig = torch.from_numpy(np.ascontiguousarray(np.tile(sc["starts"][:1], (nb, 1)).T)).to(dev)
            obig = torch.zeros_like(sbig)
            cbig = torch.zeros(nb, dtype=torch.uint8, device=dev)
            best = 1e30
            for rep in range(4):
                sim.ResetStatistics()
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record()
                abi.check(lib.upc_forward_simulate_device(h, C.byref(pb), nb, sbig.data_ptr(), n_targets, d_targets.data_ptr(), tb.data_ptr(),
                                                          obig.data_ptr(), cbig.data_ptr(), stream))
                b1.record()
                torch.cuda.synchronize()
                if rep:
                    best = min(best, b0.elapsed_time(b1))
            bst = sim.GetStatistics()
            big_bytes = (bst["particle_steps"] + bst["resolver_iterations"]) * P * bpp
            big_gbs = big_bytes / (best * 1e-3) / 1e9
            roofline["big_batch"] = {"particles": nb, "controller_intervals": sb, "kernel_ms": best, "particle_steps_per_s": nb * sb / (best * 1e-3),
                                     "achieved": big_gbs, "unit": "GB/s", "frac": big_gbs / hbm_peak, "frac_of_l2_gather": big_gbs / max(l2_gbs.value, 1e-9),
                                     "tape_gbs": tb.numel() * 8 / (best * 1e-3) / 1e9}
            del tb, sbig, obig, cbig
        # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import binding as oracle
            oracle.set_num_threads(os.cpu_count() or 1)
            sample = args.cpu_sample or auto_cpu_sample(sc, cp)
            ctape = abi.fill_noise_tape(sc["seed"], robot, params, sample)
            t_sim, t_cl, steps_done, ncl = cpu_path(sc, cp, sample, ctape)
            cpu = {"value": sample * S / (t_sim + t_cl), "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                   "sample": "first %d of %d particles of the same edge, %d controller intervals, sim + clustering, 1 pass" % (sample, n, S),
                   "sim_value": sample * S / t_sim, "clustered_per_sec": sample / max(t_cl, 1e-12)}
            # SURVEY 8d extras: a 1-thread figure, and the distance pass in the reference's own shape (std::function per pair +
            # a heap-allocated delta vector per pair, uncertainty_contact_planning.hpp:1955-1960) next to the plain loop
            all_threads = oracle.num_threads()
            oracle.set_num_threads(1)
            small = min(256, sample)
            t1 = time.perf_counter()
            cfg1, _, _ = oracle.forward_simulate(env, robot, params, sc["starts"][:small], sc["targets"], np.ascontiguousarray(ctape[..., :small]))
            cpu["sim_value_1_thread"] = small * S / (time.perf_counter() - t1)
            oracle.set_num_threads(all_threads)
            m = min(2048, sample)
            fin = oracle.forward_simulate(env, robot, params, sc["starts"][:m], sc["targets"], np.ascontiguousarray(ctape[..., :m]))[0]
            t1 = time.perf_counter(); oracle.pair_distances(robot, fin); t2 = time.perf_counter()
            oracle.pair_distances(robot, fin, reference_style=True); t3 = time.perf_counter()
            cpu["distance_matrix_pairs_per_sec"] = {"plain_loop": m * (m - 1) / 2 / (t2 - t1), "reference_style": m * (m - 1) / 2 / (t3 - t2), "n": m}
        h2d = e2e_stats["h2d_bytes"] / args.steps
        d2h = e2e_stats["d2h_bytes"] / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (seeded boxes SDF, replicated start, truncated-normal noise tape)",
            "config": {"workload": sc["label"], "edges_per_step": world, "particles_per_gpu": n,
                       "sharding": "one attempt of the edge per GPU (same endpoints, independent noise), SDF replicated, one all-gather of outcome records per step, overlapped with the next step's simulation",
                       "clustering": "CRS signatures (thr 0.125) + distance pass (thr 0.1)",
                       "lanes_per_particle": args.lanes or "auto",
                       "l2": "ring of %d distinct %.0f MB noise tapes (%.0f MB > 126 MB L2) so the streamed input is never cached; the %.0f MiB SDF is "
                             "deliberately L2-resident across steps, as it is across the edges of a planner run" % (ring, tape_bytes / 1e6, ring * tape_bytes / 1e6, env.sdf.nbytes / 2 ** 20)},
            "sim": {"value": particle_steps / (sim_ms * 1e-3), "unit": UNIT, "ms_per_step": sim_ms / args.steps},
            "clustering": {"value": n * args.steps * world / (cl_ms * 1e-3), "unit": "particles clustered/s", "ms_per_step": cl_ms / args.steps,
                           "clusters": int(d_ncl.item()), "fallback_sets": int(stats["cluster_fallback_calls"])},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(stats["kernel_launches"]),
            "contact": {"collided_fraction": stats["collided_particles"] / max(stats["particles_simulated"], 1),
                        "resolver_iterations_per_particle_step": stats["resolver_iterations"] / max(stats["particle_steps"], 1)},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    barrier()
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    # torchrun exports OMP_NUM_THREADS=1; the host-side pieces (tape generation, the CPU arm) are OpenMP code and the
    # CPU arm is meant to use every host core ("with all the host threads it can use")
    _world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    _cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(_cores if a.impl == "reference" else max(1, _cores // max(1, _world)))
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
