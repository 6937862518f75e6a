#!/usr/bin/env python
"""Benchmark of the particle hot path (BASELINE.json: particle-steps/sec for sim + SDF collision, and particles
clustered/sec, at 1/2/4/8 B200).

A "step" is one pass of the hot path over one planner batch.  Default workload = BASELINE config 5, the configuration the
metric is quoted on: 4,096 candidate RRT edges x 1,000 particles on the SE(2) environment of config 1 - ONE
ForwardSimulateRobots call over 4,096,000 particles (the targets.size() == N form, uncertainty_contact_planning.hpp:1826;
64 controller intervals of noisy actuation, point-vs-SDF collision, contact resolution) followed by ONE clustering pass
over the 4,096 outcome sets (ClusterParticles, :1986-2034).  Noise is drawn inside the kernel (counter-based).
With --gpus N the edges are split into contiguous blocks (strong scaling), the SDF is replicated, every rank simulates and
clusters its block straight into its slot of the gather buffer, and the step ends with ONE NCCL all-gather of the outcome
records issued by the library's own communicator (upc_comm_*); torch.distributed (gloo) only carries the 128-byte NCCL id,
the barriers and the max-over-ranks of the timings.  Batches are pipelined over two streams: the simulation of batch k overlaps
the clustering (and exchange) of batch k - 1 - K simulations and K clusterings inside the timed region, pipeline drained at both
ends; `sim` / `clustering` / `pipeline.ms_per_step_back_to_back` come from a second pass without the overlap.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3|c1]     # the CUDA path
  python bench.py --impl reference ...                                             # the CPU path (oracle) on the host cores

Prints ONE JSON line (DESIGN.md section 7 explains every key).
"""
import argparse
import ctypes as C
import json
import os
import struct
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"
SEED = 20261020


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "c3", "c5"])
    ap.add_argument("--edges", type=int, default=0, help="override the number of edges of the workload")
    ap.add_argument("--particles", type=int, default=0, help="override the particles per edge")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per particle (0 = built-in choice)")
    ap.add_argument("--cpu-edges", type=int, default=0, help="edges in the CPU sample (0 = sized for ~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads (C2, C3) of the `also` key")
    ap.add_argument("--no-plugin", action="store_true", help="skip the C++ plugin driver (e2e_plugin)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- workloads

def make_workload(name, edges=0, particles=0):
    """dict(env, robot, params, edge_starts [E, C], edge_targets [E, C], particles_per_edge, label): every workload is an edge
    batch; C1 / C2 / C3 are batches of one edge."""
    from uncertainty_planning_core_b200 import scenarios
    if name == "c5":
        E, P = edges or 4096, particles or 1000
        sc = scenarios.c5_batch(edges=E, particles=1)
        label = "C5 planner batch: %d edges x %d particles, SE(2), 256x256 SDF, P=64 body points, S=64" % (E, P)
        es, et = sc["starts"], sc["targets"]
    else:
        if name == "c1":
            sc = scenarios.c1_se2(n=1)
            P, label = particles or 1000, "C1 SE(2), %d particles/edge, 256x256 SDF, P=64, S=64"
        elif name == "c3":
            sc = scenarios.c3_linked(n=1)
            P, label = particles or 100000, "C3 7-DoF arm, %d particles/edge, 256^3 SDF, P=128, S=64, contact resolution"
        else:
            sc = scenarios.c2_se3(n=1)
            P, label = particles or 10000, "C2 SE(3), %d particles/edge, 128^3 SDF, P=128, S=64, contact resolution"
        label = label % P
        es, et = sc["starts"][:1], sc["targets"][:1]
    return dict(name=name, env=sc["env"], robot=sc["robot"], params=sc["params"], edge_starts=np.ascontiguousarray(es),
                edge_targets=np.ascontiguousarray(et), particles_per_edge=P, label=label)


def cluster_params():
    from uncertainty_planning_core_b200 import abi
    cp = abi.ClusterParams()
    cp.feature_type = abi.FEATURE_CRS
    cp.signature_matching_threshold = 0.125
    cp.distance_clustering_threshold = 0.1
    return cp


def sdf_bytes_per_point(wl):
    """one trilinear cell fetch = 8 corners x 4 B (4 corners for the 2-D grid)."""
    return 16 if wl["env"].sdf.shape[2] == 1 else 32


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


def host_cores():
    """(logical CPUs this process may use, physical cores among them)"""
    try:
        logical = len(os.sched_getaffinity(0))
    except AttributeError:
        logical = os.cpu_count() or 1
    physical = None
    try:
        import psutil
        physical = psutil.cpu_count(logical=False)
    except Exception:
        pass
    return logical, physical


# ---------------------------------------------------------------------------------------------- CPU arm

def cpu_pass(wl, cp, n_edges):
    """One pass of the CPU path (oracle, all host threads) over the first n_edges edges of the workload: the oracle's own
    counter-based tape for those particles, ForwardSimulateRobots, then ClusterParticles per edge.
    Returns (seconds sim, seconds clustering, seconds tape, particle-steps executed)."""
    from oracle import binding as oracle
    P = wl["particles_per_edge"]
    n = n_edges * P
    starts = np.repeat(wl["edge_starts"][:n_edges], P, axis=0)
    targets = np.repeat(wl["edge_targets"][:n_edges], P, axis=0)
    t0 = time.perf_counter()
    tape = oracle.fill_noise_tape_counter(SEED, wl["robot"], wl["params"], n)
    t1 = time.perf_counter()
    cfg, coll, stats = oracle.forward_simulate(wl["env"], wl["robot"], wl["params"], starts, targets, tape)
    t2 = time.perf_counter()
    oracle.cluster_particles(wl["env"], wl["robot"], cp, cfg, n_edges)
    t3 = time.perf_counter()
    return t2 - t1, t3 - t2, t1 - t0, stats["particle_steps"]


def size_cpu_sample(wl, cp, want_seconds):
    """Edges in the CPU sample: probe one edge (or a slice of a big single edge), scale linearly (edges are independent)."""
    E = wl["edge_starts"].shape[0]
    if E == 1:
        return 1
    threads = host_cores()[0]
    probe = min(E, 2 * threads)      # sets are taken whole per thread from 2 x threads sets on; mixes contact and free-space edges
    a, b, c, _ = cpu_pass(wl, cp, probe)
    per_edge = max((a + b + c) / probe, 1e-6)
    return int(max(probe, min(E, want_seconds / per_edge)))


def cpu_sample_workload(wl, cp, want_seconds):
    """A single big edge (C2 / C3) is sampled by particles instead of by edges: clustering grows faster than linearly with the
    set size, so the sample is bounded at 4,096 particles."""
    if wl["edge_starts"].shape[0] > 1:
        return wl, size_cpu_sample(wl, cp, want_seconds)
    sub = dict(wl)
    probe = dict(wl)
    probe["particles_per_edge"] = min(128, wl["particles_per_edge"])
    a, b, c, _ = cpu_pass(probe, cp, 1)
    per_particle = max((a + c) / probe["particles_per_edge"], 1e-9)
    sub["particles_per_edge"] = int(max(128, min(wl["particles_per_edge"], 4096, want_seconds * 0.5 / per_particle)))
    return sub, 1


def cpu_measure(wl, cp, want_seconds, passes):
    """`passes` passes over a bounded sample; returns the cpu_baseline dict (value = median pass)."""
    from oracle import binding as oracle
    logical, physical = host_cores()
    oracle.set_num_threads(logical)
    sample_wl, n_edges = cpu_sample_workload(wl, cp, want_seconds / max(passes, 1))
    P, S = sample_wl["particles_per_edge"], wl["params"].num_steps
    rows = [cpu_pass(sample_wl, cp, n_edges) for _ in range(passes)]
    total = np.array([r[0] + r[1] for r in rows])
    sim = np.array([r[0] for r in rows])
    cl = np.array([r[1] for r in rows])
    units = n_edges * P * S
    if wl["edge_starts"].shape[0] > 1:
        desc = "first %d of %d edges x %d particles, %d controller intervals, sim + CRS/distance clustering per edge" % (n_edges, wl["edge_starts"].shape[0], P, S)
    else:
        desc = "first %d of %d particles of the edge, %d controller intervals, sim + CRS/distance clustering" % (P, wl["particles_per_edge"], S)
    return {"value": units / float(np.median(total)), "unit": UNIT, "cores": physical or logical, "logical_cpus": logical, "threads": oracle.num_threads(),
            "kind": "port", "sample": desc + ", %d passes (value = median pass)" % passes,
            "value_best_pass": units / float(total.min()), "value_worst_pass": units / float(total.max()),
            "sim_value": units / float(np.median(sim)), "sim_value_best_pass": units / float(sim.min()),
            "clustered_per_sec": n_edges * P / max(float(np.median(cl)), 1e-12),
            "noise_generation_s_per_pass": float(np.median([r[2] for r in rows])),
            "note": "noise tape generation (oracle, counter-based) is outside the timed part of a pass: the CPU path replays a tape"}, float(total.sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as oracle
    oracle.build()
    wl = make_workload(args.workload, args.edges, args.particles)
    cp = cluster_params()
    passes = max(1, args.steps)
    logical, physical = host_cores()
    oracle.set_num_threads(logical)
    sample_wl, n_edges = cpu_sample_workload(wl, cp, 2.0) if not args.cpu_edges else (wl, args.cpu_edges)
    for _ in range(min(args.warmup, 1)):
        cpu_pass(sample_wl, cp, n_edges)
    rows = [cpu_pass(sample_wl, cp, n_edges) for _ in range(passes)]
    P, S = sample_wl["particles_per_edge"], wl["params"].num_steps
    units = n_edges * P * S
    elapsed = sum(r[0] + r[1] for r in rows)
    value = units * passes / elapsed
    if wl["edge_starts"].shape[0] > 1:
        desc = "first %d of %d edges x %d particles, %d controller intervals, sim + CRS/distance clustering per edge, per step" % (n_edges, wl["edge_starts"].shape[0], P, S)
    else:
        desc = "first %d of %d particles of the edge, %d controller intervals, sim + CRS/distance clustering, per step" % (P, wl["particles_per_edge"], S)
    per = np.array([r[0] + r[1] for r in rows])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": passes,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / passes, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (seeded boxes SDF, seeded edges, counter-based truncated-normal noise)",
        "config": {"workload": wl["label"], "sample": desc, "threads": oracle.num_threads()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": physical or logical, "logical_cpus": logical, "kind": "port", "sample": desc,
                         "value_best_pass": units / float(per.min()), "value_median_pass": units / float(np.median(per)),
                         "sim_value": units * passes / sum(r[0] for r in rows), "clustered_per_sec": n_edges * P * passes / max(sum(r[1] for r in rows), 1e-12)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm

def write_workload_file(wl, cp, path):
    """the raw POD descriptors + arrays for scripts/plugin_bench.cpp"""
    env, robot = wl["env"], wl["robot"]
    with open(path, "wb") as f:
        f.write(b"UPCWL01\0")
        f.write(bytes(env.desc))
        f.write(bytes(robot.desc))
        f.write(bytes(wl["params"]))
        f.write(struct.pack("<6q", 1 if env.convex_segment is not None else 0, 1 if env.occupancy is not None else 0, wl["edge_starts"].shape[0],
                            wl["edge_targets"].shape[0], wl["particles_per_edge"], robot.width))
        f.write(struct.pack("<2d", cp.signature_matching_threshold, cp.distance_clustering_threshold))
        f.write(env.sdf.tobytes())
        if env.convex_segment is not None:
            f.write(env.convex_segment.tobytes())
        if env.occupancy is not None:
            f.write(env.occupancy.tobytes())
        f.write(robot.points.tobytes())
        f.write(np.ascontiguousarray(wl["edge_starts"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(wl["edge_targets"], dtype=np.float64).tobytes())


def run_plugin_driver(wl, cp, mode, steps, warmup, gpus=1):
    exe = os.path.join(ROOT, "scripts", "plugin_bench")
    if not os.path.exists(exe):
        return {"unavailable": "scripts/plugin_bench is not built (run __graft_entry__.build())"}
    path = tempfile.mktemp(suffix=".upcwl")
    try:
        write_workload_file(wl, cp, path)
        env = dict(os.environ)
        env["OMP_NUM_THREADS"] = str(host_cores()[0])
        out = subprocess.run([exe, path, mode, str(steps), str(warmup), str(gpus)], capture_output=True, text=True, timeout=900, env=env)
        if out.returncode != 0:
            return {"unavailable": "plugin_bench exited %d: %s" % (out.returncode, (out.stderr or out.stdout).strip()[-300:])}
        return json.loads(out.stdout.strip().splitlines()[-1])
    finally:
        if os.path.exists(path):
            os.unlink(path)


class DeviceBatch:
    """Device-resident state of one rank for a workload: handle, (optional) communicator, replicated edge arrays, two outcome
    buffers, two streams.  step(k) issues one planner-batch step asynchronously.  Pipelined (default): the simulation of batch k
    runs on stream A while the clustering (and, with several GPUs, the exchange) of batch k - 1 runs on stream B - the
    simulation fills the machine, the clustering is a handful of small latency-bound launches with host decisions in between, so
    the two overlap almost for free; finish() drains the pipeline.  step(k, ev) with timing events runs the two halves back to back
    on stream A instead (the per-phase breakdown)."""

    def __init__(self, wl, cp, world, rank, local_rank, nccl_id, lanes=0, pipelined=True):
        import torch
        from uncertainty_planning_core_b200 import B200Simulator, abi
        self.torch, self.abi, self.lib = torch, abi, abi.load_library()
        self.wl, self.cp, self.world, self.rank = wl, cp, world, rank
        self.pipelined = pipelined
        self.dev = torch.device("cuda", local_rank)
        self.sim = B200Simulator(wl["env"], wl["robot"], device=local_rank)
        if lanes:
            self.sim.SetLanesPerParticle(lanes)
        self.h = self.sim.handle
        self.E, self.P = wl["edge_starts"].shape[0], wl["particles_per_edge"]
        self.width = wl["robot"].width
        self.lo, self.hi = abi.shard_bounds(self.E, rank, world)
        cap = (self.E + world - 1) // world
        self.lay = abi.outcome_layout(self.width, cap * self.P, cap)
        self.comm = None
        if world > 1:
            comm = C.c_void_p()
            abi.check(self.lib.upc_comm_create(self.h, world, rank, nccl_id.ctypes.data, C.byref(comm)))
            self.comm = comm
        self.d_starts = torch.from_numpy(np.ascontiguousarray(wl["edge_starts"].T)).to(self.dev)
        self.d_targets = torch.from_numpy(np.ascontiguousarray(wl["edge_targets"].T)).to(self.dev)
        # outcome buffers: two on one GPU (simulate k + 1 while k is clustered); three with several GPUs, so that the exchange of batch
        # k - 1 - a collective: it ends when the slowest rank has clustered - has a whole further step before its buffer is needed again
        self.nbuf = 3 if world > 1 else 2
        self.blocks = [torch.zeros(self.lay.block_bytes * world, dtype=torch.uint8, device=self.dev) for _ in range(self.nbuf)]
        self.stream_obj = torch.cuda.Stream(device=self.dev)        # A: simulation
        # B: clustering + exchange of the batch before.  High priority: the simulation kernel holds every register of every SM, and the
        # clustering's small CTAs must get the slots that free up as simulation CTAs retire instead of queueing behind the whole launch
        self.stream_b_obj = torch.cuda.Stream(device=self.dev, priority=-1)
        self.stream, self.stream_b = self.stream_obj.cuda_stream, self.stream_b_obj.cuda_stream
        assert self.stream != 0 and self.stream_b != 0
        self.ev_sim = [torch.cuda.Event() for _ in range(self.nbuf)]    # simulation of the batch in buffer i has finished
        self.ev_done = [torch.cuda.Event() for _ in range(self.nbuf)]   # clustering of the batch in buffer i has finished (and, with a
                                                                         # communicator, the exchange of the batch before it: stream B waits for it)
        self.pending = None                                          # batch simulated but not yet clustered
        self.issued = 0

    def local_particles(self):
        return (self.hi - self.lo) * self.P

    def _simulate(self, k, stream_obj):
        abi, lib, lay = self.abi, self.lib, self.lay
        blocks = self.blocks[k % self.nbuf]
        if self.comm is not None:
            abi.check(lib.upc_sharded_edge_batch_simulate(self.comm, C.byref(self.wl["params"]), SEED + k, self.E, self.P, self.d_starts.data_ptr(),
                                                          self.d_targets.data_ptr(), blocks.data_ptr(), stream_obj.cuda_stream))
            return
        base = blocks.data_ptr()
        abi.check(lib.upc_forward_simulate_seeded_device(self.h, C.byref(self.wl["params"]), SEED + k, 0, self.E * self.P, self.E, self.d_starts.data_ptr(), self.E,
                                                         self.d_targets.data_ptr(), base + lay.configs_offset, base + lay.collided_offset, stream_obj.cuda_stream))

    def _cluster(self, k, stream_obj):
        abi, lib, lay = self.abi, self.lib, self.lay
        blocks = self.blocks[k % self.nbuf]
        if self.comm is not None:
            abi.check(lib.upc_sharded_edge_batch_finish(self.comm, C.byref(self.cp), self.E, self.P, blocks.data_ptr(), stream_obj.cuda_stream))
            return
        base = blocks.data_ptr()
        abi.check(lib.upc_cluster_particles_device(self.h, C.byref(self.cp), self.E, self.P, base + lay.configs_offset, base + lay.labels_offset,
                                                   base + lay.n_clusters_offset, stream_obj.cuda_stream))

    def step(self, k, ev=None):
        A, B = self.stream_obj, self.stream_b_obj
        if ev or not self.pipelined:
            # the two halves back to back on stream A (events around each: the per-phase breakdown)
            self._drain()
            if ev:
                ev[0].record(A)
            self._simulate(k, A)
            if ev:
                ev[1].record(A)
            self._cluster(k, A)
            if ev:
                ev[2].record(A)
            return
        # the buffer was last used by batch k - nbuf: its clustering must be over before it is overwritten, and so must its exchange -
        # which stream B waited for before it recorded the `done` event of the batch AFTER it (upc_comm: an exchange is ordered behind
        # the one before), i.e. of batch k - 2 when there are three buffers
        nb = self.nbuf
        if self.issued >= nb:
            A.wait_event(self.ev_done[k % nb])
            if self.comm is not None:
                if nb >= 3:
                    A.wait_event(self.ev_done[(k - nb + 1) % nb])
                else:
                    self.abi.check(self.lib.upc_comm_wait(self.comm, A.cuda_stream))
        self._simulate(k, A)
        self.ev_sim[k % nb].record(A)
        self.issued += 1
        if self.pending is not None:
            self._finish_pending()
        self.pending = k

    def _finish_pending(self):
        k, B = self.pending, self.stream_b_obj
        B.wait_event(self.ev_sim[k % self.nbuf])
        self._cluster(k, B)          # host decisions inside only ever wait for stream B
        self.ev_done[k % self.nbuf].record(B)
        self.pending = None

    def _drain(self):
        if self.pending is not None:
            self._finish_pending()
        if self.issued:
            for ev in self.ev_done:
                self.stream_obj.wait_event(ev)
            self.issued = 0

    def finish(self):
        """everything issued so far (clustering and exchanges included) is complete when stream A reaches this point"""
        self._drain()
        if self.comm is not None:
            self.abi.check(self.lib.upc_comm_wait(self.comm, self.stream))

    def cluster_counts(self, k=0):
        lay = self.lay
        blk = self.blocks[k % self.nbuf].cpu().numpy()
        out = []
        for r in range(self.world):
            lo, hi = self.abi.shard_bounds(self.E, r, self.world)
            off = r * lay.block_bytes + lay.n_clusters_offset
            out.append(blk[off:off + 4 * (hi - lo)].view(np.int32))
        return np.concatenate(out)

    def close(self):
        if self.comm is not None:
            self.lib.upc_comm_destroy(self.comm)
        self.sim.close()


def timed_device_steps(batch, steps, warmup, barrier, sampler=None, phase_events=False):
    """W warm-up steps, then exactly K steps between barriers; returns (total ms, sim ms, clustering ms, stats, clocks)"""
    torch = batch.torch
    for k in range(warmup):
        batch.step(k)
    batch.finish()
    barrier()
    batch.sim.ResetStatistics()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)] if phase_events else None
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_begin()
    t_begin.record(batch.stream_obj)
    for k in range(steps):
        batch.step(warmup + k, evs[k] if evs else None)
    batch.finish()
    t_end.record(batch.stream_obj)
    barrier()
    if sampler:
        sampler.mark_end()
    total_ms = t_begin.elapsed_time(t_end)
    sim_ms = sum(e[0].elapsed_time(e[1]) for e in evs) if evs else None
    cl_ms = sum(e[1].elapsed_time(e[2]) for e in evs) if evs else None
    return total_ms, sim_ms, cl_ms, batch.sim.GetStatistics()


def secondary_workload(name, steps, warmup, local_rank):
    """the `also` key: another BASELINE configuration as a device-resident step (seeded noise, sim + clustering) on one GPU"""
    import torch
    wl = make_workload(name)
    cp = cluster_params()
    # one edge per step: the planner needs this edge's clusters before it can extend the next one, so these steps are NOT pipelined -
    # simulation and clustering back to back on one stream, events around each
    batch = DeviceBatch(wl, cp, 1, 0, local_rank, None, pipelined=False)
    try:
        def barrier():
            torch.cuda.synchronize()
        total_ms, sim_ms, cl_ms, stats = timed_device_steps(batch, steps, warmup, barrier, phase_events=True)
        n, S = wl["particles_per_edge"], wl["params"].num_steps
        P, bpp = wl["robot"].num_points, sdf_bytes_per_point(wl)
        alg = (stats["particle_steps"] + stats["resolver_iterations"]) / steps * P * bpp
        return {"workload": wl["label"], "value": n * S * steps / (total_ms * 1e-3), "unit": UNIT, "ms_per_step": total_ms / steps,
                "sim_ms": sim_ms / steps, "clustering_ms": cl_ms / steps, "clusters": int(batch.cluster_counts()[0]),
                "sim_value": n * S * steps / (sim_ms * 1e-3), "clustered_per_sec": n * steps / (cl_ms * 1e-3),
                "algorithmic_sdf_gather_gbs": alg / (sim_ms / steps * 1e-3) / 1e9, "lanes_per_particle": "auto",
                "collided_fraction": stats["collided_particles"] / max(stats["particles_simulated"], 1),
                "resolver_iterations_per_particle_step": stats["resolver_iterations"] / max(stats["particle_steps"], 1)}
    finally:
        batch.close()


def run_b200(args):
    import torch
    import torch.distributed as dist
    from uncertainty_planning_core_b200 import abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = abi.load_library()
    nccl_id = np.zeros(128, dtype=np.uint8)
    nccl_version = None
    if world > 1:
        # plumbing only: the NCCL communicator of the data path is created by the library from this id
        dist.init_process_group("gloo")
        if rank == 0:
            abi.check(lib.upc_comm_get_unique_id(nccl_id.ctypes.data))
        t = torch.from_numpy(nccl_id)
        dist.broadcast(t, 0)
        v = C.c_int32(0)
        abi.check(lib.upc_comm_nccl_version(C.byref(v)))
        nccl_version = v.value

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        t = torch.tensor(values, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    wl = make_workload(args.workload, args.edges, args.particles)
    cp = cluster_params()
    params, robot, env = wl["params"], wl["robot"], wl["env"]
    E, P, S = wl["edge_starts"].shape[0], wl["particles_per_edge"], params.num_steps
    n_total = E * P
    width = robot.width
    batch = DeviceBatch(wl, cp, world, rank, local_rank, nccl_id, args.lanes)

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps between barriers, max over ranks
    sampler = ClockSampler(local_rank) if rank == 0 else None
    total_ms, _, _, stats = timed_device_steps(batch, args.steps, args.warmup, barrier, sampler)
    clocks = sampler.stop() if sampler else None
    ncl_all = batch.cluster_counts(args.warmup + args.steps - 1)   # outcome of the last timed batch (the passes below reuse its buffer)
    b2b_ms = None
    if world == 1:
        # per-phase times: a second, shorter pass with the two halves back to back on one stream and events around each
        kb = min(args.steps, 5)
        b2b_ms, sim_ms, cl_ms, _ = timed_device_steps(batch, kb, 1, barrier, phase_events=True)
        b2b_ms, sim_ms, cl_ms = b2b_ms * args.steps / kb, sim_ms * args.steps / kb, cl_ms * args.steps / kb
    if world > 1:
        # per-phase times of this rank's shard, measured in a second, un-timed pass without the exchange
        solo = DeviceBatch(dict(wl, edge_starts=wl["edge_starts"][batch.lo:batch.hi], edge_targets=wl["edge_targets"][batch.lo:batch.hi]), cp, 1, 0, local_rank, None, args.lanes)
        _, sim_ms, cl_ms, _ = timed_device_steps(solo, min(args.steps, 3), 1, lambda: torch.cuda.synchronize(), phase_events=True)
        sim_ms, cl_ms = sim_ms * args.steps / min(args.steps, 3), cl_ms * args.steps / min(args.steps, 3)
        solo.close()
    total_ms, sim_ms, cl_ms = max_over_ranks([total_ms, sim_ms, cl_ms])
    particle_steps = n_total * S * args.steps
    value = particle_steps / (total_ms * 1e-3)

    # ---- end to end through the host-pointer C ABI: pinned host buffers in, every result back on the host, every step
    lo, hi = batch.lo, batch.hi
    n_local = (hi - lo) * P
    h_es = torch.from_numpy(np.ascontiguousarray(wl["edge_starts"][lo:hi].T)).pin_memory()
    h_et = torch.from_numpy(np.ascontiguousarray(wl["edge_targets"][lo:hi].T)).pin_memory()
    h_out = torch.empty((width, n_local), dtype=torch.float64).pin_memory()
    h_coll = torch.empty(n_local, dtype=torch.uint8).pin_memory()
    h_labels = torch.empty(n_local, dtype=torch.int32).pin_memory()
    h_ncl = torch.empty(hi - lo, dtype=torch.int32).pin_memory()

    def e2e_step(k):
        abi.check(lib.upc_simulate_and_cluster_edges(batch.h, C.byref(params), C.byref(cp), SEED + k, lo * P, hi - lo, P, hi - lo, h_es.data_ptr(), h_et.data_ptr(),
                                                     h_out.data_ptr(), h_coll.data_ptr(), h_labels.data_ptr(), h_ncl.data_ptr()))
        return int(h_ncl[0])

    for k in range(min(args.warmup, 3)):
        e2e_step(k)
    barrier()
    # the one-call form, a few steps: every step waits for its own results before the next begins
    k1 = min(args.steps, 3)
    t0 = time.perf_counter()
    for k in range(k1):
        e2e_step(args.warmup + k)
    torch.cuda.synchronize()
    e2e_one_call_s = (time.perf_counter() - t0) * args.steps / k1
    # the submit / collect form, two batches in flight: batch k + 1 is uploaded and simulated while batch k is clustered and downloaded.
    # Two sets of pinned result buffers: a collected batch stays readable while the next one lands.
    h_res = [(h_out, h_coll, h_labels, h_ncl),
             (torch.empty_like(h_out).pin_memory(), torch.empty_like(h_coll).pin_memory(), torch.empty_like(h_labels).pin_memory(), torch.empty_like(h_ncl).pin_memory())]

    def submit(k):
        abi.check(lib.upc_edge_batch_submit(batch.h, C.byref(params), SEED + k, lo * P, hi - lo, P, hi - lo, h_es.data_ptr(), h_et.data_ptr(), k % 2))

    def collect(k):
        o, c, l, m = h_res[k % 2]
        abi.check(lib.upc_edge_batch_collect(batch.h, C.byref(cp), k % 2, o.data_ptr(), c.data_ptr(), l.data_ptr(), m.data_ptr()))
        return int(m[0])

    submit(0); submit(1); collect(0); collect(1)       # warm-up of the second set of staging buffers
    barrier()
    batch.sim.ResetStatistics()
    t0 = time.perf_counter()
    first = args.warmup
    submit(first)
    for k in range(args.steps):
        if k + 1 < args.steps:
            submit(first + k + 1)
        collect(first + k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h_ncl = h_res[(first + args.steps - 1) % 2][3]
    e2e_stats = batch.sim.GetStatistics()
    e2e_s, e2e_one_call_s, h2d, d2h = max_over_ranks([e2e_s, e2e_one_call_s, e2e_stats["h2d_bytes"] / args.steps, e2e_stats["d2h_bytes"] / args.steps])
    e2e_value = particle_steps / e2e_s
    e2e_same = bool(np.array_equal(h_ncl.numpy(), ncl_all[lo:hi]))   # same seeds as the device-resident pass: same outcome

    if rank == 0:
        # ---- roofline of the dominant kernel (forward_simulate_kernel): algorithmic SDF gather bytes / measured launch time
        hbm_peak, peak_src = measured_peaks()
        Pn = robot.num_points
        bpp = sdf_bytes_per_point(wl)
        ps_launch = stats["particle_steps"] / args.steps
        ri_launch = stats["resolver_iterations"] / args.steps
        alg_bytes = (ps_launch + ri_launch) * Pn * bpp            # M*P*corners*4 per particle-step + R_used*P*corners*4 per resolver iteration
        sim_ms_launch = sim_ms / args.steps
        achieved = alg_bytes / (sim_ms_launch * 1e-3) / 1e9
        table_bytes = 8 * 4 * (env.sdf.shape[0] + 1) * (env.sdf.shape[1] + 1) * (env.sdf.shape[2] + 1)
        gather_buffer = table_bytes if table_bytes <= 80 * 2 ** 20 else env.sdf.nbytes
        l2_gbs = C.c_double(0.0)
        fp64 = C.c_double(0.0)
        abi.check(lib.upc_measure_l2_gather_gbs(batch.h, gather_buffer, C.byref(l2_gbs)))
        abi.check(lib.upc_measure_fp64_tflops(batch.h, C.byref(fp64)))
        # compulsory HBM traffic of a seeded launch: edge endpoints in, one configuration + one flag per particle out
        compulsory = (2 * width * (hi - lo) * 8 + (width * 8 + 1) * n_local) / (sim_ms_launch * 1e-3) / 1e9
        model = {abi.ROBOT_SE2: "Se2Model", abi.ROBOT_SE3: "Se3Model"}.get(robot.kind, "LinkedModelSmall")
        traffic, traffic_source = None, None
        tpath = os.path.join(ROOT, "profiles", "r02_sim_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath)).get(args.workload)
                if tj and not args.edges and not args.particles:
                    traffic, traffic_source = tj["traffic"], "committed ncu capture (%s), not measured in this run" % tj.get("source", tpath)
            except Exception:
                pass
        roofline = {"bound": "l2_gather", "achieved": achieved, "peak": l2_gbs.value, "unit": "GB/s", "frac": achieved / max(l2_gbs.value, 1e-9),
                    "traffic": traffic, "traffic_source": traffic_source,
                    "kernel": "forward_simulate_kernel<%s, lanes=%s, seeded>" % (model, args.lanes or "auto"), "kernel_ms": sim_ms_launch,
                    "peak_source": "random 32-byte-sector gather over a %.1f MiB buffer (the size of the SDF table the kernel reads), measured live on this GPU (upc_measure_l2_gather_gbs)" % (gather_buffer / 2 ** 20),
                    "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_per_particle_step": Pn * bpp,
                    "hbm_peak_gbs": hbm_peak, "hbm_peak_source": peak_src, "frac_of_hbm_peak": achieved / hbm_peak,
                    "compulsory_hbm_gbs": compulsory, "fp64_fma_peak_tflops": fp64.value,
                    "note": "algorithmic gathers = SURVEY 8d nominal bytes; conservative culls skip most of them, so `achieved` is a work rate (see DESIGN.md 7)",
                    "limiter": "not bandwidth: ncu of this kernel (profiles/r02l_sim_c5_metrics.txt, r02g_* before) - the 140 KB per-particle program is bound by "
                               "instruction fetch (SM instruction cache 78 % -> 92 % hits with the per-interval barrier) and, after that, by fixed-latency dependencies at "
                               "16 warps per SM; L1 serves 99 % of the SDF sectors, DRAM carries a few MB per launch"}
        # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = cpu_measure(wl, cp, want_seconds=20.0, passes=5)
        # ---- secondary workloads and the C++ plugin driver (rank 0, N = 1 only)
        also, e2e_plugin = None, None
        if world == 1 and not args.no_also and args.workload == "c5" and not args.edges and not args.particles:
            also = {}
            for name, st, wu in (("c2", 10, 3), ("c3", 3, 1)):
                try:
                    also[name] = secondary_workload(name, st, wu, local_rank)
                except Exception as exc:  # a secondary measurement must not cost the headline line
                    also[name] = {"unavailable": repr(exc)[:200]}
        if world == 1 and not args.no_plugin:
            e2e_plugin = run_plugin_driver(wl, cp, "edges" if E > 1 else "single", max(2, min(args.steps, 5)), 1)
            if isinstance(e2e_plugin, dict) and "ms_per_step" in e2e_plugin:
                e2e_plugin["unit"] = UNIT
                e2e_plugin["ratio_to_device_step"] = e2e_plugin["ms_per_step"] / (total_ms / args.steps)
            if also is not None and "c2" in also and "value" in also["c2"]:
                c2 = make_workload("c2")
                p2 = run_plugin_driver(c2, cp, "single", 10, 3)
                if isinstance(p2, dict) and "ms_per_step" in p2:
                    p2["ratio_to_device_step"] = p2["ms_per_step"] / also["c2"]["ms_per_step"]
                also["c2"]["e2e_plugin"] = p2
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (seeded boxes SDF, seeded edges, counter-based truncated-normal noise drawn in the kernel)",
            "config": {"workload": wl["label"], "edges": E, "particles_per_edge": P, "edges_per_gpu": hi - lo,
                       "sharding": "contiguous blocks of edges per GPU (upc_shard_bounds), SDF + robot tables replicated, ONE in-place NCCL all-gather of outcome records per step "
                                   "issued by the library's communicator (three gather buffers: clustering + exchange of step k overlap the simulation of steps k+1 and k+2)" if world > 1 else "single GPU",
                       "nccl_version": nccl_version,
                       "clustering": "CRS signatures (thr 0.125) + distance pass (thr 0.1), one pass over all edges",
                       "lanes_per_particle": args.lanes or "auto",
                       "l2": "inputs larger than L2 where they stream: every step writes %.0f MB of outcome records and re-reads them for clustering (126 MB L2); the %.1f MiB SDF table is "
                             "deliberately cache-resident across steps, as it is across the batches of a planner run; every step uses a new noise seed" % (batch.lay.block_bytes * world / 1e6, gather_buffer / 2 ** 20)},
            "pipeline": {"overlap": "the simulation of batch k (stream A) overlaps the clustering%s of batch k - 1 (stream B); K simulations and K clusterings inside the timed region, "
                                    "the pipeline drained at both ends" % (" and the exchange" if world > 1 else ""),
                         "ms_per_step_back_to_back": (b2b_ms / args.steps) if b2b_ms else None},
            "sim": {"value": particle_steps / (sim_ms * 1e-3), "unit": UNIT, "ms_per_step": sim_ms / args.steps,
                    "scope": "whole batch on this GPU, measured back to back (not overlapped)" if world == 1 else "slowest rank's shard, measured back to back without the exchange"},
            "clustering": {"value": n_total * args.steps / (cl_ms * 1e-3), "unit": "particles clustered/s", "ms_per_step": cl_ms / args.steps,
                           "sets_per_sec": E * args.steps / (cl_ms * 1e-3), "clusters_total": int(ncl_all.sum()), "clusters_max_per_edge": int(ncl_all.max()),
                           "fallback_sets": int(stats["cluster_fallback_calls"])},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "call": "upc_edge_batch_submit / upc_edge_batch_collect (host pointers, pinned), two batches in flight: every step the edge endpoints go up and "
                            "configurations + flags + labels + cluster counts come down; batch k + 1 is uploaded and simulated while batch k is clustered and downloaded",
                    "ms_per_step_one_call": 1e3 * e2e_one_call_s / args.steps,
                    "one_call": "upc_simulate_and_cluster_edges: the same step as one blocking call per batch (measured over %d steps)" % min(args.steps, 3),
                    "same_outcome_as_device_pass": e2e_same},
            "e2e_plugin": e2e_plugin,
            "gpu_launches": int(stats["kernel_launches"]),
            "contact": {"collided_fraction": stats["collided_particles"] / max(stats["particles_simulated"], 1),
                        "resolver_iterations_per_particle_step": stats["resolver_iterations"] / max(stats["particle_steps"], 1)},
            "roofline": roofline, "cpu_baseline": cpu, "also": also, "clocks": clocks,
        }
        print(json.dumps(line))
    barrier()
    batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is OpenMP code and is meant to use every host core
    _world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    _cores = host_cores()[0]
    os.environ["OMP_NUM_THREADS"] = str(_cores if a.impl == "reference" else max(1, _cores // max(1, _world)))
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
