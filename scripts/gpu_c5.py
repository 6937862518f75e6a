"""BASELINE config 5: a planner batch of E candidate edges x 1,000 particles in ONE simulation call (targets.size() == starts.size())
and ONE clustering call (n_sets = E), device-resident, CUDA events.  A few edges are re-run through the oracle for parity."""
import ctypes as C, sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from uncertainty_planning_core_b200 import B200Simulator, abi, scenarios
from oracle import binding as oracle

edges = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
P1 = 1000
t0 = time.time()
sc = scenarios.c5_batch(edges=edges, particles=P1)
env, robot, params = sc["env"], sc["robot"], sc["params"]
n = sc["starts"].shape[0]
S = params.num_steps
tape = abi.fill_noise_tape(sc["seed"], robot, params, n)
print("workload built in %.1f s: %d edges x %d particles, tape %.2f GB" % (time.time() - t0, edges, P1, tape.nbytes / 1e9), flush=True)
lib = abi.load_library()
dev = torch.device("cuda:0")
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
sim = B200Simulator(env, robot)
d_tape = torch.from_numpy(tape).to(dev)
d_s = torch.from_numpy(np.ascontiguousarray(sc["starts"].T)).to(dev)
d_t = torch.from_numpy(np.ascontiguousarray(sc["targets"].T)).to(dev)
d_o = torch.zeros_like(d_s); d_c = torch.zeros(n, dtype=torch.uint8, device=dev)
d_labels = torch.zeros(n, dtype=torch.int32, device=dev); d_ncl = torch.zeros(edges, dtype=torch.int32, device=dev)
cp = abi.ClusterParams(); cp.feature_type = abi.FEATURE_CRS; cp.signature_matching_threshold = 0.125; cp.distance_clustering_threshold = 0.1
best_sim = best_cl = 1e9
for rep in range(4):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    abi.check(lib.upc_forward_simulate_device(sim.handle, C.byref(params), n, d_s.data_ptr(), n, d_t.data_ptr(), d_tape.data_ptr(), d_o.data_ptr(), d_c.data_ptr(), side.cuda_stream))
    e1.record()
    abi.check(lib.upc_cluster_particles_device(sim.handle, C.byref(cp), edges, P1, d_o.data_ptr(), d_labels.data_ptr(), d_ncl.data_ptr(), side.cuda_stream))
    e2.record(); torch.cuda.synchronize()
    if rep:
        best_sim = min(best_sim, e0.elapsed_time(e1)); best_cl = min(best_cl, e1.elapsed_time(e2))
st = sim.GetStatistics()
cfg = d_o.cpu().numpy().T; coll = d_c.cpu().numpy().astype(bool); labels = d_labels.cpu().numpy(); ncl = d_ncl.cpu().numpy()
# parity on a few edges (oracle, same tape columns)
checked = 0
for e in (0, 1, edges // 2, edges - 1):
    sl = slice(e * P1, (e + 1) * P1)
    ocfg, ocoll, _ = oracle.forward_simulate(env, robot, params, sc["starts"][sl], sc["targets"][sl], np.ascontiguousarray(tape[..., sl]))
    assert np.array_equal(ocfg, cfg[sl]) and np.array_equal(ocoll, coll[sl]), "simulation parity, edge %d" % e
    ol, on, _ = oracle.cluster_particles(env, robot, cp, ocfg)
    assert np.array_equal(ol, labels[sl]) and on[0] == ncl[e], "clustering parity, edge %d" % e
    checked += 1
P = robot.desc.num_points
steps = n * S
gather_bytes = steps * P * 16.0   # planar grid: 4 corners x 4 B per point lookup, one check per particle-step
rec = dict(workload="C5: %d edges x %d particles, SE(2), 256x256 SDF, %d controller intervals, one simulation call + one clustering call" % (edges, P1, S),
           sim_ms=best_sim, particle_steps_per_s=steps / best_sim * 1e3, clustering_ms=best_cl, sets_clustered_per_s=edges / best_cl * 1e3,
           particles_clustered_per_s=n / best_cl * 1e3, step_ms=best_sim + best_cl, whole_step_particle_steps_per_s=steps / (best_sim + best_cl) * 1e3,
           collided_fraction=float(coll.mean()), clusters_mean=float(ncl.mean()), clusters_max=int(ncl.max()),
           resolver_iterations_per_particle_step=st["resolver_iterations"] / max(st["particle_steps"], 1),
           algorithmic_sdf_gather_gbs=gather_bytes / best_sim / 1e6, tape_gbs=tape.nbytes / best_sim / 1e6, oracle_checked_edges=checked)
print(json.dumps(rec), flush=True)
json.dump(rec, open("gpurun_out/c5.json", "w"), indent=1)
