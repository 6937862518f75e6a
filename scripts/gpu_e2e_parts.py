"""Where the end-to-end time of the host-pointer path goes (C2): simulate and cluster calls timed separately on the host clock."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uncertainty_planning_core_b200 import B200Simulator, abi, scenarios
sc = scenarios.c2_se3()
env, robot, params = sc["env"], sc["robot"], sc["params"]
n = sc["starts"].shape[0]; width = robot.width
lib = abi.load_library()
sim = B200Simulator(env, robot); h = sim.handle
tapes = [torch.from_numpy(abi.fill_noise_tape(sc["seed"] + k, robot, params, n)).pin_memory() for k in range(4)]
h_starts = torch.from_numpy(np.ascontiguousarray(sc["starts"].T)).pin_memory()
h_targets = torch.from_numpy(np.ascontiguousarray(sc["targets"].T)).pin_memory()
h_out = torch.empty((width, n), dtype=torch.float64).pin_memory(); h_coll = torch.empty(n, dtype=torch.uint8).pin_memory()
h_labels = torch.empty(n, dtype=torch.int32).pin_memory(); h_ncl = torch.empty(1, dtype=torch.int32).pin_memory()
cp = abi.ClusterParams(); cp.feature_type = abi.FEATURE_CRS; cp.signature_matching_threshold = 0.125; cp.distance_clustering_threshold = 0.1
ts = tc = 0.0
K = 30
for k in range(K + 5):
    t0 = time.perf_counter()
    abi.check(lib.upc_forward_simulate(h, C.byref(params), n, h_starts.data_ptr(), 1, h_targets.data_ptr(), tapes[k % 4].data_ptr(), h_out.data_ptr(), h_coll.data_ptr()))
    t1 = time.perf_counter()
    abi.check(lib.upc_cluster_particles(h, C.byref(cp), 1, n, h_out.data_ptr(), h_labels.data_ptr(), h_ncl.data_ptr()))
    t2 = time.perf_counter()
    if k >= 5:
        ts += t1 - t0; tc += t2 - t1
print("upc_forward_simulate (host pointers): %.3f ms   upc_cluster_particles (host pointers): %.3f ms" % (1e3 * ts / K, 1e3 * tc / K))
