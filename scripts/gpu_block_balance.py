"""Simulation time of each of the 8 contiguous 512-edge blocks of C5 (what 8 ranks would each run), generator order against edges dealt to
the blocks by the clearance of their end points."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uncertainty_planning_core_b200 import B200Simulator, abi
lib = abi.load_library()
dev = torch.device("cuda", 0)
side = torch.cuda.Stream(device=dev); torch.cuda.set_stream(side)
for deal in (False, True):
    wl = bench.make_workload("c5", deal=deal)
    sim = B200Simulator(wl["env"], wl["robot"])
    P, E = 1000, 512
    n = E * P
    d_out = torch.zeros((3, n), dtype=torch.float64, device=dev)
    d_coll = torch.zeros(n, dtype=torch.uint8, device=dev)
    times = []
    for b in range(8):
        d_s = torch.from_numpy(np.ascontiguousarray(wl["edge_starts"][b * E:(b + 1) * E].T)).to(dev)
        d_t = torch.from_numpy(np.ascontiguousarray(wl["edge_targets"][b * E:(b + 1) * E].T)).to(dev)
        best = 1e9
        for rep in range(3):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            abi.check(lib.upc_forward_simulate_seeded_device(sim.handle, C.byref(wl["params"]), 7, b * n, n, E, d_s.data_ptr(), E, d_t.data_ptr(), d_out.data_ptr(), d_coll.data_ptr(), side.cuda_stream))
            e.record(); torch.cuda.synchronize()
            if rep: best = min(best, a.elapsed_time(e))
        times.append(round(best, 3))
    print(json.dumps({"dealt": deal, "block_ms": times, "max": max(times), "mean": round(sum(times) / 8, 3)}), flush=True)
    sim.close()
