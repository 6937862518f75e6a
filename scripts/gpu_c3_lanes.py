"""C3 (7-DoF arm, 100,000 particles) simulation time per lanes-per-particle setting: with >= 4 lanes the link frames live in shared memory
instead of per-thread local memory."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uncertainty_planning_core_b200 import B200Simulator, abi, scenarios
lib = abi.load_library()
dev = torch.device("cuda", 0)
c3 = scenarios.c3_linked(n=1)
sim = B200Simulator(c3["env"], c3["robot"])
n = 100000
params = c3["params"]
side = torch.cuda.Stream(device=dev); torch.cuda.set_stream(side)
d_s = torch.from_numpy(np.ascontiguousarray(c3["starts"].T)).to(dev)
d_t = torch.from_numpy(np.ascontiguousarray(c3["targets"].T)).to(dev)
d_out = torch.zeros((c3["robot"].width, n), dtype=torch.float64, device=dev)
d_coll = torch.zeros(n, dtype=torch.uint8, device=dev)
ref = None
for lanes in (1, 2, 4, 8, 16):
    sim.SetLanesPerParticle(lanes)
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        abi.check(lib.upc_forward_simulate_seeded_device(sim.handle, C.byref(params), 7, 0, n, 1, d_s.data_ptr(), 1, d_t.data_ptr(), d_out.data_ptr(), d_coll.data_ptr(), side.cuda_stream))
        b.record(); torch.cuda.synchronize()
        if rep: best = min(best, a.elapsed_time(b))
    chk = float(d_out.sum().item())
    if ref is None: ref = chk
    print(json.dumps({"lanes": lanes, "ms": best, "same_result": chk == ref}), flush=True)
