"""BASELINE config 4 at its largest point (N = 65,536, 8 blobs, distance pass) for the three robots, adjacency pass forced
(`UPC_B200_NO_PIVOT_PASS=1`) and pivot pass.  `once` = one call per variant (for ncu), default = best of three after a warm-up.
Reports the pair tests per second of the adjacency pass next to the measured FP64 FMA peak (the bound SURVEY 8d names)."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cases
from uncertainty_planning_core_b200 import B200Simulator, abi, scenarios

once = len(sys.argv) > 1 and sys.argv[1] == "once"
lib = abi.load_library()
dev = torch.device("cuda:0")
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
N = 65536
for kind, case, name in ((abi.ROBOT_SE2, cases.cl_se2_blobs, "SE2"), (abi.ROBOT_SE3, cases.cl_se3_blobs, "SE3"), (abi.ROBOT_LINKED, cases.cl_linked_blobs, "7-DoF")):
    c = case()
    sim = B200Simulator(c["env"], c["robot"])
    if kind == abi.ROBOT_SE3:
        cfg0, owner0 = scenarios.c4_blobs(kind, 16384, k=8, seed=77)
        cfg = np.tile(cfg0, (N // 16384, 1)); owner = np.tile(owner0, N // 16384)
    else:
        cfg, owner = scenarios.c4_blobs(kind, N, k=8, seed=77)
    d_cfg = torch.from_numpy(np.ascontiguousarray(cfg.T)).to(dev)
    d_labels = torch.zeros(N, dtype=torch.int32, device=dev); d_ncl = torch.zeros(1, dtype=torch.int32, device=dev)
    for path in ("adjacency", "pivot"):
        if path == "adjacency":
            os.environ["UPC_B200_NO_PIVOT_PASS"] = "1"
        else:
            os.environ.pop("UPC_B200_NO_PIVOT_PASS", None)
        cp = cases._cparams(abi.FEATURE_NONE, 0.125, 0.1)
        best = 1e9
        for rep in range(1 if once else 4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            abi.check(lib.upc_cluster_particles_device(sim.handle, C.byref(cp), 1, N, d_cfg.data_ptr(), d_labels.data_ptr(), d_ncl.data_ptr(), side.cuda_stream))
            e1.record(); torch.cuda.synchronize()
            if rep or once: best = min(best, e0.elapsed_time(e1))
        lab = d_labels.cpu().numpy()
        # the partition must be the generating blobs: one label per owner and vice versa
        pairs = set(zip(lab.tolist(), owner.tolist()))
        ok = int(d_ncl.item()) == 8 and len(pairs) == 8
        print(json.dumps(dict(robot=name, n=N, path=path, ms=best, particles_per_s=N / best * 1e3,
                              pair_tests_per_s=N * (N + 1) / 2 / best * 1e3, clusters=int(d_ncl.item()), partition_is_the_blobs=bool(ok))), flush=True)
    if kind == abi.ROBOT_LINKED:
        fp64 = C.c_double(0)
        abi.check(lib.upc_measure_fp64_tflops(sim.handle, C.byref(fp64)))
        print(json.dumps(dict(fp64_fma_peak_tflops=fp64.value)))
    sim.close()
os.environ.pop("UPC_B200_NO_PIVOT_PASS", None)
