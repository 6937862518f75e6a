"""BASELINE config 4: outcome-clustering sweep N = 1k..64k (8 well-separated blobs -> unambiguous partition),
distance pass only (feature NONE) and, for SE(2), CRS + distance.  Device-resident inputs, CUDA events."""
import ctypes as C, sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cases
from uncertainty_planning_core_b200 import B200Simulator, abi, scenarios
from common import canonical_partition

lib = abi.load_library()
dev = torch.device("cuda:0")
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
out = []
for kind, case, name in ((abi.ROBOT_SE2, cases.cl_se2_blobs, "SE2"), (abi.ROBOT_SE3, cases.cl_se3_blobs, "SE3"), (abi.ROBOT_LINKED, cases.cl_linked_blobs, "7-DoF")):
    c = case()
    sim = B200Simulator(c["env"], c["robot"])
    for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
        if kind == abi.ROBOT_SE3 and n > 16384:
            cfg0, owner0 = scenarios.c4_blobs(kind, 16384, k=8, seed=77)   # python-side generation of SE3 blobs is slow: tile a 16k set
            reps = n // 16384
            cfg = np.tile(cfg0, (reps, 1)); owner = np.tile(owner0, reps)
        else:
            cfg, owner = scenarios.c4_blobs(kind, n, k=8, seed=77)
        d_cfg = torch.from_numpy(np.ascontiguousarray(cfg.T)).to(dev)
        d_labels = torch.zeros(n, dtype=torch.int32, device=dev); d_ncl = torch.zeros(1, dtype=torch.int32, device=dev)
        for feat in ((abi.FEATURE_NONE, abi.FEATURE_CRS) if kind == abi.ROBOT_SE2 else (abi.FEATURE_NONE,)):
          for path in ("pivot pass (default)", "adjacency pass (UPC_B200_NO_PIVOT_PASS=1)"):
            if path.startswith("adjacency"):
                os.environ["UPC_B200_NO_PIVOT_PASS"] = "1"
            else:
                os.environ.pop("UPC_B200_NO_PIVOT_PASS", None)
            cp = cases._cparams(feat, 0.125, 0.1)
            best = 1e9
            for rep in range(4):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                abi.check(lib.upc_cluster_particles_device(sim.handle, C.byref(cp), 1, n, d_cfg.data_ptr(), d_labels.data_ptr(), d_ncl.data_ptr(), side.cuda_stream))
                e1.record(); torch.cuda.synchronize()
                if rep: best = min(best, e0.elapsed_time(e1))
            ok = int(d_ncl.item()) == 8 and (n > 16384 or canonical_partition(d_labels.cpu().numpy()) == canonical_partition(owner))
            pairs = n * (n - 1) / 2
            rec = dict(robot=name, n=n, feature="CRS+distance" if feat else "distance", path=path, ms=best, particles_per_s=n / best * 1e3, pair_distances_per_s=pairs / best * 1e3, clusters=int(d_ncl.item()), partition_ok=bool(ok))
            out.append(rec); print(json.dumps(rec), flush=True)
    sim.close()
os.environ.pop("UPC_B200_NO_PIVOT_PASS", None)
json.dump(out, open("gpurun_out/cluster_sweep.json", "w"), indent=1)
