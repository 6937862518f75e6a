"""Simulation-kernel probe: device-resident kernel time for tape replay vs seeded noise on C5 (a quarter of the batch), C2 and C3,
and - when the library was built with -DUPC_PHASE_TIMING (UPC_B200_LIBRARY=...) - cycles per particle-step in each phase.
  python scripts/gpu_sim_probe.py [ab] [phases]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from uncertainty_planning_core_b200 import B200Simulator, abi, scenarios

lib = abi.load_library()
dev = torch.device("cuda", 0)
NAMES = ["control", "estimate", "apply", "check", "resolver", "reached", "noise/tape", "-", "res:points", "res:replay", "res:solve", "contributing points (count)"]


def workloads():
    c5 = scenarios.c5_batch(edges=1024, particles=1)
    yield "c5 (1024 edges x 1000)", c5["env"], c5["robot"], c5["params"], c5["starts"], c5["targets"], 1000
    c2 = scenarios.c2_se3(n=1)
    yield "c2 (10000)", c2["env"], c2["robot"], c2["params"], c2["starts"], c2["targets"], 10000
    c3 = scenarios.c3_linked(n=1)
    yield "c3 (100000)", c3["env"], c3["robot"], c3["params"], c3["starts"], c3["targets"], 100000


def run(mode):
    out = {}
    for name, env, robot, params, es, et, P in workloads():
        sim = B200Simulator(env, robot)
        E = es.shape[0]
        n = E * P
        d_s = torch.from_numpy(np.ascontiguousarray(es.T)).to(dev)
        d_t = torch.from_numpy(np.ascontiguousarray(et.T)).to(dev)
        d_full_s = torch.from_numpy(np.ascontiguousarray(np.repeat(es, P, axis=0).T)).to(dev)
        d_full_t = torch.from_numpy(np.ascontiguousarray(np.repeat(et, P, axis=0).T)).to(dev)
        d_out = torch.zeros((robot.width, n), dtype=torch.float64, device=dev)
        d_coll = torch.zeros(n, dtype=torch.uint8, device=dev)
        side = torch.cuda.Stream(device=dev)      # a non-NULL handle: NULL would select the library's own stream, which torch events do not see
        torch.cuda.set_stream(side)
        stream = side.cuda_stream
        res = {}
        variants = [("seeded", None)]
        tape_bytes = int(np.prod(abi.tape_shape(params, robot.dof, n))) * 8
        if mode == "ab" and tape_bytes < 8e9:
            tape = torch.from_numpy(abi.fill_noise_tape_seeded(7, robot, params, n)).to(dev)
            variants.append(("tape", tape))
        for label, tape in variants:
            times = []
            for rep in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                sim.ResetStatistics()
                a.record()
                if tape is None:
                    abi.check(lib.upc_forward_simulate_seeded_device(sim.handle, C.byref(params), 7, 0, n, E, d_s.data_ptr(), E, d_t.data_ptr(),
                                                                     d_out.data_ptr(), d_coll.data_ptr(), stream))
                else:
                    abi.check(lib.upc_forward_simulate_device(sim.handle, C.byref(params), n, d_full_s.data_ptr(), n, d_full_t.data_ptr(), tape.data_ptr(),
                                                              d_out.data_ptr(), d_coll.data_ptr(), stream))
                b.record()
                torch.cuda.synchronize()
                if rep:
                    times.append(a.elapsed_time(b))
            st = sim.GetStatistics()
            res[label] = {"ms": min(times), "particle_steps_per_s": n * params.num_steps / (min(times) * 1e-3), "checksum": float(d_out.sum().item()),
                          "resolver_iterations_per_particle_step": st["resolver_iterations"] / max(st["particle_steps"], 1)}
            if mode == "phases":
                cyc = np.zeros(12, dtype=np.int64)
                lib.upc_debug_phase_cycles.argtypes = [C.c_void_p, C.c_void_p]
                sim.ResetStatistics()
                lib.upc_debug_phase_cycles(sim.handle, cyc.ctypes.data)
                if tape is None:
                    abi.check(lib.upc_forward_simulate_seeded_device(sim.handle, C.byref(params), 7, 0, n, E, d_s.data_ptr(), E, d_t.data_ptr(),
                                                                     d_out.data_ptr(), d_coll.data_ptr(), stream))
                torch.cuda.synchronize()
                lib.upc_debug_phase_cycles(sim.handle, cyc.ctypes.data)
                steps = max(sim.GetStatistics()["particle_steps"], 1)
                res[label]["cycles_per_particle_step"] = {nm: float(cyc[k] / steps) for k, nm in enumerate(NAMES) if nm != "-"}
            del tape
        out[name] = res
        print(name, json.dumps(res), flush=True)
        sim.close()
        del d_out, d_full_s, d_full_t
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "ab"
    result = run(mode)
    print(json.dumps({"mode": mode, "library": abi.library_path(), "result": result}))
