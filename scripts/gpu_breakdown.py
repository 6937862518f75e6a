"""Times the forward-simulation kernel on C2 variants to separate state math / point work / contact resolution."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uncertainty_planning_core_b200 import B200Simulator, abi, scenarios

def time_sim(env, robot, params, starts, targets, lanes, reps=5):
    n = starts.shape[0]
    sim = B200Simulator(env, robot)
    sim.SetLanesPerParticle(lanes)
    lib = abi.load_library()
    dev = torch.device("cuda:0")
    tape = torch.from_numpy(abi.fill_noise_tape(2, robot, params, n)).to(dev)
    d_s = torch.from_numpy(np.ascontiguousarray(starts.T)).to(dev); d_t = torch.from_numpy(np.ascontiguousarray(targets.T)).to(dev)
    d_o = torch.zeros_like(d_s); d_c = torch.zeros(n, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream(); torch.cuda.set_stream(side)
    best = 1e9
    for r in range(reps + 2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        abi.check(lib.upc_forward_simulate_device(sim.handle, C.byref(params), n, d_s.data_ptr(), targets.shape[0], d_t.data_ptr(), tape.data_ptr(), d_o.data_ptr(), d_c.data_ptr(), side.cuda_stream))
        e1.record(); torch.cuda.synchronize()
        if r >= 2: best = min(best, e0.elapsed_time(e1))
    st = sim.GetStatistics(); sim.close()
    return best, st

sc = scenarios.c2_se3()
free_env = abi.Environment(np.full(sc["env"].sdf.shape, 5.0, dtype=np.float32), sc["env"].resolution)
small = scenarios.se3_robot(lattice=(2, 2, 2))
for lanes in (1, 2, 4, 8, 16, 32):
    a, st = time_sim(sc["env"], sc["robot"], sc["params"], sc["starts"], sc["targets"], lanes)
    b, _ = time_sim(free_env, sc["robot"], sc["params"], sc["starts"], sc["targets"], lanes)
    c, _ = time_sim(free_env, small, sc["params"], sc["starts"], sc["targets"], min(lanes, 8))
    print("lanes %2d: contact %.3f ms | free space P=128 %.3f ms | free space P=8 %.3f ms | resolver iters/step %.2f" % (lanes, a, b, c, st["resolver_iterations"] / max(st["particle_steps"], 1)), flush=True)
for n in (100000, 1000000):
    starts = np.tile(sc["starts"][:1], (n, 1))
    p = abi.make_sim_params(16, 0.01, sc["env"].resolution)
    for lanes in (1, 2, 4):
        a, st = time_sim(sc["env"], sc["robot"], p, starts, sc["targets"], lanes, reps=2)
        print("n %d steps 16 lanes %d: %.3f ms -> %.3g particle-steps/s" % (n, lanes, a, n * 16 / a * 1e3), flush=True)
