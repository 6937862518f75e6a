"""Per-phase cycle breakdown of the simulation program (needs a library built with -DUPC_PHASE_TIMING)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uncertainty_planning_core_b200 import B200Simulator, abi, scenarios
names = ["control", "estimate", "apply", "check", "resolver", "reached", "tape/misc", "-", "res:points", "res:replay", "res:solve", "contributing points (count)"]
lib = abi.load_library()
lib.upc_debug_phase_cycles.argtypes = [C.c_void_p, C.c_void_p]
sc = scenarios.c2_se3()
n = sc["starts"].shape[0]
tape = abi.fill_noise_tape(2, sc["robot"], sc["params"], n)
for lanes in (1, 4, 8, 16):
    sim = B200Simulator(sc["env"], sc["robot"]); sim.SetLanesPerParticle(lanes)
    sim.ForwardSimulateRobots(sc["starts"], sc["targets"], tape, sc["params"])
    out = np.zeros(12, dtype=np.int64)
    lib.upc_debug_phase_cycles(sim.handle, out.ctypes.data)
    st = sim.GetStatistics()
    steps = n * sc["params"].num_steps
    print("lanes %2d: cycles per particle-step: " % lanes + "  ".join("%s %.0f" % (nm, out[k] / steps) for k, nm in enumerate(names)) + "  | total %.0f | resolver iters/step %.2f" % (out[:7].sum() / steps, st["resolver_iterations"] / steps), flush=True)
    sim.close()
