"""where the end-to-end time of the Lyapunov workload goes (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode_lyap
B = 100_000
ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False, seed=1)
ODE.set_integrator("dopri5")
Y0 = W.fhn_lyap_ensemble(B)
times = np.arange(1, 21) * 10.0
for rep in range(4):
	torch.cuda.synchronize(); t0 = time.perf_counter()
	ODE.set_initial_value(Y0, 0.0)
	torch.cuda.synchronize(); t1 = time.perf_counter()
	states, exponents = ODE.integrate_many(times, record_states=False)
	torch.cuda.synchronize(); t2 = time.perf_counter()
	print(f"set_initial_value {1e3*(t1-t0):.1f} ms, integrate_many {1e3*(t2-t1):.1f} ms, result {type(exponents).__name__} {getattr(exponents, 'shape', None)}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
ODE.set_initial_value(Y0, 0.0); states, exponents = ODE.integrate_many(times, record_states=False)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
