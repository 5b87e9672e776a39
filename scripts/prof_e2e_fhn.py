"""where the end-to-end time of the Lyapunov workload goes (development aid)"""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
wl = bench.workload("fhn_lyap", 100_000)
from jitcode_b200 import jitcode_lyap
ODE = jitcode_lyap(wl["system"]["f_sym"], n_lyap=4, verbose=False, seed=1)
ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12)
Y0 = torch.as_tensor(wl["Y0"]).pin_memory()
def sync(): torch.cuda.synchronize()
for rep in range(4):
	sync(); t0 = time.perf_counter()
	ODE.set_initial_value(Y0, 0.0)
	sync(); t1 = time.perf_counter()
	states = ODE.integrator.integrate_many(wl["times"], False)
	sync(); t2 = time.perf_counter()
	lyaps = ODE.integrator.last_lyaps
	host = ODE.integrator._to_host(lyaps)
	sync(); t3 = time.perf_counter()
	print(f"set_initial_value {1e3*(t1-t0):.2f} ms, integrate_many {1e3*(t2-t1):.2f} ms, exponents to host {1e3*(t3-t2):.2f} ms", lyaps.shape, lyaps.is_contiguous())
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
	ODE.set_initial_value(Y0, 0.0); s, l = ODE.integrate_many(wl["times"], record_states=False)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
print("---- steady state, as bench.py's host_step")
keep = None
for rep in range(8):
	sync(); t0 = time.perf_counter()
	ODE.set_initial_value(Y0, 0.0)
	t1 = time.perf_counter()
	s, l = ODE.integrate_many(wl["times"], record_states=False)
	t2 = time.perf_counter()
	keep = l
	sync(); t3 = time.perf_counter()
	print(f"iteration {rep}: set_initial_value {1e3*(t1-t0):.2f}  integrate_many {1e3*(t2-t1):.2f}  total {1e3*(t3-t0):.2f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(4):
	ODE.set_initial_value(Y0, 0.0); s, keep = ODE.integrate_many(wl["times"], record_states=False)
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(10)
