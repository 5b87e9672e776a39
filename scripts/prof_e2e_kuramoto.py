"""where the end-to-end time of config 5 goes (development aid): python scripts/prof_e2e_kuramoto.py"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench

wl = bench.workload("kuramoto", 4096)
from jitcode_b200 import jitcode
ODE = jitcode.from_sine_coupling(*wl["system"]["sine_coupling"])
ODE.set_integrator(wl["integrator"], rtol=wl["rtol"], atol=wl["atol"])
Y0 = torch.as_tensor(wl["Y0"]).pin_memory()
times = wl["times"]
def clock(label, f, reps=3):
	for _ in range(reps):
		torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
		print(f"{label}: {1e3 * (time.perf_counter() - t0):8.1f} ms", flush=True)
	return r
def device_call():
	ODE.set_initial_value(Y0.cuda(), 0.0)
	return ODE.integrate_many(times)
def host_call():
	ODE.set_initial_value(Y0, 0.0)
	return ODE.integrate_many(times)
clock("device-resident", device_call)
clock("host buffers   ", host_call)
shape = (len(times), 4096, ODE.n)
clock("pinned alloc   ", lambda: torch.empty(shape, dtype=torch.float64, pin_memory=True))
dev = torch.empty(shape, dtype=torch.float64, device="cuda")
host = torch.empty(shape, dtype=torch.float64, pin_memory=True)
clock("d2h 1.6 GB     ", lambda: host.copy_(dev, non_blocking=True))
clock("device alloc   ", lambda: torch.empty(shape, dtype=torch.float64, device="cuda"))
