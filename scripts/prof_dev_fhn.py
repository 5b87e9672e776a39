"""where the device-resident step of the Lyapunov workload spends its time outside the integrate kernel (development aid)"""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
wl = bench.workload("fhn_lyap", 100_000)
from jitcode_b200 import jitcode_lyap
ODE = jitcode_lyap(wl["system"]["f_sym"], n_lyap=4, verbose=False, seed=1)
ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12)
Y0 = torch.as_tensor(wl["Y0"]).cuda()
def sync(): torch.cuda.synchronize()
for rep in range(6):
	sync(); t0 = time.perf_counter()
	ODE.set_initial_value(Y0, 0.0)
	sync(); t1 = time.perf_counter()
	ODE.integrator.kernel_events = []
	states, lyaps = ODE.integrate_many(wl["times"], record_states=False)
	sync(); t2 = time.perf_counter()
	kernel = sum(a.elapsed_time(b) for a, b in ODE.integrator.kernel_events)
	print(f"set_initial_value {1e3*(t1-t0):.2f} ms, integrate_many {1e3*(t2-t1):.2f} ms of which kernel {kernel:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
	for _ in range(3):
		ODE.set_initial_value(Y0, 0.0)
		ODE.integrate_many(wl["times"], record_states=False)
	sync()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
