"""a few samples of the Kuramoto config for ncu launch lists / captures (development aid):
    python scripts/prof_kuramoto.py [dop853|dopri5] [samples] [B]"""
import sys
import torch
sys.path.insert(0, ".")
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "dop853"
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
wl = bench.workload("kuramoto" if name == "dop853" else "kuramoto_dopri5", B)
from jitcode_b200 import jitcode
ODE = jitcode.from_sine_coupling(*wl["system"]["sine_coupling"])
import os
ODE.set_integrator(wl["integrator"], rtol=wl["rtol"], atol=wl["atol"], interleave=int(os.environ.get("INTERLEAVE", "0")))
Y0 = torch.as_tensor(wl["Y0"]).cuda()
for rep in range(2):
	ODE.set_initial_value(Y0, 0.0)
	ODE.integrator.reset_step_counts()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record(); ODE.integrate_many(wl["times"][:samples], record_states=False); e1.record(); torch.cuda.synchronize()
	attempts = sum(ODE.step_counts())
	print(f"{name} B={B} samples={samples}: {e0.elapsed_time(e1):.2f} ms, attempts {attempts}, rounds {ODE.integrator.rounds}, {attempts/e0.elapsed_time(e1)*1e3:.4e} steps/s")
