"""examples/kuramoto_network.py as it is (n = 100, q = 0.2): which kernels handle it, and how fast (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
n, B = 100, 65536
system = W.kuramoto(n)
Y0 = torch.as_tensor(W.kuramoto_ensemble(B, n)).cuda()
ref = None
for storage in (None, "warp", "global", "dense"):
	t0 = time.time()
	ODE = jitcode(system["f_sym"], n=n, verbose=False)
	try:
		ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-9, storage=storage)
	except Exception as error:
		print(storage, "FAILED", str(error)[-300:]); continue
	build = time.time() - t0
	for rep in range(2):
		ODE.set_initial_value(Y0, 0.0)
		ODE.integrator.reset_step_counts()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record(); out = ODE.integrate(10.0); e1.record(); torch.cuda.synchronize()
	acc, rej = ODE.step_counts()
	if ref is None: ref = out
	info = ODE.integrator.module.info
	print(f"storage={storage} -> {info.storage} threads={info.threads_per_block} build {build:.1f}s {e0.elapsed_time(e1):8.1f} ms {(acc+rej)/e0.elapsed_time(e1)*1e3:.3e} steps/s maxdiff {float((out-ref).abs().max()):.2e}", flush=True)
