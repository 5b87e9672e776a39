"""Rössler networks of other sizes: which storage is chosen and what it delivers (development aid)."""
import sys, time
import torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
for N, B in [(50, 131072), (200, 65536), (500, 16384)]:
	system = W.roessler_network(N)
	Y0, pars = W.roessler_ensemble(B, N=N)
	for storage in ((None,) if "warp-only" in sys.argv else (None, "global")):
		t0 = time.time()
		ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)
		ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12, storage=storage)
		build = time.time() - t0
		ODE.set_parameters(torch.as_tensor(pars).cuda())
		for rep in range(2):
			ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
			ODE.integrator.reset_step_counts()
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			e0.record(); out = ODE.integrate(5.0); e1.record(); torch.cuda.synchronize()
		acc, rej = ODE.step_counts()
		info = ODE.integrator.module.info
		print(f"N={N} n={3*N} B={B} storage={info.storage} threads={info.threads_per_block} smem={info.shared_bytes} build {build:.1f}s  {e0.elapsed_time(e1):8.1f} ms {(acc+rej)/e0.elapsed_time(e1)*1e3:.3e} steps/s  {(acc+rej)/e0.elapsed_time(e1)*1e3*280*3*N/1e9:.0f} GB/s-equivalent", flush=True)
