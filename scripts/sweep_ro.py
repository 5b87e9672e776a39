"""sweep of warp-storage configurations on the Rössler network (development aid)."""
import sys, itertools
import torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = 10.0
system = W.roessler_network(100)
Y0, pars = W.roessler_ensemble(B)
Yd, Pd = torch.as_tensor(Y0).cuda(), torch.as_tensor(pars).cuda()
configs = [("global", None, None)] + [("warp", k, th) for k, th in [(0, None), (2, None), (4, None), (5, None), (6, None), (7, None), (7, 192), (7, 128), (6, 256), (5, 256), (4, 256), (0, 256)]]
if len(sys.argv) > 2:
	configs = [("warp", int(a.split(":")[0]), int(a.split(":")[1]) or None) for a in sys.argv[2:]]
ref = None
for storage, kregs, threads in configs:
	ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)
	try:
		ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12, storage=storage, kregs=kregs, threads=threads)
	except Exception as error:
		print(storage, kregs, threads, "FAILED", str(error)[-300:]); continue
	ODE.set_parameters(Pd)
	best = 1e9
	try:
		for rep in range(3):
			ODE.set_initial_value(Yd, 0.0)
			ODE.integrator.reset_step_counts()
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			e0.record(); out = ODE.integrate(T); e1.record(); torch.cuda.synchronize()
			best = min(best, e0.elapsed_time(e1))
	except Exception as error:
		print(storage, kregs, threads, "FAILED", str(error)[-200:]); continue
	acc, rej = ODE.step_counts()
	if ref is None: ref = out
	info = ODE.integrator.module.info
	print(f"{storage:6s} kregs={kregs} threads={info.threads_per_block:4d} smem={info.shared_bytes:6d} {best:8.2f} ms {(acc+rej)/best*1e3:.3e} steps/s  maxdiff {float((out-ref).abs().max()):.2e}", flush=True)
