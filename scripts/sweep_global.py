"""register-cap sweep for the thread-global kernel (development aid)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode, jitcode_lyap
which = sys.argv[1]
for storage, threads, cap in [("global", 64, 8), ("shared", 160, 1), ("shared", 128, 1), ("shared", 96, 1), ("shared", 64, 1), ("shared", 64, 2), ("shared", 32, 4)]:
	for _ in (0,):
		args = []
		if which == "ro":
			system = W.roessler_network(100)
			B, T = 65536, 5.0
			Y0, pars = W.roessler_ensemble(B)
			ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=300, control_pars=system["control_pars"], verbose=False)
			ODE.compile_cuda(extra_compile_args=args)
			ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12, storage="global", threads=threads, min_blocks=cap)
			ODE.set_parameters(torch.as_tensor(pars).cuda())
			init = lambda: ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
			run = lambda: ODE.integrate(T)
		else:
			B = 100_000
			ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False, seed=1)
			ODE.compile_cuda(extra_compile_args=args)
			ODE.set_integrator("dopri5", storage=storage, threads=threads, min_blocks=cap)
			Y0 = torch.as_tensor(W.fhn_lyap_ensemble(B)).cuda()
			init = lambda: ODE.set_initial_value(Y0, 0.0)
			run = lambda: ODE.integrate_many(np.arange(1, 21) * 10.0)
		best = 1e9
		for rep in range(2):
			init(); ODE.integrator.reset_step_counts()
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			e0.record(); run(); e1.record(); torch.cuda.synchronize()
			best = min(best, e0.elapsed_time(e1))
		acc, rej = ODE.step_counts()
		print(f"{which} {storage} cap={cap} threads={threads} {best:9.2f} ms {(acc+rej)/best*1e3:.3e} steps/s", flush=True)
