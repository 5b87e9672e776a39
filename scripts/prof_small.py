"""one launch of a register-storage (lv) or global-storage (fhn_lyap) kernel for ncu (development aid)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode, jitcode_lyap
which = sys.argv[1]
if which == "lv":
	B = 1 << 20
	ODE = jitcode(W.lotka_volterra()["f_sym"], n=2, verbose=False)
	ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12)
	Y0 = torch.as_tensor(W.lotka_volterra_ensemble(B)[0]).cuda()
	run = lambda: ODE.integrate(100.0)
else:
	B = 100_000
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False, seed=1)
	ODE.set_integrator("dopri5")
	Y0 = torch.as_tensor(W.fhn_lyap_ensemble(B)).cuda()
	run = lambda: ODE.integrate_many(np.arange(1, 21) * 10.0, record_states=False)
for rep in range(2):
	ODE.set_initial_value(Y0, 0.0)
	ODE.integrator.reset_step_counts()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record(); run(); e1.record(); torch.cuda.synchronize()
	print(which, e0.elapsed_time(e1), "ms", sum(ODE.step_counts()), "attempts")
