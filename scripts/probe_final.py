"""Rössler networks of several sizes and both driver families with the default configuration (development aid)."""
import sys
import torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
compile_only = not torch.cuda.is_available()
for N, B, name in [(50, 262144, "dopri5"), (100, 131072, "RK45"), (100, 131072, "dop853"), (200, 65536, "dopri5"), (500, 16384, "dopri5")]:
	system = W.roessler_network(N)
	ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)
	if compile_only:
		from jitcode_b200 import _codegen, _jitcode
		ODE.compile_cuda(method=_jitcode.INTEGRATORS[name][0], driver=_codegen.JB_DRV_ODE if _jitcode.INTEGRATORS[name][1] == "ode" else _codegen.JB_DRV_IVP_DENSE)
		continue
	ODE.set_integrator(name, rtol=1e-6, atol=1e-12)
	Y0, pars = W.roessler_ensemble(B, N=N)
	ODE.set_parameters(torch.as_tensor(pars).cuda())
	for rep in range(2):
		ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
		ODE.integrator.reset_step_counts()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record(); out = ODE.integrate(5.0); e1.record(); torch.cuda.synchronize()
	acc, rej = ODE.step_counts()
	info = ODE.integrator.module.info
	print(f"N={N} n={3*N} B={B} {name} storage={info.storage} threads={info.threads_per_block} smem={info.shared_bytes} {e0.elapsed_time(e1):8.1f} ms {(acc+rej)/e0.elapsed_time(e1)*1e3:.3e} steps/s", flush=True)
