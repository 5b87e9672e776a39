"""one Rössler-network launch for ncu (development aid)."""
import sys
import torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
T = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
name = sys.argv[3] if len(sys.argv) > 3 else "dopri5"
system = W.roessler_network(100)
Y0, pars = W.roessler_ensemble(B)
ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)
import os
ODE.set_integrator(name, rtol=1e-6, atol=1e-12, kregs=(int(os.environ["KREGS"]) if "KREGS" in os.environ else None))
ODE.set_parameters(torch.as_tensor(pars).cuda())
for rep in range(2):
	ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
	ODE.integrator.reset_step_counts()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record(); ODE.integrate(T); e1.record(); torch.cuda.synchronize()
	acc, rej = ODE.step_counts()
	print(f"B={B} T={T} {e0.elapsed_time(e1):.2f} ms steps={acc+rej} {(acc+rej)/e0.elapsed_time(e1)*1e3:.3e} steps/s")
