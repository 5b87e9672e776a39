"""lane refill of the register kernels: same results, how much faster? (development aid)"""
import os, sys
import torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
system = W.lotka_volterra()
Y0, _ = W.lotka_volterra_ensemble(B)
Yd = torch.as_tensor(Y0).cuda() if torch.cuda.is_available() else None
results = {}
variants = sys.argv[2:] or ["default"]      # numbers: JB_REFILL_LANES
for name in ("dopri5",) if len(variants) > 1 else ("dopri5", "dop853"):
	for refill in ("off", *variants):
		os.environ.pop("JITCODE_B200_NO_LANE_REFILL", None)
		if refill == "off":
			os.environ["JITCODE_B200_NO_LANE_REFILL"] = "1"
		ODE = jitcode(system["f_sym"], n=2, verbose=False)
		if refill.isdigit():
			ODE._compile_options["extra_compile_args"] = (f"-DJB_REFILL_LANES={refill}",)
		if not torch.cuda.is_available():
			from jitcode_b200 import _codegen
			ODE.compile_cuda(method=_codegen.JB_DP5 if name == "dopri5" else _codegen.JB_DOP853); continue
		ODE.set_integrator(name, rtol=1e-6, atol=1e-9)
		best = 1e9
		for rep in range(3):
			ODE.set_initial_value(Yd, 0.0)
			ODE.integrator.reset_step_counts()
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			e0.record(); out = ODE.integrate(100.0); e1.record(); torch.cuda.synchronize()
			best = min(best, e0.elapsed_time(e1))
		acc, rej = ODE.step_counts()
		results[(name, refill)] = out.clone()
		print(f"{name} refill={refill} {best:8.2f} ms {(acc+rej)/best*1e3:.3e} steps/s accepted {acc} rejected {rej}", flush=True)
	if torch.cuda.is_available():
		print(name, "identical results:", all(bool((results[(name, "off")] == results[(name, v)]).all()) for v in variants), flush=True)
