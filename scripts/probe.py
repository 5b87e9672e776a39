"""quick GPU probe: throughput of the main configurations (development aid, not the benchmark)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode

def run(label, system, Y0, pars, name, T, **kw):
	ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)
	ODE.set_integrator(name, rtol=1e-6, atol=1e-12, storage=storage if "roessler" in label else None, **kw)
	if pars is not None and pars.shape[1]:
		ODE.set_parameters(torch.as_tensor(pars).cuda())
	Yd = torch.as_tensor(Y0).cuda()
	for rep in range(2):
		ODE.set_initial_value(Yd, 0.0)
		ODE.integrator.reset_step_counts()
		torch.cuda.synchronize()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		ODE.integrate(T)
		e1.record(); torch.cuda.synchronize()
		ms = e0.elapsed_time(e1)
		acc, rej = ODE.step_counts()
		print(f"{label:28s} {name:7s} B={len(Y0):8d} T={T:6.1f} {ms:10.2f} ms  steps={acc+rej:12d} (rej {rej})  {(acc+rej)/ms*1e3:.3e} steps/s", flush=True)

which = sys.argv[1:] or ["lv", "ro"]
storage = next((a[8:] for a in which if a.startswith("storage=")), None)
if "lv" in which:
	Y0, _ = W.lotka_volterra_ensemble(1_000_000)
	for name in ("dopri5", "RK45"):
		run("lotka_volterra", W.lotka_volterra(), Y0, None, name, 100.0)
if "ro" in which:
	B = int(next((a[2:] for a in which if a.startswith("B=")), 65536))
	Y0, pars = W.roessler_ensemble(B)
	for name in ("dopri5", "RK45"):
		run("roessler_100", W.roessler_network(100), Y0, pars, name, 20.0)
if "fhn" in which:
	from jitcode_b200 import jitcode_lyap
	B = 100_000
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False, seed=1)
	ODE.set_integrator("dopri5")
	ODE.set_initial_value(torch.as_tensor(W.fhn_lyap_ensemble(B)).cuda(), 0.0)
	torch.cuda.synchronize(); t0 = time.perf_counter()
	ODE.integrate_many(np.arange(1, 21)*10.0)
	torch.cuda.synchronize(); dt = time.perf_counter() - t0
	acc, rej = ODE.step_counts()
	print(f"fhn_lyap n=20 B={B} 20 calls: {dt*1e3:.1f} ms, {(acc+rej)/dt:.3e} steps/s")
