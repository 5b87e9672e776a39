import os, sys
import numpy as np
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
name, interpolate = sys.argv[1], sys.argv[2] == "1"
system = W.lotka_volterra(gamma_as_parameter=True)
B = 3001
Y0, pars = W.lotka_volterra_ensemble(B, seed=9, gamma_as_parameter=True)
out = {}
for refill in (True, False):
	os.environ.pop("JITCODE_B200_NO_LANE_REFILL", None)
	if not refill:
		os.environ["JITCODE_B200_NO_LANE_REFILL"] = "1"
	ODE = jitcode(system["f_sym"], n=2, control_pars=system["control_pars"], verbose=False)
	ODE.set_integrator(name, interpolate=interpolate, rtol=1e-6, atol=1e-12)
	ODE.set_parameters(pars)
	ODE.set_initial_value(Y0, 0.0)
	res = [ODE.integrate(7.0), ODE.integrate(20.0), ODE.integrate(20.0 + 1e-9)]
	res.append(ODE.integrate_many([20.5, 24.0, 24.0 + 1e-7, 31.0]))
	res.append(ODE.y)
	res.append(np.array(ODE.step_counts()))
	out[refill] = res
for k, (x, y) in enumerate(zip(out[True], out[False])):
	d = np.abs(np.asarray(x, dtype=float) - np.asarray(y, dtype=float))
	print(k, np.asarray(x).shape, "max diff", d.max(), "n differing", int((d > 0).sum()), "first at", np.argwhere(d > 0)[:3].tolist())
