"""a tiny run of the three storage modes and the dense engine for compute-sanitizer (development aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode, jitcode_lyap
which = sys.argv[1] if len(sys.argv) > 1 else "warp"
if which in ("warp", "team"):
	system = W.roessler_network(100)
	Y0, pars = W.roessler_ensemble(40, seed=1)
	for name in ("dopri5", "RK45"):
		ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=300, control_pars=system["control_pars"], verbose=False)
		ODE.set_integrator(name, rtol=1e-6, atol=1e-12, storage="warp", team_warps=(2 if which == "team" else None))
		ODE.set_parameters(pars[:, 0])
		ODE.set_initial_value(Y0, 0.0)
		print(name, np.abs(ODE.integrate(0.3)).max(), np.abs(ODE.integrate(0.5)).max())
elif which == "lyapwarp":
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False, seed=1)
	ODE.set_integrator("dopri5", storage="warp")
	ODE.set_initial_value(np.tile(W.FHN_Y0, (33, 1)), 0.0)
	print(ODE.integrate(5.0)[1][:2])
elif which == "dense":
	system = W.kuramoto(40)
	ODE = jitcode(system["f_sym"], n=40, verbose=False)
	ODE.set_integrator("dopri5", storage="dense")
	ODE.set_initial_value(W.kuramoto_ensemble(70, 40), 0.0)
	print(np.abs(ODE.integrate(1.0)).max())
else:
	system = W.fhn()
	ODE = jitcode(system["f_sym"], verbose=False)
	ODE.set_integrator("RK45", storage=which)
	ODE.set_initial_value(np.tile(W.FHN_Y0, (70, 1)), 0.0)
	print(ODE.integrate(1.0)[0], ODE.integrate(1.01)[0])
