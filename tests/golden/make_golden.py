"""
Generates the golden fixtures under tests/golden/ by running the REFERENCE's own hot path
in this container (it cannot travel to the GPU box):

  * right-hand sides compiled inside the reference's jitcode/jitced_template.c
    (oracle/build_ref.py renders and compiles it where it lies),
  * driven by the reference's own jitcode/integrator_tools.py, imported standalone from
    /root/reference (it needs only NumPy + SciPy), i.e. ODE_wrapper / IVP_wrapper /
    IVP_wrapper_no_interpolation exactly as jitcode.set_integrator creates them
    (_jitcode.py:599-617; `nsteps=10**6` for the `ode` backend),
  * SciPy 1.18.1 underneath.

Run from the repo root:  python tests/golden/make_golden.py
Inputs are seeded (workloads.py); outputs are small .npz files committed next to this script.
"""

import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import workloads as W  # noqa: E402
from oracle import build_ref, c_rhs  # noqa: E402

GOLDEN = Path(__file__).resolve().parent

spec = importlib.util.spec_from_file_location("ref_integrator_tools", "/root/reference/jitcode/integrator_tools.py")
tools = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tools)

TOL = {"rtol": 1e-6, "atol": 1e-12}
VARIANTS = [  # label, integrator name, interpolate
	("dopri5", "dopri5", True),
	("dop853", "dop853", True),
	("RK45", "RK45", True),
	("RK45_land", "RK45", False),
	("DOP853", "DOP853", True),
	("DOP853_land", "DOP853", False),
	("RK23", "RK23", True),
	("RK23_land", "RK23", False),
]

def reference_backend(name, interpolate, f, **params):
	"""what jitcode.set_integrator builds (_jitcode.py:599-617)."""
	if tools.integrator_info(name)["backend"] == "ode":
		integrator = tools.ODE_wrapper(f)
		integrator.set_integrator(name, nsteps=10**6, **params)
		return integrator
	cls = tools.IVP_wrapper if interpolate else tools.IVP_wrapper_no_interpolation
	return cls(name, f, **params)

def run(handle, name, interpolate, Y0, pars, times, **params):
	out = np.empty((len(times),) + Y0.shape)
	for b in range(len(Y0)):
		if pars is not None and pars.shape[1]:
			handle.mod.initialise(*pars[b])
		integrator = reference_backend(name, interpolate, handle.f, **params)
		integrator.set_initial_value(Y0[b].copy(), 0.0)
		for i, T in enumerate(times):
			out[i, b] = integrator.integrate(T)
	return out

def system_handle(system):
	handle = build_ref.build_ref_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"])
	assert handle is not None, "needs /root/reference"
	return handle

def main():
	rng = np.random.default_rng(2026)

	# ---- FHN pair: f at random states, y(1) with every explicit method (tests/scenarios.py goldens alongside)
	fhn = system_handle(W.fhn())
	X = rng.uniform(-1, 1, (32, 4))
	fX = np.array([fhn.f(0.3*i, X[i]) for i in range(len(X))])
	fhn_out = {"X": X, "t": 0.3*np.arange(len(X)), "fX": fX}
	for label, name, interp in VARIANTS:
		fhn_out[f"y1_{label}"] = run(fhn, name, interp, W.FHN_Y0[None], None, [1.0])[0, 0]
		fhn_out[f"y1_{label}_tol"] = run(fhn, name, interp, W.FHN_Y0[None], None, [1.0], **TOL)[0, 0]
	np.savez(GOLDEN / "fhn.npz", **fhn_out)

	# ---- C1/C2 Lotka–Volterra: 48 members of the seeded ensemble, coarse and fine sampling
	lv = system_handle(W.lotka_volterra())
	Y0, _ = W.lotka_volterra_ensemble(48)
	times = np.array([2.5, 10.0, 30.0, 100.0])
	fine = np.arange(0.1, 3.05, 0.1)   # sampling finer than the step: dense-output path of IVP_wrapper
	lv_out = {"Y0": Y0, "times": times, "fine_times": fine}
	for label, name, interp in VARIANTS:
		lv_out[label] = run(lv, name, interp, Y0, None, times, **TOL)
		lv_out[f"{label}_fine"] = run(lv, name, interp, Y0[:8], None, fine, **TOL)
	# C1 exactly as the example: R0=.2, B0=.5, dopri5 defaults, times=arange(0,100,0.1) (first 300 samples)
	c1_times = np.arange(0.0, 100, 0.1)[:300]
	c1 = run(lv, "dopri5", True, np.array([[W.LV_C1_INITIAL["R"], W.LV_C1_INITIAL["B"]]]), None, c1_times)
	lv_out["c1_times"], lv_out["c1_dopri5"] = c1_times, c1[:, 0]
	np.savez(GOLDEN / "lotka_volterra.npz", **lv_out)

	# ---- control parameter variant (γ per trajectory)
	lvg = system_handle(W.lotka_volterra(gamma_as_parameter=True))
	Y0g, gam = W.lotka_volterra_ensemble(16, gamma_as_parameter=True)
	np.savez(GOLDEN / "lotka_volterra_gamma.npz", Y0=Y0g, pars=gam, times=times,
			dopri5=run(lvg, "dopri5", True, Y0g, gam, times, **TOL),
			RK45=run(lvg, "RK45", True, Y0g, gam, times, **TOL))

	# ---- C3 Rössler network N=100: 6 members (k, Y0 per trajectory), chaotic → short horizon
	ro = system_handle(W.roessler_network(100))
	Y0r, kr = W.roessler_ensemble(6)
	tr = np.array([10.0, 20.0, 50.0])
	Xr = rng.random((4, 300))
	tX = np.array([0.0, 0.7, 1.9, 33.0])
	ro.mod.initialise(0.0123)
	fX = np.array([ro.f(tt, x) for tt, x in zip(tX, Xr)])
	np.savez(GOLDEN / "roessler_100.npz", Y0=Y0r, pars=kr, times=tr,
			dopri5=run(ro, "dopri5", True, Y0r, kr, tr, **TOL),
			RK45=run(ro, "RK45", True, Y0r, kr, tr, **TOL),
			X=Xr, tX=tX, kX=np.array([0.0123]), fX=fX)

	# ---- jitcode_lyap's 20-dimensional system (tests/test_generation.py:85-97 evaluates it at rng(42).random(20))
	full, helpers = c_rhs.lyapunov_system(W.fhn()["f_sym"], 4)
	ly = build_ref.build_ref_rhs(full, helpers, (), 20)
	x20 = np.random.default_rng(42).random(20)
	np.savez(GOLDEN / "fhn_lyap_rhs.npz", x=x20, fx=ly.f(0.0, x20))
	print("golden fixtures written to", GOLDEN)

if __name__ == "__main__":
	main()
