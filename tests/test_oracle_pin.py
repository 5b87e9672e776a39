"""
Pins the CPU oracle (oracle/) before anything is compared against it:

  * the reference test-suite's golden vectors for the FHN pair (tests/scenarios.py:14-38 of the
    reference, restated in workloads.py),
  * the fixtures under tests/golden/, which tests/golden/make_golden.py produced by running the
    reference's own jitced_template.c + integrator_tools.py + SciPy in the build container,
  * SciPy itself, step for step (number of RHS evaluations), for the NumPy restatements of the
    two step-size controllers.
"""

import numpy as np
import pytest
from numpy.testing import assert_allclose

import workloads as W
from oracle import c_rhs, ensemble, rk_numpy, wrappers

TOL = {"rtol": 1e-6, "atol": 1e-12}
VARIANTS = [
	("dopri5", "dopri5", True), ("dop853", "dop853", True), ("RK45", "RK45", True),
	("RK45_land", "RK45", False), ("DOP853", "DOP853", True), ("DOP853_land", "DOP853", False),
	("RK23", "RK23", True), ("RK23_land", "RK23", False),
]


def build(system):
	return c_rhs.build_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"])


@pytest.fixture(scope="module")
def fhn():
	return build(W.fhn())


@pytest.fixture(scope="module")
def lv():
	return build(W.lotka_volterra())


def test_f_of_y0_golden(fhn):
	# reference tests/test_generation.py:57
	assert_allclose(fhn.f(0.0, W.FHN_Y0), W.FHN_F_OF_Y0, rtol=1e-5)


def test_f_matches_reference_container(fhn, golden):
	g = golden("fhn")
	assert_allclose(fhn.f_many(g["t"], g["X"]), g["fX"], rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("label,name,interpolate", VARIANTS)
def test_y1_golden(fhn, golden, label, name, interpolate):
	# reference tests/test_jitcode.py:93,130 and test_integrator_tools.py:38-40
	backend = wrappers.make_backend(name, fhn.f, interpolate=interpolate)
	backend.set_initial_value(W.FHN_Y0, 0.0)
	y1 = backend.integrate(1.0)
	# reference tests/test_jitcode.py: rtol 1e-4 (:93), 1e-3 in the sweep over all integrators (:130), which RK23 needs
	assert_allclose(y1, W.FHN_Y1, rtol=1e-3 if name == "RK23" else 1e-4)
	assert_allclose(y1, golden("fhn")[f"y1_{label}"], rtol=1e-9)


def test_wrapper_contract(fhn):
	# reference tests/test_integrator_tools.py:52-69
	rng = np.random.default_rng(42)
	for name, interpolate in [("dopri5", True), ("dop853", True), ("RK45", True), ("RK45", False)]:
		backend = wrappers.make_backend(name, fhn.f, interpolate=interpolate)
		initial = rng.random(4)
		backend.set_initial_value(initial)
		assert_allclose(initial, backend.integrate(0))
		if not (name == "RK45" and interpolate):
			with pytest.raises(ValueError):
				backend.integrate(-1)


def test_forced_failure():
	# reference tests/test_integrator_tools.py:59-63: time-reversed system, tolerance 1e-12
	system = W.fhn()
	back = c_rhs.build_rhs([-entry for entry in system["f_sym"]])
	for name, interpolate in [("dopri5", True), ("RK45", True), ("RK45", False)]:
		backend = wrappers.make_backend(name, back.f, interpolate=interpolate, atol=1e-12, rtol=1e-12)
		backend.set_initial_value([0, 1, 2, 3], 0)
		with pytest.raises(wrappers.UnsuccessfulIntegration):
			backend.integrate(1)


@pytest.mark.parametrize("label,name,interpolate", VARIANTS)
def test_port_reproduces_reference_runs(lv, golden, label, name, interpolate):
	g = golden("lotka_volterra")
	res = ensemble.run_ensemble(lv, name, g["Y0"][:12], g["times"], interpolate=interpolate, processes=1, **TOL)
	assert_allclose(res["y"], g[label][:, :12], rtol=1e-10)
	fine = ensemble.run_ensemble(lv, name, g["Y0"][:4], g["fine_times"], interpolate=interpolate, processes=1, **TOL)
	assert_allclose(fine["y"], g[f"{label}_fine"][:, :4], rtol=1e-10)


def test_control_parameters(golden):
	g = golden("lotka_volterra_gamma")
	handle = build(W.lotka_volterra(gamma_as_parameter=True))
	res = ensemble.run_ensemble(handle, "dopri5", g["Y0"], g["times"], pars=g["pars"], processes=2, **TOL)
	assert_allclose(res["y"], g["dopri5"], rtol=1e-10)


def test_roessler_rhs_and_short_run(golden):
	g = golden("roessler_100")
	handle = build(W.roessler_network(100))
	handle.initialise(*g["kX"])
	assert_allclose(handle.f_many(g["tX"], g["X"]), g["fX"], rtol=1e-12, atol=1e-15)
	res = ensemble.run_ensemble(handle, "dopri5", g["Y0"][:2], g["times"][:1], pars=g["pars"][:2], processes=1, **TOL)
	assert_allclose(res["y"], g["dopri5"][:1, :2], rtol=1e-7)


def test_lyapunov_rhs(golden):
	# the 20-dimensional system of jitcode_lyap (reference tests/test_generation.py:85-97)
	g = golden("fhn_lyap_rhs")
	full, helpers = c_rhs.lyapunov_system(W.fhn()["f_sym"], 4)
	handle = c_rhs.build_rhs(full, helpers)
	assert_allclose(handle.f(0.0, g["x"]), g["fx"], rtol=1e-13)


def _vectorised(handle):
	return lambda t, Y: handle.f_many(np.ascontiguousarray(t), np.ascontiguousarray(Y))


@pytest.mark.parametrize("method", ["RK45", "DOP853"])
@pytest.mark.parametrize("interpolate", [True, False])
def test_numpy_restatement_of_scipy_ivp(lv, golden, method, interpolate):
	g = golden("lotka_volterra")
	Y0 = g["Y0"][:10]
	stepper = rk_numpy.ScipyRK(_vectorised(lv), 0.0, Y0, method, **TOL)
	ref = ensemble.run_ensemble(lv, method, Y0, g["times"], interpolate=interpolate, processes=1, **TOL)
	for i, T in enumerate(g["times"]):
		out = stepper.integrate_dense(T) if interpolate else stepper.integrate_land(T)
		assert_allclose(out, ref["y"][i], rtol=1e-9)
	assert np.array_equal(stepper.nfev, ref["nfev"])   # identical step sequences
	label = method if interpolate else method + "_land"
	assert_allclose(out, g[label][-1, :10], rtol=1e-9)


@pytest.mark.parametrize("method", ["dopri5", "dop853"])
def test_numpy_restatement_of_hairer(lv, golden, method):
	g = golden("lotka_volterra")
	Y0 = g["Y0"][:10]
	solver = rk_numpy.HairerDopri(_vectorised(lv), method, **TOL)
	ref = ensemble.run_ensemble(lv, method, Y0, g["times"], processes=1, **TOL)
	Y, t, nfev = Y0, 0.0, 0
	for i, T in enumerate(g["times"]):
		result = solver.integrate(t, Y, T)
		Y, t, nfev = result["y"], T, nfev + result["nfev"]
		assert not result["status"].any()
		assert_allclose(Y, ref["y"][i], rtol=1e-9)
	if method == "dopri5":
		assert np.array_equal(nfev, ref["nfev"])
	else:
		# scipy's compiled dop853 sometimes re-evaluates f once per step depending on what ran before
		# (observed: 12 extra calls in one segment with bit-identical step times); results above are equal
		assert np.all(nfev <= ref["nfev"]) and np.all(ref["nfev"] - nfev <= 0.2*nfev)
	assert_allclose(Y, g[method][-1, :10], rtol=1e-9)


def test_lyapunov_spectrum_short():
	# reference tests/test_jitcode.py:132-162 runs 9 999 calls; here 1 200 calls × 2 trajectories
	full, helpers = c_rhs.lyapunov_system(W.fhn()["f_sym"], 4)
	handle = c_rhs.build_rhs(full, helpers)
	rng = np.random.default_rng(1)
	Y0 = np.empty((2, 20))
	Y0[:, :4] = W.FHN_Y0
	tangents = rng.normal(size=(2, 4, 4))
	Y0[:, 4:] = (tangents / np.linalg.norm(tangents, axis=2, keepdims=True)).reshape(2, 16)
	res = ensemble.run_ensemble(handle, "dopri5", Y0, range(10, 12010, 10), lyap=(4, 4), processes=2)
	lyaps = res["lyaps"][200:].mean(axis=(0, 1))
	assert_allclose(lyaps, W.FHN_LYAPS, atol=0.015)


def _stiff(lam):
	"""fast relaxation towards cos(t) plus a slow decay: stiff for the explicit pairs once lam·h reaches their stability limit"""
	def scalar(t, y):
		return [-lam * (y[0] - np.cos(t)), -0.5 * y[1]]
	def batched(t, Y):
		return np.stack([-lam * (Y[:, 0] - np.cos(t)), -0.5 * Y[:, 1]], axis=1)
	return scalar, batched


@pytest.mark.parametrize("lam,T,expect_stiff", [(1e4, 20.0, True), (1e3, 20.0, False)])
def test_stiffness_detection_dopri5_matches_scipy(lam, T, expect_stiff):
	# Hairer's stiffness test (dopri5.f; IDID = −4 → scipy "problem is probably stiff (interrupted)"), which
	# ODE_wrapper.integrate turns into UnsuccessfulIntegration (integrator_tools.py:137-140): the restatement
	# stops at the same time, in the same state, after the same number of right-hand-side evaluations
	import warnings
	from scipy.integrate import ode
	scalar, batched = _stiff(lam)
	calls = [0]
	def counted(t, y):
		calls[0] += 1
		return scalar(t, y)
	solver = ode(counted).set_integrator("dopri5", nsteps=10**6, **TOL)
	solver.set_initial_value([0.0, 1.0], 0.0)
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		y_scipy = solver.integrate(T)
	assert (solver.get_return_code() == -4) == expect_stiff
	result = rk_numpy.HairerDopri(batched, "dopri5", nsteps=10**6, **TOL).integrate(0.0, np.array([[0.0, 1.0]]), T)
	assert result["status"][0] == (3 if expect_stiff else 0)
	assert result["nfev"][0] == calls[0]
	assert_allclose(result["t"][0], solver.t, rtol=1e-9)
	assert_allclose(result["y"][0], y_scipy, rtol=1e-7)


def test_stiffness_detection_dop853():
	# dop853.f (limit 6.1): at the stability limit the error estimate amplifies rounding differences, so the step
	# sequences of SciPy's build and of the restatement drift apart after ~20 steps; both give up as stiff, in the
	# same window of accepted steps (1000 + 15 k) and at comparable times
	import warnings
	from scipy.integrate import ode
	scalar, batched = _stiff(1e6)
	solver = ode(scalar).set_integrator("dop853", nsteps=10**6, **TOL)
	solver.set_initial_value([0.0, 1.0], 0.0)
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		solver.integrate(1.0)
	assert solver.get_return_code() == -4
	result = rk_numpy.HairerDopri(batched, "dop853", nsteps=10**6, **TOL).integrate(0.0, np.array([[0.0, 1.0]]), 1.0)
	assert result["status"][0] == 3
	assert_allclose(result["t"][0], solver.t, rtol=0.2)


def test_network_rhs_equals_the_written_out_form():
	"""oracle/network_rhs.py (the Kuramoto sum as a loop over adjacency lists, for networks too large to write out) against
	the written-out form of examples/kuramoto_network.py:19-26 compiled like every other oracle system"""
	from oracle import network_rhs
	system = W.kuramoto(40)
	written_out = c_rhs.build_rhs(system["f_sym"], n=40)
	loop = network_rhs.NetworkRhsHandle(system["A"], system["omega"], system["c"])
	for state in W.kuramoto_ensemble(6, 40, seed=2):
		assert_allclose(loop.f(0.0, state), written_out.f(0.0, state), rtol=1e-13, atol=1e-15)
	backend = wrappers.make_backend("dop853", loop.f, atol=1e-6, rtol=0.0)
	reference = wrappers.make_backend("dop853", written_out.f, atol=1e-6, rtol=0.0)
	state = W.kuramoto_ensemble(1, 40, seed=3)[0]
	backend.set_initial_value(state, 0.0)
	reference.set_initial_value(state, 0.0)
	assert_allclose(backend.integrate(5.0), reference.integrate(5.0), rtol=1e-10, atol=1e-12)
