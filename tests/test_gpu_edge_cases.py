"""
Edge cases of the hot path on the GPU, each against the CPU oracle (the reference's integrators):
ragged batch sizes in every storage mode, per-component atol, max_step / first_step, the step limit,
parameter changes between calls, integrator swaps that keep the state, NaN inputs.
"""

import numpy as np
import pytest
import sympy
import torch
from numpy.testing import assert_allclose

import workloads as W
from jitcode_b200 import UnsuccessfulIntegration, jitcode, t, y
from oracle import c_rhs, ensemble, wrappers

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-6, 1e-12


def within_tolerance(result, expected, factor=10.0):
	excess = np.abs(result - expected) - factor * (ATOL + RTOL * np.abs(expected))
	assert np.all(excess <= 0), f"max excess {excess.max():.3e}"


def make(system):
	return jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)


def oracle(system, name, Y0, times, pars=None, **params):
	handle = c_rhs.build_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"])
	return ensemble.run_ensemble(handle, name, Y0, times, pars=pars, processes=2, **params)["y"]


@pytest.mark.parametrize("B", [1, 31, 33, 65])
@pytest.mark.parametrize("storage", ["registers", "global"])
def test_ragged_batches_thread_storage(B, storage):
	system = W.fhn()
	Y0 = W.FHN_Y0 + 0.01 * np.random.default_rng(B).normal(size=(B, 4))
	ODE = make(system)
	ODE.set_integrator("dopri5", storage=storage, rtol=RTOL, atol=ATOL)
	ODE.set_initial_value(Y0, 0.0)
	result = ODE.integrate(30.0)
	assert result.shape == (B, 4)
	within_tolerance(result, oracle(system, "dopri5", Y0, [30.0], rtol=RTOL, atol=ATOL)[0])


@pytest.mark.parametrize("B", [1, 7, 148 * 13 + 5])
def test_ragged_batches_warp_storage(B):
	system = W.roessler_network(100)
	Y0, pars = W.roessler_ensemble(B, seed=B)
	ODE = make(system)
	ODE.set_integrator("dopri5", storage="warp", rtol=RTOL, atol=ATOL)
	ODE.set_parameters(pars[:, 0])
	ODE.set_initial_value(Y0, 0.0)
	result = ODE.integrate(2.0)
	assert result.shape == (B, 300)
	ref = oracle(system, "dopri5", Y0[:5], [2.0], pars=pars[:5], rtol=RTOL, atol=ATOL)[0]
	assert_allclose(result[:5], ref, rtol=1e-7, atol=1e-9)
	# every trajectory was integrated (none left at its initial value), also in the last partial block
	assert np.all(np.abs(result - Y0).max(axis=1) > 1e-3)


@pytest.mark.parametrize("name,interpolate", [("dopri5", True), ("RK45", True), ("RK45", False)])
def test_vector_atol_and_max_step(name, interpolate):
	system = W.lotka_volterra()
	Y0, _ = W.lotka_volterra_ensemble(40, seed=3)
	# SciPy 1.18.1's compiled dopri5 only takes a scalar atol ("only length-1 arrays …"); solve_ivp takes one per component
	atol = np.array([1e-9, 1e-9]) if name == "dopri5" else np.array([1e-9, 1e-5])
	params = {"rtol": 1e-7, "atol": atol, "max_step": 0.37}
	ODE = make(system)
	ODE.set_integrator(name, interpolate=interpolate, **params)
	ODE.set_initial_value(Y0, 0.0)
	result = ODE.integrate(12.0)
	handle = c_rhs.build_rhs(system["f_sym"], n=2)
	if name == "dopri5":
		params["atol"] = 1e-9
	ref = ensemble.run_ensemble(handle, name, Y0, [12.0], interpolate=interpolate, processes=2, **params)
	bound = 10 * (atol + 1e-7 * np.abs(ref["y"][0]))
	assert np.all(np.abs(result - ref["y"][0]) <= bound)
	attempts = sum(ODE.step_counts())
	assert attempts >= 40 * int(12.0 / 0.37)                      # max_step is honoured
	assert abs(attempts - ref["attempts"]) <= 0.05 * ref["attempts"]


def test_first_step_and_step_limit():
	system = W.lotka_volterra()
	Y0, _ = W.lotka_volterra_ensemble(8, seed=4)
	ODE = make(system)
	ODE.set_integrator("dopri5", rtol=RTOL, atol=ATOL, first_step=1e-3)
	ODE.set_initial_value(Y0, 0.0)
	within_tolerance(ODE.integrate(5.0), oracle(system, "dopri5", Y0, [5.0], rtol=RTOL, atol=ATOL, first_step=1e-3)[0])
	# nsteps: the reference raises UnsuccessfulIntegration when the limit is hit (integrator_tools.py:137-140)
	ODE.set_integrator("dopri5", nsteps=5, rtol=RTOL, atol=ATOL)
	ODE.set_initial_value(Y0, 0.0)
	with pytest.raises(UnsuccessfulIntegration) as failure:
		ODE.integrate(100.0)
	assert len(failure.value.indices) == 8 and set(failure.value.status) == {2}
	backend = wrappers.make_backend("dopri5", c_rhs.build_rhs(system["f_sym"], n=2).f, nsteps=5, rtol=RTOL, atol=ATOL)
	backend.set_initial_value(Y0[0], 0.0)
	with pytest.raises(wrappers.UnsuccessfulIntegration):
		backend.integrate(100.0)


@pytest.mark.parametrize("name,interpolate", [("dopri5", True), ("dop853", True), ("RK45", True), ("RK45", False), ("DOP853", True), ("RK23", False)])
def test_lane_refill_changes_nothing(name, interpolate, monkeypatch):
	"""large ensembles in register storage run on persistent warps whose lanes take trajectories from a queue (all three
	drivers, any number of sample times): step counts and failure reports are those of the one-thread-per-trajectory
	launch, results agree to rounding (the two kernels are separate instantiations of the same source, and the compiler
	decides per instantiation where the generated right-hand side contracts a·b + c into one FMA), and those are the
	oracle's within tolerance"""
	system = W.lotka_volterra(gamma_as_parameter=True)
	B = 3001                                          # ragged: the last warps run dry at different times
	Y0, pars = W.lotka_volterra_ensemble(B, seed=9, gamma_as_parameter=True)
	hairer = name in ("dopri5", "dop853")
	outcomes = {}
	for refill in (True, False):
		if refill:
			monkeypatch.delenv("JITCODE_B200_NO_LANE_REFILL", raising=False)
		else:
			monkeypatch.setenv("JITCODE_B200_NO_LANE_REFILL", "1")
		ODE = make(system)
		ODE.set_integrator(name, interpolate=interpolate, rtol=RTOL, atol=ATOL)
		ODE.set_parameters(pars)
		ODE.set_initial_value(Y0, 0.0)
		first = ODE.integrate(7.0)
		second = ODE.integrate(20.0)                 # a second call continues from the first (solve_ivp names: the same stepper)
		third = ODE.integrate(20.0 + 1e-9) if not hairer else second      # dense output: a sample inside the last step, no stepping at all
		# several sample times in one launch, two of them so close that the dense output serves both from one step
		many = ODE.integrate_many([20.5, 24.0, 24.0 + 1e-7, 31.0])
		counts = ODE.step_counts()
		failed = ()
		if hairer:
			ODE.set_integrator(name, nsteps=40, rtol=RTOL, atol=ATOL)
			ODE.set_initial_value(Y0, 0.0)
			with pytest.raises(UnsuccessfulIntegration) as failure:
				ODE.integrate(60.0)
			failed = tuple(failure.value.indices)
			assert len(failed) > 0                   # trajectories that hit the step limit
		outcomes[refill] = (first, second, third, many, counts, failed, ODE.y)
	for with_refill, without in zip(outcomes[True], outcomes[False]):
		if isinstance(with_refill, np.ndarray):
			assert_allclose(with_refill, without, rtol=1e-12, atol=0, equal_nan=True)
		else:
			assert with_refill == without
	expected = oracle(system, name, Y0[:64], [7.0, 20.0], pars=pars[:64], interpolate=interpolate, rtol=RTOL, atol=ATOL)
	within_tolerance(outcomes[True][0][:64], expected[0])
	within_tolerance(outcomes[True][1][:64], expected[1])


def test_parameters_can_change_between_calls():
	system = W.lotka_volterra(gamma_as_parameter=True)
	Y0, gam = W.lotka_volterra_ensemble(12, seed=6, gamma_as_parameter=True)
	for name in ("dopri5", "RK45"):
		ODE = make(system)
		ODE.set_integrator(name, rtol=RTOL, atol=ATOL)
		ODE.set_parameters(gam[:, 0])
		ODE.set_initial_value(Y0, 0.0)
		middle = ODE.integrate(5.0)
		ODE.set_parameters(0.5)                              # one value for all trajectories from here on
		final = ODE.integrate(9.0)
		first = oracle(system, name, Y0, [5.0], pars=gam, rtol=RTOL, atol=ATOL)[0]
		within_tolerance(middle, first)
		handle = c_rhs.build_rhs(system["f_sym"], control_pars=system["control_pars"], n=2)
		second = ensemble.run_ensemble(handle, name, first, [9.0], pars=np.full((12, 1), 0.5), t0=5.0, processes=1, rtol=RTOL, atol=ATOL)["y"][0]
		within_tolerance(final, second, factor=30.0)


def test_integrator_swap_keeps_state_and_time():
	# reference _jitcode.py:619-626
	system = W.lotka_volterra()
	Y0, _ = W.lotka_volterra_ensemble(10, seed=8)
	ODE = make(system)
	ODE.set_initial_value(Y0, 0.0)                       # before any integrator exists (empty_integrator holds it)
	ODE.set_integrator("dopri5", rtol=RTOL, atol=ATOL)
	middle = ODE.integrate(4.0)
	ODE.set_integrator("RK45", rtol=RTOL, atol=ATOL)
	assert ODE.t == 4.0
	assert_allclose(ODE.y, middle, rtol=0, atol=0)
	final = ODE.integrate(8.0)
	within_tolerance(final, oracle(system, "dopri5", Y0, [8.0], rtol=RTOL, atol=ATOL)[0], factor=30.0)
	assert set(ODE.y_dict) == {y(0), y(1)}


def test_explicit_time_dependence_and_helpers():
	# reference tests/test_sympy_input.py: dy/dt = cos(t)·y  →  y(t) = y0·exp(sin t); plus a helper
	h = sympy.Symbol("h")
	ODE = jitcode([h * y(0)], helpers=[(h, sympy.cos(t))], verbose=False)
	ODE.set_integrator("dopri5", rtol=1e-10, atol=1e-12)
	start = np.linspace(0.5, 2.0, 37).reshape(-1, 1)
	ODE.set_initial_value(start, 0.0)
	assert_allclose(ODE.integrate(10.0), start * np.exp(np.sin(10.0)), rtol=1e-8)


def test_non_finite_input_is_reported_not_hung():
	system = W.lotka_volterra()
	Y0, _ = W.lotka_volterra_ensemble(6, seed=9)
	Y0[2, 0] = np.nan
	for name in ("dopri5", "RK45"):
		ODE = make(system)
		ODE.set_integrator(name, rtol=RTOL, atol=ATOL)
		ODE.set_initial_value(Y0, 0.0)
		with pytest.raises(UnsuccessfulIntegration) as failure:
			ODE.integrate(3.0)
		assert list(failure.value.indices) == [2]
		healthy = ODE.integrator.y_device.cpu().numpy()[[0, 1, 3, 4, 5]]
		assert np.all(np.isfinite(healthy))


def test_torch_cpu_round_trip_and_out_buffer():
	system = W.lotka_volterra()
	Y0, _ = W.lotka_volterra_ensemble(100, seed=10)
	ODE = make(system)
	ODE.set_integrator("dopri5", rtol=RTOL, atol=ATOL)
	ODE.set_initial_value(torch.as_tensor(Y0), 0.0)
	result = ODE.integrate(3.0)
	assert isinstance(result, torch.Tensor) and not result.is_cuda
	out = torch.empty((100, 2), dtype=torch.float64).pin_memory()
	ODE.set_initial_value(torch.as_tensor(Y0).pin_memory(), 0.0)
	assert ODE.integrate(3.0, out=out) is out
	assert_allclose(out.numpy(), result.numpy(), rtol=0, atol=0)


@pytest.mark.parametrize("case", ["lotka_volterra", "roessler", "fhn_rk45_dense"])
def test_streamed_host_ensemble_equals_plain_calls(case):
	"""integrate_streamed (chunked, copies overlapped with kernels) gives exactly what set_initial_value + integrate give."""
	if case == "roessler":
		system, (Y0, pars), T, name = W.roessler_network(100), W.roessler_ensemble(1000, seed=5), 1.5, "dopri5"
	elif case == "lotka_volterra":
		system, (Y0, pars), T, name = W.lotka_volterra(gamma_as_parameter=True), W.lotka_volterra_ensemble(5003, seed=5, gamma_as_parameter=True), 9.0, "dopri5"
	else:
		system, T, name = W.fhn(), 25.0, "RK45"
		Y0, pars = W.FHN_Y0 + 0.01 * np.random.default_rng(1).normal(size=(777, 4)), np.empty((777, 0))
	ODE = make(system)
	ODE.set_integrator(name, rtol=RTOL, atol=ATOL)
	if pars.shape[1]:
		ODE.set_parameters(pars[:, 0])
	ODE.set_initial_value(Y0, 0.0)
	plain = ODE.integrate(T)
	counts = ODE.step_counts()
	ODE.integrator.reset_step_counts()
	host = torch.as_tensor(Y0).pin_memory()
	out = torch.empty_like(host).pin_memory()
	assert ODE.integrate_streamed(host, T, out, chunks=5) is out
	assert_allclose(out.numpy(), plain, rtol=0, atol=0)
	assert ODE.step_counts() == counts and ODE.t == T
	assert_allclose(ODE.y.numpy(), plain, rtol=0, atol=0)


@pytest.mark.parametrize("name,interpolate", [("RK45", True), ("RK45", False), ("DOP853", True)])
def test_warp_storage_options(name, interpolate):
	"""per-component atol (goes through the component permutation of warp storage), max_step, and the loop over
	sample times as one launch — all against the thread-per-trajectory kernel"""
	system = W.roessler_network(100)
	Y0, pars = W.roessler_ensemble(50, seed=21)
	atol = 10.0 ** np.random.default_rng(3).uniform(-12, -7, 300)
	times = [0.3, 0.35, 1.0, 2.5]
	results = {}
	for storage in ("warp", "global"):
		ODE = make(system)
		ODE.set_integrator(name, interpolate=interpolate, storage=storage, rtol=1e-7, atol=atol, max_step=0.05)
		ODE.set_parameters(pars[:, 0])
		ODE.set_initial_value(Y0, 0.0)
		results[storage] = ODE.integrate_many(times)
		assert sum(ODE.step_counts()) >= 50 * 50              # max_step 0.05 over t = 0 … 2.5
		ODE.set_initial_value(Y0, 0.0)
		looped = np.stack([ODE.integrate(T) for T in times])
		assert_allclose(results[storage], looped, rtol=1e-12, atol=1e-14)
	assert_allclose(results["warp"], results["global"], rtol=1e-9, atol=1e-12)


def _relaxation_system():
	"""y0 relaxes towards cos(t) at rate lam (control parameter), y1 decays slowly: stiff for large lam"""
	lam = sympy.Symbol("lam")
	return {"f_sym": [-lam * (y(0) - sympy.cos(t)), -0.5 * y(1)], "helpers": [], "control_pars": [lam], "n": 2}


@pytest.mark.parametrize("B,storage", [(40, None), (600, None), (40, "global"), (40, "shared")])
def test_stiffness_detection(B, storage):
	"""Hairer's stiffness test (dopri5.f: every 1000 accepted steps; 15 suspicious steps in a row → IDID −4 →
	UnsuccessfulIntegration, integrator_tools.py:137-140), per trajectory: the stiff members of a mixed ensemble stop
	where SciPy's dopri5 stops, the others arrive.  B = 600 takes the lane-refill kernel."""
	from oracle import rk_numpy
	system = _relaxation_system()
	rates = np.where(np.arange(B) % 3 == 0, 1e4, 1e3)
	Y0 = np.tile([0.0, 1.0], (B, 1)) + 0.01 * np.random.default_rng(2).normal(size=(B, 2))
	ODE = make(system)
	ODE.set_integrator("dopri5", rtol=RTOL, atol=ATOL, **({"storage": storage} if storage else {}))
	ODE.set_parameters(rates)
	ODE.set_initial_value(Y0, 0.0)
	with pytest.raises(UnsuccessfulIntegration) as failure:
		ODE.integrate(20.0)
	stiff = np.flatnonzero(rates == 1e4)
	assert np.array_equal(failure.value.indices, stiff) and set(failure.value.status) == {3}
	state = ODE.integrator.y_device.cpu().numpy()
	times = ODE.integrator.tt[:B].cpu().numpy()
	sample = np.r_[stiff[:4], np.flatnonzero(rates == 1e3)[:4]]
	for b in sample:
		def batched(tt, Y, lam=rates[b]):
			return np.stack([-lam * (Y[:, 0] - np.cos(tt)), -0.5 * Y[:, 1]], axis=1)
		ref = rk_numpy.HairerDopri(batched, "dopri5", nsteps=10**6, rtol=RTOL, atol=ATOL).integrate(0.0, Y0[b:b+1], 20.0)
		assert ref["status"][0] == (3 if rates[b] == 1e4 else 0)
		assert_allclose(times[b], ref["t"][0], rtol=1e-6)
		assert_allclose(state[b], ref["y"][0], rtol=1e-5, atol=1e-9)
	accepted = ODE.integrator.n_accepted[:B].cpu().numpy()
	assert np.all(accepted[stiff] == 1014)            # the 1000th accepted step raises the first suspicion, the 1014th the 15th


def test_sample_at_the_current_time_for_large_ensembles():
	"""integrate_many([t_now, …]) with enough trajectories for the lane-refill kernel: the sample at the current time
	is the current state (ODE_wrapper.integrate returns self._y for t == self.t, integrator_tools.py:142-143)"""
	system = W.lotka_volterra()
	Y0, _ = W.lotka_volterra_ensemble(1024, seed=12)
	ODE = make(system)
	ODE.set_integrator("dopri5", rtol=RTOL, atol=ATOL)
	ODE.set_initial_value(Y0, 0.0)
	assert_allclose(ODE.integrate_many([0.0])[0], Y0, rtol=0, atol=0)
	both = ODE.integrate_many([0.0, 3.0])
	assert_allclose(both[0], Y0, rtol=0, atol=0)
	within_tolerance(both[1][:32], oracle(system, "dopri5", Y0[:32], [3.0], rtol=RTOL, atol=ATOL)[0])
	with pytest.raises(ValueError):
		ODE.integrate_many([5.0, 4.0])                  # sample times must not decrease


def test_work_queue_of_warp_storage_changes_nothing():
	"""warp storage with more trajectories than resident warps draws them from a queue; without the caller's counter
	(queue = NULL) the warps stride statically: same results bit for bit"""
	system = W.roessler_network(100)
	B = 148 * 14 + 77
	Y0, pars = W.roessler_ensemble(B, seed=31)
	results = []
	for use_queue in (True, False):
		ODE = make(system)
		ODE.set_integrator("dopri5", storage="warp", rtol=RTOL, atol=ATOL)
		ODE.set_parameters(pars[:, 0])
		ODE.set_initial_value(Y0, 0.0)
		if not use_queue:
			ODE.integrator.queue = None
		results.append((ODE.integrate(1.0), ODE.step_counts()))
	assert np.array_equal(results[0][0], results[1][0]) and results[0][1] == results[1][1]
