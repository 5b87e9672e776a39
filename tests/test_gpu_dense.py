"""
Dense sine-coupled phase oscillators (BASELINE config 5, reference examples/kuramoto_network.py) on the FP64
tensor cores: the GEMM itself, the right-hand side, and the lock-step DOPRI5 against the generic kernels, the
CPU oracle (the reference's written-out form) and size-independent properties at n = 1000.
"""

import ctypes

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

import workloads as W
from jitcode_b200 import _dense, jitcode

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-6, 1e-12


def within_tolerance(result, expected, factor=10.0, rtol=RTOL, atol=ATOL):
	excess = np.abs(result - expected) - factor * (atol + rtol * np.abs(expected))
	assert np.all(excess <= 0), f"max excess {excess.max():.3e}"


@pytest.mark.parametrize("M,N,K", [(128, 64, 16), (256, 128, 64), (300, 200, 250), (1000, 192, 1000), (7, 5, 3)])
def test_dmma_gemm(M, N, K):
	lib = _dense.dense_library()
	rng = np.random.default_rng(M + N + K)
	A = torch.as_tensor(rng.normal(size=(M, K))).cuda()
	B = torch.as_tensor(rng.normal(size=(K, N))).cuda()
	C = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
	stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
	assert lib.jb_dense_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, stream) == 0
	expected = A.cpu().numpy() @ B.cpu().numpy()
	assert_allclose(C.cpu().numpy(), expected, rtol=1e-12, atol=1e-12 * np.sqrt(K))


@pytest.mark.parametrize("shape", [0, 1, 2, 3])
@pytest.mark.parametrize("M,N,K", [(1000, 192, 1000), (250, 130, 62), (113, 64, 16)])
def test_dmma_gemm_every_tile_shape(shape, M, N, K):
	"""the pipelined GEMM is compiled for four block-tile shapes (launch_gemm picks the one whose last wave of blocks is
	fullest); each of them on sizes with ragged edges in M, N and K"""
	lib = _dense.dense_library()
	rng = np.random.default_rng(shape + M + N + K)
	A = torch.as_tensor(rng.normal(size=(M, K))).cuda()
	B = torch.as_tensor(rng.normal(size=(K, N))).cuda()
	C = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
	stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
	assert lib.jb_dense_gemm_shape(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, shape, stream) == 0
	assert_allclose(C.cpu().numpy(), A.cpu().numpy() @ B.cpu().numpy(), rtol=1e-12, atol=1e-12 * np.sqrt(K))


def test_dense_kuramoto_1000_against_the_oracle():
	"""config 5 at its full system size against the CPU oracle: 32 trajectories of the 1000-oscillator network under the
	example's integrator (dop853, atol 1e-6, rtol 0; examples/kuramoto_network.py:29) and under dopri5, sampled like the
	example's loop; the oracle evaluates the reference's sum edge by edge (oracle/network_rhs.py), not the GEMM form."""
	from oracle import ensemble, network_rhs
	data = W.kuramoto_data(1000)
	n = data["n"]
	Wm = data["c"] / (n - 1) * data["A"].T.astype(float)
	np.fill_diagonal(Wm, 0.0)
	handle = network_rhs.NetworkRhsHandle(data["A"], data["omega"], data["c"])
	Y0 = W.kuramoto_ensemble(32, n, seed=44)
	assert_allclose(jitcode.from_sine_coupling(Wm, data["omega"]).f(0.0, Y0[:4]), handle.f_many(0.0, Y0[:4]), rtol=1e-11, atol=1e-12)
	times = [1.0, 2.0, 3.0, 10.0]
	for name, tol in (("dop853", {"atol": 1e-6, "rtol": 0.0}), ("dopri5", {"atol": 1e-9, "rtol": 1e-6})):
		ODE = jitcode.from_sine_coupling(Wm, data["omega"])
		ODE.set_integrator(name, **tol)
		ODE.set_initial_value(Y0, 0.0)
		result = ODE.integrate_many(times)
		ref = ensemble.run_ensemble(handle, name, Y0, times, **tol)
		within_tolerance(result, ref["y"], factor=10.0, **tol)
		# the oracle counts right-hand-side evaluations (÷ 12 or 6), which include the two of every fresh start
		evaluations = {"dop853": 12, "dopri5": 6}[name] * sum(ODE.step_counts()) + 2 * len(times) * len(Y0)
		assert abs(evaluations - ref["nfev"].sum()) <= 0.03 * ref["nfev"].sum()


def kuramoto_rhs(Wm, omega, Y):
	diff = Y[:, None, :] - Y[:, :, None]            # [b, i, j] = y_j − y_i
	return omega + np.einsum("ij,bij->bi", Wm, np.sin(diff))


def test_structure_detection_and_rhs():
	n = 48
	system = W.kuramoto(n)
	ODE = jitcode(system["f_sym"], n=n, verbose=False)
	Wm, omega = ODE._dense_coupling()
	expected = system["c"] / (n - 1) * system["A"].T.astype(float)
	np.fill_diagonal(expected, 0.0)                   # sin(y_i − y_i) vanishes symbolically
	assert_allclose(Wm, expected, rtol=1e-14)
	assert_allclose(omega, system["omega"], rtol=1e-14)
	Y = W.kuramoto_ensemble(70, n)
	ODE.compile_cuda()
	generic = ODE.f(0.0, Y)                           # generated code of the written-out form
	ODE.set_integrator("dopri5", storage="dense")
	dense = ODE.f(0.0, Y)
	assert_allclose(dense, kuramoto_rhs(Wm, omega, Y), rtol=1e-12, atol=1e-13)
	assert_allclose(dense, generic, rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("name", ["dopri5", "dop853"])
def test_dense_integration_matches_generic_kernels_and_oracle(name):
	from oracle import c_rhs, ensemble
	n, B = 40, 150
	system = W.kuramoto(n)
	Y0 = W.kuramoto_ensemble(B, n)
	times = [1.0, 5.0]
	results, counts = {}, {}
	for storage in ("dense", "global"):
		ODE = jitcode(system["f_sym"], n=n, verbose=False)
		ODE.set_integrator(name, rtol=RTOL, atol=ATOL, storage=storage)
		ODE.set_initial_value(Y0, 0.0)
		results[storage] = np.stack([ODE.integrate(T) for T in times])
		counts[storage] = ODE.step_counts()
		with pytest.raises(ValueError):
			ODE.integrate(0.5)
	within_tolerance(results["dense"], results["global"])
	assert abs(sum(counts["dense"]) - sum(counts["global"])) <= 0.02 * sum(counts["global"])
	auto = jitcode(system["f_sym"], n=n, verbose=False)       # 20 % density: the dense engine is the default
	auto.set_integrator(name, rtol=RTOL, atol=ATOL)
	assert type(auto.integrator).__name__ == "CudaDenseSineIntegrator"
	handle = c_rhs.build_rhs(system["f_sym"], n=n)
	ref = ensemble.run_ensemble(handle, name, Y0[:24], times, rtol=RTOL, atol=ATOL)
	within_tolerance(results["dense"][:, :24], ref["y"])


@pytest.mark.parametrize("name,interpolate", [("RK45", True), ("RK45", False), ("DOP853", True), ("DOP853", False)])
def test_dense_solve_ivp_drivers(name, interpolate):
	"""SciPy's solve_ivp stepper inside the dense engine (IVP_wrapper / IVP_wrapper_no_interpolation,
	integrator_tools.py:98-128): persistent between calls, dense output of the step that passed the target time or last
	step clipped to it — against the generic kernels (the same drivers, golden-pinned in test_gpu_parity.py), against the
	CPU oracle (SciPy's own stepper under the reference's wrappers) and step for step in the number of attempts.
	Sample times closer than a step (several samples from one interpolant), one call per time and all in one call."""
	from oracle import c_rhs, ensemble
	n, B = 40, 150
	system = W.kuramoto(n)
	Y0 = W.kuramoto_ensemble(B, n)
	times = [0.5, 0.51, 0.52, 1.0, 5.0]
	results, counts = {}, {}
	for storage in ("dense", "global"):
		ODE = jitcode(system["f_sym"], n=n, verbose=False)
		ODE.set_integrator(name, interpolate=interpolate, rtol=RTOL, atol=ATOL, storage=storage)
		assert (type(ODE.integrator).__name__ == "CudaDenseSineIntegrator") == (storage == "dense")
		ODE.set_initial_value(Y0, 0.0)
		results[storage] = np.stack([ODE.integrate(T) for T in times])
		counts[storage] = ODE.step_counts()
		if interpolate:
			# IVP_wrapper.integrate reads an earlier time off the last step's interpolant (no step, no error)
			earlier = ODE.integrate(4.9999)
			assert ODE.step_counts() == counts[storage] and np.abs(earlier - results[storage][-1]).max() < 1e-2
		else:
			with pytest.raises(ValueError):
				ODE.integrate(0.5)
	within_tolerance(results["dense"], results["global"])
	# the same stepper: the same number of attempts, up to decisions that fall differently by round-off (GEMM against term by term)
	assert abs(sum(counts["dense"]) - sum(counts["global"])) <= 0.01 * sum(counts["global"])
	handle = c_rhs.build_rhs(system["f_sym"], n=n)
	ref = ensemble.run_ensemble(handle, name, Y0[:24], times, interpolate=interpolate, rtol=RTOL, atol=ATOL)
	within_tolerance(results["dense"][:, :24], ref["y"])
	# all sample times in one call; a second initial value re-creates the steppers
	ODE = jitcode(system["f_sym"], n=n, verbose=False)
	ODE.set_integrator(name, interpolate=interpolate, rtol=RTOL, atol=ATOL, storage="dense")
	ODE.set_initial_value(Y0[:70], 0.0)
	ODE.integrate(0.3)
	ODE.set_initial_value(Y0, 0.0)
	many = ODE.integrate_many(times)
	assert_allclose(many, results["dense"], rtol=0, atol=0)
	assert ODE.step_counts()[0] >= counts["dense"][0]


def test_dense_solve_ivp_with_coupling_parameter_vector_atol_and_first_step():
	"""the solve_ivp drivers of the dense engine with everything the Hairer drivers have: the coupling strength as a
	per-trajectory control parameter, one atol per component, a given first step — against the generic kernels"""
	import sympy
	from jitcode_b200 import y
	n, B = 40, 96
	data = W.kuramoto_data(n)
	kappa = sympy.Symbol("c")
	f = [data["omega"][i] + kappa / (n - 1) * sympy.Add(*[sympy.sin(y(j) - y(i)) for j in range(n) if data["A"][j, i]]) for i in range(n)]
	Y0 = W.kuramoto_ensemble(B, n, seed=8)
	strengths = np.random.default_rng(2).uniform(0.5, 4.0, B)
	atol = np.geomspace(1e-12, 1e-9, n)
	results, counts = {}, {}
	for storage in ("dense", "global"):
		ODE = jitcode(f, n=n, control_pars=[kappa], verbose=False)
		ODE.set_integrator("RK45", rtol=1e-7, atol=atol, first_step=1e-3, storage=storage)
		ODE.set_parameters(strengths)
		ODE.set_initial_value(Y0, 0.0)
		results[storage] = np.stack([ODE.integrate(T) for T in (0.7, 2.0)])
		counts[storage] = ODE.step_counts()
	excess = np.abs(results["dense"] - results["global"]) - 10 * (atol + 1e-7 * np.abs(results["global"]))
	assert np.all(excess <= 0), f"max excess {excess.max():.3e}"
	assert abs(sum(counts["dense"]) - sum(counts["global"])) <= 0.01 * sum(counts["global"])


def test_dense_samples_streamed_to_the_host():
	"""integrate_many with a host-resident ensemble copies every sample to page-locked memory on a second stream while the
	next one is integrated (large outputs only): the same numbers as the device-resident call"""
	n, B = 40, 4096
	system = W.kuramoto(n)
	Y0 = W.kuramoto_ensemble(B, n, seed=12)
	times = [0.25 * k for k in range(1, 9)]
	ODE = jitcode(system["f_sym"], n=n, verbose=False)
	ODE.set_integrator("dop853", atol=1e-6, rtol=0, storage="dense")
	ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
	on_device = ODE.integrate_many(times).cpu().numpy()
	ODE.set_initial_value(Y0, 0.0)
	on_host = ODE.integrate_many(times)
	assert isinstance(on_host, np.ndarray) and on_host.shape == (8, B, n)
	assert_allclose(on_host, on_device, rtol=0, atol=0)
	ODE.set_initial_value(Y0, 0.0)
	last = ODE.integrate_many(times, record_states=False)
	assert_allclose(last[0], on_device[-1], rtol=0, atol=0)


def test_dense_example_settings():
	"""the example's own integrator settings (examples/kuramoto_network.py:29): dop853, atol 1e-6, rtol 0"""
	from oracle import c_rhs, ensemble
	n, B = 40, 64
	system = W.kuramoto(n)
	Y0 = W.kuramoto_ensemble(B, n, seed=3)
	ODE = jitcode(system["f_sym"], n=n, verbose=False)
	ODE.set_integrator("dop853", atol=1e-6, rtol=0, storage="dense")
	ODE.set_initial_value(Y0, 0.0)
	result = ODE.integrate(10.0)
	handle = c_rhs.build_rhs(system["f_sym"], n=n)
	ref = ensemble.run_ensemble(handle, "dop853", Y0, [10.0], atol=1e-6, rtol=0)
	assert np.abs(result - ref["y"][0]).max() <= 10 * 1e-6
	# the oracle counts right-hand-side calls / 12; SciPy's dop853 spends 11 on a rejected attempt and now and then one
	# more at the start of a call (tests/test_oracle_pin.py), so the comparison of attempts is loose
	attempts = sum(ODE.step_counts())
	assert abs(attempts - ref["attempts"]) <= 0.15 * ref["attempts"]


def test_dense_kuramoto_1000():
	"""config 5 at full system size: n = 1000, dense adjacency; properties that need no reference run."""
	data = W.kuramoto_data(1000)
	n = data["n"]
	Wm = data["c"] / (n - 1) * data["A"].T.astype(float)
	np.fill_diagonal(Wm, 0.0)
	ODE = jitcode.from_sine_coupling(Wm, data["omega"])
	Y0 = W.kuramoto_ensemble(192, n)
	assert_allclose(ODE.f(0.0, Y0[:16]), kuramoto_rhs(Wm, data["omega"], Y0[:16]), rtol=1e-11, atol=1e-12)
	ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-9)
	ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
	coarse = ODE.integrate(2.0).cpu().numpy()
	accepted, rejected = ODE.step_counts()
	rounds = ODE.integrator.rounds
	assert accepted >= 192 and rounds * 192 >= accepted + rejected        # lock-step: rounds ≥ attempts of the slowest
	assert (accepted + rejected) / (rounds * 192) > 0.5                   # … and most slots do useful work
	# a ten times tighter run is the yardstick for the looser one
	ODE.set_integrator("dopri5", rtol=1e-8, atol=1e-11)
	ODE.set_initial_value(Y0, 0.0)
	fine = ODE.integrate(2.0)
	assert np.abs(coarse - fine).max() < 1e-4
	# the mean phase velocity is the mean frequency plus the antisymmetric-part imbalance; for every trajectory
	# Σ_i dy_i/dt = Σω + Σ_ij W_ij sin(y_j − y_i) holds — check f at the final state against NumPy once more
	assert_allclose(ODE.f(0.0, fine[:8]), kuramoto_rhs(Wm, data["omega"], fine[:8]), rtol=1e-11, atol=1e-12)


def test_dense_stiffness_detection():
	"""Hairer's stiffness test in the dense engine (control block C_NACC / C_STIFF / C_STNUM / C_STDEN): strongly coupled
	oscillators lock within a fraction of a time unit and the locked state is stiff for an explicit pair — SciPy's
	dopri5 gives up after 1000 + 14 accepted steps (IDID −4 → UnsuccessfulIntegration, integrator_tools.py:137-140);
	so does every trajectory here, at the oracle's time and state; a weakly coupled ensemble of the same size arrives."""
	from jitcode_b200 import UnsuccessfulIntegration
	from oracle import rk_numpy
	n, B = 32, 6
	rng = np.random.default_rng(11)
	omega = rng.uniform(-0.5, 0.5, n)
	Y0 = rng.uniform(0, 0.3, (B, n))
	for strength, stiff in ((1e4, True), (1.0, False)):
		Wm = np.full((n, n), strength / n)
		np.fill_diagonal(Wm, 0.0)
		ODE = jitcode.from_sine_coupling(Wm, omega)
		ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12)
		assert ODE.integrator.module.info.storage == 4
		ODE.set_initial_value(Y0, 0.0)
		ref = rk_numpy.HairerDopri(lambda t, Y: kuramoto_rhs(Wm, omega, Y), "dopri5", rtol=1e-6, atol=1e-12, nsteps=10**6).integrate(0.0, Y0, 2.0)
		if stiff:
			with pytest.raises(UnsuccessfulIntegration) as failure:
				ODE.integrate(2.0)
			assert list(failure.value.indices) == list(range(B)) and set(failure.value.status) == {3}
			assert set(ref["status"]) == {3}
			assert np.all(ODE.integrator.n_accepted[:B].cpu().numpy() == 1014)
		else:
			ODE.integrate(2.0)
			assert set(ref["status"]) == {0}
		# at the stability limit the step sequence amplifies rounding differences (tests/test_oracle_pin.py): the time at
		# which the test gives up agrees to a few per cent there, and to rounding where the problem is not stiff
		assert_allclose(ODE.integrator.tt[:B].cpu().numpy(), ref["t"], rtol=0.05 if stiff else 1e-12)
		assert_allclose(ODE.integrator.y_device.cpu().numpy(), ref["y"], rtol=1e-6, atol=0.05 if stiff else 1e-9)


@pytest.mark.parametrize("name", ["dopri5", "dop853"])
def test_dense_coupling_parameter_and_vector_atol(name):
	"""the coupling strength as a per-trajectory control parameter (dy_i/dt = ω_i + c·Σ_j W_ij sin(y_j − y_i), recognised
	in the symbolic input) and one atol per component: dense engine against the generic kernels, which compile the
	written-out form with `c` as an ordinary control parameter, and against the oracle"""
	import sympy
	from jitcode_b200 import y
	from oracle import c_rhs, ensemble
	n, B = 40, 70
	data = W.kuramoto_data(n)
	A, omega = data["A"], data["omega"]
	c = sympy.Symbol("c")
	f = [omega[i] + c / (n - 1) * sympy.Add(*[sympy.sin(y(j) - y(i)) for j in range(n) if A[j, i]]) for i in range(n)]
	Y0 = W.kuramoto_ensemble(B, n, seed=8)
	strengths = np.random.default_rng(9).uniform(0.5, 4.0, B)
	atol = 10.0 ** np.random.default_rng(10).uniform(-11, -8, n)
	times = [1.0, 2.5, 6.0]
	results = {}
	for storage in ("dense", None):
		ODE = jitcode(f, n=n, control_pars=[c], verbose=False)
		ODE.set_integrator(name, rtol=1e-7, atol=atol, **({"storage": storage} if storage else {"storage": "global"}))
		assert (ODE.integrator.module.info.storage == 4) == (storage == "dense")
		ODE.set_parameters(strengths)
		ODE.set_initial_value(Y0, 0.0)
		results[storage] = ODE.integrate_many(times)
	bound = 10 * (atol + 1e-7 * np.abs(results[None]))
	assert np.all(np.abs(results["dense"] - results[None]) <= bound)
	# the default choice for such a system is the dense engine, and one value serves all trajectories
	ODE = jitcode(f, n=n, control_pars=[c], verbose=False)
	ODE.set_integrator(name, rtol=1e-7, atol=atol)
	assert ODE.integrator.module.info.storage == 4
	ODE.set_parameters(2.0)
	ODE.set_initial_value(Y0[:16], 0.0)
	single = ODE.integrate(2.0)
	handle = c_rhs.build_rhs(f, control_pars=[c], n=n)
	# (SciPy's compiled dopri5/dop853 take a scalar atol: the oracle runs with the smallest one)
	ref = ensemble.run_ensemble(handle, name, Y0[:16], [2.0], pars=np.full((16, 1), 2.0), rtol=1e-7, atol=float(atol.min()))["y"][0]
	assert np.all(np.abs(single - ref) <= 30 * (atol + 1e-7 * np.abs(ref)))
	assert_allclose(ODE.f(0.3, Y0[:9]), handle_f_many(handle, 2.0, Y0[:9]), rtol=1e-11, atol=1e-12)


def handle_f_many(handle, strength, states):
	handle.initialise(strength)
	return np.stack([handle.f(0.3, state) for state in states])


@pytest.mark.parametrize("chains", [1, 2, 3, 4])
def test_dense_interleaved_sub_ensembles(chains):
	"""large ensembles run as several sub-ensembles whose launch sequences overlap on separate streams; every
	trajectory is integrated exactly once by exactly one of them: results and step counts do not depend on the number
	of sub-ensembles (ragged ensemble size: the last one is shorter, and its tiles are partly empty)"""
	n, B = 48, 4096 + 1000 + 37
	system = W.kuramoto(n)
	Wm = system["c"] / (n - 1) * system["A"].T.astype(float)
	np.fill_diagonal(Wm, 0.0)
	Y0 = torch.as_tensor(W.kuramoto_ensemble(B, n, seed=3)).cuda()
	outcome = []
	for k in (chains, 1):
		ODE = jitcode.from_sine_coupling(Wm, system["omega"])
		ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-9, interleave=k)
		ODE.set_initial_value(Y0, 0.0)
		first = ODE.integrate(1.5)
		second = ODE.integrate(4.0)
		outcome.append((first.cpu().numpy(), second.cpu().numpy(), ODE.step_counts()))
	assert np.array_equal(outcome[0][0], outcome[1][0]) and np.array_equal(outcome[0][1], outcome[1][1])
	assert outcome[0][2] == outcome[1][2]
