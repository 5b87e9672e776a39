"""
Parity of the CUDA path with the reference's own integrators (GPU box, `-m gpu`).

Every test drives the product through its public API (jitcode_b200.jitcode → generated CUDA module →
C ABI) and compares with
  * the golden vectors of the reference's test-suite (workloads.FHN_*; reference tests/scenarios.py:14-38),
  * the fixtures under tests/golden/ (made by tests/golden/make_golden.py from the reference's own
    jitced_template.c + integrator_tools.py + SciPy),
  * the CPU oracle (oracle/) run on the same seeded inputs.
Tolerance for non-chaotic states: 10·(atol + rtol·|y|) (north star); chaotic ones as stated per test.
"""

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

import workloads as W
from jitcode_b200 import UnsuccessfulIntegration, jitcode, jitcode_lyap, y

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-6, 1e-12
TOL = {"rtol": RTOL, "atol": ATOL}
VARIANTS = [  # label in the golden files, integrator name, interpolate
	("dopri5", "dopri5", True), ("dop853", "dop853", True), ("RK45", "RK45", True),
	("RK45_land", "RK45", False), ("DOP853", "DOP853", True), ("DOP853_land", "DOP853", False),
	("RK23", "RK23", True), ("RK23_land", "RK23", False),
]


def within_tolerance(result, expected, factor=10.0):
	bound = factor * (ATOL + RTOL * np.abs(expected))
	excess = np.abs(result - expected) - bound
	assert np.all(excess <= 0), f"max excess over 10(atol+rtol|y|): {excess.max():.3e} at {np.unravel_index(excess.argmax(), excess.shape)}"


def make(system, **kwargs):
	return jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False, **kwargs)


# ------------------------------------------------------------------ the generated right-hand side

def test_f_of_y0_golden(golden):
	ODE = make(W.fhn())
	assert_allclose(ODE.f(0.0, W.FHN_Y0), W.FHN_F_OF_Y0, rtol=1e-5)    # reference tests/test_generation.py:57
	g = golden("fhn")
	assert_allclose(ODE.f(g["t"], g["X"]), g["fX"], rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("storage", ["warp", "global"])
def test_f_roessler_golden(golden, storage):
	g = golden("roessler_100")
	ODE = make(W.roessler_network(100))
	ODE.compile_cuda(storage=storage)
	ODE.set_parameters(g["kX"][0])
	assert_allclose(ODE.f(g["tX"], g["X"]), g["fX"], rtol=1e-12, atol=1e-15)


def test_f_lyapunov_system_golden(golden):
	g = golden("fhn_lyap_rhs")     # reference tests/test_generation.py:85-97
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False)
	assert_allclose(jitcode.f(ODE, 0.0, g["x"]), g["fx"], rtol=1e-13)


# ------------------------------------------------------------------ golden final states

@pytest.mark.parametrize("label,name,interpolate", VARIANTS)
def test_y1_golden(golden, label, name, interpolate):
	# reference tests/test_jitcode.py:93,130 and tests/test_integrator_tools.py:38-40
	g = golden("fhn")
	ODE = make(W.fhn())
	ODE.set_integrator(name, interpolate=interpolate)
	ODE.set_initial_value(W.FHN_Y0, 0.0)
	y1 = ODE.integrate(1.0)
	assert y1.shape == (4,)
	# reference tests/test_jitcode.py: rtol 1e-4 (:93), 1e-3 in the sweep over all integrators (:130), which RK23 needs
	assert_allclose(y1, W.FHN_Y1, rtol=1e-3 if name == "RK23" else 1e-4)
	assert_allclose(y1, g[f"y1_{label}"], rtol=1e-7)
	ODE.set_integrator(name, interpolate=interpolate, **TOL)
	ODE.set_initial_value(W.FHN_Y0, 0.0)
	within_tolerance(ODE.integrate(1.0), g[f"y1_{label}_tol"])


@pytest.mark.parametrize("label,name,interpolate", VARIANTS)
def test_lotka_volterra_ensemble_golden(golden, label, name, interpolate):
	g = golden("lotka_volterra")
	ODE = make(W.lotka_volterra())
	ODE.set_integrator(name, interpolate=interpolate, **TOL)
	ODE.set_initial_value(g["Y0"], 0.0)
	for i, T in enumerate(g["times"]):
		result = ODE.integrate(T)
		assert result.shape == g["Y0"].shape
		within_tolerance(result, g[label][i])
	# sampling finer than the step size (dense-output path of IVP_wrapper, integrator_tools.py:98-105)
	ODE.set_initial_value(g["Y0"][:8], 0.0)
	for i, T in enumerate(g["fine_times"]):
		within_tolerance(ODE.integrate(T), g[f"{label}_fine"][i])
	# the same loop as one launch
	ODE.set_initial_value(g["Y0"][:8], 0.0)
	within_tolerance(ODE.integrate_many(g["fine_times"]), g[f"{label}_fine"])


def test_lotka_volterra_example_c1(golden):
	# examples/lotka_volterra.py:96-116 exactly (BASELINE config 1)
	g = golden("lotka_volterra")
	ODE = make(W.lotka_volterra())
	ODE.set_integrator("dopri5")
	ODE.set_initial_value({y(0): W.LV_C1_INITIAL["R"], y(1): W.LV_C1_INITIAL["B"]}, 0.0)
	data = np.vstack([ODE.integrate(T) for T in g["c1_times"]])
	within_tolerance(data, g["c1_dopri5"])


def test_control_parameters_per_trajectory(golden):
	g = golden("lotka_volterra_gamma")
	ODE = make(W.lotka_volterra(gamma_as_parameter=True))
	for name in ("dopri5", "RK45"):
		ODE.set_integrator(name, **TOL)
		ODE.set_parameters(g["pars"][:, 0])
		ODE.set_initial_value(g["Y0"], 0.0)
		for i, T in enumerate(g["times"]):
			within_tolerance(ODE.integrate(T), g[name][i])


@pytest.mark.parametrize("storage", ["warp", "global"])
def test_roessler_network_golden(golden, storage):
	# chaotic: agreement to 1e-3 relative (of the state's scale) up to T = 50
	g = golden("roessler_100")
	ODE = make(W.roessler_network(100))
	for name in ("dopri5", "RK45"):
		ODE.set_integrator(name, storage=storage, **TOL)
		assert ODE.integrator.module.info.storage == {"warp": 2, "global": 1}[storage]
		ODE.set_parameters(g["pars"][:, 0])
		ODE.set_initial_value(g["Y0"], 0.0)
		for i, T in enumerate(g["times"]):
			result = ODE.integrate(T)
			scale = np.abs(g[name][i]).max()
			assert np.abs(result - g[name][i]).max() < 1e-3 * scale, (name, T)
			if i == 0:
				assert_allclose(result, g[name][i], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("name,interpolate", [("dopri5", True), ("RK45", True), ("RK45", False), ("dop853", True), ("DOP853", True), ("RK23", True)])
def test_warp_storage_matches_global_storage(name, interpolate):
	"""the two large-system kernels (thread per trajectory in HBM tiles / warp per trajectory in shared memory)
	run the same controller: on a short, non-chaotic horizon they agree to rounding, with equal step counts."""
	system = W.roessler_network(100)
	Y0, pars = W.roessler_ensemble(333, seed=9)          # odd size: partial tiles and a partial last block
	results, counts = {}, {}
	for storage in ("warp", "global"):
		ODE = make(system)
		ODE.set_integrator(name, interpolate=interpolate, storage=storage, **TOL)
		ODE.set_parameters(pars[:, 0])
		ODE.set_initial_value(Y0, 0.0)
		results[storage] = np.stack([ODE.integrate(T) for T in (0.5, 1.0, 3.0)])
		counts[storage] = ODE.step_counts()
	assert_allclose(results["warp"], results["global"], rtol=1e-9, atol=1e-11)
	assert abs(sum(counts["warp"]) - sum(counts["global"])) <= 0.001 * sum(counts["global"])


@pytest.mark.parametrize("team_warps", [2, 4])
def test_teams_of_warps_match_single_warps(team_warps):
	"""warp storage with several warps per trajectory (larger systems) runs the same arithmetic as one warp per trajectory"""
	system = W.roessler_network(100)
	Y0, pars = W.roessler_ensemble(97, seed=12)
	results, counts = {}, {}
	for team in (1, team_warps):
		ODE = make(system)
		ODE.set_integrator("RK45", storage="warp", team_warps=team, **TOL)
		ODE.set_parameters(pars[:, 0])
		ODE.set_initial_value(Y0, 0.0)
		results[team] = np.stack([ODE.integrate(T) for T in (0.7, 1.0, 4.0)])
		counts[team] = ODE.step_counts()
	assert_allclose(results[team_warps], results[1], rtol=1e-10, atol=1e-12)
	assert counts[team_warps] == counts[1]


def test_large_network_uses_teams():
	"""n = 1500 (the size of the reference's examples/SW_of_Roesslers.py, N = 500): a trajectory no longer fits one
	warp's share of an SM; several warps share it.  Checked against the thread-per-trajectory kernel."""
	system = W.roessler_network(500)
	Y0, pars = W.roessler_ensemble(40, N=500, seed=13)
	results = {}
	for storage in ("warp", "global"):
		ODE = make(system)
		ODE.set_integrator("dopri5", storage=storage, **TOL)
		ODE.set_parameters(pars[:, 0])
		ODE.set_initial_value(Y0, 0.0)
		results[storage] = ODE.integrate(2.0)
		if storage == "warp":
			assert ODE.integrator.module.info.threads_per_block >= 256
	assert_allclose(results["warp"], results["global"], rtol=1e-9, atol=1e-11)


def test_full_size_roessler_sweep(golden):
	"""BASELINE config 3 at its full size — 2^20 trajectories of the 300-dimensional network — through properties
	that need no reference run: the fixture's trajectories sit at the front of the ensemble and must come out as in
	the golden file; a random sample must equal what the other large-system kernel computes for it alone (every
	trajectory is independent of its position in the batch); every trajectory is finite and took a plausible number
	of steps."""
	g = golden("roessler_100")
	B = 1 << 20
	system = W.roessler_network(100)
	Y0, pars = W.roessler_ensemble(B, seed=77)
	Y0[:6], pars[:6] = g["Y0"], g["pars"]
	ODE = make(system)
	ODE.set_integrator("dopri5", **TOL)
	assert ODE.integrator.module.info.storage == 2
	ODE.set_parameters(torch.as_tensor(pars).cuda())
	ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
	result = ODE.integrate(10.0)
	assert result.shape == (B, 300) and bool(torch.isfinite(result).all())
	assert_allclose(result[:6].cpu().numpy(), g["dopri5"][0], rtol=1e-6, atol=1e-8)
	attempts = (ODE.integrator.n_accepted[:B] + ODE.integrator.n_rejected[:B]).cpu().numpy()
	assert attempts.min() > 40 and attempts.max() < 400
	sample = np.random.default_rng(5).choice(B, 300, replace=False)
	check = make(system)
	check.set_integrator("dopri5", storage="global", **TOL)
	check.set_parameters(pars[sample, 0])
	check.set_initial_value(Y0[sample], 0.0)
	assert_allclose(result[torch.as_tensor(sample).cuda()].cpu().numpy(), check.integrate(10.0), rtol=1e-8, atol=1e-10)


# ------------------------------------------------------------------ against the CPU oracle on seeded ensembles

@pytest.mark.parametrize("name,interpolate", [("dopri5", True), ("RK45", True), ("RK45", False), ("dop853", True)])
def test_lotka_volterra_vs_oracle(name, interpolate):
	from oracle import c_rhs, ensemble
	system = W.lotka_volterra()
	Y0, _ = W.lotka_volterra_ensemble(400, seed=7)
	times = [10.0, 40.0]
	handle = c_rhs.build_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"])
	ref = ensemble.run_ensemble(handle, name, Y0, times, interpolate=interpolate, **TOL)
	ODE = make(system)
	ODE.set_integrator(name, interpolate=interpolate, **TOL)
	ODE.set_initial_value(Y0, 0.0)
	for i, T in enumerate(times):
		within_tolerance(ODE.integrate(T), ref["y"][i])
	accepted, rejected = ODE.step_counts()
	# same controller → (almost) the same number of step attempts; the oracle counts right-hand-side calls
	attempts = ref["attempts"]
	assert abs((accepted + rejected) - attempts) < 0.05 * attempts


def test_device_resident_interface():
	"""CUDA tensors in → CUDA tensors out, no host copies."""
	system = W.lotka_volterra()
	Y0, _ = W.lotka_volterra_ensemble(1000, seed=3)
	ODE = make(system)
	ODE.set_integrator("dopri5", **TOL)
	ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
	result = ODE.integrate(5.0)
	assert result.is_cuda and result.shape == (1000, 2)
	ODE.set_initial_value(Y0, 0.0)
	assert_allclose(result.cpu().numpy(), ODE.integrate(5.0), rtol=0, atol=0)


def test_lotka_volterra_invariant_large_ensemble():
	"""size-independent property at a large batch: V = νB − ω ln B + φR − γ ln R is conserved (LV first integral)."""
	B = 200_000
	Y0, _ = W.lotka_volterra_ensemble(B, seed=11)
	ODE = make(W.lotka_volterra())
	ODE.set_integrator("RK45", **TOL)
	ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
	result = ODE.integrate(50.0).cpu().numpy()
	def invariant(state):
		R, Bp = state[:, 0], state[:, 1]
		return 0.5*Bp - 0.5*np.log(Bp) + 1.0*R - 0.6*np.log(R)
	assert_allclose(invariant(result), invariant(Y0), rtol=2e-4)
	accepted, rejected = ODE.step_counts()
	assert accepted > 50 * B and rejected < accepted


# ------------------------------------------------------------------ the wrappers' contract (reference tests/test_integrator_tools.py:33-69)

@pytest.mark.parametrize("name,interpolate", [("dopri5", True), ("dop853", True), ("RK45", True), ("RK45", False), ("DOP853", True), ("DOP853", False), ("RK23", True), ("RK23", False)])
def test_wrapper_contract(name, interpolate):
	rng = np.random.default_rng(42)
	ODE = make(W.fhn())
	ODE.set_integrator(name, interpolate=interpolate)
	initial = rng.random(4)
	ODE.set_initial_value(initial, 0.0)
	assert_allclose(ODE.integrate(0), initial, rtol=1e-12)      # zero-length integration
	ODE.set_initial_value(W.FHN_Y0, 0.0)                        # re-initialisation
	assert_allclose(ODE.integrate(1.0), W.FHN_Y1, rtol=1e-3 if name == "RK23" else 1e-4)
	if not (name in ("RK45", "DOP853", "RK23") and interpolate):
		with pytest.raises(ValueError):
			ODE.integrate(0.5)                                  # backwards


@pytest.mark.parametrize("name,interpolate", [("dopri5", True), ("RK45", True), ("RK45", False)])
def test_forced_failure(name, interpolate):
	# reference tests/test_integrator_tools.py:59-63: time-reversed system, tolerances 1e-12
	system = W.fhn()
	ODE = jitcode([-entry for entry in system["f_sym"]], verbose=False)
	ODE.set_integrator(name, interpolate=interpolate, atol=1e-12, rtol=1e-12)
	ODE.set_initial_value(np.array([[0.0, 1, 2, 3], [0.0, 0.1, 0.0, 0.1]]), 0.0)
	with pytest.raises(UnsuccessfulIntegration) as failure:
		ODE.integrate(1.0)
	assert 0 in failure.value.indices


# ------------------------------------------------------------------ Lyapunov exponents

def _lyap_initial(B, seed):
	rng = np.random.default_rng(seed)
	Y0 = np.empty((B, 20))
	Y0[:, :4] = W.FHN_Y0 + 1e-3*rng.normal(size=(B, 4))    # near the attractor: no exponent is lost to round-off
	tangents = rng.normal(size=(B, 4, 4))
	Y0[:, 4:] = (tangents / np.linalg.norm(tangents, axis=2, keepdims=True)).reshape(B, 16)
	return Y0


def test_lyapunov_step_vs_oracle():
	"""one jitcode_lyap.integrate call (_jitcode.py:751-775) with given tangent vectors: state, local
	exponents and the spans of the re-orthonormalised vectors agree with the oracle."""
	from oracle import c_rhs, ensemble
	B = 64
	Y0 = _lyap_initial(B, 5)
	full, helpers = c_rhs.lyapunov_system(W.fhn()["f_sym"], 4)
	handle = c_rhs.build_rhs(full, helpers)
	ref = ensemble.run_ensemble(handle, "dopri5", Y0, [10.0, 20.0], lyap=(4, 4))
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False)
	ODE.set_integrator("dopri5")
	jitcode.set_initial_value(ODE, Y0, 0.0)
	for i, T in enumerate([10.0, 20.0]):
		state, lyaps, vectors = ODE.integrate(T)
		assert state.shape == (B, 4) and lyaps.shape == (B, 4) and vectors.shape == (B, 4, 4)
		assert_allclose(state, ref["y"][i][:, :4], rtol=1e-4, atol=1e-7)
		assert_allclose(lyaps, ref["lyaps"][i], rtol=1e-4, atol=1e-6)
		expected = ref["y"][i][:, 4:].reshape(B, 4, 4)
		signs = np.sign(np.einsum("bki,bki->bk", vectors, expected))
		assert_allclose(vectors, signs[:, :, None] * expected, atol=1e-5)
		assert_allclose(np.einsum("bki,bli->bkl", vectors, vectors), np.broadcast_to(np.eye(4), (B, 4, 4)), atol=1e-12)


def test_lyapunov_spectrum_golden():
	"""reference tests/test_jitcode.py:132-162 with the ensemble replacing the long time average:
	256 trajectories × 400 calls (Δt = 10), first 100 discarded; spectrum within max(5·SEM, 0.003) of the golden one."""
	B, calls, discard = 256, 400, 100
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False, seed=1)
	ODE.set_integrator("dopri5")
	ODE.set_initial_value(np.tile(W.FHN_Y0, (B, 1)) + 1e-3*np.random.default_rng(2).normal(size=(B, 4)), 0.0)
	last, lyaps = ODE.integrate_many(np.arange(1, calls+1) * 10.0, discard=discard, record_states=False)
	assert last.shape == (1, B, 4) and lyaps.shape == (calls, B, 4)
	kept = lyaps[discard:]                       # (calls-discard, B, 4)
	mean = kept.mean(axis=(0, 1))
	per_trajectory = kept.mean(axis=0)
	sem = per_trajectory.std(axis=0, ddof=1) / np.sqrt(B)
	for k in range(4):
		assert abs(mean[k] - W.FHN_LYAPS[k]) < max(5*sem[k], 0.003), (k, mean, sem)
