"""
jitcode_restricted_lyap and jitcode_transversal_lyap on the GPU, compared with each other like the reference's
tests/test_transversal_lyap.py:26-144 does — with the five couplings of that test as ONE ensemble (the coupling
strength is a per-trajectory control parameter) and all 999 `integrate` calls as one launch.
"""

import numpy as np
import pytest
import sympy
from numpy.testing import assert_allclose
from scipy.stats import sem

import workloads as W
from jitcode_b200 import jitcode, jitcode_lyap, jitcode_restricted_lyap, jitcode_transversal_lyap, y

pytestmark = pytest.mark.gpu

a, b, c = -0.025794, 0.01, 0.02
k = sympy.Symbol("k")

SCENARIOS = {
	"grouped by variables": {
		"f": [
			y(0)*(a-y(0))*(y(0)-1.0) - y(3) + k*(y(1)-y(0)),
			y(1)*(a-y(1))*(y(1)-1.0) - y(4) + k*(y(2)-y(1)),
			y(2)*(a-y(2))*(y(2)-1.0) - y(5) + k*(y(0)-y(2)),
			b*y(0) - c*y(3),
			b*y(1) - c*y(4),
			b*y(2) - c*y(5),
		],
		"vectors": [[1., 1., 1., 0., 0., 0.], [0., 0., 0., 1., 1., 1.]],
		"groups": ([0, 1, 2], [3, 4, 5]),
	},
	"grouped by oscillators": {
		"f": [
			y(0)*(a-y(0))*(y(0)-1.0) - y(1) + k*(y(2)-y(0)),
			b*y(0) - c*y(1),
			y(2)*(a-y(2))*(y(2)-1.0) - y(3) + k*(y(4)-y(2)),
			b*y(2) - c*y(3),
			y(4)*(a-y(4))*(y(4)-1.0) - y(5) + k*(y(0)-y(4)),
			b*y(4) - c*y(5),
		],
		"vectors": [[1., 0., 1., 0., 1., 0.], [0., 1., 0., 1., 0., 1.]],
		"groups": ([0, 2, 4], [1, 3, 5]),
	},
}
COUPLINGS = [(-0.1, 1), (-0.2, 1), (0.1, -1), (0.2, -1), (0.0, 0)]


@pytest.mark.parametrize("average_dynamics", [False, True])
@pytest.mark.parametrize("name", list(SCENARIOS))
def test_restricted_and_transversal_agree(name, average_dynamics):
	scenario = SCENARIOS[name]
	n = len(scenario["f"])
	rng = np.random.default_rng(42)
	ks = np.array([coupling for coupling, _ in COUPLINGS])

	ODE1 = jitcode_restricted_lyap(scenario["f"], vectors=scenario["vectors"], verbose=False, control_pars=[k], seed=1)
	ODE1.set_integrator("dopri5")
	ODE2 = jitcode_transversal_lyap(scenario["f"], groups=scenario["groups"], average_dynamics=average_dynamics,
			verbose=False, control_pars=[k], seed=2)
	ODE2.set_integrator("dopri5")
	ODE1.set_parameters(ks)
	ODE2.set_parameters(ks)

	initial = np.empty((len(COUPLINGS), n))
	for row, (_, sign) in enumerate(COUPLINGS):
		if sign < 0:
			initial[row] = rng.random(n)
		else:       # start on the synchronisation manifold
			single = rng.random(2)
			for j, group in enumerate(scenario["groups"]):
				initial[row, list(group)] = single[j]
	ODE1.set_initial_value(initial, 0.0)
	ODE2.set_initial_value(rng.random((len(COUPLINGS), 2)), 0.0)

	times = np.arange(100, 100000, 100)
	states1, lyaps1 = ODE1.integrate_many(times)
	states2, lyaps2 = ODE2.integrate_many(times)
	assert lyaps1.shape == (len(times), len(COUPLINGS), 1) and lyaps2.shape == (len(times), len(COUPLINGS))
	assert states2.shape == (len(times), len(COUPLINGS), 2)

	checked = 0
	for row, (coupling, sign) in enumerate(COUPLINGS):
		final = states1[-1, row]
		on_manifold = all(final[i] == final[j] for group in scenario["groups"] for i in group for j in group)
		if not on_manifold:
			# the reference's test has the same escape hatch (:124-134): it only works where the compiled arithmetic
			# happens to be exactly symmetric between the synchronised units
			continue
		L1, L2 = lyaps1[500:, row, 0], lyaps2[500:, row]
		mean1, mean2 = L1.mean(), L2.mean()
		margin1, margin2 = sem(L1), sem(L2)
		sign1 = np.sign(mean1) if abs(mean1) > margin1 else 0
		sign2 = np.sign(mean2) if abs(mean2) > margin2 else 0
		assert sign1 == sign and sign2 == sign, (coupling, mean1, margin1, mean2, margin2)
		assert abs(mean1 - mean2) < max(margin1, margin2, 1e-4), (coupling, mean1, margin1, mean2, margin2)
		checked += 1
	assert checked >= 2, "the dynamics left the synchronisation manifold for almost every coupling"


# ------------------------------------------------------------------ warp storage: Gram–Schmidt as shuffle reductions

def _with_tangents(states, n_lyap, seed):
	rng = np.random.default_rng(seed)
	B, n = states.shape
	tangents = rng.normal(size=(B, n_lyap, n))
	tangents /= np.linalg.norm(tangents, axis=2, keepdims=True)
	return np.hstack([states, tangents.reshape(B, n_lyap * n)])


@pytest.mark.parametrize("name", ["dopri5", "RK45"])
def test_lyapunov_qr_in_warp_storage_small(name):
	"""the FHN pair with 4 tangent vectors (n = 20) forced into warp storage against the thread-per-trajectory kernel"""
	B = 70
	full = _with_tangents(W.FHN_Y0 + 1e-3 * np.random.default_rng(1).normal(size=(B, 4)), 4, 2)
	results = {}
	for storage in ("warp", "global"):
		ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False)
		ODE.set_integrator(name, storage=storage)
		jitcode.set_initial_value(ODE, full, 0.0)
		results[storage] = [ODE.integrate(T) for T in (5.0, 10.0)]
		assert ODE.integrator.module.info.storage == {"warp": 2, "global": 1}[storage]
	for (s1, l1, v1), (s2, l2, v2) in zip(results["warp"], results["global"]):
		assert_allclose(s1, s2, rtol=1e-9, atol=1e-12)
		assert_allclose(l1, l2, rtol=1e-8, atol=1e-10)
		assert_allclose(v1, v2, rtol=1e-7, atol=1e-9)


def test_lyapunov_of_a_network_uses_warp_storage():
	"""two Lyapunov exponents of a 40-oscillator Rössler network: a 360-dimensional system that vectorises (the
	tangent equations of all oscillators and of both vectors have the same shapes) and therefore runs in warp storage"""
	system = W.roessler_network(40, k_as_parameter=False)
	B = 24
	Y0, _ = W.roessler_ensemble(B, N=40, seed=4)
	full = _with_tangents(Y0, 2, 5)
	results = {}
	for storage in (None, "global"):
		ODE = jitcode_lyap(system["f_sym"], n_lyap=2, helpers=system["helpers"], n=120, verbose=False)
		ODE.set_integrator("dopri5", storage=storage, rtol=1e-6, atol=1e-12)
		jitcode.set_initial_value(ODE, full, 0.0)
		results[storage] = ODE.integrate(1.0)
		if storage is None:
			assert ODE.integrator.module.info.storage == 2
	(s1, l1, v1), (s2, l2, v2) = results[None], results["global"]
	assert_allclose(s1, s2, rtol=1e-8, atol=1e-10)
	assert_allclose(l1, l2, rtol=1e-6, atol=1e-9)
	assert_allclose(np.einsum("bki,bli->bkl", v1, v1), np.broadcast_to(np.eye(2), (B, 2, 2)), atol=1e-12)
	assert_allclose(v1, v2, rtol=1e-6, atol=1e-8)


def test_restricted_and_transversal_in_warp_storage():
	scenario = SCENARIOS["grouped by oscillators"]
	ks = np.array([0.1, 0.2, -0.1])
	rng = np.random.default_rng(3)
	initial = rng.random((3, 6))
	tangent = rng.normal(size=(3, 6))
	tangent /= np.linalg.norm(tangent, axis=1, keepdims=True)
	outcomes = {}
	for storage in ("warp", "global"):
		ODE1 = jitcode_restricted_lyap(scenario["f"], vectors=scenario["vectors"], verbose=False, control_pars=[k])
		ODE1.set_integrator("dopri5", storage=storage)
		ODE1.set_parameters(ks)
		jitcode.set_initial_value(ODE1, np.hstack([initial, tangent]), 0.0)
		ODE2 = jitcode_transversal_lyap(scenario["f"], groups=scenario["groups"], verbose=False, control_pars=[k], seed=7)
		ODE2.set_integrator("dopri5", storage=storage)
		ODE2.set_parameters(ks)
		ODE2.set_initial_value(initial[:, :2], 0.0)
		outcomes[storage] = (ODE1.integrate(50.0), ODE2.integrate(50.0))
	for got, expected in zip(outcomes["warp"][0], outcomes["global"][0]):
		assert_allclose(got, expected, rtol=1e-8, atol=1e-10)
	for got, expected in zip(outcomes["warp"][1], outcomes["global"][1]):
		assert_allclose(got, expected, rtol=1e-8, atol=1e-10)


def test_non_finite_norms_warn_like_the_reference():
	"""a tangent vector that overflows between two renormalisations: the reference warns (jitcode_lyap.norms,
	_jitcode.py:747-748; restricted :944-945; transversal :890-891) and returns non-finite exponents; so does this"""
	# dy/dt = 40·y: the tangent vector grows like exp(40 t); after Δt = 11.5 its components are ≈ 1e200 — finite, so the
	# integration itself succeeds, but their squares overflow and the norm is not finite
	ODE = jitcode_lyap([40.0 * y(0)], n_lyap=1, verbose=False, seed=3)
	ODE.set_integrator("dopri5", nsteps=10**6)
	states = np.full((5, 1), 1e-300)
	ODE.set_initial_value(states, 0.0)
	state, lyaps, vectors = ODE.integrate(1.0)
	assert_allclose(lyaps, 40.0, rtol=1e-5)                       # finite interval: no warning, exponent 40
	with pytest.warns(UserWarning, match="out of numerical bounds"):
		state, lyaps, vectors = ODE.integrate(12.5)
	assert not np.any(np.isfinite(lyaps))
	ODE = jitcode_restricted_lyap([40.0 * y(0), -y(1)], vectors=[[0.0, 1.0]], verbose=False, seed=3)
	ODE.set_integrator("dopri5", nsteps=10**6)
	ODE.set_initial_value(np.full((3, 2), 1e-300), 0.0)
	with pytest.warns(UserWarning, match="Norm of perturbation vector"):
		ODE.integrate_many([11.5, 12.0])


def test_lyapunov_call_that_does_not_advance_time():
	"""integrate(t_now) on the ode back end launches nothing; the reference's ln(1)/0 is NaN (with a warning)"""
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=2, verbose=False, seed=5)
	ODE.set_integrator("dopri5")
	ODE.set_initial_value(np.tile(W.FHN_Y0, (3, 1)), 0.0)
	with pytest.warns(UserWarning, match="equals the current time"):
		state, lyaps, vectors = ODE.integrate(0.0)
	assert state.shape == (3, 4) and lyaps.shape == (3, 2) and vectors.shape == (3, 2, 4)
	assert np.all(np.isnan(lyaps))
	assert_allclose(state, np.tile(W.FHN_Y0, (3, 1)), rtol=0, atol=0)
	state, lyaps, vectors = ODE.integrate(5.0)
	assert np.all(np.isfinite(lyaps))


def _tangent_ensemble(B, n_lyap, seed):
	rng = np.random.default_rng(seed)
	base = W.FHN_Y0 + 1e-3 * rng.normal(size=(B, 4))
	tangents = rng.normal(size=(B, n_lyap, 4))
	tangents /= np.linalg.norm(tangents, axis=2, keepdims=True)
	return np.hstack([base, tangents.reshape(B, 4 * n_lyap)])


@pytest.mark.parametrize("n_lyap", [2, 3, 4])
@pytest.mark.parametrize("name", ["dopri5", "RK45"])
def test_quad_storage_equals_thread_storage(n_lyap, name):
	"""quad storage (a group of lanes per trajectory, lane j integrates the state and tangent vector j in registers, norms
	and Gram–Schmidt by shuffles within the group; n_lyap = 3: one idle lane per group) against the
	thread-per-trajectory kernel with the whole system in shared memory: the same arithmetic up to the order of the sums
	over components.  B = 101: groups of the last warp leave early."""
	B, times = 101, [10.0, 20.0, 25.0]
	Y0 = _tangent_ensemble(B, n_lyap, seed=n_lyap)
	atol = np.r_[np.full(4, 1e-12), 10.0 ** np.random.default_rng(1).uniform(-12, -10, 4 * n_lyap)]
	outcome = {}
	for storage in ("quad", "shared"):
		ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=n_lyap, verbose=False)
		ODE.set_integrator(name, storage=storage, rtol=1e-6, atol=atol)
		assert ODE.integrator.module.info.storage == {"quad": 5, "shared": 3}[storage]
		jitcode.set_initial_value(ODE, Y0, 0.0)
		states, lyaps = ODE.integrate_many(times)
		tangents = ODE.y[:, 4:].reshape(B, n_lyap, 4)
		outcome[storage] = (states, lyaps, tangents, ODE.step_counts())
		assert_allclose(np.einsum("bki,bli->bkl", tangents, tangents), np.broadcast_to(np.eye(n_lyap), (B, n_lyap, n_lyap)), atol=1e-12)
		# f of the full system through the module's own evaluation kernel
		full = np.hstack([states[-1], tangents.reshape(B, -1)])
		outcome[storage] += (ODE.f(25.0, full),)
	quad, shared = outcome["quad"], outcome["shared"]
	assert_allclose(quad[0], shared[0], rtol=1e-9, atol=1e-12)
	assert_allclose(quad[1], shared[1], rtol=1e-7, atol=1e-9)
	assert_allclose(quad[2], shared[2], rtol=0, atol=1e-7)
	assert abs(sum(quad[3]) - sum(shared[3])) <= 0.01 * sum(shared[3])
	assert_allclose(quad[4], shared[4], rtol=1e-10, atol=1e-14)


def test_quad_storage_is_the_default_for_small_lyapunov_systems():
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=4, verbose=False, seed=1)
	ODE.set_integrator("dopri5")
	assert ODE.integrator.module.info.storage == 5
	# 8th-order pair: 16 stage rows × 8 components do not fit a thread's registers
	ODE.set_integrator("dop853")
	assert ODE.integrator.module.info.storage != 5
	# accumulated exponents (lyap_acc) of a quad launch equal the mean of the recorded ones
	ODE.set_integrator("dopri5")
	ODE.set_initial_value(np.tile(W.FHN_Y0, (37, 1)), 0.0)
	ODE.integrator.lyap_acc.zero_()
	states, lyaps = ODE.integrate_many([10.0 * k for k in range(1, 9)], discard=2)
	from jitcode_b200._integrator import padded
	acc = ODE.integrator._unpack(ODE.integrator.lyap_acc[:4 * padded(37)], 4).cpu().numpy()
	assert_allclose(acc, lyaps[2:].sum(axis=0), rtol=1e-12, atol=1e-14)
