"""
Synthetic workloads of BASELINE.json / SURVEY.md §8(d), as neutral data.

This module belongs neither to the product (``jitcode_b200``) nor to the oracle
(``oracle/``): it only states the five benchmark systems symbolically (SymPy) and
draws the seeded ensembles of initial conditions / control parameters.  ``bench.py``,
``tests/`` and the oracle drivers all consume the same definitions, so that the CUDA
path and the CPU reference path are guaranteed to see identical inputs.

Sources (reference repo, read-only, cited for parity):
  * C1/C2 Lotka–Volterra ............ examples/lotka_volterra.py:96-116
  * FHN pair (+golden vectors) ...... tests/scenarios.py:12-55, examples/double_fhn_lyapunov.py:8-32
  * C3 Rössler small-world network .. examples/SW_of_Roesslers.py:37-111
  * C5 Kuramoto network ............. examples/kuramoto_network.py:7-35
"""

import numpy as np
import sympy

y = sympy.Function("y")
t = sympy.Symbol("t", real=True)


# --------------------------------------------------------------------------
# C1 / C2: Lotka–Volterra (examples/lotka_volterra.py:98-108)
# --------------------------------------------------------------------------

def lotka_volterra(gamma_as_parameter=False):
	"""y(0)=R (predator), y(1)=B (prey); γ=.6 φ=1 ω=.5 ν=.5."""
	phi, omega, nu = 1.0, 0.5, 0.5
	if gamma_as_parameter:
		gamma = sympy.Symbol("gamma")
		control_pars = [gamma]
	else:
		gamma = 0.6
		control_pars = []
	R, B = y(0), y(1)
	f = {
		B:  gamma*B - phi*R*B,
		R: -omega*R + nu*R*B,
	}
	return {"f_sym": f, "n": 2, "control_pars": control_pars, "helpers": []}

def lotka_volterra_ensemble(B, seed=42, gamma_as_parameter=False):
	"""Y0 ~ U[0.1,2.0]^2; optional γ ~ U[0.4,0.8]."""
	rng = np.random.default_rng(seed)
	Y0 = rng.uniform(0.1, 2.0, (B, 2))
	pars = rng.uniform(0.4, 0.8, (B, 1)) if gamma_as_parameter else np.empty((B, 0))
	return Y0, pars

LV_C1_INITIAL = {"R": 0.2, "B": 0.5}  # examples/lotka_volterra.py:111


# --------------------------------------------------------------------------
# FitzHugh–Nagumo pair (tests/scenarios.py:41-53) + golden vectors (:14-38)
# --------------------------------------------------------------------------

FHN = {"a": -0.025794, "b1": 0.0065, "b2": 0.0135, "c": 0.02, "k": 0.128}

def fhn(b1=None, b2=None):
	a, c, k = FHN["a"], FHN["c"], FHN["k"]
	b1 = FHN["b1"] if b1 is None else b1
	b2 = FHN["b2"] if b2 is None else b2
	f = [
		y(0) * (a-y(0)) * (y(0)-1.0) - y(1) + k * (y(2) - y(0)),
		b1*y(0) - c*y(1),
		y(2) * (a-y(2)) * (y(2)-1.0) - y(3) + k * (y(0) - y(2)),
		b2*y(2) - c*y(3),
	]
	return {"f_sym": f, "n": 4, "control_pars": [], "helpers": []}

# golden vectors of the reference's test-suite (tests/scenarios.py:14-38), 6-digit y0
FHN_Y0 = np.array([-0.00338158, -0.00223185, 0.01524253, -0.00613449])
FHN_F_OF_Y0 = np.array([0.0045396904008868, 0.00002265673, 0.0043665702488807, 0.000328463955])
FHN_JAC_OF_Y0 = np.array([
	[-0.1088290163008492, -1.0, 0.128, 0.0],
	[0.0065, -0.02, 0.0, 0.0],
	[0.128, 0.0, -0.0732042758000427, -1.0],
	[0.0, 0.0, 0.0135, -0.02],
])
FHN_Y1 = np.array([0.0011789485114731, -0.0021947158873226, 0.0195744683782066, -0.0057801623466600])
FHN_LYAPS = np.array([0.00710, 0.0, -0.0513, -0.187])

def fhn_lyap_ensemble(B, seed=42):
	"""C4: Y0 ~ (1,2,3,4)+U[-.5,.5]^4."""
	rng = np.random.default_rng(seed)
	return np.array([1.0, 2.0, 3.0, 4.0]) + rng.uniform(-0.5, 0.5, (B, 4))


# --------------------------------------------------------------------------
# C3: small-world network of Rössler oscillators (examples/SW_of_Roesslers.py)
# --------------------------------------------------------------------------

def small_world_network(rng, number_of_nodes, nearest_neighbours, rewiring_probability):
	"""Watts–Strogatz-like ring with symmetric rewiring; same construction and the same
	order of rng draws as examples/SW_of_Roesslers.py:38-60, so seed 42 gives the same graph."""
	n = number_of_nodes
	m = nearest_neighbours // 2
	A = np.zeros((n, n), dtype=bool)
	ring = np.arange(-m, m+1)
	for i in range(n):
		A[i, (i+ring) % n] = True
	for i in range(n):
		for j in range(i):
			if A[i, j] and rng.random() < rewiring_probability:
				A[j, i] = A[i, j] = False
				while True:
					i_new, j_new = rng.integers(0, n, 2)
					if not (A[i_new, j_new] or i_new == j_new):
						A[j_new, i_new] = A[i_new, j_new] = True
						break
	return A

def erdos_renyi_network(rng, number_of_nodes, mean_degree):
	"""symmetric random graph without self-loops, every pair linked with probability mean_degree/(n−1)"""
	n = number_of_nodes
	upper = np.triu(rng.random((n, n)) < mean_degree / (n - 1), k=1)
	return upper | upper.T

def roessler_network(N=100, seed=42, k_as_parameter=True, topology="small_world"):
	"""3N-dimensional system; data layout x0,v0,z0,x1,… (SW_of_Roesslers.py:110).
	Draw order as in the example (:69-84): ω first, then the adjacency matrix.
	topology: "small_world" — the example's small_world_network(N, 20, 0.1); "random" — the same construction with
	rewiring probability 1 (no lattice left); "erdos_renyi" — G(N, p) of the same mean degree 20.  The two variants
	exist to measure the engine where the lattice-specific code paths of the vectoriser do not apply."""
	rng = np.random.default_rng(seed)
	omega = rng.uniform(0.8, 1.0, N)
	a, b, c = 0.165, 0.2, 10.0
	if topology == "small_world":
		A = small_world_network(rng, N, 20, 0.1)
	elif topology == "random":
		A = small_world_network(rng, N, 20, 1.0)
	elif topology == "erdos_renyi":
		A = erdos_renyi_network(rng, N, 20)
	else:
		raise ValueError(topology)
	if k_as_parameter:
		k = sympy.Symbol("k")
		control_pars = [k]
	else:
		k = 0.01
		control_pars = []
	sum_z = sympy.Symbol("sum_z")
	helpers = [(sum_z, sympy.Add(*[y(3*j+2) for j in range(N)]))]

	def f():
		for i in range(N):
			coupling_sum = sympy.Add(*[y(3*j)-y(3*i) for j in range(N) if A[i, j]])
			coupling_term = k * sympy.sin(t) * coupling_sum
			yield -omega[i] * y(3*i+1) - y(3*i+2) + coupling_term
			yield omega[i] * y(3*i) + a*y(3*i+1)
			coupling_term_2 = k * (sum_z - N*y(3*i+2))
			yield b + y(3*i+2) * (y(3*i) - c) + coupling_term_2

	return {"f_sym": f, "n": 3*N, "control_pars": control_pars, "helpers": helpers,
			"adjacency": A, "omega": omega}

def roessler_ensemble(B, N=100, seed=43):
	"""Y0 ~ U[0,1]^{3N}, k ~ U[0,0.02]."""
	rng = np.random.default_rng(seed)
	pars = rng.uniform(0.0, 0.02, (B, 1))
	Y0 = rng.random((B, 3*N))
	return Y0, pars


# --------------------------------------------------------------------------
# C5: Kuramoto network with dense adjacency (examples/kuramoto_network.py)
# --------------------------------------------------------------------------

def kuramoto_data(n=100, q=0.2, c=3.0, seed=42):
	"""Draw order as in kuramoto_network.py:7-17: A, then ω (sorted)."""
	rng = np.random.default_rng(seed)
	A = rng.choice([1, 0], size=(n, n), p=[q, 1-q])
	omega = rng.uniform(-0.5, 0.5, n)
	omega.sort()
	return {"A": A, "omega": omega, "c": c, "n": n}

def kuramoto(n=100, q=0.2, c=3.0, seed=42):
	"""The reference's written-out form: n·deg sine terms (kuramoto_network.py:19-26)."""
	data = kuramoto_data(n, q, c, seed)
	A, omega = data["A"], data["omega"]

	def f():
		for i in range(n):
			coupling_sum = sympy.Add(*[sympy.sin(y(j)-y(i)) for j in range(n) if A[j, i]])
			yield omega[i] + c/(n-1)*coupling_sum

	return {"f_sym": f, "n": n, "control_pars": [], "helpers": [], **data}

def kuramoto_ensemble(B, n=100, seed=44):
	rng = np.random.default_rng(seed)
	return rng.uniform(0, 2*np.pi, (B, n))
