"""
The reference's API behaviour on the GPU: call orders (reference tests/test_order.py, tests/test_jitcode.py
TestOrders :15-95), symbol provenance (tests/test_sympy_input.py), saved modules (tests/test_jitcode.py:77-84),
the installation self-test (`jitcode.test`, _jitcode.py:948-965).
"""

from itertools import permutations

import numpy as np
import pytest
import sympy
from numpy.testing import assert_allclose

import jitcode_b200
import jitcode_b200.sympy_symbols
import workloads as W
from jitcode_b200 import jitcode, jitcode_lyap, t, y

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("integrator", ["dopri5", "RK45"])
def test_call_orders_with_parameters(integrator):
	# reference tests/test_order.py:10-36: every order either works or says that parameters are missing
	p = sympy.Symbol("p")
	steps = {
		"set_integrator": lambda ODE: ODE.set_integrator(integrator),
		"set_parameters": lambda ODE: ODE.set_parameters(1),
		"set_initial_value": lambda ODE: ODE.set_initial_value([0], 0),
	}
	working = 0
	for order in permutations(steps):
		ODE = jitcode([1/p], control_pars=[p], verbose=False)
		try:
			for name in order:
				steps[name](ODE)
			result = ODE.integrate(1.0)
		except RuntimeError as exception:
			assert str(exception).startswith("Something needs parameters to be set")
		else:
			assert abs(result[0] - 1) < 1e-8
			working += 1
	assert working == 6      # parameters travel with every launch here: no order can miss them


@pytest.mark.parametrize("integrator", ["dopri5", "RK45"])
def test_missing_parameters_are_reported(integrator):
	p = sympy.Symbol("p")
	ODE = jitcode([1/p], control_pars=[p], verbose=False)
	ODE.set_integrator(integrator)
	ODE.set_initial_value([0], 0)
	with pytest.raises(RuntimeError, match="Something needs parameters to be set"):
		ODE.integrate(1.0)


@pytest.mark.parametrize("integrator", ["dopri5", "RK45"])
def test_call_orders_without_parameters(integrator):
	for order in permutations(["set_integrator", "set_initial_value"]):
		ODE = jitcode([1], verbose=False)
		for name in order:
			if name == "set_integrator":
				ODE.set_integrator(integrator)
			else:
				ODE.set_initial_value([0], 0)
		assert abs(ODE.integrate(1.0)[0] - 1) < 1e-8


def test_symbol_provenance_gives_identical_results():
	# reference tests/test_sympy_input.py:44-58
	variants = [
		(sympy.Symbol("t", real=True), sympy.Function("y", real=True)),
		(jitcode_b200.t, jitcode_b200.y),
		(jitcode_b200.sympy_symbols.t, jitcode_b200.sympy_symbols.y),
		(jitcode_b200.sympy_symbols.t, sympy.Function("y")),
		(sympy.Symbol("t"), jitcode_b200.y),
	]
	results = set()
	for time, state in variants:
		ODE = jitcode([sympy.cos(time) * state(0)], verbose=False)
		ODE.set_integrator("dopri5")
		ODE.set_initial_value([1.0], 0.0)
		results.add(float(ODE.integrate(10)[0]))
	assert len(results) == 1
	assert_allclose(results.pop(), np.exp(np.sin(10.0)), rtol=1e-5)


def test_generation_steps_in_any_order():
	# reference tests/test_jitcode.py TestOrders: explicit generation / compilation steps before set_integrator
	system = W.fhn()
	recipes = [
		lambda ODE: None,
		lambda ODE: ODE.generate_f_C(),
		lambda ODE: (ODE.generate_f_C(do_cse=True), ODE.compile_C()),
		lambda ODE: ODE.compile_C(extra_compile_args=["-O2"]),
		lambda ODE: (ODE.check(), ODE.generate_f_C(simplify=True)),
	]
	for recipe in recipes:
		ODE = jitcode(system["f_sym"], verbose=False)
		recipe(ODE)
		ODE.set_integrator("dopri5")
		ODE.set_initial_value(W.FHN_Y0, 0.0)
		assert_allclose(ODE.f(0.0, W.FHN_Y0), W.FHN_F_OF_Y0, rtol=1e-5)
		assert_allclose(ODE.integrate(1.0), W.FHN_Y1, rtol=1e-4)
		assert set(ODE.y_dict) == {y(i) for i in range(4)}


def test_saved_module_integrates(tmp_path):
	# reference tests/test_jitcode.py:77-84 and :144-149
	system = W.fhn()
	ODE = jitcode(system["f_sym"], verbose=False)
	ODE.set_integrator("dopri5")
	path = ODE.save_compiled(str(tmp_path / "fhn_module.so"), overwrite=True)
	loaded = jitcode(n=4, module_location=path, verbose=False)
	loaded.set_integrator("dopri5")
	loaded.set_initial_value(W.FHN_Y0, 0.0)
	assert_allclose(loaded.integrate(1.0), W.FHN_Y1, rtol=1e-4)
	with pytest.raises(RuntimeError):
		loaded.set_integrator("RK45")         # another configuration would have to be compiled, and there is no f_sym


def test_installation_self_test():
	jitcode_b200.test()


def test_lyap_single_trajectory_interface():
	# the reference's return types for one trajectory: (n,), (n_lyap,), (n_lyap, n)
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=3, verbose=False, seed=3)
	ODE.set_integrator("dopri5")
	ODE.set_initial_value(W.FHN_Y0, 0.0)
	state, lyaps, vectors = ODE.integrate(10.0)
	assert state.shape == (4,) and lyaps.shape == (3,) and vectors.shape == (3, 4)
	assert_allclose(vectors @ vectors.T, np.eye(3), atol=1e-12)
	assert set(ODE.y_dict) == {y(i) for i in range(4)}
	assert ODE.t == 10.0
