"""
CPU-side tests of the product (no GPU): symbolic front-end, CUDA code generator, run-time build + C ABI
export table, and the error behaviour of the reference's API (reference tests/test_jitcode.py:164-205,
tests/scenarios.py:48-110 input styles).
"""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import sympy

import workloads as W
from jitcode_b200 import _build, _codegen, _symbolic, jitcode, jitcode_lyap, jitcode_transversal_lyap, t, y
from jitcode_b200._transversal import GroupHandler

ROOT = Path(__file__).resolve().parents[1]
a, b1, b2, c, k = (W.FHN[x] for x in ("a", "b1", "b2", "c", "k"))


def fhn_list():
	return W.fhn()["f_sym"]

def fhn_generator():
	yield from fhn_list()

def fhn_dict():
	return {y(i): entry for i, entry in reversed(list(enumerate(fhn_list())))}

def fhn_with_helpers():
	# reference tests/scenarios.py:70-91: shuffled helpers that depend on each other
	p, q, r, s = sympy.symbols("p q r s")
	helpers = [
		(r, y(2) * q),
		(p, y(0) * (a - y(0)) * (y(0) - 1.0)),
		(q, (a - y(2)) * (y(2) - 1.0)),
		(s, k * (y(2) - y(0))),
	]
	f = [p - y(1) + s, b1*y(0) - c*y(1), r - y(3) - s, b2*y(2) - c*y(3)]
	return f, helpers


def evaluate(f, helpers, state, time=0.0):
	"""plain SymPy evaluation of what the generator would print."""
	subs = {y(i): v for i, v in enumerate(state)}
	subs[t] = time
	values = {}
	for symbol, definition in _symbolic.sort_helpers(helpers):
		values[symbol] = float(definition.subs(values).subs(subs))
	return np.array([float(sympy.sympify(entry).subs(values).subs(subs)) for entry in f])


def evaluate_large(f, helpers, state, time=0.0):
	"""the same for large systems: exact replacement of the atoms instead of the general substitution."""
	values = {y(i): sympy.Float(v) for i, v in enumerate(state)}
	values[t] = sympy.Float(time)
	for symbol, definition in _symbolic.sort_helpers(helpers):
		values[symbol] = sympy.Float(float(definition.xreplace(values)))
	return np.array([float(sympy.sympify(entry).xreplace(values)) for entry in f])


@pytest.mark.parametrize("style", [fhn_list, lambda: fhn_generator, fhn_dict])
def test_input_styles(style):
	f, n = _symbolic.handle_input(style())
	assert n == 4
	np.testing.assert_allclose(evaluate(f, [], W.FHN_Y0), W.FHN_F_OF_Y0, rtol=1e-5)


def test_helpers_sorted_and_equivalent():
	f, helpers = fhn_with_helpers()
	ordered = _symbolic.sort_helpers(helpers)
	seen = set()
	symbols = {s for s, _ in ordered}
	for symbol, definition in ordered:
		assert not ((definition.free_symbols & symbols) - seen)
		seen.add(symbol)
	np.testing.assert_allclose(evaluate(f, helpers, W.FHN_Y0), W.FHN_F_OF_Y0, rtol=1e-5)
	with pytest.raises(ValueError):
		p, q = sympy.symbols("p q")
		_symbolic.sort_helpers([(p, q + 1), (q, p + 1)])


def test_jacobian_golden_with_and_without_helpers():
	# reference tests/scenarios.py:24-29, tests/test_generation.py:57-59,66-83
	subs = {y(i): v for i, v in enumerate(W.FHN_Y0)}
	plain = sympy.Matrix(list(_symbolic.jacobian_rows(_symbolic.handle_input(fhn_list())[0], [], 4))).subs(subs)
	np.testing.assert_allclose(np.array(plain, dtype=float), W.FHN_JAC_OF_Y0, rtol=1e-5)
	f, helpers = fhn_with_helpers()
	helpers = _symbolic.sort_helpers(helpers)
	rows = list(_symbolic.jacobian_rows(_symbolic.handle_input(f)[0], helpers, 4))
	values = {}
	for symbol, definition in helpers:
		values[symbol] = definition.subs(values)
	with_helpers = sympy.Matrix(rows).subs(values).subs(subs)
	np.testing.assert_allclose(np.array(with_helpers, dtype=float), W.FHN_JAC_OF_Y0, rtol=1e-5)


def test_lyapunov_system_matches_golden(golden):
	# the 20-dimensional system of jitcode_lyap evaluated at the reference's test point (tests/test_generation.py:85-97)
	g = golden("fhn_lyap_rhs")
	f, helpers = fhn_with_helpers()
	helpers = _symbolic.sort_helpers(helpers)
	for basic, hs in ((_symbolic.handle_input(fhn_list())[0], []), (_symbolic.handle_input(f)[0], helpers)):
		full = _symbolic.lyapunov_system(basic, hs, 4, 4)
		assert len(full) == 20
		np.testing.assert_allclose(evaluate(full, hs, g["x"]), g["fx"], rtol=1e-12)


def test_symbol_provenance():
	# reference tests/test_sympy_input.py: own symbols, SymPy-made symbols, mixtures
	own = [sympy.cos(t) * y(0)]
	foreign_y, foreign_t = sympy.Function("y"), sympy.Symbol("t", real=True)
	plain_t = sympy.Symbol("t")
	for f in ([sympy.cos(foreign_t) * foreign_y(0)], [sympy.cos(plain_t) * foreign_y(0)], [sympy.cos(t) * foreign_y(0)]):
		assert _symbolic.handle_input(f)[0] == _symbolic.handle_input(own)[0]


def test_cuda_printer():
	printer = _codegen.CudaPrinter()
	x = sympy.Symbol("x")
	assert printer.doprint(y(3)**2) == "y(3)*y(3)"
	assert printer.doprint((x + 1)**3) == "(x + 1.0)*(x + 1.0)*(x + 1.0)"
	assert printer.doprint(x**-2) == "1.0/(x*x)"
	assert "pow" in printer.doprint(x**sympy.Rational(5, 3))
	assert printer.doprint(sympy.Rational(3, 4) * x) in ("(3.0/4.0)*x", "3.0/4.0*x", "0.75*x")
	assert printer.doprint(2 * y(0)) == "2.0*y(0)"


def test_statements_share_repeated_calls():
	system = W.roessler_network(20)
	f, n = _symbolic.handle_input(system["f_sym"], system["n"])
	lines = _codegen.rhs_statements(f, _symbolic.sort_helpers(system["helpers"]), system["control_pars"])
	text = "\n".join(lines)
	assert text.count("sin(t)") == 1            # sin(t) appears in 20 equations but is evaluated once
	assert len([line for line in lines if line.startswith("dy.set(")]) == n
	assert "par.v[0]" in text and "jb_helper_0" in text


@pytest.mark.parametrize("name", ["roessler", "fhn", "lotka_volterra", "kuramoto"])
def test_vectoriser_plan_matches_direct_evaluation(name):
	"""the SIMT-vectorised form (groups of equally shaped statements + tables + loops) computes the same f."""
	from jitcode_b200 import _vectorise
	system, pars = {
		"roessler": (W.roessler_network(24), [0.013]), "fhn": (W.fhn(), []),
		"lotka_volterra": (W.lotka_volterra(gamma_as_parameter=True), [0.55]), "kuramoto": (W.kuramoto(24), []),
	}[name]
	f, n = _symbolic.handle_input(system["f_sym"], system["n"])
	helpers = _symbolic.sort_helpers(system["helpers"])
	plan = _vectorise.make_plan(f, helpers, system["control_pars"])
	state = np.random.default_rng(3).random(n)
	subs = dict(zip(system["control_pars"], pars))
	expected = evaluate([entry.subs(subs) for entry in f], [(s, d.subs(subs)) for s, d in helpers], state, time=0.7)
	np.testing.assert_allclose(_vectorise.evaluate_plan(plan, 0.7, state, pars), expected, rtol=1e-13, atol=1e-15)
	lines, tables, component_of = _vectorise.warp_statements(plan)
	assert sorted(component_of) == list(range(n))
	assert sum(line.lstrip().startswith("dy.set(") for line in lines) == len(plan.groups)
	if name == "roessler":
		assert [group.size for group in plan.groups] == [24, 24, 24] and plan.efficiency == 0.75
		assert len(plan.helpers) == 1 and plan.helpers[0][1].loops[0].start == [0, 24]


@pytest.mark.parametrize("name,team_warps", [("roessler100", 1), ("roessler100", 2), ("roessler50", 1), ("roessler24", 1), ("kuramoto40", 1), ("roessler200", 1), ("roessler40lyap", 1)])
def test_printed_vectorised_code_on_host_lanes(name, team_warps, tmp_path):
	"""the code printed for the GPU (blocked ring order with register windows, diagonals, pair tables, permuted
	components), compiled for the host and run lane by lane, equals direct evaluation of the symbolic input."""
	from jitcode_b200 import _vectorise
	from host_lanes import evaluate_printed
	if name.startswith("roessler"):
		system, pars = W.roessler_network(int(name[8:].replace("lyap", ""))), [0.013]
	else:
		system, pars = W.kuramoto(int(name[8:])), []
	f, n = _symbolic.handle_input(system["f_sym"], system["n"])
	helpers = _symbolic.sort_helpers(system["helpers"])
	if name.endswith("lyap"):       # tangent equations: neighbour sums with a coefficient
		f, n = _symbolic.lyapunov_system(f, helpers, n, 1), 2 * n
	plan = _vectorise.make_plan(f, helpers, system["control_pars"])
	blockings = _vectorise.ring_blockings(plan, 32 * team_warps)
	assert blockings == {("roessler40lyap", 1): {40: (20, 2)}, ("roessler100", 1): {100: (25, 4)}, ("roessler100", 2): {100: (50, 2)}, ("roessler50", 1): {50: (25, 2)}, ("roessler200", 1): {200: (25, 8)}}.get((name, team_warps), {})
	state = np.random.default_rng(5).random(n)
	subs = dict(zip(system["control_pars"], pars))
	expected = evaluate_large([entry.xreplace(subs) for entry in f], [(s, d.xreplace(subs)) for s, d in helpers], state, time=0.7)
	for blocked in (True, False):
		lines, tables, component_of = _vectorise.warp_statements(plan, team_warps, blocked_rings=blocked)
		assert ("switch (pass)" in "\n".join(lines)) == (blocked and bool(blockings))
		got = evaluate_printed(lines, tables, component_of, len(pars), 0.7, state, pars, 32 * team_warps, 1 + len(helpers), tmp_path)
		np.testing.assert_allclose(got, expected, rtol=1e-12, atol=1e-14)


def test_stragglers_join_the_large_groups(tmp_path):
	"""a node with one or two neighbours has no run of equally shaped terms; read as a short loop its equation has the
	shape of all the others and joins their group (the random graph of bench.py: 98 + 2 statements became 100)"""
	from jitcode_b200 import _vectorise
	from host_lanes import evaluate_printed
	system = W.roessler_network(100, topology="random")
	f, n = _symbolic.handle_input(system["f_sym"], system["n"])
	helpers = _symbolic.sort_helpers(system["helpers"])
	plan = _vectorise.make_plan(f, helpers, system["control_pars"])
	assert [group.size for group in plan.groups] == [100, 100, 100]
	lengths = [plan.groups[0].loops[0].start[m+1] - plan.groups[0].loops[0].start[m] for m in range(100)]
	assert min(lengths) < _vectorise.LOOP_MIN <= max(lengths)
	state = np.random.default_rng(6).random(n)
	subs = dict(zip(system["control_pars"], [0.013]))
	expected = evaluate_large([entry.xreplace(subs) for entry in f], [(s, d.xreplace(subs)) for s, d in helpers], state, time=0.4)
	np.testing.assert_allclose(_vectorise.evaluate_plan(plan, 0.4, state, [0.013]), expected, rtol=1e-12, atol=1e-14)
	lines, tables, component_of = _vectorise.warp_statements(plan)
	got = evaluate_printed(lines, tables, component_of, 1, 0.4, state, [0.013], 32, 2, tmp_path)
	np.testing.assert_allclose(got, expected, rtol=1e-12, atol=1e-14)
	# a tiny system in which nothing is "large" stays as it was
	small = _vectorise.make_plan([y(1) + y(2), y(0) * 2.0, y(0) + y(1) + y(2) + y(0)**2], [], [])
	assert len(small.groups) == 3


@pytest.mark.parametrize("rewiring,expected", [(0.0, "complete"), (0.1, "subtract"), (0.3, "masks")])
def test_printed_ring_variants_on_host_lanes(rewiring, expected, tmp_path):
	"""the three ways a blocked ring is summed: complete ring (no masks, no corrections), nearly complete (all
	neighbours summed, absent ones subtracted through the signed pair table), ragged (a mask bit per neighbour)"""
	from jitcode_b200 import _vectorise
	from host_lanes import evaluate_printed
	n = 64
	A = W.small_world_network(np.random.default_rng(11), n, 8, rewiring)
	kappa = sympy.Symbol("kappa")
	f = [-0.3 * y(i) + kappa * sympy.Add(*[sympy.sin(y(j)) for j in range(n) if A[i, j]]) for i in range(n)]
	plan = _vectorise.make_plan(f, [], [kappa])
	assert _vectorise.ring_blockings(plan, 32) == {64: (32, 2)}
	lines, tables, component_of = _vectorise.warp_statements(plan)
	text = "\n".join(lines)
	assert ("flip_sign" in text, "_present0" in text) == {"complete": (False, False), "subtract": (True, False), "masks": (False, True)}[expected]
	assert text.count("sin(") < 40        # the sine of a neighbour is taken once per window element, not once per use
	state = np.random.default_rng(5).random(n)
	direct = evaluate_large([entry.xreplace({kappa: 0.7}) for entry in f], [], state)
	got = evaluate_printed(lines, tables, component_of, 1, 0.0, state, [0.7], 32, 1, tmp_path)
	np.testing.assert_allclose(got, direct, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("n,neighbours,rewiring,team_warps,seed", [(36, 4, 0.1, 1, 1), (128, 12, 0.2, 1, 4), (250, 20, 0.1, 2, 5), (100, 40, 0.1, 1, 9)])
def test_printed_code_on_varied_networks(n, neighbours, rewiring, team_warps, seed, tmp_path):
	"""two equations per node (a weighted neighbour sum; a mean field through a helper) on small-world rings of various
	sizes, degrees and raggedness, single warps and teams: printed code on host lanes ≡ the plan's interpreter"""
	from jitcode_b200 import _vectorise
	from host_lanes import evaluate_printed
	A = W.small_world_network(np.random.default_rng(seed), n, neighbours, rewiring)
	kappa, mean = sympy.symbols("kappa mean")
	helpers = [(mean, sympy.Add(*[y(2*j + 1) for j in range(n)]) / n)]
	f = []
	for i in range(n):
		f.append(-0.3 * y(2*i) + y(2*i + 1) + kappa * sympy.Add(*[1.5 * y(2*j) for j in range(n) if A[i, j]]))
		f.append(0.1 * y(2*i) - 0.2 * y(2*i + 1) * mean)
	f, dimension = _symbolic.handle_input(f, 2 * n)
	plan = _vectorise.make_plan(f, _symbolic.sort_helpers(helpers), [kappa])
	assert n in _vectorise.ring_blockings(plan, 32 * team_warps)
	assert _vectorise._ring_split(33, 32) is None and _vectorise._ring_split(101, 32) is None     # 11 lanes × 3; a prime
	lines, tables, component_of = _vectorise.warp_statements(plan, team_warps)
	state = np.random.default_rng(seed).random(dimension)
	expected = _vectorise.evaluate_plan(plan, 0.3, state, [0.7])
	got = evaluate_printed(lines, tables, component_of, 1, 0.3, state, [0.7], 32 * team_warps, 2, tmp_path)
	np.testing.assert_allclose(got, expected, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("n,neighbours,rewiring,idle", [(100, 20, 0.1, 7), (104, 20, 0.1, 6), (112, 16, 0.12, 4), (120, 20, 0.1, 2), (124, 20, 0.1, 1)])
def test_idle_lanes_help_with_the_gathers(n, neighbours, rewiring, idle, tmp_path):
	"""blocked rings that leave lanes of the warp idle: what does not fit the first slots of an over-long statement is
	gathered by an idle lane and fetched with a shuffle (_helper_plan); with a single idle lane there is nobody to help
	(its empty row serves the statements that fit).  Printed code on host lanes (shuffles emulated round by round) ≡ the
	plan's interpreter; the mean field's sum is printed as butterfly steps inside the window when all lanes get there."""
	from jitcode_b200 import _vectorise
	from host_lanes import evaluate_printed
	A = W.small_world_network(np.random.default_rng(n), n, neighbours, rewiring)
	kappa, mean = sympy.symbols("kappa mean")
	helpers = [(mean, sympy.Add(*[y(2*j + 1) for j in range(n)]) / n)]
	f = []
	for i in range(n):
		f.append(-0.3 * y(2*i) + y(2*i + 1) + kappa * sympy.Add(*[y(2*j) - y(2*i) for j in range(n) if A[i, j]]))
		f.append(0.1 * y(2*i) - 0.2 * y(2*i + 1) * mean)
	f, dimension = _symbolic.handle_input(f, 2 * n)
	plan = _vectorise.make_plan(f, _symbolic.sort_helpers(helpers), [kappa])
	assert _vectorise.ring_blockings(plan, 32) == {n: (32 - idle, 4)}
	lines, tables, component_of = _vectorise.warp_statements(plan)
	text = "\n".join(lines)
	assert ("jb::shfl(" in text) == (idle > 1), "helpers where there are idle lanes to spare"
	assert ("butterfly<4>" in text) and (("if (lane < " + str(32 - idle) + ") {") in text) == (idle <= 1)
	state = np.random.default_rng(n + 1).random(dimension)
	expected = _vectorise.evaluate_plan(plan, 0.3, state, [0.7])
	got = evaluate_printed(lines, tables, component_of, 1, 0.3, state, [0.7], 32, 2, tmp_path)
	np.testing.assert_allclose(got, expected, rtol=1e-12, atol=1e-13)


def test_warp_configuration_rule():
	"""cached rows / threads / team size of a warp-storage module follow the measured rule"""
	from jitcode_b200._jitcode import _warp_configuration
	ode, dense = _codegen.JB_DRV_ODE, _codegen.JB_DRV_IVP_DENSE
	assert _warp_configuration(300, _codegen.JB_DP5, ode, 3000) == (3, 384, 1)        # 12 resident warps suffice: more rows in registers
	assert _warp_configuration(150, _codegen.JB_DP5, ode, 3000) == (7, 384, 1)
	assert _warp_configuration(600, _codegen.JB_DP5, ode, 3000) == (3, 256, 1)        # shared memory limits: the most trajectories first
	assert _warp_configuration(1500, _codegen.JB_DP5, ode, 3000)[2] == 8              # teams: one trajectory per 8 warps
	kregs, threads, team = _warp_configuration(300, _codegen.JB_DP5, dense, 3000, kregs=0)
	assert kregs == 0 and threads % 32 == 0 and team == 1
	for n in (64, 150, 300, 600, 1500):
		for method in (_codegen.JB_DP5, _codegen.JB_DOP853, _codegen.JB_RK23):
			kregs, threads, team = _warp_configuration(n, method, ode, 3000)
			assert threads % (32 * team) == 0 and threads <= 1024 and kregs >= 0


def test_warp_module_builds():
	system = W.roessler_network(100)
	ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=300, control_pars=system["control_pars"], verbose=False)
	module = ODE.compile_cuda()
	assert module.info.storage == _codegen.JB_STORE_WARP and module.info.work_vectors == 0
	assert module.info.threads_per_block % 32 == 0 and 0 < module.info.shared_bytes <= 227 * 1024
	thread_module = ODE.compile_cuda(storage="global")
	assert thread_module.info.storage == _codegen.JB_STORE_GLOBAL and thread_module.info.work_vectors == 8
	# small Lyapunov systems: a group of lanes per trajectory, lane j with the state and tangent vector j in registers
	lyap_ode = jitcode_lyap(fhn_list(), n_lyap=4, verbose=False)
	lyap = lyap_ode.compile_cuda()
	assert lyap.info.storage == _codegen.JB_STORE_QUAD and lyap.info.work_vectors == 0 and lyap.info.n == 20 and lyap.info.n_lyap == 4
	assert lyap_ode._quad_group(_codegen.JB_DP5, _codegen.JB_DRV_ODE) == 4
	assert jitcode_lyap(fhn_list(), n_lyap=3, verbose=False)._quad_group(_codegen.JB_DP5, _codegen.JB_DRV_ODE) == 4      # padded to a power of two
	assert jitcode_lyap(fhn_list(), n_lyap=1, verbose=False)._quad_group(_codegen.JB_DP5, _codegen.JB_DRV_ODE) == 0      # one vector: plain registers
	assert lyap_ode._quad_group(_codegen.JB_DOP853, _codegen.JB_DRV_ODE) == 0                                           # 16 rows x 8 do not fit
	# … and with the whole system per thread in shared memory when asked for
	shared = lyap_ode.compile_cuda(storage="shared")
	assert shared.info.storage == _codegen.JB_STORE_SHARED and shared.info.work_vectors == 0 and shared.info.shared_bytes == 9 * 20 * 160 * 8


def test_module_builds_loads_and_exports_every_symbol():
	"""the C ABI: every function declared in include/jitcode_b200.h is exported by a compiled module."""
	header = (ROOT / "include" / "jitcode_b200.h").read_text()
	declared = set(re.findall(r"\b(jb_[a-z_]+)\s*\(", header))
	from jitcode_b200 import _dense
	assert declared == set(_build.EXPORTS) | set(_dense.DENSE_EXPORTS)
	dense = _dense.dense_library()                    # the dense sine-coupling library: built once, loaded without a GPU
	for name in _dense.DENSE_EXPORTS:
		assert hasattr(dense, name), name
	assert dense.jb_dense_work_vectors() == 30 and dense.jb_dense_control_scalars() == 19 and dense.jb_dense_result_vector() == 29
	declared -= set(_dense.DENSE_EXPORTS)
	ODE = jitcode(W.lotka_volterra(gamma_as_parameter=True)["f_sym"], n=2, control_pars=W.lotka_volterra(gamma_as_parameter=True)["control_pars"], verbose=False)
	module = ODE.compile_cuda()
	lib = ctypes.CDLL(module.path)
	for name in declared:
		assert hasattr(lib, name), name
	assert module.info.abi_version == _build.JB_ABI_VERSION
	assert (module.info.n, module.info.n_params, module.info.n_basic, module.info.n_lyap) == (2, 1, 2, 0)
	assert module.info.storage == _codegen.JB_STORE_REGISTERS and module.info.work_vectors == 0
	assert ctypes.sizeof(_build.IntegrateArgs) % 8 == 0
	# argument validation happens before any CUDA call
	assert lib.jb_integrate(None) == -2
	assert b"ABI" in ctypes.cast(lib.jb_last_error, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_dense_struct_layout_and_detection():
	import subprocess, tempfile
	from jitcode_b200 import _dense
	probe = r'''
#include <stdio.h>
#include <stddef.h>
#include "jitcode_b200.h"
int main(void) {
	printf("%zu %zu %zu %zu %zu\n", sizeof(jb_dense_args), offsetof(jb_dense_args, W), offsetof(jb_dense_args, t), offsetof(jb_dense_args, host_flag), offsetof(jb_dense_args, stream));
	return 0;
}'''
	with tempfile.TemporaryDirectory() as tmp:
		src = Path(tmp) / "probe.c"
		src.write_text(probe)
		subprocess.run(["gcc", f"-I{ROOT/'include'}", str(src), "-o", str(Path(tmp)/"probe")], check=True)
		out = subprocess.run([str(Path(tmp)/"probe")], capture_output=True, text=True, check=True).stdout.split()
	A = _dense.DenseArgs
	assert [int(v) for v in out] == [ctypes.sizeof(A), A.W.offset, A.t.offset, A.host_flag.offset, A.stream.offset]
	# structure detection: the reference's written-out Kuramoto network → (W, omega); anything else → None
	system = W.kuramoto(12)
	f, _ = _symbolic.handle_input(system["f_sym"], 12)
	Wm, omega = _dense.detect_sine_coupling(f)
	expected = system["c"] / 11 * system["A"].T.astype(float)
	np.fill_diagonal(expected, 0.0)
	np.testing.assert_allclose(Wm, expected, rtol=1e-14)
	np.testing.assert_allclose(omega, system["omega"], rtol=1e-14)
	assert _dense.detect_sine_coupling(_symbolic.handle_input(fhn_list())[0]) is None
	assert _dense.detect_sine_coupling([sympy.sin(y(0) - y(1)) + y(0), sympy.sin(y(0) - y(1))]) is None


def test_struct_layout_matches_header():
	"""ctypes mirror and C struct agree (compiled probe prints sizeof/offsetof)."""
	import subprocess, tempfile
	probe = r'''
#include <stdio.h>
#include <stddef.h>
#include "jitcode_b200.h"
int main(void) {
	printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(jb_integrate_args), offsetof(jb_integrate_args, times), offsetof(jb_integrate_args, nsteps),
		offsetof(jb_integrate_args, t), offsetof(jb_integrate_args, status), offsetof(jb_integrate_args, stream), sizeof(jb_module_info));
	return 0;
}'''
	with tempfile.TemporaryDirectory() as tmp:
		src = Path(tmp) / "probe.c"
		src.write_text(probe)
		subprocess.run(["gcc", f"-I{ROOT/'include'}", str(src), "-o", str(Path(tmp)/"probe")], check=True)
		out = subprocess.run([str(Path(tmp)/"probe")], capture_output=True, text=True, check=True).stdout.split()
	A = _build.IntegrateArgs
	expected = [ctypes.sizeof(A), A.times.offset, A.nsteps.offset, A.t.offset, A.status.offset, A.stream.offset, ctypes.sizeof(_build.ModuleInfo)]
	assert [int(v) for v in out] == expected


def test_storage_choice():
	from jitcode_b200._jitcode import _storage_for
	assert _storage_for(2, _codegen.JB_DP5, _codegen.JB_DRV_ODE) == _codegen.JB_STORE_REGISTERS
	assert _storage_for(8, _codegen.JB_DP5, _codegen.JB_DRV_IVP_DENSE) == _codegen.JB_STORE_REGISTERS
	assert _storage_for(20, _codegen.JB_DP5, _codegen.JB_DRV_ODE) == _codegen.JB_STORE_SHARED      # 160 trajectories' stages per block
	assert _storage_for(48, _codegen.JB_DP5, _codegen.JB_DRV_ODE) == _codegen.JB_STORE_GLOBAL
	assert _storage_for(4, _codegen.JB_DOP853, _codegen.JB_DRV_IVP_DENSE) == _codegen.JB_STORE_REGISTERS
	assert _storage_for(300, _codegen.JB_DP5, _codegen.JB_DRV_ODE) != _codegen.JB_STORE_REGISTERS


# ------------------------------------------------------------------ error behaviour (reference tests/test_jitcode.py:164-205)

def test_duplicate_module_name():
	ODE = jitcode(fhn_list(), verbose=False)
	ODE.compile_cuda(modulename="foo_b200")
	with pytest.raises(NameError):
		jitcode(fhn_list(), verbose=False).compile_cuda(modulename="foo_b200")

def test_wrong_n():
	with pytest.raises(ValueError):
		jitcode(fhn_list(), n=5, verbose=False)

def test_dimension_mismatch():
	ODE = jitcode(fhn_list(), verbose=False)
	with pytest.raises(ValueError):
		ODE.set_initial_value(np.array([1.0, 2.0, 3.0]), 0.0)

def test_check_index_negative():
	with pytest.raises(ValueError):
		jitcode([y(0) - y(-1)], verbose=False).check()

def test_check_index_too_high():
	with pytest.raises(ValueError):
		jitcode([y(0) - y(1)], verbose=False).check()

def test_check_undefined_variable():
	with pytest.raises(ValueError):
		jitcode([y(0) * sympy.Symbol("x")], verbose=False).check()

def test_no_integrator():
	ODE = jitcode(fhn_list(), verbose=False)
	ODE.set_initial_value(W.FHN_Y0, 0.0)
	with pytest.raises(RuntimeError):
		ODE.integrate(1.0)
	with pytest.raises(RuntimeError):
		jitcode(fhn_list(), verbose=False).t

def test_implicit_methods_are_refused():
	ODE = jitcode(fhn_list(), verbose=False)
	for name in ("LSODA", "vode", "lsoda", "BDF", "Radau"):
		with pytest.raises(NotImplementedError):
			ODE.set_integrator(name)

def test_callbacks_are_refused():
	with pytest.raises(NotImplementedError):
		jitcode(fhn_list(), callback_functions=[(sympy.Function("cb"), lambda state: 0.0, 0)], verbose=False)

def test_wrong_number_of_parameters():
	system = W.lotka_volterra(gamma_as_parameter=True)
	ODE = jitcode(system["f_sym"], control_pars=system["control_pars"], verbose=False)
	with pytest.raises(ValueError):
		ODE.set_parameters(1.0, 2.0)

def test_save_compiled_and_module_location(tmp_path):
	ODE = jitcode(fhn_list(), verbose=False)
	destination = ODE.save_compiled(str(tmp_path / "saved.so"), overwrite=True)
	with pytest.raises(OSError):
		ODE.save_compiled(destination)
	loaded = jitcode(n=4, module_location=destination, verbose=False)
	assert next(iter(loaded._modules.values())).info.n == 4
	with pytest.raises(ValueError):
		jitcode(n=5, module_location=destination, verbose=False)


# ------------------------------------------------------------------ transversal bookkeeping

def test_group_handler_round_trip():
	groups = [[0, 2, 4], [1, 3]]
	handler = GroupHandler(groups)
	assert handler.main_indices == [0, 3] and handler.tangent_indices == [1, 2, 4]
	assert handler.map_to_main(4) == 0 and handler.map_to_main(3) == 3
	z = sympy.symbols("z0:5")
	back = handler.back_transform(list(z))
	# differences of consecutive members reproduce the difference coordinates; group sums vanish
	assert sympy.simplify(back[0] - back[2] - z[1]) == 0
	assert sympy.simplify(back[2] - back[4] - z[2]) == 0
	assert sympy.simplify(back[1] - back[3] - z[4]) == 0
	assert sympy.simplify(back[0] + back[2] + back[4]) == 0 and sympy.simplify(back[1] + back[3]) == 0

def test_transversal_system_dimension():
	# two identical, mutually coupled oscillators: one group per component
	f = [y(1), -y(0) + 0.1*(y(2) - y(0)), y(3), -y(2) + 0.1*(y(0) - y(2))]
	ODE = jitcode_transversal_lyap(f, groups=[[0, 2], [1, 3]], verbose=False)
	assert ODE.n == 4 and ODE.main_indices == [0, 2] and ODE.tangent_indices == [1, 3]
	module = ODE.compile_cuda()
	assert module.info.n_tangent == 2 and module.info.n_lyap == 1

def test_lyap_module_layout():
	ODE = jitcode_lyap(fhn_list(), n_lyap=2, verbose=False)
	assert ODE.n == 12 and ODE.n_basic == 4
	module = ODE.compile_cuda()
	assert (module.info.n, module.info.n_basic, module.info.n_lyap) == (12, 4, 2)
