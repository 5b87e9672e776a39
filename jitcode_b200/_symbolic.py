"""
Symbolic front-end: what the reference does with SymEngine and jitcxde_common before any code
is generated, done here with SymPy (the input flavour named by the north star).

  * dynamical variable `y` and time `t` ..................... jitcode/sympy_symbols.py:5,8; _jitcode.py:20,23
  * input styles: iterable / generator function / dict ...... jitcxde._handle_input as used at _jitcode.py:98,687,806
  * helpers: sympify + topological sort ..................... sort_helpers / sympify_helpers, _jitcode.py:102,695
  * derivative through helpers (chain rule) ................. find_dependent_helpers + _jac_from_f_with_helpers, _jitcode.py:31-47
  * tangent-vector equations of jitcode_lyap ................ f_lyap, _jitcode.py:697-714
  * arguments with which `y` is called ...................... collect_arguments, _jitcode.py:135
"""

import sympy
from sympy.core.function import AppliedUndef

y = sympy.Function("y")
t = sympy.Symbol("t", real=True)


def is_dynvar(expr):
	"""y(i) from whatever module the user took `y` from (tests/test_sympy_input.py of the reference)."""
	return isinstance(expr, AppliedUndef) and expr.func.__name__ == "y"


def canonical(expr):
	"""sympify and make `y`, `t` independent of their provenance."""
	expr = sympy.sympify(expr)
	repl = {}
	for atom in expr.atoms(AppliedUndef):
		if atom.func.__name__ == "y" and atom.func is not y:
			repl[atom] = y(*atom.args)
	for sym in expr.free_symbols:
		if sym.name == "t" and sym != t:
			repl[sym] = t
	return expr.xreplace(repl) if repl else expr


def handle_input(f_sym, n=None):
	"""Returns (list of expressions, n).  Accepts an iterable, a generator function or a
	dictionary {y(i): expression} with exactly the keys y(0) … y(n-1)."""
	if isinstance(f_sym, dict):
		entries = {}
		for key, value in f_sym.items():
			key = canonical(key)
			if not (is_dynvar(key) and len(key.args) == 1 and key.args[0].is_Integer):
				raise ValueError("If f_sym is a dictionary, its keys must be y(0), y(1), …")
			entries[int(key.args[0])] = value
		if sorted(entries) != list(range(len(entries))):
			raise ValueError("If f_sym is a dictionary, its keys must be exactly y(0) … y(n-1).")
		f = [entries[i] for i in range(len(entries))]
	elif callable(f_sym):
		f = list(f_sym())
	else:
		f = list(f_sym)
	if n is not None and n != len(f):
		raise ValueError("len(f_sym) and n do not match.")
	return [canonical(entry) for entry in f], len(f)


def sort_helpers(helpers):
	"""sympify the (symbol, definition) pairs and order them so that each one only uses earlier ones."""
	helpers = [(sympy.sympify(symbol), canonical(definition)) for symbol, definition in (helpers or ())]
	symbols = {symbol for symbol, _ in helpers}
	done, ordered, pending = set(), [], list(helpers)
	while pending:
		ready = [item for item in pending if not ((item[1].free_symbols & symbols) - done)]
		if not ready:
			raise ValueError("Helpers have cyclic dependencies.")
		for item in ready:
			ordered.append(item)
			done.add(item[0])
			pending.remove(item)
	return ordered


def helper_derivatives(helpers, var):
	"""{helper symbol: d helper / d var}, including dependence through earlier helpers; zero entries omitted."""
	result = {}
	for symbol, definition in helpers:
		derivative = definition.diff(var)
		for other, dother in result.items():
			partial = definition.diff(other)
			if partial != 0:
				derivative += partial * dother
		if derivative != 0:
			result[symbol] = derivative
	return result


def jacobian_rows(f, helpers, n):
	"""generator of rows of the Jacobian of f (w.r.t. y(0) … y(n-1)) with the chain rule through helpers."""
	dhelpers = [helper_derivatives(helpers, y(k)) for k in range(n)]
	for entry in f:
		row = []
		for k in range(n):
			derivative = entry.diff(y(k))
			for symbol, dsymbol in dhelpers[k].items():
				partial = entry.diff(symbol)
				if partial != 0:
					derivative += partial * dsymbol
			row.append(derivative)
		yield row


def lyapunov_system(f, helpers, n_basic, n_lyap):
	"""f extended by n_lyap tangent vectors: component k of tangent vector i is y(n_basic*(i+1)+k) and
	obeys Σ_j J_kj(y) · y(n_basic*(i+1)+j), only non-zero Jacobian entries being kept (_jitcode.py:697-714)."""
	rows = list(jacobian_rows(f, helpers, n_basic))
	full = list(f)
	for i in range(n_lyap):
		for row in rows:
			full.append(sympy.Add(*[entry * y(j + (i+1)*n_basic) for j, entry in enumerate(row) if entry != 0]))
	return full


def dynvar_arguments(expr):
	"""all argument tuples with which y is called in expr."""
	return {atom.args for atom in sympy.sympify(expr).atoms(AppliedUndef) if atom.func.__name__ == "y"}
