"""
TEST INFRASTRUCTURE — CPU oracle, not product code.  Only tests/, __graft_entry__.smoke()
and bench.py's CPU legs may import this.

Symbolic → C → CPython extension, restating the reference's build chain for the right-hand
side with SymPy in place of SymEngine/jitcxde_common (neither is installable here):

  * input styles (iterable / generator function / dict) ...... jitcxde._handle_input, used at _jitcode.py:98
  * helpers sorted so that each only uses earlier ones ........ sort_helpers, _jitcode.py:102
  * control parameters → C symbols `parameter_<name>` ......... _jitcode.py:123-126,206
  * helpers → `set_general_helper(i, expr)` statements ........ generate_helpers_C, _jitcode.py:332-360
  * f → `set_dy(i, expr)` statements .......................... generate_f_C, _jitcode.py:186-249
  * tangent equations for Lyapunov exponents .................. _jac_from_f_with_helpers :31-47, f_lyap :697-714
  * compiler flags family of DEFAULT_COMPILE_ARGS ............. SURVEY.md App. A (`-std=c11 -Ofast -g0 -march=native -mtune=native`)

Two containers can hold the printed statements:
  * oracle/rhs_module.c.in — our own restatement of the extension (always available; "port");
  * the reference's own jitcode/jitced_template.c, rendered with Jinja2 where it lies
    under /root/reference and compiled into oracle/_ref/ (see build_ref.py; "reference").
Both see byte-identical `f.c` bodies.
"""

import hashlib
import importlib.machinery
import importlib.util
import os
import subprocess
import sysconfig
from pathlib import Path

import numpy as np
import sympy
from sympy.printing.c import C99CodePrinter

ORACLE_DIR = Path(__file__).resolve().parent
BUILD_DIR = ORACLE_DIR / "_build"
COMPILE_ARGS = ["-std=c11", "-Ofast", "-g0", "-march=native", "-mtune=native", "-Wno-unknown-pragmas"]
# for files that are built here and shipped to another host (oracle/_ref): no -march=native
PORTABLE_COMPILE_ARGS = ["-std=c11", "-Ofast", "-g0", "-march=x86-64-v3", "-mtune=generic", "-Wno-unknown-pragmas"]


# ---------------------------------------------------------------- symbolic side

def _is_dynvar(expr):
	return getattr(getattr(expr, "func", None), "__name__", None) == "y" and len(expr.args) == 1

def _canonical(expr):
	"""Make `y` and `t` independent of their provenance (tests/test_sympy_input.py)."""
	expr = sympy.sympify(expr)
	y = sympy.Function("y")
	repl = {}
	for atom in expr.atoms(sympy.core.function.AppliedUndef):
		if _is_dynvar(atom):
			repl[atom] = y(atom.args[0])
	for sym in expr.free_symbols:
		if sym.name == "t":
			repl[sym] = sympy.Symbol("t", real=True)
	return expr.xreplace(repl) if repl else expr

def normalise_f(f_sym, n=None):
	"""list of n expressions from any of the three input styles."""
	if isinstance(f_sym, dict):
		entries = {}
		for key, value in f_sym.items():
			assert _is_dynvar(key), "dictionary keys must be y(i)"
			entries[int(key.args[0])] = value
		assert sorted(entries) == list(range(len(entries)))
		f = [entries[i] for i in range(len(entries))]
	elif callable(f_sym):
		f = list(f_sym())
	else:
		f = list(f_sym)
	if n is not None and n != len(f):
		raise ValueError("len(f_sym) and n do not match.")
	return [_canonical(entry) for entry in f]

def sort_helpers(helpers):
	"""topological order: every helper only depends on earlier ones."""
	helpers = [(sympy.sympify(s), _canonical(e)) for s, e in helpers]
	symbols = {s for s, _ in helpers}
	done, result, pending = set(), [], list(helpers)
	while pending:
		progress = False
		for item in list(pending):
			if not ((item[1].free_symbols & symbols) - done):
				result.append(item); done.add(item[0]); pending.remove(item); progress = True
		if not progress:
			raise ValueError("Helpers have cyclic dependencies.")
	return result

def total_derivative(expr, helpers, var):
	"""d expr / d var including the dependence through helpers (chain rule); this is what
	find_dependent_helpers + the inner loop of _jac_from_f_with_helpers compute (_jitcode.py:31-44)."""
	dhelper = {}
	for symbol, definition in helpers:
		d = definition.diff(var)
		for other, dother in dhelper.items():
			if dother != 0:
				d += definition.diff(other) * dother
		dhelper[symbol] = d
	result = expr.diff(var)
	for symbol, d in dhelper.items():
		if d != 0:
			result += expr.diff(symbol) * d
	return result

def lyapunov_system(f_sym, n_lyap, helpers=(), n=None):
	"""the n·(n_lyap+1)-dimensional system of jitcode_lyap (_jitcode.py:697-721)."""
	y = sympy.Function("y")
	f = normalise_f(f_sym, n)
	helpers = sort_helpers(helpers)
	nb = len(f)
	full = list(f)
	for i in range(n_lyap):
		for entry in f:
			line = [total_derivative(entry, helpers, y(k)) for k in range(nb)]
			full.append(sympy.Add(*[d * y(k + (i+1)*nb) for k, d in enumerate(line) if d != 0]))
	return full, helpers


# ---------------------------------------------------------------- printing

def print_statements(f_sym, helpers=(), control_pars=(), n=None):
	"""returns (f_statements, helper_statements, n, parameter_names), each statement one C line."""
	f = normalise_f(f_sym, n)
	helpers = sort_helpers(helpers)
	subs = {par: sympy.Symbol("parameter_" + par.name) for par in control_pars}
	get_helper = sympy.Function("get_general_helper")
	for i, (symbol, _) in enumerate(helpers):
		subs[symbol] = get_helper(i)
	printer = C99CodePrinter({"allow_unknown_functions": True})
	helper_lines = [
		f"set_general_helper({i}, {printer.doprint(definition.xreplace(subs))});"
		for i, (_, definition) in enumerate(helpers)
	]
	f_lines = [f"set_dy({i}, {printer.doprint(entry.xreplace(subs))});" for i, entry in enumerate(f)]
	return f_lines, helper_lines, len(f), [par.name for par in control_pars]


# ---------------------------------------------------------------- building

def _host_tag():
	try:
		with open("/proc/cpuinfo") as cpuinfo:
			flags = next(line for line in cpuinfo if line.startswith("flags"))
	except (OSError, StopIteration):
		flags = "unknown"
	return hashlib.sha1(flags.encode()).hexdigest()[:8]

def _python_build_flags():
	include = sysconfig.get_paths()["include"]
	candidates = [include, "/usr/include/python3.12"]
	includes = [f"-I{path}" for path in candidates if os.path.exists(os.path.join(path, "Python.h"))]
	includes.append(f"-I{np.get_include()}")
	return includes

def render_port_source(f_lines, helper_lines, n, par_names, module):
	template = (ORACLE_DIR / "rhs_module.c.in").read_text()
	repl = {
		"@DIM@": str(n),
		"@NPAR@": str(len(par_names)),
		"@NHELP@": str(len(helper_lines)),
		"@PARAMETER_DECLARATIONS@": "\n".join(f"static double parameter_{name};" for name in par_names),
		"@PARAMETER_ADDRESSES@": ", ".join(f"&parameter_{name}" for name in par_names),
		"@HELPER_STATEMENTS@": "\n".join("\t" + line for line in helper_lines),
		"@F_STATEMENTS@": "\n".join("\t" + line for line in f_lines),
		"@MODULE@": module,
	}
	for key, value in repl.items():
		template = template.replace(key, value)
	return template

def load_extension(path, module):
	loader = importlib.machinery.ExtensionFileLoader(module, str(path))
	spec = importlib.util.spec_from_file_location(module, str(path), loader=loader)
	mod = importlib.util.module_from_spec(spec)
	loader.exec_module(mod)
	return mod

def compile_extension(source_path, target, extra_includes=(), portable=False):
	cmd = ["gcc", "-shared", "-fPIC", *(PORTABLE_COMPILE_ARGS if portable else COMPILE_ARGS), *_python_build_flags(),
			*[f"-I{inc}" for inc in extra_includes], str(source_path), "-o", str(target), "-lm"]
	subprocess.run(cmd, check=True, capture_output=True)

def build_rhs(f_sym, helpers=(), control_pars=(), n=None, build_dir=None):
	"""Compile the system and return a handle: .f(t,y), .f_many(t,Y), .initialise(*pars), .path, .module_name.
	Content-addressed, so repeated calls (and multiprocessing workers) reuse the file."""
	f_lines, helper_lines, n, par_names = print_statements(f_sym, helpers, control_pars, n)
	key = hashlib.sha1(("\n".join(f_lines + helper_lines + par_names) + _host_tag()).encode()).hexdigest()[:16]
	module = f"jbo_{key}"
	build_dir = Path(build_dir or BUILD_DIR)
	build_dir.mkdir(parents=True, exist_ok=True)
	target = build_dir / f"{module}.so"
	if not target.exists():
		source = build_dir / f"{module}.c"
		source.write_text(render_port_source(f_lines, helper_lines, n, par_names, module))
		tmp = build_dir / f"{module}.{os.getpid()}.tmp.so"
		compile_extension(source, tmp)
		os.replace(tmp, target)
	return RhsHandle(target, module, n, len(par_names))


class RhsHandle:
	"""picklable reference to a built extension (workers re-import it by path)."""
	def __init__(self, path, module_name, n, n_par):
		self.path, self.module_name, self.n, self.n_par = str(path), module_name, n, n_par
		self._mod = None

	def __getstate__(self):
		return {"path": self.path, "module_name": self.module_name, "n": self.n, "n_par": self.n_par}

	def __setstate__(self, state):
		self.__dict__.update(state)
		self._mod = None

	@property
	def mod(self):
		if self._mod is None:
			self._mod = load_extension(self.path, self.module_name)
		return self._mod

	@property
	def f(self):
		return self.mod.f

	@property
	def f_many(self):
		return self.mod.f_many

	def initialise(self, *pars):
		self.mod.initialise(*[float(p) for p in pars])
