"""
The JiTCODE classes on top of the B200 ensemble engine.

Same public surface as the reference (jitcode/_jitcode.py): `jitcode` (:49-669), `jitcode_lyap`
(:671-779), `jitcode_transversal_lyap` (:782-919), `jitcode_restricted_lyap` (:922-946), `test`
(:948-965); SymPy input instead of SymEngine.  What differs is what happens below the surface
and that every state has a batch axis:

  * `set_initial_value` takes one state (n,) like the reference or an ensemble (B, n);
    `set_parameters` takes scalars or one value per trajectory; `integrate(t)` advances all B
    trajectories with ONE kernel launch and returns (B, n) (or (n,) for a single state);
  * the code generator emits CUDA (generate_f_cuda, alias generate_f_C), nvcc compiles it for
    sm_100a into a cached shared object (compile_cuda, alias compile_C) that is loaded through
    ctypes (include/jitcode_b200.h); the Runge–Kutta steppers are part of that module;
  * only explicit Runge–Kutta pairs exist: "dopri5"/"dop853" (Hairer's controller, the reference's
    `ode` back end) and "RK45"/"DOP853"/"RK23" (SciPy's solve_ivp controller, with or without
    interpolation); no lambdified fallback, no Python callbacks, no implicit solvers.
"""

import shutil
from pathlib import Path
from warnings import warn

import numpy as np
import sympy
import torch

from . import _build, _codegen, _dense, _vectorise
from ._integrator import (
	JB_POST_LYAP_QR, JB_POST_LYAP_RESTRICTED, JB_POST_LYAP_TRANSVERSAL, JB_POST_NONE,
	CudaEnsembleIntegrator, UnsuccessfulIntegration, empty_integrator, require_cuda,
)
from ._symbolic import dynvar_arguments, handle_input, jacobian_rows, lyapunov_system, sort_helpers, t, y
from ._transversal import GroupHandler

# name → (method, backend); the reference resolves names in integrator_info (integrator_tools.py:15-37)
INTEGRATORS = {
	"dopri5": (_codegen.JB_DP5, "ode"),
	"dop853": (_codegen.JB_DOP853, "ode"),
	"RK45": (_codegen.JB_DP5, "ivp"),
	"DOP853": (_codegen.JB_DOP853, "ivp"),
	"RK23": (_codegen.JB_RK23, "ivp"),
}
NOT_REBUILT = ("Radau", "BDF", "LSODA", "vode", "lsoda", "zvode")

_used_module_names = set()


SHARED_MEMORY_PER_BLOCK = 227 * 1024      # sm_100a: opt-in maximum of dynamic shared memory per block
STORAGE_BY_NAME = {"registers": _codegen.JB_STORE_REGISTERS, "global": _codegen.JB_STORE_GLOBAL, "warp": _codegen.JB_STORE_WARP,
		"shared": _codegen.JB_STORE_SHARED, "quad": _codegen.JB_STORE_QUAD, "dense": "dense"}
REGISTER_DOUBLES = 88                     # state + stages a thread keeps in registers (measured: 80 doubles compile without spills)
# The dense engine costs 4·n² flops per right-hand side and trajectory, the written-out form ≈ 40 per sine term:
# measured on the reference's own example (examples/kuramoto_network.py: n = 100, 20 % density) the GEMM form is
# 4.5× faster than the generated code, so it is chosen from 10 % density on.
DENSE_MINIMUM_N = 32
DENSE_MINIMUM_DENSITY = 0.1


def _stage_rows(method, driver):
	if method == _codegen.JB_RK23:
		return 4
	return 7 if method == _codegen.JB_DP5 else (16 if driver == _codegen.JB_DRV_IVP_DENSE else 13)


def _shared_threads(n, method, driver):
	"""threads per block of shared storage: every thread keeps its trajectory's state and stages in shared memory"""
	rows = 2 + _stage_rows(method, driver) + (1 if driver == _codegen.JB_DRV_IVP_DENSE else 0)
	return min(256, SHARED_MEMORY_PER_BLOCK // (rows * n * 8) // 32 * 32)


def _storage_for(n, method, driver):
	"""registers while state + stages fit a thread's register file; the block's shared memory while at least four
	warps of trajectories fit it (mid-size systems: the stages stay on chip, 30 instead of 250 cycles away); else the
	tiled global workspace (warp storage is chosen on top of this when the system vectorises well, see jitcode._module)."""
	rows = _stage_rows(method, driver) + 3
	if rows * n <= REGISTER_DOUBLES:
		return _codegen.JB_STORE_REGISTERS
	return _codegen.JB_STORE_SHARED if _shared_threads(n, method, driver) >= 128 else _codegen.JB_STORE_GLOBAL


def _warp_configuration(n, method, driver, table_bytes, kregs=None, threads=None, team_warps=None):
	"""(kregs, threads, team_warps) of a warp-storage module.

	team_warps : how many warps (1, 2, 4, 8) integrate ONE trajectory together.  One is best while many trajectories
		fit an SM's shared memory; larger systems get more warps per trajectory so that the SM still has enough
		warps in flight to hide the latency of the shared-memory gathers.
	kregs : how many stage rows K[0 … kregs) — plus the state — every thread keeps in registers (its share of the
		components) instead of shared memory.  Shared memory is the bottleneck of this kernel, so rows are taken
		out of it as long as that does not cost resident warps.
	threads : threads per block (one persistent block per SM) = 32 · team_warps · trajectories per block."""
	stages = _stage_rows(method, driver)
	dense = 1 if driver == _codegen.JB_DRV_IVP_DENSE else 0
	row_bytes = _vectorise.row_length(n) * 8
	available = SHARED_MEMORY_PER_BLOCK - table_bytes - 8 * n - 64
	last_stage = {_codegen.JB_DP5: 6, _codegen.JB_DOP853: 12, _codegen.JB_RK23: 3}[method]      # K[S]: its row doubles as the staging row while k <= S

	def teams(k, w):
		"""trajectories per SM with k cached rows and w warps per trajectory"""
		passes = (n + 32 * w - 1) // (32 * w)
		rows = 1 + (1 if (k == 0 or k > last_stage) else 0) + dense + stages - k       # WarpStore::ROWS
		by_shared = available // (rows * row_bytes + 64)
		registers = 80 + 2 * passes * (1 + k) if k else 96       # measured on the Rössler network (profiles/)
		if registers > 255:
			return 0
		by_registers = 65536 // (32 * w * registers)
		limit = 15 if w > 1 else 16                              # named barriers 1 … 15 / warps per block
		return max(0, min(limit, 1024 // (32 * w), by_shared, by_registers))

	def best_rows(w):
		"""cached rows for teams of w warps: most trajectories per SM first, then most rows out of shared memory; with
		one warp per trajectory 12 resident warps count as enough.  Measured on Rössler networks with the blocked ring order: n = 300 runs 6 % faster with
		3 rows and 12 warps than with 2 rows and 14; n = 150 equally fast with 7 rows and 12 warps as with 3 and 16."""
		options = [max(0, min(int(k), stages)) for k in ((kregs,) if kregs is not None else range(stages + 1))]
		enough = max(teams(k, w) for k in options)
		if w == 1:
			enough = min(enough, 12)
		return max(k for k in options if teams(k, w) >= enough)

	if team_warps is None:
		# Measured on Rössler networks (profiles/README.md): a single warp per trajectory wins as long as it keeps ~6+
		# warps resident (n = 600: 8 single warps beat 7 teams of 2 by 15 %) — block-level barriers and coarser passes
		# cost more than the extra warps hide.  Beyond that, the team size that keeps most warps resident (up to 16),
		# the smaller team on a tie (n = 1500: 2 teams of 8 warps run 10× faster than the thread-per-trajectory
		# kernel; single warps do not fit at all).
		choices = {w: teams(best_rows(w), w) * w for w in (1, 2, 4, 8)}
		team_warps = 1 if choices[1] >= 6 else max(choices, key=lambda w: (min(choices[w], 16), -w))
	kregs = best_rows(team_warps)
	if threads is None:
		threads = 32 * team_warps * teams(kregs, team_warps)
	return kregs, threads, team_warps


class jitcode:
	"""
	Parameters (as in the reference, _jitcode.py:49-95)
	----------
	f_sym : iterable of SymPy expressions, generator function, or dictionary {y(i): expression}
	helpers : list of (symbol, expression) pairs evaluated before the derivative
	wants_jacobian : accepted for compatibility; the explicit methods never use a Jacobian
	n : length of f_sym (checked)
	control_pars : symbols that are set after compilation with `set_parameters` — per trajectory here
	callback_functions : not supported on a GPU (must be empty)
	verbose : progress reports
	module_location : path of a module saved with `save_compiled`; then f_sym may be omitted but n must be given
	"""

	dynvar = y

	def __init__(self, f_sym=(), *, helpers=None, wants_jacobian=False, n=None, control_pars=(),
			callback_functions=(), verbose=True, module_location=None):
		if callback_functions:
			raise NotImplementedError("Python callbacks cannot run inside a CUDA kernel.")
		self.verbose = verbose
		self._f_list, self.n = self._handle_input(f_sym, n, module_location)
		self.helpers = sort_helpers(helpers)
		self.control_pars = list(control_pars)
		self.control_par_values = None
		self._wants_jacobian = wants_jacobian
		self.integrator = empty_integrator()
		self._statements = None
		self._compile_options = {"extra_compile_args": (), "extra_link_args": (), "verbose": False}
		self._modules = {}
		self._integrator_config = None
		self._module_location = module_location
		self._f_module = None
		if module_location is not None:
			module = _build.CudaModule(module_location)
			if module.info.n != self.n:
				raise ValueError("The module at module_location was compiled for a different dimension.")
			self._modules[(module.info.method, module.info.driver, module.info.storage)] = module
			self._modules[(module.info.method, module.info.driver, None)] = module
		# layout constants of the module; the Lyapunov subclasses overwrite them
		self._n_basic_module = self.n
		self._n_lyap_module = 0
		self._tangent_indices_module = []

	@classmethod
	def from_sine_coupling(cls, W, omega, coupling_parameter=None, verbose=False):
		"""
		The system  dy_i/dt = omega[i] + c·Σ_j W[i, j]·sin(y_j − y_i)  given by its data instead of by n·deg symbolic
		sine terms (which SymPy needs minutes to build for a dense network of 1000 oscillators).  Equivalent to
		`jitcode(f_sym)` with that f_sym — the symbolic input is recognised and ends up in the same kernels — and
		only available with `set_integrator("dopri5")` and `set_integrator("dop853")`.
		coupling_parameter : a symbol to make c a control parameter (`set_parameters`: one value or one per trajectory);
			None: c = 1.
		"""
		W, omega = np.asarray(W, dtype=float), np.asarray(omega, dtype=float)
		if W.shape != (len(omega), len(omega)):
			raise ValueError("W must be (n, n) for omega of length n.")
		self = cls.__new__(cls)
		jitcode.__init__(self, [y(0)], n=1, verbose=verbose)        # placeholder system; replaced by the coupling data
		self._f_list, self.n = [], len(omega)
		self._n_basic_module = self.n
		self.control_pars = [coupling_parameter] if coupling_parameter is not None else []
		self._sine_coupling = (W, omega, coupling_parameter) if coupling_parameter is not None else (W, omega)
		return self

	def _dense_coupling(self):
		"""(W, omega[, c]) if the system is a network of sine-coupled phase oscillators (optionally with the coupling
		strength c as its one control parameter), else None (cached)."""
		if not hasattr(self, "_sine_coupling"):
			self._sine_coupling = None
			if self._f_list and not self.helpers and len(self.control_pars) <= 1 and self._post == JB_POST_NONE:
				self._sine_coupling = _dense.detect_sine_coupling(self._f_list, self.control_pars)
		return self._sine_coupling

	# ------------------------------------------------------------------ input handling and checks

	def _handle_input(self, f_sym, n, module_location=None):
		if module_location is not None and not f_sym:
			if n is None:
				raise ValueError("If no f_sym is given, n must be given.")
			return [], n
		return handle_input(f_sym, n)

	def f_sym(self):
		return iter(self._f_list)

	def report(self, message):
		if self.verbose:
			print(message)

	def _check_assert(self, condition, message):
		if not condition:
			self._check_failures.append(message)

	def check(self, fail_fast=True):
		"""Runs the consistency checks of the reference (_jitcode.py:128-154); raises ValueError on the first
		(fail_fast) or on all collected problems."""
		self._check_failures = []
		checks = (self._check_non_empty, self._check_valid_arguments, self._check_valid_symbols)
		for checker in checks:
			checker()
			if self._check_failures and fail_fast:
				break
		if self._check_failures:
			raise ValueError("\n".join(self._check_failures))

	def _check_non_empty(self):
		self._check_assert(self._f_list, "f_sym is empty.")

	def _check_valid_arguments(self):
		for i, entry in enumerate(self._f_list):
			for argument in dynvar_arguments(entry):
				self._check_assert(len(argument) == 1 and argument[0].is_Integer, f"y is called with invalid arguments {argument} in equation {i}.")
				if len(argument) == 1 and argument[0].is_Integer:
					self._check_assert(argument[0] >= 0, f"y is called with a negative argument ({argument[0]}) in equation {i}.")
					self._check_assert(argument[0] < self.n, f"y is called with an argument ({argument[0]}) higher than the system’s dimension ({self.n}) in equation {i}.")

	def _check_valid_symbols(self):
		valid = {t} | {helper[0] for helper in self.helpers} | set(self.control_pars)
		for i, entry in enumerate(self._f_list):
			for symbol in entry.free_symbols:
				self._check_assert(symbol in valid, f"Invalid symbol ({symbol.name}) in equation {i}.")

	# ------------------------------------------------------------------ symbolic Jacobian (tangent equations use it)

	@property
	def jac_sym(self):
		if not hasattr(self, "_jac_sym"):
			self._jac_sym = list(jacobian_rows(self._f_list, self.helpers, self.n))
			self.report("generated symbolic Jacobian")
		return self._jac_sym

	# ------------------------------------------------------------------ code generation and compilation

	def generate_f_cuda(self, simplify=None, do_cse=False, chunk_size=None):
		"""
		Translates the derivative to CUDA statements (counterpart of generate_f_C, _jitcode.py:186-249).

		simplify : apply SymPy's `simplify(ratio=1)` to every component first (default: off — nvcc optimises
			the arithmetic anyway and SymPy's simplify is slow; the reference defaults to n<=10)
		do_cse : full common-subexpression elimination with sympy.cse (repeated function calls such as
			sin(t) are always shared)
		chunk_size : accepted for compatibility; the global-storage kernels call one out-of-line copy of the
			right-hand side, which bounds code size the way chunking does for the C compiler
		"""
		self.check()
		f = self._f_list
		if simplify:
			f = [entry.simplify(ratio=1.0) for entry in f]
		self._statements = _codegen.rhs_statements(f, self.helpers, self.control_pars, do_cse=do_cse)
		self._modules = {key: module for key, module in self._modules.items() if key[2] == _codegen.JB_STORE_WARP}
		self._f_module = None
		self.report("generated CUDA code for f")

	generate_f_C = generate_f_cuda

	def generate_helpers_C(self, chunk_size=None):
		"""helpers are emitted together with f (generate_f_cuda); kept for call compatibility."""

	def generate_jac_C(self, *args, **kwargs):
		raise NotImplementedError("Only explicit Runge–Kutta methods exist here; none of them uses a Jacobian.")

	def generate_lambdas(self, *args, **kwargs):
		raise NotImplementedError("There is no CPU fallback: the derivative only exists as CUDA code.")

	def compile_cuda(self, extra_compile_args=None, extra_link_args=None, verbose=False, modulename=None, omp=None,
			method=_codegen.JB_DP5, driver=_codegen.JB_DRV_ODE, storage=None, threads=None, kregs=None, min_blocks=None, team_warps=None):
		"""
		Compiles (or fetches from the cache) and loads the module for one integrator configuration
		(counterpart of compile_C, _jitcode.py:367-417).  `set_integrator` calls this itself.
		"""
		if isinstance(storage, str):
			storage = STORAGE_BY_NAME[storage]
		if modulename is not None:
			if modulename in _used_module_names:
				raise NameError("Module name has already been used in this instance of Python.")
			_used_module_names.add(modulename)
		if extra_compile_args is not None:
			self._compile_options["extra_compile_args"] = tuple(extra_compile_args)
		if extra_link_args is not None:
			self._compile_options["extra_link_args"] = tuple(extra_link_args)
		self._compile_options["verbose"] = verbose
		return self._module(method, driver, storage, threads, kregs, min_blocks, team_warps)

	compile_C = compile_cuda

	blocked_rings = True      # warp storage: ring neighbourhoods served from register windows (_vectorise.ring_blockings)
	fuse_groups = True        # warp storage: equations of equal count share one loop (loads of common operands merge)

	def _warp_plan(self, team_warps=1):
		"""the vectorised form of the right-hand side, printed for teams of `team_warps` warps per trajectory
		(None if the system cannot be vectorised)."""
		if not hasattr(self, "_plans"):
			self._plans = {}
		if team_warps not in self._plans:
			self._plans[team_warps] = None
			if self._f_list:
				try:
					if not hasattr(self, "_plan_tree"):
						self._plan_tree = _vectorise.make_plan(self._f_list, self.helpers, self.control_pars)
					lines, tables, component_of = _vectorise.warp_statements(self._plan_tree, team_warps, self.blocked_rings, self.fuse_groups)
					self._plans[team_warps] = (self._plan_tree, lines, tables, component_of)
				except NotImplementedError:
					pass
		return self._plans[team_warps]

	def _quad_group(self, method, driver):
		"""lanes per trajectory of quad storage, or 0 if it does not apply: jitcode_lyap with at least two tangent vectors
		whose state plus ONE tangent vector fit a thread's registers (lane j of the group integrates the state and
		tangent vector j; norms and Gram–Schmidt are shuffles within the group)."""
		k = self._n_lyap_module
		if self._post != JB_POST_LYAP_QR or k < 2 or k > 32 or not getattr(self, "_f_basic", None):
			return 0
		if (_stage_rows(method, driver) + 3) * 2 * self._n_basic_module > REGISTER_DOUBLES:
			return 0
		return 1 << (k - 1).bit_length()

	def _choose_storage(self, method, driver):
		storage = _storage_for(self.n, method, driver)
		if storage != _codegen.JB_STORE_REGISTERS and self._quad_group(method, driver):
			return _codegen.JB_STORE_QUAD
		if storage != _codegen.JB_STORE_REGISTERS and self.n >= 64:
			plan = self._warp_plan()
			if plan is not None and plan[0].efficiency >= 0.5:
				if _warp_configuration(self.n, method, driver, plan[2].nbytes)[1] >= 128:
					storage = _codegen.JB_STORE_WARP
		return storage

	def _module(self, method, driver, storage=None, threads=None, kregs=None, min_blocks=None, team_warps=None):
		if storage is None:
			if (method, driver, None) in self._modules:
				return self._modules[(method, driver, None)]
			storage = self._choose_storage(method, driver) if self._f_list else None
			if storage is None:
				raise RuntimeError("This object was created from a saved module for another integrator configuration and has no f_sym to compile.")
		key = (method, driver, storage) if (threads is None and kregs is None and min_blocks is None and team_warps is None) else (method, driver, storage, threads, kregs, min_blocks, team_warps)
		if key not in self._modules:
			if not self._f_list:
				raise RuntimeError("This object was created from a saved module for another integrator configuration and has no f_sym to compile.")
			tables = component_of = None
			group = 1
			if storage == _codegen.JB_STORE_QUAD:
				group = self._quad_group(method, driver)
				if not group:
					raise ValueError("Quad storage needs a jitcode_lyap system with 2 … 32 tangent vectors whose state plus one tangent vector fit a thread's registers.")
				self.check()
				# the right-hand side every lane runs: the state and ONE tangent vector (_jitcode.py:697-714 with n_lyap = 1)
				local = lyapunov_system(self._f_basic, self.helpers, self._n_basic_module, 1)
				statements = _codegen.rhs_statements(local, self.helpers, self.control_pars)
				if threads is None:
					threads = 128
				if min_blocks is None:
					# measured on the FHN pair with four tangent vectors (234 registers uncapped): 12 warps per SM at 168
					# registers with a few spills run 7 % faster than 8 warps without
					min_blocks = max(1, 384 // threads)
			elif storage == _codegen.JB_STORE_WARP:
				plan = self._warp_plan()
				if plan is None:
					raise NotImplementedError("This system cannot be vectorised for warp storage.")
				self.check()
				kregs, threads, team_warps = _warp_configuration(self.n, method, driver, plan[2].nbytes, kregs, threads, team_warps)
				if threads < 32 * team_warps:
					raise ValueError("The system is too large for warp storage: one trajectory does not fit shared memory.")
				_, statements, tables, component_of = self._warp_plan(team_warps)
			else:
				if self._statements is None:
					self.generate_f_cuda()
				statements = self._statements
				if threads is None:
					threads = {_codegen.JB_STORE_REGISTERS: 128, _codegen.JB_STORE_SHARED: _shared_threads(self.n, method, driver)}.get(storage, 64)
				if storage == _codegen.JB_STORE_SHARED and not 32 <= threads <= _shared_threads(self.n, method, driver):
					raise ValueError("Shared storage: the stages of that many trajectories do not fit one block's shared memory.")
				if min_blocks is None and storage == _codegen.JB_STORE_GLOBAL:
					min_blocks = 8      # 16 warps/SM at ≤ 128 registers: measured +60 % over 8 warps at 255 (latency-bound streaming)
				if min_blocks is None and storage == _codegen.JB_STORE_REGISTERS:
					# small systems: cap the registers so that 16 (12) warps stay resident — the delivery code of the
					# persistent kernel (dense output, write-back) otherwise inflates the allocation of the step loop
					# (Lotka–Volterra under RK45: 142 registers uncapped = 12 warps, measured 10 % slower than 16)
					doubles = (_stage_rows(method, driver) + 3) * self.n
					min_blocks = max(1, (4 if doubles <= 40 else 3 if doubles <= 64 else 2) * 128 // threads)
			source = _codegen.module_source(
				statements, self.n, len(self.control_pars),
				self._n_basic_module, self._n_lyap_module, self._tangent_indices_module,
				method, driver, storage, threads, tables=tables, component_of=component_of, kregs=kregs or 0,
				min_blocks=min_blocks or 1, team_warps=team_warps or 1, group=group,
			)
			path = _build.compile_module(
				source, self._compile_options["extra_compile_args"], self._compile_options["extra_link_args"],
				verbose=self._compile_options["verbose"],
			)
			self._modules[key] = _build.CudaModule(path)
			self.report(f"compiled and loaded {path.name}")
		return self._modules[key]

	def save_compiled(self, destination="", overwrite=False):
		"""Copies the module of the current integrator (compiling the default one if there is none) to
		`destination` and returns the path, to be passed as `module_location` (tests/test_jitcode.py:77-84)."""
		if self._integrator_config is not None:
			module = self._module(*self._integrator_config)
		elif self._modules:
			module = next(iter(self._modules.values()))
		else:
			module = self._module(_codegen.JB_DP5, _codegen.JB_DRV_ODE)
		source = Path(module.path)
		destination = Path(destination) if destination else Path.cwd()
		if destination.is_dir() or str(destination).endswith("/"):
			destination = destination / source.name
		if destination.exists() and not overwrite:
			raise OSError(f"Target File already exists and “overwrite” is set to False: {destination}")
		destination.parent.mkdir(parents=True, exist_ok=True)
		shutil.copyfile(source, destination)
		return str(destination)

	# ------------------------------------------------------------------ f(t, y), as the reference exposes jitced.f

	def f(self, time, state):
		"""Evaluates the compiled derivative on the GPU: state (n,) → (n,), or (B, n) → (B, n)."""
		if isinstance(self.integrator, _dense.CudaDenseSineIntegrator) or not self._f_list:
			coupling = self._dense_coupling()
			helper = _dense.CudaDenseSineIntegrator(*coupling[:2], coupling_parameter=len(coupling) == 3)
			if len(coupling) == 3:
				helper.set_parameters(self._parameter_rows())
			return helper.eval_f(state)
		module = self._f_module or (next(iter(self._modules.values())) if self._modules else self._module(_codegen.JB_DP5, _codegen.JB_DRV_ODE))
		self._f_module = module
		helper = CudaEnsembleIntegrator(module, device=getattr(self.integrator, "device", None))
		rows, kind, single = helper._to_device_rows(state, self.n, "state")
		helper._alloc(rows.shape[0])
		helper._out_kind, helper._single = kind, single
		if self.control_pars:
			helper.set_parameters(self._parameter_rows())
			helper._upload_parameters()
		helper._pack(rows, helper.Y, self.n)
		times = np.broadcast_to(np.asarray(time, dtype=float), (helper.B,)).copy()
		helper.tt[:helper.B].copy_(torch.as_tensor(times))
		out = torch.empty_like(helper.Y)
		module.eval_f(helper.tt.data_ptr(), helper.Y.data_ptr(), helper.P.data_ptr(), out.data_ptr(), helper.B, helper.ld, helper.stream)
		return helper._export(helper._unpack(out, self.n))

	# ------------------------------------------------------------------ state, parameters, integration

	@property
	def t(self):
		return self.integrator.t

	@property
	def _y(self):
		return self.integrator._y

	@property
	def y(self):
		return self.integrator._y

	@property
	def y_dict(self):
		state = self.y
		return {self.dynvar(i): state[..., i] for i in range(self.n)}

	def _list_from_dynvar_dict(self, dictionary, name, n):
		keys = {}
		for key, value in dictionary.items():
			key = sympy.sympify(key)
			if not (getattr(key.func, "__name__", None) == "y" and len(key.args) == 1):
				raise ValueError(f"The keys of the {name} dictionary must be y(0), y(1), …")
			keys[int(key.args[0])] = value
		if sorted(keys) != list(range(n)):
			raise ValueError(f"The {name} dictionary must have exactly the keys y(0) … y({n-1}).")
		return [keys[i] for i in range(n)]

	def set_initial_value(self, initial_value, time=0.0):
		"""One state (list / array of length n, or dictionary {y(i): value}) or an ensemble: (B, n) NumPy
		array, torch tensor (CPU or CUDA).  The kind of the input decides the kind of what `integrate` returns."""
		if isinstance(initial_value, dict):
			initial_value = self._list_from_dynvar_dict(initial_value, "initial value", self.n)
		shape = tuple(initial_value.shape) if hasattr(initial_value, "shape") else (len(initial_value),)
		if shape[-1] != self.n or len(shape) > 2:
			raise ValueError("The dimension of the initial value does not match the dimension of your differential equations.")
		self.integrator.set_initial_value(initial_value, time)

	def _parameter_rows(self):
		"""(1 or B, p) array from whatever set_parameters got: scalars and/or one vector per parameter."""
		if self.control_par_values is None:
			raise RuntimeError("Something needs parameters to be set. Try calling `set_parameters` earlier.")
		values = self.control_par_values
		if isinstance(values, torch.Tensor) or (isinstance(values, np.ndarray) and values.ndim == 2):
			return values
		columns = [v if isinstance(v, torch.Tensor) else np.asarray(v, dtype=float) for v in values]
		if all(np.ndim(c) == 0 for c in columns):
			return np.array([[float(c) for c in columns]])
		if any(isinstance(c, torch.Tensor) for c in columns):
			B = max(c.shape[0] for c in columns if np.ndim(c))
			columns = [torch.as_tensor(c, dtype=torch.float64).expand(B) if np.ndim(c) == 0 else c.to(torch.float64) for c in columns]
			device = next(c.device for c in columns if c.is_cuda) if any(c.is_cuda for c in columns) else "cpu"
			return torch.stack([c.to(device) for c in columns], dim=1)
		B = max(c.shape[0] for c in columns if c.ndim)
		return np.stack([np.broadcast_to(c, (B,)) for c in columns], axis=1)

	def set_parameters(self, *args):
		"""
		As in the reference (_jitcode.py:638-655): the values as separate arguments or as one sequence.
		Each value may be a number (shared by all trajectories) or a vector with one entry per trajectory;
		a (B, p) array or tensor is taken as is.
		"""
		p = len(self.control_pars)
		if len(args) == 1 and isinstance(args[0], (torch.Tensor, np.ndarray)) and np.ndim(args[0]) == 2:
			values = args[0]
			count = values.shape[1]
		elif len(args) == 1 and p == 1 and np.ndim(args[0]) == 1 and len(args[0]) != 1:
			values = (args[0],)          # one control parameter, one value per trajectory
			count = 1
		else:
			try:
				values = tuple(args[0])
			except (TypeError, IndexError):
				values = args
			else:
				if len(args) > 1:
					if len(args) == p:
						values = args    # one vector per control parameter
					else:
						raise TypeError("Argument must either be a single sequence or multiple numbers.")
			count = len(values)
		if count != len(self.control_pars):
			raise ValueError("Number of values does not match number of control parameters.")
		self.control_par_values = values
		if isinstance(self.integrator, (CudaEnsembleIntegrator, _dense.CudaDenseSineIntegrator)) and count:
			self.integrator.set_parameters(self._parameter_rows())

	def set_f_params(self, *args):
		warn("This function has been replaced by `set_parameters`", stacklevel=2)
		self.set_parameters(*args)

	def set_jac_params(self, *args):
		warn("This function has been replaced by `set_parameters`", stacklevel=2)
		self.set_parameters(*args)

	_post = JB_POST_NONE
	_proj = None

	def set_integrator(self, name, nsteps=10**6, interpolate=True, *, device=None, storage=None, threads=None, kregs=None, min_blocks=None, team_warps=None, **integrator_params):
		"""
		As in the reference (_jitcode.py:560-626).

		name : "dopri5" | "dop853" (Hairer's codes as scipy.integrate.ode runs them: every `integrate` call is a
			fresh start that lands exactly on the target time) or "RK45" | "DOP853" | "RK23" (SciPy's solve_ivp stepper,
			persistent between calls; with `interpolate` the sample is the dense output of the step that passed
			the target time, without it the last step is clipped to the target time).
		nsteps : limit of step attempts per `integrate` call for the `ode` names.
		integrator_params : rtol, atol (scalar or one per component), max_step, first_step, and for the `ode`
			names safety, ifactor, dfactor, beta — passed to the kernel with SciPy's defaults.
		device / storage ("registers" | "global" | "warp") / threads / kregs : where and how the kernel runs
			(engine-specific; see _storage_for, _choose_storage and _warp_configuration).
		"""
		if name in NOT_REBUILT:
			raise NotImplementedError(f"The B200 engine only implements explicit Runge–Kutta pairs (dopri5, dop853, RK45, DOP853, RK23), not {name!r}.")
		if name not in INTEGRATORS:
			raise RuntimeError("There is no integrator with that name.")
		method, backend = INTEGRATORS[name]
		if backend == "ode":
			driver = _codegen.JB_DRV_ODE
		else:
			driver = _codegen.JB_DRV_IVP_DENSE if interpolate else _codegen.JB_DRV_IVP_LAND
		if isinstance(storage, str):
			storage = STORAGE_BY_NAME[storage]
		old_integrator = self.integrator
		dense = storage == "dense"
		dense_capable = method in (_codegen.JB_DP5, _codegen.JB_DOP853)
		if storage is None and dense_capable and self.n >= DENSE_MINIMUM_N and self._dense_coupling() is not None:
			dense = np.count_nonzero(self._dense_coupling()[0]) >= DENSE_MINIMUM_DENSITY * self.n**2
		if dense:
			if not dense_capable or self._dense_coupling() is None:
				raise NotImplementedError("The dense engine integrates sine-coupled phase oscillators with \"dopri5\", \"dop853\", \"RK45\", or \"DOP853\" only.")
			_dense.dense_library()                                       # compiling needs no GPU …
			require_cuda(device)                                         # … running does
			integrator_params.pop("verbosity", None)
			self._integrator_config = None
			coupling = self._dense_coupling()
			self.integrator = _dense.CudaDenseSineIntegrator(*coupling[:2], method=method, driver=driver, device=device, nsteps=nsteps,
					coupling_parameter=len(coupling) == 3, **integrator_params)
			if len(coupling) == 3 and self.control_par_values is not None:
				self.integrator.set_parameters(self._parameter_rows())
			self._restore_state(old_integrator)
			return
		module = self._module(method, driver, storage, threads, kregs, min_blocks, team_warps)   # compiling needs no GPU …
		require_cuda(device)                                      # … running does
		self._integrator_config = (method, driver, storage, threads, kregs, min_blocks, team_warps)
		self._f_module = module
		integrator_params.pop("verbosity", None)
		self.integrator = CudaEnsembleIntegrator(
			module, device=device, nsteps=nsteps if backend == "ode" else 0,
			post=self._post, proj=self._proj, **integrator_params,
		)
		if self.control_pars and self.control_par_values is not None:
			self.integrator.set_parameters(self._parameter_rows())
		self._restore_state(old_integrator)

	def _restore_state(self, old_integrator):
		"""restore state and time from the previous integrator object, if applicable (_jitcode.py:619-626)"""
		on_device = hasattr(old_integrator, "y_device")
		try:
			old_y = old_integrator.y_device if on_device else old_integrator._y
			old_t = old_integrator.t
		except (AttributeError, RuntimeError):
			return
		if old_y is not None and len(old_y):
			self.integrator.set_initial_value(old_y, old_t)
			if on_device:
				self.integrator._out_kind, self.integrator._single = old_integrator._out_kind, old_integrator._single

	def integrate(self, *args, **kwargs):
		return self.integrator.integrate(*args, **kwargs)

	def integrate_many(self, times, record_states=True):
		"""`[integrate(T) for T in times]` in one kernel launch: (len(times), B, n); with record_states=False only the
		last state is kept and returned as (1, B, n)."""
		if not hasattr(self.integrator, "integrate_many"):
			raise RuntimeError("You must call set_integrator first.")
		return self.integrator.integrate_many(times, record_states)

	def integrate_streamed(self, initial_value, target_time, out, time=0.0, chunks=None):
		"""`set_initial_value` + `integrate` for an ensemble in pinned host memory, uploads and downloads overlapped with
		the kernels (see CudaEnsembleIntegrator.integrate_streamed)."""
		if not isinstance(self.integrator, CudaEnsembleIntegrator):
			raise RuntimeError("You must call set_integrator first (generic kernels only).")
		if initial_value.shape[-1] != self.n:
			raise ValueError("The dimension of the initial value does not match the dimension of your differential equations.")
		return self.integrator.integrate_streamed(initial_value, target_time, out, time, chunks)

	def successful(self):
		raise NotImplementedError("JiTCODE doesn’t support ODE’s `successful`. Instead the exception `UnsuccessfulIntegration` is raised in case of an unsuccessful integration (its `indices` attribute names the failing trajectories).")

	def step_counts(self):
		"""(accepted, rejected) step attempts of the whole ensemble so far — the benchmark's trajectory-steps."""
		return self.integrator.step_counts()


NORMS_WARNING = "Norms of perturbation vectors for Lyapunov exponents out of numerical bounds. You probably waited too long before renormalising and should call integrate with smaller intervals between steps (as renormalisations happen once with every call of integrate)."
NORM_WARNING = "Norm of perturbation vector for Lyapunov exponent out of numerical bounds. You probably waited too long before renormalising and should call integrate with smaller intervals between steps (as renormalisations happen once with every call of integrate)."


def _warn_if_not_finite(lyaps, message, stacklevel=3):
	"""the reference warns when a tangent-vector norm is not finite (_jitcode.py:747-748, 890-891, 944-945); here the
	norms stay on the device and the local exponents ln(norm)/Δt carry the same information."""
	if lyaps is None or bool(torch.isfinite(lyaps).all().item()):
		return
	bad = (~torch.isfinite(lyaps)).any(dim=-1)                 # (B,) or (samples, B)
	if bad.dim() == 2:
		bad = bad.any(dim=0)
	bad = bad.nonzero().flatten()
	warn(f"{message} (trajectories: {bad[:8].tolist()}{' …' if bad.numel() > 8 else ''})", stacklevel=stacklevel)


def random_directions(B, n, generator=None, device="cpu"):
	"""B isotropic unit vectors of dimension n (random_direction of jitcxde_common, _jitcode.py:729,873)."""
	vectors = torch.randn((B, n), dtype=torch.float64, generator=generator, device=device)
	return vectors / vectors.norm(dim=1, keepdim=True)


class _Generators:
	"""one seeded torch generator per device: tangent vectors of a CUDA-resident ensemble are drawn on the GPU"""
	def __init__(self, seed):
		self.seed, self.generators = seed, {}

	def on(self, device):
		if self.seed is None:
			return None
		device = torch.device(device)
		if device not in self.generators:
			self.generators[device] = torch.Generator(device=device).manual_seed(self.seed)
		return self.generators[device]


class jitcode_lyap(jitcode):
	"""
	As `jitcode`, plus (_jitcode.py:671-779):

	n_lyap : number of Lyapunov exponents; negative or larger than the dimension → all
	simplify : simplify the tangent-vector equations with SymPy (default: off, see generate_f_cuda)

	`integrate(t)` returns (state (B, n), local Lyapunov exponents (B, n_lyap), tangent vectors (B, n_lyap, n));
	the tangent vectors are integrated alongside the state in the same kernel and re-orthonormalised on the
	device after every call.  `seed` makes the random initial tangent vectors reproducible.
	"""

	_post = JB_POST_LYAP_QR

	def __init__(self, f_sym=(), n_lyap=-1, simplify=None, seed=None, **kwargs):
		n_basic = kwargs.pop("n", None)
		f_basic, self.n_basic = handle_input(f_sym, n_basic)
		self._n_lyap = n_lyap if 0 <= n_lyap <= self.n_basic else self.n_basic
		if self._n_lyap > 10:
			warn(f"You are about to calculate {self._n_lyap} Lyapunov exponents. This is very likely more than you need and may lead to severe difficulties with compilation and integration.", stacklevel=2)
		helpers = sort_helpers(kwargs.pop("helpers", None))
		full = lyapunov_system(f_basic, helpers, self.n_basic, self._n_lyap)
		self._f_basic = list(f_basic)
		if simplify:
			full = full[:self.n_basic] + [entry.simplify(ratio=1.0) for entry in full[self.n_basic:]]
		super().__init__(full, helpers=helpers, n=self.n_basic*(self._n_lyap+1), **kwargs)
		self._n_basic_module = self.n_basic
		self._n_lyap_module = self._n_lyap
		self._generators = _Generators(seed)

	def _with_tangent_vectors(self, state):
		if isinstance(state, dict):
			state = self._list_from_dynvar_dict(state, "initial value", self.n_basic)
		is_tensor = isinstance(state, torch.Tensor)
		rows = state if is_tensor else torch.as_tensor(np.ascontiguousarray(state, dtype=float))
		single = rows.dim() == 1
		rows2 = rows.reshape(1, -1) if single else rows
		if rows2.shape[1] != self.n_basic:
			raise ValueError("The dimension of the initial value does not match the dimension of your differential equations.")
		B = rows2.shape[0]
		# large host-resident ensembles: the tangent vectors are drawn where they are needed (100k × 16 normal deviates
		# cost 25 ms on the host, none to speak of on the GPU); what `integrate` returns keeps the kind of the input
		device = getattr(self.integrator, "device", None)
		kind = None
		if device is not None and not rows2.is_cuda and B >= 4096:
			kind = "torch" if is_tensor else "numpy"
			rows2 = rows2.to(device)
		generator = self._generators.on(rows2.device)
		tangents = torch.cat([random_directions(B, self.n_basic, generator, rows2.device) for _ in range(self._n_lyap)], dim=1)
		full = torch.cat([rows2.to(torch.float64), tangents], dim=1)
		if single:
			full = full[0]
		return (full if is_tensor or kind else full.numpy()), kind

	def set_initial_value(self, y, time=0.0):
		full, kind = self._with_tangent_vectors(y)
		super().set_initial_value(full, time)
		if kind:
			self.integrator._out_kind = kind

	def set_integrator(self, name, interpolate=None, **kwargs):
		"""
		As for `jitcode`.  The reference passes `interpolate` on positionally into the `nsteps` slot
		(_jitcode.py:741 vs :560), so that in effect the `ode` names run with SciPy's default step limit and the
		solve_ivp names always interpolate; that effective behaviour is kept (SURVEY.md §7 "reference quirks").
		"""
		kwargs.setdefault("nsteps", 0)
		super().set_integrator(name, interpolate=True, **kwargs)

	def _split(self, state):
		n, k = self.n_basic, self._n_lyap
		tangents = state[..., n:].reshape(*state.shape[:-1], k, n)
		return state[..., :n], tangents

	_warning = NORMS_WARNING

	def _local_exponents(self, integrator, target_time, time_before):
		"""device tensor (B, n_out) of the local exponents of the call that just ended.  A call that does not advance
		the time launches nothing: the reference then divides ln(norm) = ln(1) = 0 by Δt = 0 (_jitcode.py:771), i.e.
		returns NaN with NumPy's warning — the same is returned here."""
		if float(target_time) == time_before and integrator.driver == _codegen.JB_DRV_ODE:
			warn("The target time equals the current time: the local Lyapunov exponents are 0/0.", stacklevel=3)
			return torch.full((integrator.B, integrator.n_out), float("nan"), dtype=torch.float64, device=integrator.device)
		_warn_if_not_finite(integrator.last_lyaps, self._warning, stacklevel=4)
		return integrator.last_lyaps

	def integrate(self, target_time, *args, **kwargs):
		integrator = self.integrator
		time_before = integrator.t
		state = integrator.integrate(target_time, *args, **kwargs)
		lyaps = integrator._export(self._local_exponents(integrator, target_time, time_before))
		basic, tangents = self._split(state)
		return basic, lyaps, tangents

	def integrate_many(self, times, discard=0, record_states=True):
		"""All calls of `integrate` for `times` in one launch: returns (states (T, B, n), local exponents
		(T, B, n_lyap)); with record_states=False only the last state is returned, as (1, B, n)."""
		integrator = self.integrator
		integrator.lyap_discard = discard
		states = integrator.integrate_many(times, record_states)
		lyaps = integrator.last_lyaps
		_warn_if_not_finite(lyaps, self._warning)
		if integrator._single:
			lyaps = lyaps[:, 0]
		if integrator._out_kind != "cuda":
			lyaps = integrator._to_host(lyaps)
		return states[..., :self.n_basic], lyaps

	@property
	def y_dict(self):
		state = self.y
		return {self.dynvar(i): state[..., i] for i in range(self.n_basic)}


class jitcode_restricted_lyap(jitcode_lyap):
	"""
	Largest Lyapunov exponent orthogonal to a given plane (_jitcode.py:922-946).

	vectors : basis of the plane whose projection is removed from the tangent vector after every call
	"""

	_post = JB_POST_LYAP_RESTRICTED
	_warning = NORM_WARNING

	def __init__(self, f_sym=(), vectors=(), **kwargs):
		kwargs["n_lyap"] = 1
		super().__init__(f_sym, **kwargs)
		self.vectors = [np.asarray(vector, dtype=float)/np.linalg.norm(vector) for vector in vectors]
		self._proj = np.array(self.vectors, dtype=float).reshape(len(self.vectors), self.n_basic) if self.vectors else None


class jitcode_transversal_lyap(jitcode, GroupHandler):
	"""
	Largest Lyapunov exponent transversal to a synchronisation manifold (_jitcode.py:782-919): the system is
	transformed to one representative per group of synchronised variables plus difference coordinates of the
	tangent vector, which shrinks it considerably.

	groups : iterable of iterables of indices of variables that are identical on the manifold
	simplify : simplify the transformed equations with SymPy (default: off)
	average_dynamics : average the dynamics over each group instead of taking its first member

	`integrate(t)` returns (one value per group (B, n_groups), local exponent (B,)).
	"""

	_post = JB_POST_LYAP_TRANSVERSAL

	def __init__(self, f_sym=(), groups=(), simplify=None, average_dynamics=False, seed=None, **kwargs):
		GroupHandler.__init__(self, groups)
		n = kwargs.pop("n", None)
		f_list, n = handle_input(f_sym, n)
		if n != self.n_members:
			raise ValueError("Every dynamical variable must belong to exactly one group.")
		helpers = sort_helpers(kwargs.pop("helpers", None))
		z = sympy.Function("z")
		tangent_vector = self.back_transform([z(i) for i in range(n)])
		tangent_f = [
			sympy.Add(*[entry * tangent_vector[k] for k, entry in enumerate(row) if entry != 0])
			for row in jacobian_rows(f_list, helpers, n)
		]
		to_main = {y(i): z(self.map_to_main(i)) for i in range(n)}

		def finalise(entry):
			entry = sympy.sympify(entry).xreplace(to_main)
			if simplify:
				entry = entry.simplify(ratio=1.0)
			return entry.replace(z, y)

		f_lyap = []
		for entry in self.iterate(tangent_f):
			if isinstance(entry, int):
				group = self.groups[entry]
				if average_dynamics:
					f_lyap.append(sympy.Add(*[finalise(f_list[i]) for i in group]) / len(group))
				else:
					f_lyap.append(finalise(f_list[group[0]]))
			else:
				f_lyap.append(finalise(entry[0] - entry[1]))
		helpers = [(symbol, finalise(definition)) for symbol, definition in helpers]
		jitcode.__init__(self, f_lyap, helpers=helpers, n=n, **kwargs)
		self._n_basic_module = n
		self._n_lyap_module = 1
		self._tangent_indices_module = list(self.tangent_indices)
		self._generators = _Generators(seed)

	def set_initial_value(self, y, time=0.0):
		if isinstance(y, dict):
			raise NotImplementedError("Dictionary input is not supported for transversal Lyapunov exponents")
		is_tensor = isinstance(y, torch.Tensor)
		rows = y if is_tensor else torch.as_tensor(np.ascontiguousarray(y, dtype=float))
		single = rows.dim() == 1
		rows2 = (rows.reshape(1, -1) if single else rows).to(torch.float64)
		if rows2.shape[1] != len(self.groups):
			raise ValueError("Provide exactly one value per synchronisation group.")
		B = rows2.shape[0]
		full = torch.empty((B, self.n), dtype=torch.float64, device=rows2.device)
		full[:, self.main_indices] = rows2
		full[:, self.tangent_indices] = random_directions(B, len(self.tangent_indices), self._generators.on(rows2.device), rows2.device)
		if single:
			full = full[0]
		super().set_initial_value(full if is_tensor else full.numpy(), time)

	def set_integrator(self, name, interpolate=False, **kwargs):
		"""As for `jitcode`; like the reference (_jitcode.py:884) the positional hand-over makes the `ode` names
		use SciPy's default step limit and the solve_ivp names interpolate."""
		kwargs.setdefault("nsteps", 0)
		super().set_integrator(name, interpolate=True, **kwargs)

	_warning = NORM_WARNING
	_local_exponents = jitcode_lyap._local_exponents

	def integrate(self, target_time, *args, **kwargs):
		integrator = self.integrator
		time_before = integrator.t
		state = integrator.integrate(target_time, *args, **kwargs)
		lyap = integrator._export(self._local_exponents(integrator, target_time, time_before))[..., 0]
		return state[..., self.main_indices], lyap

	def integrate_many(self, times, discard=0):
		"""All calls of `integrate` for `times` in one launch: (states (T, B, n_groups), local exponents (T, B))."""
		integrator = self.integrator
		integrator.lyap_discard = discard
		states = integrator.integrate_many(times)
		_warn_if_not_finite(integrator.last_lyaps, self._warning)
		lyaps = integrator.last_lyaps[..., 0]
		if integrator._single:
			lyaps = lyaps[:, 0]
		if integrator._out_kind != "cuda":
			lyaps = integrator._to_host(lyaps)
		return states[..., self.main_indices], lyaps

	@property
	def y_dict(self):
		raise NotImplementedError("Dictionary output is not supported for transversal Lyapunov exponents")


def test(omp=True, sympy=True):
	"""Quick sanity check of the installation (reference :948-965): generate, compile with nvcc, load, and —
	on a machine with a GPU — integrate a harmonic oscillator."""
	ODE = jitcode([y(1), -y(0)], verbose=False)
	ODE.generate_f_cuda(chunk_size=1)
	ODE.compile_cuda()
	ODE.set_integrator("dopri5")
	ODE.set_initial_value([1, 2])
	ODE.integrate(0.1)
