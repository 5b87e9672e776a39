"""
Dense sine-coupled phase oscillators — BASELINE config 5 (reference examples/kuramoto_network.py:7-37).

    dy_i/dt = ω_i + Σ_j W_ij sin(y_j − y_i)

The reference writes the coupling out term by term (n·deg sines of generated C); for a dense network that
is the one place where the right-hand side really is a matrix product (csrc/jb_dense.cu):
sin(y_j − y_i) = sin y_j cos y_i − cos y_j sin y_i turns it into W·[sin Y | cos Y] for the whole ensemble,
computed on the FP64 tensor cores.

  * `detect_sine_coupling` recognises the structure in the symbolic input (so that the reference's API keeps
    working unchanged: `jitcode(kuramoto_f, n=n)` + `set_integrator("dopri5")`);
  * `CudaDenseSineIntegrator` is the integrator object (protocol of integrator_tools.py, like
    `_integrator.CudaEnsembleIntegrator`) on top of the C ABI `jb_dense_*` of include/jitcode_b200.h.
"""

import ctypes

import numpy as np
import sympy
import torch

from . import _build
from ._integrator import STATUS_TEXT, UnsuccessfulIntegration, require_cuda
from ._symbolic import is_dynvar


def detect_sine_coupling(f, control_pars=()):
	"""(W, omega) if every equation is  constant + Σ_j number · sin(±(y(j) − y(i))),  else None.
	With control parameters: (W, omega, c) if ONE of them multiplies every coupling term (and nothing else contains any) —
	dy_i/dt = ω_i + c · Σ_j W_ij sin(y_j − y_i), the coupling strength as a per-trajectory parameter."""
	n = len(f)
	W = np.zeros((n, n))
	omega = np.zeros(n)
	control_pars = list(control_pars)
	factor = None
	found = False
	for i, expr in enumerate(f):
		expr = sympy.expand(expr) if control_pars else expr
		for term in sympy.Add.make_args(expr):
			coefficient, rest = term.as_coeff_Mul()
			if not coefficient.is_Number:
				return None
			if rest is sympy.S.One:
				omega[i] += float(coefficient)
				continue
			if control_pars and rest.is_Mul:
				# c · sin(…): split off the control parameter
				symbols = [a for a in rest.args if a in control_pars]
				others = [a for a in rest.args if a not in control_pars]
				if len(symbols) != 1 or len(others) != 1 or (factor is not None and symbols[0] != factor):
					return None
				factor, rest = symbols[0], others[0]
			elif control_pars:
				return None           # a coupling term without the parameter, next to terms with it: not this form
			if rest.func is not sympy.sin:
				return None
			parts = sympy.Add.make_args(rest.args[0])
			if len(parts) != 2:
				return None
			(c1, v1), (c2, v2) = parts[0].as_coeff_Mul(), parts[1].as_coeff_Mul()
			if not (is_dynvar(v1) and is_dynvar(v2)) or {c1, c2} != {sympy.S.One, sympy.S.NegativeOne}:
				return None
			plus, minus = (v1, v2) if c1 == 1 else (v2, v1)
			a, b = int(plus.args[0]), int(minus.args[0])
			if b == i:        # sin(y_a − y_i)
				W[i, a] += float(coefficient)
			elif a == i:      # sin(y_i − y_b) = −sin(y_b − y_i)
				W[i, b] -= float(coefficient)
			else:
				return None
			found = True
	if not found or (control_pars and (factor is None or len(control_pars) != 1)):
		return None
	return (W, omega, factor) if control_pars else (W, omega)


class DenseArgs(ctypes.Structure):
	_fields_ = [
		("struct_size", ctypes.c_uint32), ("n", ctypes.c_int32),
		("B", ctypes.c_int64), ("ld", ctypes.c_int64),
		("W", ctypes.c_void_p), ("omega", ctypes.c_void_p), ("coupling", ctypes.c_void_p), ("atol_vec", ctypes.c_void_p),
		("T", ctypes.c_double),
		("rtol", ctypes.c_double), ("atol", ctypes.c_double), ("max_step", ctypes.c_double), ("first_step", ctypes.c_double),
		("nsteps", ctypes.c_int64),
		("safety", ctypes.c_double), ("ifactor", ctypes.c_double), ("dfactor", ctypes.c_double), ("beta", ctypes.c_double),
		("t", ctypes.c_void_p), ("Y", ctypes.c_void_p), ("work", ctypes.c_void_p), ("ctrl", ctypes.c_void_p),
		("status", ctypes.c_void_p), ("n_accepted", ctypes.c_void_p), ("n_rejected", ctypes.c_void_p),
		("scratch", ctypes.c_void_p), ("host_flag", ctypes.c_void_p),
		("rounds_per_check", ctypes.c_int32), ("method_flags", ctypes.c_int32),
		("rounds_out", ctypes.c_void_p),
		("stream", ctypes.c_void_p),
	]

DENSE_EXPORTS = ("jb_dense_work_vectors", "jb_dense_control_scalars", "jb_dense_result_vector", "jb_dense_last_error", "jb_dense_launch_count",
		"jb_dense_integrate", "jb_dense_eval_f", "jb_dense_gemm", "jb_dense_gemm_shape", "jb_dense_transpose")

_library = None

def dense_library():
	"""compiles (once, cached like every module) and loads csrc/jb_dense.cu."""
	global _library
	if _library is None:
		path = _build.compile_module((_build.CSRC_DIR / "jb_dense.cu").read_text())
		lib = ctypes.CDLL(str(path))
		for name in DENSE_EXPORTS:
			if not hasattr(lib, name):
				raise _build.CudaModuleError(f"{path} does not export {name}")
		lib.jb_dense_last_error.restype = ctypes.c_char_p
		lib.jb_dense_launch_count.restype = ctypes.c_int64
		lib.jb_dense_integrate.argtypes = [ctypes.POINTER(DenseArgs)]
		lib.jb_dense_eval_f.argtypes = [ctypes.POINTER(DenseArgs), ctypes.c_void_p]
		lib.jb_dense_gemm.argtypes = [ctypes.c_void_p]*3 + [ctypes.c_int]*3 + [ctypes.c_void_p]
		lib.jb_dense_gemm_shape.argtypes = [ctypes.c_void_p]*3 + [ctypes.c_int]*4 + [ctypes.c_void_p]
		lib.jb_dense_transpose.argtypes = [ctypes.c_void_p]*2 + [ctypes.c_int64]*4 + [ctypes.c_void_p]
		lib.path = str(path)
		_library = lib
	return _library


def _check(lib, code, what):
	if code != 0:
		raise _build.CudaModuleError(f"{what} failed with code {code}: {lib.jb_dense_last_error().decode(errors='replace')}")


class DenseInfo:
	"""what CudaEnsembleIntegrator users read from module.info"""
	def __init__(self, n):
		self.n, self.n_params, self.n_basic, self.n_lyap, self.storage = n, 0, n, 0, 4
		self.threads_per_block, self.shared_bytes = 256, 0


class DenseModule:
	"""stands where a generated module stands (launch counter, info), for bench.py and tests"""
	def __init__(self, n):
		self.lib = dense_library()
		self.path = self.lib.path
		self.info = DenseInfo(n)

	def launch_count(self):
		return int(self.lib.jb_dense_launch_count())


class CudaDenseSineIntegrator:
	"""B trajectories of dy_i/dt = ω_i + Σ_j W_ij sin(y_j − y_i) on one GPU; Hairer's DOPRI5 ("dopri5") or DOP853 ("dop853")
	(driver 0: every `integrate` call is a fresh start that lands on the target time, ODE_wrapper), or SciPy's solve_ivp
	stepper with the same pairs ("RK45", "DOP853"): persistent between calls, driver 1 = the sample is the dense output
	of the step that passed the target time (IVP_wrapper), driver 2 = the last step is clipped to the target time
	(IVP_wrapper_no_interpolation)."""

	def __init__(self, W, omega, *, method=0, driver=0, device=None, rtol=1e-6, atol=1e-12, max_step=0.0, first_step=0.0, nsteps=10**6,
			safety=0.0, ifactor=0.0, dfactor=0.0, beta=0.0, verbatim_dop853_reject=True, rounds_per_check=0, coupling_parameter=False,
			interleave=0):
		self.device = require_cuda(device)
		self.n = len(omega)
		self.module = DenseModule(self.n)
		self.lib = self.module.lib
		self.W = torch.as_tensor(np.ascontiguousarray(W, dtype=float)).to(self.device)
		self.omega = torch.as_tensor(np.ascontiguousarray(omega, dtype=float)).to(self.device)
		if np.ndim(atol) == 0:
			self.atol, self.atol_vec = float(atol), None
		else:
			atol = np.asarray(atol, dtype=float)
			if atol.shape != (self.n,):
				raise ValueError("atol must be a scalar or have one entry per component.")
			self.atol, self.atol_vec = 0.0, torch.as_tensor(atol, device=self.device)
		self.rtol = float(rtol)
		# dy_i/dt = ω_i + c·Σ_j W_ij sin(y_j − y_i) with c a control parameter (one value per trajectory)
		self.n_params = 1 if coupling_parameter else 0
		self._pars = None
		self.max_step = float(max_step) if max_step and np.isfinite(max_step) else 0.0
		self.first_step = float(first_step) if first_step else 0.0
		self.nsteps = int(nsteps)
		self.controller = (float(safety), float(ifactor), float(dfactor), float(beta))
		self.rounds_per_check = int(rounds_per_check)
		# interleave: number of sub-ensembles whose launch sequences overlap on separate streams (0: default — up to
		# four of ≥ 1024 trajectories each; 1: none; csrc/jb_dense.cu, integrate)
		self.driver = int(driver)
		self.method_flags = int(method) | (256 if verbatim_dop853_reject else 0) | ((int(interleave) & 7) << 9) | (self.driver << 12)
		self._fresh = True      # the solve_ivp stepper has to be created by the next call (a new state was set)
		self.B = None
		self.time = None
		self._rows = None
		self._out_kind, self._single = "numpy", False
		self.kernel_events = None
		self.rounds = 0

	@property
	def stream(self):
		return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

	def _alloc(self, B):
		self.B, self.ld = B, (B + 63) // 64 * 64
		n, ld, dev = self.n, self.ld, self.device
		f64 = {"dtype": torch.float64, "device": dev}
		self.Y = torch.zeros(n * ld, **f64)
		self.tt = torch.zeros(ld, **f64)
		self.work = torch.zeros(self.lib.jb_dense_work_vectors() * n * ld, **f64)
		self.ctrl = torch.zeros(self.lib.jb_dense_control_scalars() * ld, **f64)
		self.status = torch.zeros(ld, dtype=torch.int32, device=dev)
		self.n_accepted = torch.zeros(ld, dtype=torch.int64, device=dev)
		self.n_rejected = torch.zeros(ld, dtype=torch.int64, device=dev)
		self.scratch = torch.zeros(ld // 32 + 8, dtype=torch.int32, device=dev)
		self.host_flag = torch.zeros(4, dtype=torch.int32).pin_memory()
		self.rounds_host = np.zeros(1, dtype=np.int64)
		self.coupling = torch.ones(ld, **f64) if self.n_params else None
		self._pars_on_device = False

	def _args(self, T):
		a = DenseArgs()
		a.struct_size = ctypes.sizeof(DenseArgs)
		a.n, a.B, a.ld = self.n, self.B, self.ld
		a.W, a.omega = self.W.data_ptr(), self.omega.data_ptr()
		a.coupling = None if self.coupling is None else self.coupling.data_ptr()
		a.atol_vec = None if self.atol_vec is None else self.atol_vec.data_ptr()
		a.T = T
		a.rtol, a.atol, a.max_step, a.first_step = self.rtol, self.atol, self.max_step, self.first_step
		a.nsteps = self.nsteps
		a.safety, a.ifactor, a.dfactor, a.beta = self.controller
		a.t, a.Y, a.work, a.ctrl = self.tt.data_ptr(), self.Y.data_ptr(), self.work.data_ptr(), self.ctrl.data_ptr()
		a.status, a.n_accepted, a.n_rejected = self.status.data_ptr(), self.n_accepted.data_ptr(), self.n_rejected.data_ptr()
		a.scratch, a.host_flag = self.scratch.data_ptr(), self.host_flag.data_ptr()
		a.rounds_per_check = self.rounds_per_check
		a.method_flags = self.method_flags | ((1 << 14) if (self.driver and self._fresh) else 0)
		a.rounds_out = self.rounds_host.ctypes.data
		a.stream = self.stream
		return a

	def _to_device_rows(self, array):
		kind = "numpy"
		if isinstance(array, torch.Tensor):
			kind = "cuda" if array.is_cuda else "torch"
			rows = array.to(device=self.device, dtype=torch.float64, non_blocking=True)
		else:
			rows = torch.as_tensor(np.ascontiguousarray(array, dtype=float)).to(self.device, non_blocking=True)
		single = rows.dim() == 1
		if single:
			rows = rows.reshape(1, -1)
		if rows.dim() != 2 or rows.shape[1] != self.n:
			raise ValueError("The dimension of the initial value does not match the dimension of your differential equations.")
		return rows.contiguous(), kind, single

	def _export(self, rows, out=None):
		if self._single:
			rows = rows[0]
		if out is not None:
			out.copy_(rows, non_blocking=True)
			torch.cuda.current_stream(self.device).synchronize()
			return out
		if self._out_kind == "cuda":
			return rows
		return self._to_host(rows)

	def _to_host(self, rows):
		"""CPU tensor or NumPy array (the kind of the initial value); large results go through page-locked memory"""
		if rows.numel() >= 1 << 20 and rows.is_cuda:
			host = torch.empty(rows.shape, dtype=rows.dtype, pin_memory=True)
			host.copy_(rows, non_blocking=True)
			torch.cuda.current_stream(self.device).synchronize()
		else:
			host = rows.cpu()
		return host if self._out_kind == "torch" else host.numpy()

	def _rows_from(self, soa, rows=None):
		if rows is None:
			rows = torch.empty((self.B, self.n), dtype=torch.float64, device=self.device)
		_check(self.lib, self.lib.jb_dense_transpose(soa.data_ptr(), rows.data_ptr(), self.n, self.B, self.ld, self.n, self.stream), "jb_dense_transpose")
		return rows

	# ---- protocol of integrator_tools

	def set_parameters(self, pars):
		"""the coupling strength: one value for all trajectories or one per trajectory ((B,), (B, 1), (1,) …)"""
		if not self.n_params:
			return
		pars = pars.detach().to(torch.float64) if isinstance(pars, torch.Tensor) else torch.as_tensor(np.asarray(pars, dtype=float))
		self._pars = pars.reshape(-1)
		self._pars_on_device = False
		if self.B is not None:
			self._upload_parameters()

	def _upload_parameters(self):
		if not self.n_params or self._pars_on_device:
			return
		if self._pars is None:
			raise RuntimeError("Something needs parameters to be set. Try calling `set_parameters` earlier.")
		values = self._pars.to(self.device)
		if values.numel() == 1:
			values = values.expand(self.B)
		if values.numel() != self.B:
			raise ValueError(f"Got parameters for {values.numel()} trajectories, but initial values for {self.B}.")
		self.coupling[:self.B].copy_(values)
		self._pars_on_device = True

	def set_initial_value(self, initial_value, time=0.0):
		rows, kind, single = self._to_device_rows(initial_value)
		if self.B != rows.shape[0]:
			self._alloc(rows.shape[0])
		self._out_kind, self._single = kind, single
		_check(self.lib, self.lib.jb_dense_transpose(rows.data_ptr(), self.Y.data_ptr(), self.B, self.n, self.n, self.ld, self.stream), "jb_dense_transpose")
		self.tt.fill_(float(time))
		self.time = float(time)
		self.status.zero_()
		self.ctrl.zero_()       # (the persistent stepper of the solve_ivp drivers lives there)
		self._fresh = True
		self._rows = rows
		if self._pars is not None:
			self._upload_parameters()

	@property
	def t(self):
		if self.time is None:
			raise RuntimeError("You must call set_initial_value first.")
		return self.time

	@property
	def _y(self):
		if self._rows is None:
			raise RuntimeError("You must call set_initial_value first.")
		return self._export(self._rows)

	@property
	def y_device(self):
		return self._rows

	def successful(self):
		return not bool((self.status[:self.B] != 0).any().item())

	def _advance(self, T, rows=None):
		"""one call of jb_dense_integrate: every trajectory to T; the state there as (B, n) device rows (`rows`: where to)"""
		self._upload_parameters()
		args = self._args(T)
		if self.kernel_events is not None:
			start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			start.record()
		_check(self.lib, self.lib.jb_dense_integrate(ctypes.byref(args)), "jb_dense_integrate")
		if self.kernel_events is not None:
			end.record()
			self.kernel_events.append((start, end))
		self.rounds += int(self.rounds_host[0])
		self.time = T
		self._fresh = False
		if self.driver == 1:      # dense output: the sample is not the solver's state
			width = self.n * self.ld
			result = self.lib.jb_dense_result_vector()
			self._rows = self._rows_from(self.work[result * width:(result + 1) * width], rows)
		else:
			self._rows = self._rows_from(self.Y, rows)

	def _check_failures(self):
		failed = (self.status[:self.B] != 0).nonzero().flatten()
		if failed.numel():
			indices = failed.cpu().numpy()
			status = self.status[:self.B][failed].cpu().numpy()
			raise UnsuccessfulIntegration(
				f"{len(indices)} of {self.B} trajectories failed (first: #{indices[0]}, {STATUS_TEXT.get(int(status[0]), status[0])}).", indices, status)

	def integrate(self, T, out=None):
		T = float(T)
		if self.B is None:
			raise RuntimeError("You must call set_initial_value first.")
		# (IVP_wrapper.integrate has no such check: an earlier time is read off the last step's interpolant, integrator_tools.py:98-105)
		if T < self.time and self.driver != 1:
			raise ValueError("Target time smaller than current time. Cannot integrate backwards in time")
		if T == self.time and (self.driver != 1 or not self._fresh):
			return self._export(self._rows, out)
		self._advance(T)
		self._check_failures()
		return self._export(self._rows, out)

	def integrate_many(self, times, record_states=True):
		"""the loop `for T in times: integrate(T)` (examples/kuramoto_network.py:34-35) without leaving the device between
		the calls: (len(times), B, n) in the kind of the initial value; record_states=False keeps only the last state (1, B, n)."""
		times = [float(T) for T in times]
		if self.B is None:
			raise RuntimeError("You must call set_initial_value first.")
		if any(later < earlier for earlier, later in zip(([] if self.driver == 1 else [self.time]) + times, times)):
			raise ValueError("Target time smaller than current time. Cannot integrate backwards in time")
		out = torch.empty((len(times) if record_states else 1, self.B, self.n), dtype=torch.float64, device=self.device)
		# results for the host: every sample starts its way there as soon as it exists, on a second stream, while the
		# ensemble integrates towards the next one (config 5: 1.6 GB of samples, a fifth of the call if copied at the end)
		stream_out = self._out_kind != "cuda" and out.numel() >= 1 << 20
		if stream_out:
			host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
			main, side = torch.cuda.current_stream(self.device), self._copy_stream()
		for i, T in enumerate(times):
			k = i if record_states else 0
			slot = out[k]
			if T == self.time and (self.driver != 1 or not self._fresh):
				slot.copy_(self._rows)
			else:
				self._advance(T, slot)
			if stream_out and (record_states or i == len(times) - 1):
				side.wait_stream(main)
				with torch.cuda.stream(side):
					host[k].copy_(slot, non_blocking=True)
		self._rows = out[-1]
		self._check_failures()
		if stream_out:
			side.synchronize()
			out.record_stream(side)
			if self._single:
				host = host[:, 0]
			return host if self._out_kind == "torch" else host.numpy()
		if self._single:
			out = out[:, 0]
		return out if self._out_kind == "cuda" else self._to_host(out)

	def _copy_stream(self):
		if getattr(self, "_side_stream", None) is None:
			self._side_stream = torch.cuda.Stream(self.device)
		return self._side_stream

	def eval_f(self, state):
		rows, kind, single = self._to_device_rows(state)
		if self.B != rows.shape[0]:
			self._alloc(rows.shape[0])
		self._out_kind, self._single = kind, single
		_check(self.lib, self.lib.jb_dense_transpose(rows.data_ptr(), self.Y.data_ptr(), self.B, self.n, self.n, self.ld, self.stream), "jb_dense_transpose")
		self.status.zero_()
		self._pars_on_device = False
		if self.n_params:
			self._upload_parameters()
		out = torch.empty_like(self.Y)
		args = self._args(0.0)
		_check(self.lib, self.lib.jb_dense_eval_f(ctypes.byref(args), out.data_ptr()), "jb_dense_eval_f")
		return self._export(self._rows_from(out))

	def step_counts(self):
		return int(self.n_accepted[:self.B].sum().item()), int(self.n_rejected[:self.B].sum().item())

	def reset_step_counts(self):
		self.n_accepted.zero_()
		self.n_rejected.zero_()
		self.rounds = 0
