"""
Integrator objects: the plug-in slot of the reference (jitcode/integrator_tools.py) with a batch axis.

`jitcode.set_integrator` creates one of the reference's wrappers — ODE_wrapper (:130-151),
IVP_wrapper (:39-108) or IVP_wrapper_no_interpolation (:110-128) — and talks to it through
`set_initial_value(y, t)`, `integrate(t)`, `._y`, `.t`.  `CudaEnsembleIntegrator` offers the same
members for B trajectories at once: one kernel launch per `integrate` call advances the whole
ensemble on the GPU, with the semantics of the respective reference wrapper selected by `driver`.
`empty_integrator` (:153-184) is the state holder used before `set_integrator`.

There is no CPU path: constructing the integrator without a CUDA device raises.
"""

import ctypes
import os

import numpy as np
import torch

from ._build import JB_TILE, IntegrateArgs
from ._codegen import JB_DRV_IVP_DENSE, JB_DRV_IVP_LAND, JB_DRV_ODE

JB_POST_NONE, JB_POST_LYAP_QR, JB_POST_LYAP_RESTRICTED, JB_POST_LYAP_TRANSVERSAL = 0, 1, 2, 3
STATUS_TEXT = {0: "ok", 1: "step size too small", 2: "too many steps", 3: "problem is probably stiff", 4: "non-finite step size"}


class UnsuccessfulIntegration(Exception):
	"""
	Raised when the integrator cannot meet the accuracy and step-size requirements for at least one
	trajectory (integrator_tools.py:9-13).  `indices` holds the failing trajectories, `status` their codes.
	"""
	def __init__(self, message="", indices=None, status=None):
		super().__init__(message)
		self.indices = indices
		self.status = status


class empty_integrator:
	"""state holder before set_integrator (integrator_tools.py:153-184)."""

	def __init__(self):
		self.params = ()
		self._y = []
		self._t = None

	@property
	def t(self):
		if self._t is None:
			raise RuntimeError("You must call set_initial_value first.")
		return self._t

	def set_integrator(self, *args, **kwargs):
		raise RuntimeError("This function must not be called")

	def set_initial_value(self, initial_value, time=0.0):
		self._y = initial_value
		self._t = time

	def set_params(self, *args):
		raise NotImplementedError("This method should not be called anymore")

	def integrate(self, *args, **kwargs):
		raise RuntimeError("You must call set_integrator first.")

	def successful(self):
		raise RuntimeError("You must call set_integrator first.")


def require_cuda(device=None):
	if not torch.cuda.is_available():
		raise RuntimeError("jitcode_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
	return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def padded(B):
	return (B + JB_TILE - 1) // JB_TILE * JB_TILE


class CudaEnsembleIntegrator:
	"""
	B trajectories of one compiled system on one GPU.

	Buffers (all torch CUDA tensors, owned here, handed to the module as raw pointers; tiled layout
	of include/jitcode_b200.h): Y [n][ld], t [ld], P [p][ld], state / sstate (persistent stepper of the
	solve_ivp drivers), work (stages, global storage only), Yres (sample of the dense driver).
	"""

	def __init__(self, module, *, device=None, rtol=None, atol=None, max_step=0.0, first_step=0.0, nsteps=10**6,
			safety=0.0, ifactor=0.0, dfactor=0.0, beta=0.0, verbatim_dop853_reject=True,
			post=JB_POST_NONE, proj=None):
		self.device = require_cuda(device)
		self.module = module
		info = module.info
		self.n, self.n_params, self.driver = info.n, info.n_params, info.driver
		self.n_basic, self.n_lyap = info.n_basic, info.n_lyap
		ode = self.driver == JB_DRV_ODE
		# defaults of the respective SciPy back end (scipy _ode.py dopri5: rtol 1e-6, atol 1e-12; solve_ivp: 1e-3, 1e-6)
		self.rtol = float(rtol if rtol is not None else (1e-6 if ode else 1e-3))
		atol = atol if atol is not None else (1e-12 if ode else 1e-6)
		if np.ndim(atol) == 0:
			self.atol, self.atol_vec = float(atol), None
		else:
			atol = np.asarray(atol, dtype=float)
			if atol.shape != (self.n,):
				raise ValueError("atol must be a scalar or have one entry per component.")
			self.atol, self.atol_vec = 0.0, torch.as_tensor(atol, device=self.device)
		self.max_step = float(max_step) if max_step and np.isfinite(max_step) else 0.0
		self.first_step = float(first_step) if first_step else 0.0
		self.nsteps = int(nsteps)
		self.controller = (float(safety), float(ifactor), float(dfactor), float(beta))
		self.verbatim_dop853_reject = bool(verbatim_dop853_reject)
		self.post = post
		self.n_out = {JB_POST_NONE: 0, JB_POST_LYAP_QR: self.n_lyap}.get(post, 1)
		self.proj = None if proj is None else torch.as_tensor(np.ascontiguousarray(proj, dtype=float), device=self.device)
		self.lyap_discard = 0
		self.B = None
		self.time = None
		self._pars = None          # (B or 1, p) host array until the batch size is known
		self._rows = None          # cached (B,n) device rows of the current state
		self._out_kind = "numpy"
		self._single = False
		self.last_lyaps = None
		self._sampled = False      # an integrate call has produced a sample since set_initial_value
		self.kernel_events = None   # set to [] to collect (start, end) CUDA events around every integrate launch

	# ------------------------------------------------------------------ plumbing

	@property
	def stream(self):
		return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

	def _alloc(self, B):
		info = self.module.info
		self.B, self.ld = B, padded(B)
		n, ld, dev = self.n, self.ld, self.device
		f64 = {"dtype": torch.float64, "device": dev}
		self.Y = torch.zeros(n * ld, **f64)
		self.tt = torch.zeros(ld, **f64)
		self.status = torch.zeros(ld, dtype=torch.int32, device=dev)
		self.n_accepted = torch.zeros(ld, dtype=torch.int64, device=dev)
		self.n_rejected = torch.zeros(ld, dtype=torch.int64, device=dev)
		self.P = torch.zeros(max(1, self.n_params) * ld, **f64)
		self.state = torch.zeros(max(1, info.state_vectors) * n * ld, **f64) if info.state_vectors else None
		self.sstate = torch.zeros(max(1, info.state_scalars) * ld, **f64) if info.state_scalars else None
		self.work = torch.empty(info.work_vectors * n * ld, **f64) if info.work_vectors else None
		self.Yres = torch.zeros(n * ld, **f64) if self.driver == JB_DRV_IVP_DENSE else None
		self.lyap_acc = torch.zeros(2 * self.n_out * ld, **f64) if self.n_out else None
		self._times_host = torch.empty(1, dtype=torch.float64).pin_memory()
		self._pars_on_device = False
		# the counter from which persistent kernels draw trajectories: one per integrator (its launches share a stream)
		self.queue = None if os.environ.get("JITCODE_B200_NO_QUEUE") else torch.zeros(1, dtype=torch.int64, device=dev)

	def _to_device_rows(self, array, width, what):
		"""(B, width) float64 contiguous CUDA tensor from numpy / torch / list input; returns (tensor, kind, single)."""
		kind = "numpy"
		if isinstance(array, torch.Tensor):
			kind = "cuda" if array.is_cuda else "torch"
			rows = array.to(device=self.device, dtype=torch.float64, non_blocking=True)
		else:
			rows = torch.as_tensor(np.ascontiguousarray(array, dtype=float)).to(self.device, non_blocking=True)
		single = rows.dim() == 1
		if single:
			rows = rows.reshape(1, -1)
		if rows.dim() != 2 or rows.shape[1] != width:
			raise ValueError(f"The dimension of the {what} does not match: expected (B, {width}), got {tuple(rows.shape)}.")
		return rows.contiguous(), kind, single

	def _pack(self, rows, tiled, width):
		self.module.pack(rows.data_ptr(), tiled.data_ptr(), rows.shape[0], width, self.ld, self.stream)

	def _unpack(self, tiled, width, rows=None):
		"""tiled device vector → (B, width) device rows; `rows`: a contiguous (B, width) CUDA tensor to write into (for
		example this rank's slot of an all-gather buffer) instead of a fresh one."""
		if rows is None:
			rows = torch.empty((self.B, width), dtype=torch.float64, device=self.device)
		elif not (rows.is_cuda and rows.dtype == torch.float64 and rows.is_contiguous() and tuple(rows.shape) == (self.B, width)):
			raise ValueError(f"The device buffer for the result must be a contiguous ({self.B}, {width}) float64 CUDA tensor.")
		self.module.unpack(tiled.data_ptr(), rows.data_ptr(), self.B, width, self.ld, self.stream)
		return rows

	def _export(self, rows, out=None):
		"""device rows → what the caller gets (mirrors the kind of the initial value); `out`: a (pinned) CPU
		tensor of the right shape to receive the copy instead of a fresh array."""
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
		"""CPU tensor or NumPy array (the kind of the initial value) of a device tensor."""
		if rows.numel() >= 1 << 20 and rows.is_cuda:
			# large results go through page-locked memory (torch caches such blocks): 64 MB of exponents arrive in 3 ms
			# instead of 30; the array that is returned keeps its block alive
			host = torch.empty(rows.shape, dtype=rows.dtype, pin_memory=True)
			host.copy_(rows, non_blocking=True)
			torch.cuda.current_stream(self.device).synchronize()
		else:
			host = rows.cpu()
		return host if self._out_kind == "torch" else host.numpy()

	# ------------------------------------------------------------------ protocol of integrator_tools

	def set_params(self, *args):
		raise NotImplementedError("This method should not be called anymore")

	def set_parameters(self, pars):
		"""pars: (p,) for all trajectories or (B, p) per trajectory (numpy / torch)."""
		if self.n_params == 0:
			return
		if isinstance(pars, torch.Tensor):
			pars = pars.detach()
		else:
			pars = np.asarray(pars, dtype=float)
		if pars.ndim == 1:
			pars = pars.reshape(1, -1)
		if pars.shape[1] != self.n_params:
			raise ValueError("Number of values does not match number of control parameters.")
		self._pars = pars
		self._pars_on_device = False
		if self.B is not None:
			self._upload_parameters()

	def _upload_parameters(self):
		if self.n_params == 0 or self._pars_on_device:
			return
		if self._pars is None:
			raise RuntimeError("Something needs parameters to be set. Try calling `set_parameters` earlier.")
		rows, _, _ = self._to_device_rows(self._pars, self.n_params, "control parameters")
		if rows.shape[0] == 1 and self.B > 1:
			rows = rows.expand(self.B, -1).contiguous()
		if rows.shape[0] != self.B:
			raise ValueError(f"Got parameters for {rows.shape[0]} trajectories, but initial values for {self.B}.")
		self._pack(rows, self.P, self.n_params)
		self._pars_on_device = True
		if self.sstate is not None:
			# the solve_ivp stepper is re-created with the new right-hand side — from the last SAMPLE (t, y(t)),
			# which for the dense driver is not where the stepper itself stands (it has stepped beyond t)
			if self.driver == JB_DRV_IVP_DENSE and self._sampled:
				self.Y.copy_(self.Yres)
				self.tt.fill_(self.time)
			self.sstate.zero_()

	def set_initial_value(self, initial_value, time=0.0):
		rows, kind, single = self._to_device_rows(initial_value, self.n, "initial value")
		if self.B != rows.shape[0]:
			self._alloc(rows.shape[0])
		self._out_kind, self._single = kind, single
		self._pack(rows, self.Y, self.n)
		self.tt.fill_(float(time))
		self.time = float(time)
		self.status.zero_()
		self._sampled = False
		if self.sstate is not None:
			self.sstate.zero_()       # IVP_wrapper.set_initial_value re-creates the stepper (:90-93)
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
		"""(B, n) CUDA tensor of the current state, whatever the kind of the initial value was."""
		return self._rows

	def successful(self):
		return not bool((self.status[:self.B] != 0).any().item())

	def _args(self, times_dev, n_times, Yrec=None, lyap_rec=None):
		a = IntegrateArgs()
		a.post = self.post
		a.B, a.ld = self.B, self.ld
		a.n_times = n_times
		a.n_proj = 0 if self.proj is None else self.proj.shape[0]
		a.times = times_dev.data_ptr()
		a.rtol, a.atol = self.rtol, self.atol
		a.atol_vec = None if self.atol_vec is None else self.atol_vec.data_ptr()
		a.max_step, a.first_step = self.max_step, self.first_step
		a.nsteps = self.nsteps
		a.safety, a.ifactor, a.dfactor, a.beta = self.controller
		a.hairer_reject_by_dfactor = int(self.verbatim_dop853_reject)
		a.lyap_discard = self.lyap_discard
		a.t, a.Y = self.tt.data_ptr(), self.Y.data_ptr()
		a.P = self.P.data_ptr()
		a.state = None if self.state is None else self.state.data_ptr()
		a.sstate = None if self.sstate is None else self.sstate.data_ptr()
		a.work = None if self.work is None else self.work.data_ptr()
		a.Yres = None if self.Yres is None else self.Yres.data_ptr()
		a.Yrec = None if Yrec is None else Yrec.data_ptr()
		a.lyap_rec = None if lyap_rec is None else lyap_rec.data_ptr()
		a.lyap_acc = None if self.lyap_acc is None else self.lyap_acc.data_ptr()
		a.proj = None if self.proj is None else self.proj.data_ptr()
		a.status = self.status.data_ptr()
		a.n_accepted, a.n_rejected = self.n_accepted.data_ptr(), self.n_rejected.data_ptr()
		a.queue = None if self.queue is None else self.queue.data_ptr()
		a.stream = self.stream
		return a

	def _check_failures(self):
		failed = (self.status[:self.B] != 0).nonzero().flatten()
		if failed.numel():
			indices = failed.cpu().numpy()
			status = self.status[:self.B][failed].cpu().numpy()
			raise UnsuccessfulIntegration(
				f"{len(indices)} of {self.B} trajectories failed (first: #{indices[0]}, {STATUS_TEXT.get(int(status[0]), status[0])}).",
				indices, status)

	def _launch(self, times, record):
		if self.B is None:
			raise RuntimeError("You must call set_initial_value first.")
		self._upload_parameters()
		times = [float(T) for T in times]
		if any(later < earlier for earlier, later in zip(times, times[1:])):
			raise ValueError("Sample times must not decrease. Cannot integrate backwards in time")
		if self.driver != JB_DRV_IVP_DENSE and times and times[0] < self.time:
			raise ValueError("Target time smaller than current time. Cannot integrate backwards in time")
		if len(times) == 1:
			self._times_host[0] = times[0]
			times_dev = self._times_host.to(self.device, non_blocking=True)
		else:
			times_dev = torch.tensor(times, dtype=torch.float64).to(self.device)
		n_times = len(times)
		Yrec = torch.empty(n_times * self.n * self.ld, dtype=torch.float64, device=self.device) if record else None
		lyap_rec = torch.empty(max(1, n_times * self.n_out) * self.ld, dtype=torch.float64, device=self.device) if self.n_out else None
		args = self._args(times_dev, n_times, Yrec, lyap_rec)
		if self.kernel_events is not None:
			start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			start.record()
			self.module.integrate(args)
			end.record()
			self.kernel_events.append((start, end))
		else:
			self.module.integrate(args)
		self.time = times[-1]
		self._sampled = True
		return Yrec, lyap_rec

	def _current_rows(self, rows=None):
		source = self.Yres if (self.driver == JB_DRV_IVP_DENSE) else self.Y
		return self._unpack(source, self.n, rows)

	def integrate(self, T, out=None, into=None):
		"""advance every trajectory to T and return the state there: (B, n), or (n,) if the initial value was 1-D.
		out: optional CPU tensor (ideally pinned) that receives the result.
		into: optional (B, n) CUDA tensor the kernel's result is written to directly (it becomes the state that is
		returned for a CUDA-resident ensemble) — e.g. this rank's slot of an all-gather buffer (_distributed.py)."""
		T = float(T)
		if self.B is None:
			raise RuntimeError("You must call set_initial_value first.")
		if T == self.time and self.driver != JB_DRV_IVP_DENSE:
			if into is not None:
				into.copy_(self._rows)
				return into
			return self._export(self._rows)       # ODE_wrapper.integrate :142-143
		_, lyap_rec = self._launch([T], record=False)
		self._rows = self._current_rows(into)
		if self.n_out:
			self.last_lyaps = self._unpack(lyap_rec, self.n_out)
		self._check_failures()
		if into is not None and out is None:
			return into                           # the caller's device buffer is the result: no copy to the host
		return self._export(self._rows, out)

	def integrate_many(self, times, record_states=True):
		"""the loop `for T in times: integrate(T)` as ONE launch; returns (len(times), B, n).
		record_states=False keeps only the last state (returned as (1, B, n)) — for Lyapunov runs, whose result is
		the exponents (`last_lyaps`), this saves recording, un-tiling and copying every intermediate state."""
		times = list(times)
		Yrec, lyap_rec = self._launch(times, record=record_states)
		n_times = len(times)
		if record_states:
			out = torch.empty((n_times, self.B, self.n), dtype=torch.float64, device=self.device)
			vec = self.n * self.ld
			for i in range(n_times):
				self.module.unpack(Yrec[i*vec:(i+1)*vec].data_ptr(), out[i].data_ptr(), self.B, self.n, self.ld, self.stream)
		else:
			out = self._current_rows()[None]
		self._rows = out[-1]
		if self.n_out:
			# one vector of n_times·n_out entries per trajectory → (B, n_times, n_out) with a single un-tiling
			flat = self._unpack(lyap_rec, n_times * self.n_out)
			self.last_lyaps = flat.view(self.B, n_times, self.n_out).permute(1, 0, 2)
		self._check_failures()
		if self._single:
			out = out[:, 0]
		if self._out_kind == "cuda":
			return out
		return self._to_host(out)

	# ------------------------------------------------------------------ host-resident ensembles: copies overlapped with kernels

	def _vector_offset(self, b0, rows):
		"""element offset of trajectory b0 (a multiple of 32) in an array of `rows`-component vectors"""
		return b0 * rows           # tiled: (b0/32)·rows·32; row layout: b0·rows

	def integrate_streamed(self, initial_value, T, out, time=0.0, chunks=None):
		"""
		`set_initial_value(initial_value, time)` followed by `integrate(T, out=out)` for an ensemble that lives in
		(pinned) host memory, with the copies hidden behind the kernels: the batch is cut into `chunks` pieces; while
		piece c is integrated, piece c+1 is uploaded and the result of piece c-1 is downloaded on separate streams.
		initial_value, out : (B, n) float64 CPU tensors (pinned for truly asynchronous copies).  Returns `out`.
		chunks : default one piece per 256 MB of state, at most 8 (measured: 2.5 GB of Rössler states gain from 8 pieces,
		16 MB of Lotka–Volterra states lose 27 % against a single piece).
		"""
		if self.post != JB_POST_NONE:
			raise NotImplementedError("integrate_streamed does not run Lyapunov post-steps.")
		host = initial_value
		if not (isinstance(host, torch.Tensor) and not host.is_cuda and host.dtype == torch.float64 and host.dim() == 2 and host.shape[1] == self.n):
			raise ValueError("integrate_streamed needs a (B, n) float64 CPU tensor.")
		T, B, n = float(T), host.shape[0], self.n
		if T < time:
			raise ValueError("Target time smaller than current time. Cannot integrate backwards in time")
		if self.B != B:
			self._alloc(B)
		self._out_kind, self._single = "torch", False
		self.tt.fill_(float(time))
		self.status.zero_()
		self._sampled = False
		if self.sstate is not None:
			self.sstate.zero_()
		self._upload_parameters()
		if chunks is None:
			chunks = max(1, min(8, (8 * B * n) >> 28))
		step = max(JB_TILE, (-(-B // max(1, chunks)) + JB_TILE - 1) // JB_TILE * JB_TILE)
		bounds = [(b0, min(B, b0 + step)) for b0 in range(0, B, step)]
		main = torch.cuda.current_stream(self.device)
		if not hasattr(self, "_copy_streams"):
			self._copy_streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
		stream_in, stream_out = self._copy_streams
		rows = torch.empty((B, n), dtype=torch.float64, device=self.device)
		result = torch.empty((B, n), dtype=torch.float64, device=self.device)
		stream_in.wait_stream(main)
		stream_out.wait_stream(main)
		self._times_host[0] = T
		times_dev = self._times_host.to(self.device, non_blocking=True)
		source = self.Yres if self.driver == JB_DRV_IVP_DENSE else self.Y
		size = ctypes.sizeof(ctypes.c_double)
		for b0, b1 in bounds:
			with torch.cuda.stream(stream_in):
				rows[b0:b1].copy_(host[b0:b1], non_blocking=True)
				uploaded = stream_in.record_event()
			main.wait_event(uploaded)
			count, ld_chunk = b1 - b0, padded(b1 - b0)
			offset = self._vector_offset(b0, n)
			self.module.pack(rows[b0:b1].data_ptr(), self.Y.data_ptr() + size * offset, count, n, ld_chunk, self.stream)
			a = self._args(times_dev, 1)
			a.B = count
			a.t, a.Y = self.tt.data_ptr() + size * b0, self.Y.data_ptr() + size * offset
			a.P = self.P.data_ptr() + size * self._vector_offset(b0, max(1, self.n_params))
			for name, tensor in (("state", self.state), ("work", self.work), ("Yres", self.Yres)):
				if tensor is not None:
					setattr(a, name, tensor.data_ptr() + size * offset)
			if self.sstate is not None:
				a.sstate = self.sstate.data_ptr() + size * b0
			a.status = self.status.data_ptr() + 4 * b0
			a.n_accepted, a.n_rejected = self.n_accepted.data_ptr() + 8 * b0, self.n_rejected.data_ptr() + 8 * b0
			if self.kernel_events is not None:
				start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
				start.record()
				self.module.integrate(a)
				end.record()
				self.kernel_events.append((start, end))
			else:
				self.module.integrate(a)
			self.module.unpack(source.data_ptr() + size * offset, result[b0:b1].data_ptr(), count, n, ld_chunk, self.stream)
			integrated = main.record_event()
			with torch.cuda.stream(stream_out):
				stream_out.wait_event(integrated)
				out[b0:b1].copy_(result[b0:b1], non_blocking=True)
		stream_out.synchronize()
		main.synchronize()
		self.time, self._rows, self._sampled = T, result, True
		self._check_failures()
		return out

	# ------------------------------------------------------------------ counters (the benchmark metric)

	def step_counts(self):
		"""(accepted, rejected) step attempts summed over the ensemble since the last reset."""
		return int(self.n_accepted[:self.B].sum().item()), int(self.n_rejected[:self.B].sum().item())

	def reset_step_counts(self):
		self.n_accepted.zero_()
		self.n_rejected.zero_()
