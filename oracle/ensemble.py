"""
TEST INFRASTRUCTURE — CPU oracle, not product code.

The reference's hot path for an ensemble: compiled C right-hand side handed to SciPy one
trajectory at a time (SURVEY.md §3 stacks B/C), parallelised over trajectories with a
process pool — the CPU baseline of BASELINE.md §3 and the row-by-row parity oracle.

`run_ensemble` returns the sampled states, the number of RHS evaluations per trajectory and
the wall time; a *trajectory-step* is nfev/6 for the 5th-order pair, nfev/12 for the
8th-order one.
"""

import multiprocessing as mp
import os
import time

import numpy as np

from . import wrappers

RHS_PER_ATTEMPT = {"dopri5": 6, "RK45": 6, "dop853": 12, "DOP853": 12, "RK23": 3}


class _CountingRhs:
	"""counts calls of the compiled f for the `ode` backends, which do not expose nfev."""
	__slots__ = ("f", "calls")

	def __init__(self, f):
		self.f, self.calls = f, 0

	def __call__(self, t, y):
		self.calls += 1
		return self.f(t, y)


def _lyap_step(state, n_basic, n_lyap, dt):
	"""jitcode_lyap.norms/integrate (_jitcode.py:743-775) with orthonormalise_qr restated as
	NumPy QR of the column matrix: norms = |diag R|."""
	vectors = state[n_basic:].reshape(n_lyap, n_basic)
	Q, R = np.linalg.qr(vectors.T)
	norms = np.abs(np.diag(R))
	state = state.copy()
	state[n_basic:] = Q.T.flatten()
	return state, np.log(norms) / dt


def _run_chunk(task):
	(handle, method, interpolate, params, Y0, pars, t0, times, lyap) = task
	B, n = Y0.shape
	out = np.empty((len(times), B, n))
	nfev = np.zeros(B, dtype=np.int64)
	failed = np.zeros(B, dtype=bool)
	lyaps = np.empty((len(times), B, lyap[1])) if lyap else None
	for b in range(B):
		if pars.shape[1]:
			handle.initialise(*pars[b])
		counter = _CountingRhs(handle.f)
		backend = wrappers.make_backend(method, counter, interpolate=interpolate, **params)
		backend.set_initial_value(Y0[b], t0)
		try:
			for i, T in enumerate(times):
				t_before = backend.t
				state = backend.integrate(T)
				if lyap:
					state, lyaps[i, b] = _lyap_step(np.asarray(state), lyap[0], lyap[1], T - t_before)
					backend.set_initial_value(state, T)
				out[i, b] = state
		except wrappers.UnsuccessfulIntegration:
			failed[b] = True
			out[i:, b] = np.nan
		nfev[b] = counter.calls
	return out, nfev, failed, lyaps


def _warm_up(handle):
	handle.f      # imports the compiled right-hand side in the worker
	return os.getpid()


def available_cores():
	return len(os.sched_getaffinity(0))


def run_ensemble(handle, method, Y0, times, pars=None, t0=0.0, interpolate=True,
		processes=None, lyap=None, **params):
	"""
	handle      : c_rhs.RhsHandle
	method      : "dopri5" | "dop853" | "RK45" | "DOP853"
	Y0          : (B, n) initial states; pars: (B, p) control parameters or None
	times       : increasing sample times (one backend.integrate call each, as in the examples)
	lyap        : None or (n_basic, n_lyap): re-orthonormalise after every call like jitcode_lyap
	returns dict(y=(len(times),B,n), nfev=(B,), failed=(B,), seconds=wall, cores=P, attempts=total)
	"""
	Y0 = np.ascontiguousarray(Y0, dtype=float)
	B = Y0.shape[0]
	pars = np.empty((B, 0)) if pars is None else np.ascontiguousarray(pars, dtype=float).reshape(B, -1)
	times = [float(T) for T in times]
	P = processes or available_cores()
	P = max(1, min(P, B))
	bounds = np.linspace(0, B, 4*P + 1).astype(int) if P > 1 else np.array([0, B])
	tasks = [
		(handle, method, interpolate, params, Y0[a:b], pars[a:b], t0, times, lyap)
		for a, b in zip(bounds[:-1], bounds[1:]) if b > a
	]
	if P == 1:
		start = time.perf_counter()
		results = [_run_chunk(task) for task in tasks]
		seconds = time.perf_counter() - start
	else:
		# forkserver: the workers descend from a clean helper process, not from this one — the caller may hold a
		# CUDA context and torch threads (GPU tests, bench.py), and forking that aborts in the children now and then
		with mp.get_context("forkserver").Pool(P) as pool:
			pool.map(_warm_up, [handle] * P, chunksize=1)     # start-up of the workers is not integration time
			start = time.perf_counter()
			results = pool.map(_run_chunk, tasks, chunksize=1)
			seconds = time.perf_counter() - start
	nfev = np.concatenate([r[1] for r in results])
	return {
		"y": np.concatenate([r[0] for r in results], axis=1),
		"nfev": nfev,
		"failed": np.concatenate([r[2] for r in results]),
		"lyaps": np.concatenate([r[3] for r in results], axis=1) if lyap else None,
		"seconds": seconds,
		"cores": P,
		"attempts": float(nfev.sum()) / RHS_PER_ATTEMPT[method],
	}
