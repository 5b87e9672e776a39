"""
TEST INFRASTRUCTURE — CPU oracle, not product code.

NumPy restatement, batched over B trajectories, of the arithmetic that lives in the
reference's third-party integrators (SciPy 1.18.1; SURVEY.md §8a "Algorithm spec"):

  * ScipyRK ....... scipy/integrate/_ivp/rk.py: rk_step :14-71, RungeKutta._step_impl :111-176,
                    RK45 tableau :538-565, DOP853 :568-712, dense outputs :715-764;
                    _ivp/common.py: norm :63-65, select_initial_step :68-134, validate_tol :43-60;
                    _ivp/base.py: OdeSolver.step (status handling).
                    Driven the way integrator_tools.py drives it: t_bound=∞ + dense output
                    (IVP_wrapper :98-105) or moving t_bound (IVP_wrapper_no_interpolation :110-128).
  * HairerDopri ... E. Hairer's DOPRI5 / DOP853 (dopri5.f / dop853.f: HINIT, DOPCOR/DP86CO) as
                    scipy.integrate.ode calls them (scipy/integrate/_ode.py:1112-1236: safety .9,
                    ifactor 10 / 6, dfactor .2 / .3, beta 0 → 0.04 for dopri5 and 0 for dop853,
                    max_step 0 → |T−t|, first_step 0 → HINIT); one fresh start per integrate(T)
                    (integrator_tools.py:134-144).  The compiled `_dop` module has no source here;
                    tests/test_oracle_pin.py pins this restatement against it step for step.

Every trajectory carries its own step size; rejected attempts are repeated per trajectory
(masked lock-step).  `fun(t[B], Y[B,n]) -> dY[B,n]` must be vectorised over the batch.
These classes are the readable specification of the CUDA kernels in
jitcode_b200/csrc/jb_rk.cuh and count attempts exactly as those do.
"""

import numpy as np

EPS = np.finfo(float).eps

# Dormand–Prince 5(4) (published coefficients; same numbers as scipy rk.py:541-565)
DP5_C = np.array([0, 1/5, 3/10, 4/5, 8/9, 1, 1])
DP5_A = np.zeros((7, 7))
DP5_A[1, :1] = [1/5]
DP5_A[2, :2] = [3/40, 9/40]
DP5_A[3, :3] = [44/45, -56/15, 32/9]
DP5_A[4, :4] = [19372/6561, -25360/2187, 64448/6561, -212/729]
DP5_A[5, :5] = [9017/3168, -355/33, 46732/5247, 49/176, -5103/18656]
DP5_B = np.array([35/384, 0, 500/1113, 125/192, -2187/6784, 11/84])
DP5_A[6, :6] = DP5_B
DP5_E = np.array([-71/57600, 0, 71/16695, -71/1920, 17253/339200, -22/525, 1/40])
DP5_P = np.array([
	[1, -8048581381/2820520608, 8663915743/2820520608, -12715105075/11282082432],
	[0, 0, 0, 0],
	[0, 131558114200/32700410799, -68118460800/10900136933, 87487479700/32700410799],
	[0, -1754552775/470086768, 14199869525/1410260304, -10690763975/1880347072],
	[0, 127303824393/49829197408, -318862633887/49829197408, 701980252875/199316789632],
	[0, -282668133/205662961, 2019193451/616988883, -1453857185/822651844],
	[0, 40617522/29380423, -110615467/29380423, 69997945/29380423],
])


def dop853_coefficients():
	"""Hairer's DOP853 coefficients as tabulated by SciPy (third-party data, not reference code)."""
	from scipy.integrate._ivp import dop853_coefficients as co
	return co


def _combine(K, coeffs, upto):
	"""Σ_j coeffs[j]·K[j] for j<upto, skipping exact zeros; K: (stages, B, n)."""
	acc = None
	for j in range(upto):
		if coeffs[j] != 0:
			term = coeffs[j] * K[j]
			acc = term if acc is None else acc + term
	return acc


def rms(x):
	"""scipy _ivp/common.py:63-65, per trajectory (last axis = components)."""
	return np.sqrt(np.sum(x*x, axis=-1) / x.shape[-1])


class ScipyRK:
	"""B independent copies of scipy's RK45 / DOP853 stepper."""

	MAX_FACTOR, MIN_FACTOR, SAFETY = 10.0, 0.2, 0.9

	def __init__(self, fun, t0, Y0, method="RK45", rtol=1e-3, atol=1e-6, max_step=np.inf, first_step=None):
		self.fun = fun
		self.Y = np.array(Y0, dtype=float)
		self.B, self.n = self.Y.shape
		self.t = np.full(self.B, float(t0))
		self.method = method
		self.rtol = max(rtol, 100*EPS)  # validate_tol, common.py:47-51
		self.atol = np.asarray(atol, dtype=float)
		self.max_step = max_step
		if method == "RK45":
			self.A, self.Bc, self.C, self.n_stages, self.err_order = DP5_A, DP5_B, DP5_C, 6, 4
		elif method == "DOP853":
			co = dop853_coefficients()
			self.co = co
			self.n_stages, self.err_order = co.N_STAGES, 7
			self.A, self.Bc, self.C = co.A, co.B, co.C
		else:
			raise ValueError(method)
		self.exponent = -1.0 / (self.err_order + 1)
		self.nfev = np.zeros(self.B, dtype=np.int64)
		self.n_accepted = np.zeros(self.B, dtype=np.int64)
		self.n_rejected = np.zeros(self.B, dtype=np.int64)
		self.failed = np.zeros(self.B, dtype=bool)
		self.F = self._f(self.t, self.Y, slice(None))
		if first_step is None:
			self.h_abs = self._select_initial_step()
		else:
			self.h_abs = np.full(self.B, float(first_step))
		self.K = np.zeros((16 if method == "DOP853" else 7, self.B, self.n))
		self.t_old = np.full(self.B, np.nan)   # nan: no step taken yet
		self.Y_old = np.zeros_like(self.Y)
		self.h_prev = np.zeros(self.B)
		self.t_bound = np.full(self.B, np.inf)

	def _f(self, t, Y, idx):
		self.nfev[idx] += 1
		return self.fun(t, Y)

	def _select_initial_step(self):
		# common.py:108-134 with t_bound=∞ (IVP_wrapper kwargs, integrator_tools.py:59-63), direction=+1
		scale = self.atol + np.abs(self.Y) * self.rtol
		d0 = rms(self.Y / scale)
		d1 = rms(self.F / scale)
		with np.errstate(divide="ignore", invalid="ignore"):
			h0 = np.where((d0 < 1e-5) | (d1 < 1e-5), 1e-6, 0.01 * d0 / d1)
		Y1 = self.Y + h0[:, None] * self.F
		F1 = self._f(self.t + h0, Y1, slice(None))
		d2 = rms((F1 - self.F) / scale) / h0
		with np.errstate(divide="ignore"):
			h1 = np.where((d1 <= 1e-15) & (d2 <= 1e-15), np.maximum(1e-6, h0*1e-3),
					(0.01 / np.maximum(d1, d2)) ** (1.0 / (self.err_order + 1)))
		return np.minimum(np.minimum(100*h0, h1), self.max_step)

	def _error_norm(self, K, h, scale):
		if self.method == "RK45":
			err = _combine(K, DP5_E, 7) * h[:, None]
			return rms(err / scale)
		co = self.co
		err5 = _combine(K, co.E5, 13) / scale   # rk.py:683-691
		err3 = _combine(K, co.E3, 13) / scale
		n5 = np.sum(err5*err5, axis=-1)
		n3 = np.sum(err3*err3, axis=-1)
		denom = n5 + 0.01*n3
		with np.errstate(invalid="ignore", divide="ignore"):
			result = np.abs(h) * n5 / np.sqrt(denom * self.n)
		return np.where((n5 == 0) & (n3 == 0), 0.0, result)

	def step(self, who):
		"""one accepted step (or failure) for the trajectories in index array `who` (rk.py:111-176)."""
		who = np.asarray(who)
		t, Y, F = self.t[who], self.Y[who], self.F[who]
		min_step = 10 * np.abs(np.nextafter(t, np.inf) - t)
		h_abs = np.clip(self.h_abs[who], min_step, self.max_step)
		rejected = np.zeros(len(who), dtype=bool)
		pending = np.arange(len(who))
		S = self.n_stages
		while len(pending):
			p = pending
			too_small = h_abs[p] < min_step[p]
			if too_small.any():
				self.failed[who[p[too_small]]] = True
				p = p[~too_small]
				if not len(p):
					break
			tp, Yp = t[p], Y[p]
			t_new = tp + h_abs[p]
			t_new = np.where(t_new - self.t_bound[who[p]] > 0, self.t_bound[who[p]], t_new)
			h = t_new - tp
			h_abs[p] = np.abs(h)
			K = np.zeros((S+1, len(p), self.n))
			K[0] = F[p]
			for s in range(1, S):
				dy = _combine(K, self.A[s], s) * h[:, None]
				K[s] = self._f(tp + self.C[s]*h, Yp + dy, who[p])
			Y_new = Yp + h[:, None] * _combine(K, self.Bc, S)
			F_new = self._f(tp + h, Y_new, who[p])
			K[S] = F_new
			scale = self.atol + np.maximum(np.abs(Yp), np.abs(Y_new)) * self.rtol
			err = self._error_norm(K, h, scale)
			ok = err < 1
			with np.errstate(divide="ignore"):
				pw = self.SAFETY * err ** self.exponent
			factor_ok = np.where(err == 0, self.MAX_FACTOR, np.minimum(self.MAX_FACTOR, pw))
			factor_ok = np.where(rejected[p], np.minimum(1.0, factor_ok), factor_ok)
			factor_bad = np.maximum(self.MIN_FACTOR, pw)
			h_abs[p] = h_abs[p] * np.where(ok, factor_ok, factor_bad)
			rejected[p[~ok]] = True
			self.n_rejected[who[p[~ok]]] += 1
			acc = p[ok]
			if len(acc):
				w = who[acc]
				self.n_accepted[w] += 1
				self.t_old[w] = tp[ok]
				self.Y_old[w] = Yp[ok]
				self.h_prev[w] = h[ok]
				self.t[w] = t_new[ok]
				self.Y[w] = Y_new[ok]
				self.F[w] = F_new[ok]
				self.K[:S+1, w] = K[:, ok]
			pending = p[~ok]
		self.h_abs[who] = h_abs

	def dense(self, T):
		"""dense output of the last step of every trajectory at time T (rk.py:178-180,693-764)."""
		h = self.h_prev
		x = (T - self.t_old) / h
		if self.method == "RK45":
			result = np.zeros_like(self.Y)
			xp = x.copy()
			for p in range(4):
				Qp = _combine(self.K, DP5_P[:, p], 7)
				result += Qp * xp[:, None]
				xp = xp * x
			return self.Y_old + h[:, None] * result
		co = self.co
		K = self.K
		for s in range(13, 16):
			dy = _combine(K, co.A[s], s) * h[:, None]
			K[s] = self._f(self.t_old + co.C[s]*h, self.Y_old + dy, slice(None))
		Fs = np.empty((co.INTERPOLATOR_POWER, self.B, self.n))
		delta = self.Y - self.Y_old
		Fs[0] = delta
		Fs[1] = h[:, None]*K[0] - delta
		Fs[2] = 2*delta - h[:, None]*(self.F + K[0])
		for r in range(co.D.shape[0]):
			Fs[3+r] = h[:, None] * _combine(K, co.D[r], 16)
		result = np.zeros_like(self.Y)
		for i, f in enumerate(reversed(Fs)):
			result += f
			result *= (x if i % 2 == 0 else 1 - x)[:, None]
		return result + self.Y_old

	# -- the two ways the reference drives the stepper

	def integrate_dense(self, T):
		"""IVP_wrapper.integrate (integrator_tools.py:98-105)."""
		while True:
			who = np.flatnonzero(((self.t < T) | np.isnan(self.t_old)) & ~self.failed)
			if not len(who):
				break
			self.step(who)
		return self.dense(T)

	def integrate_land(self, T):
		"""IVP_wrapper_no_interpolation.integrate (integrator_tools.py:110-128)."""
		self.t_bound[:] = T
		while True:
			who = np.flatnonzero((self.t < T) & ~self.failed)
			if not len(who):
				break
			self.step(who)
		return self.Y.copy()


class HairerDopri:
	"""B independent runs of Hairer's DOPRI5 / DOP853 core as configured by scipy.integrate.ode."""

	def __init__(self, fun, method="dopri5", rtol=1e-6, atol=1e-12, nsteps=10**6, max_step=0.0,
			first_step=0.0, safety=0.9, ifactor=None, dfactor=None, beta=0.0, verbatim_dop853_reject=True):
		self.fun, self.method = fun, method
		self.rtol, self.atol = rtol, np.asarray(atol, dtype=float)
		self.nmax = nsteps if nsteps > 0 else 100000
		self.max_step, self.first_step, self.safety = max_step, first_step, safety
		if method == "dopri5":
			self.ifactor = 10.0 if ifactor is None else ifactor
			self.dfactor = 0.2 if dfactor is None else dfactor
			self.beta = 0.04 if beta == 0 else max(beta, 0.0)
			self.expo1 = 0.2 - self.beta*0.75
			self.iord = 5
		elif method == "dop853":
			self.ifactor = 6.0 if ifactor is None else ifactor
			self.dfactor = 0.3 if dfactor is None else dfactor
			self.beta = max(beta, 0.0)
			self.expo1 = 1.0/8.0 - self.beta*0.2
			self.iord = 8
			self.co = dop853_coefficients()
		else:
			raise ValueError(method)
		# scipy 1.18.1's _dop.dopri853 shrinks by exactly dfactor on rejection
		self.verbatim_dop853_reject = verbatim_dop853_reject
		self.nfev = None

	def _f(self, t, Y, idx):
		self.nfev[idx] += 1
		return self.fun(t, Y)

	def _hinit(self, t, Y, F0, hmax, idx):
		sk = self.atol + self.rtol*np.abs(Y)
		dnf = np.sum((F0/sk)**2, axis=-1)
		dny = np.sum((Y/sk)**2, axis=-1)
		with np.errstate(divide="ignore", invalid="ignore"):
			h = np.where((dnf <= 1e-10) | (dny <= 1e-10), 1e-6, np.sqrt(dny/dnf)*0.01)
		h = np.minimum(h, hmax)
		F1 = self._f(t + h, Y + h[:, None]*F0, idx)
		der2 = np.sqrt(np.sum(((F1-F0)/sk)**2, axis=-1)) / h
		der12 = np.maximum(np.abs(der2), np.sqrt(dnf))
		with np.errstate(divide="ignore"):
			h1 = np.where(der12 <= 1e-15, np.maximum(1e-6, np.abs(h)*1e-3), (0.01/der12)**(1.0/self.iord))
		return np.minimum(np.minimum(100*np.abs(h), h1), hmax)

	def integrate(self, t0, Y0, T):
		"""fresh start at (t0, Y0), advance every trajectory to exactly T.
		returns dict(y, failed(status codes), n_accepted, n_rejected, nfev)."""
		Y = np.array(Y0, dtype=float)
		B, n = Y.shape
		t = np.full(B, float(t0)) if np.ndim(t0) == 0 else np.array(t0, dtype=float)
		self.nfev = np.zeros(B, dtype=np.int64)
		status = np.zeros(B, dtype=np.int64)
		n_acc = np.zeros(B, dtype=np.int64)
		n_rej = np.zeros(B, dtype=np.int64)
		nstep = np.zeros(B, dtype=np.int64)
		everyone = np.arange(B)
		hmax = np.full(B, self.max_step) if self.max_step > 0 else np.abs(T - t)
		K1 = self._f(t, Y, everyone)
		if self.first_step > 0:
			h = np.full(B, float(self.first_step))
		else:
			h = self._hinit(t, Y, K1, hmax, everyone)
		facold = np.full(B, 1e-4)
		reject = np.zeros(B, dtype=bool)
		# stiffness detection (dopri5.f / dop853.f "STIFFNESS DETECTION": every NSTIFF = 1000 accepted steps, or while
		# a suspicion is pending): h·|f(y_new) − f(last stage)| / |y_new − last stage input| estimates h·λ
		iasti = np.zeros(B, dtype=np.int64)
		nonsti = np.zeros(B, dtype=np.int64)
		hlamb = np.zeros(B)
		active = t < T
		facc1, facc2 = 1.0/self.dfactor, 1.0/self.ifactor
		dp5 = self.method == "dopri5"
		S = 7 if dp5 else 13
		while active.any():
			p = np.flatnonzero(active)
			over = nstep[p] > self.nmax
			status[p[over]] = 2
			small = (0.1*np.abs(h[p]) <= np.abs(t[p])*2.3e-16) & ~over
			status[p[small]] = 1
			active[p[over | small]] = False
			p = p[~(over | small)]
			if not len(p):
				break
			tp, Yp, hp = t[p], Y[p], h[p].copy()
			last = (tp + 1.01*hp - T) > 0
			hp = np.where(last, T - tp, hp)
			nstep[p] += 1
			K = np.zeros((S, len(p), n))
			K[0] = K1[p]
			if dp5:
				for s in range(1, 6):
					K[s] = self._f(tp + DP5_C[s]*hp, Yp + hp[:, None]*_combine(K, DP5_A[s], s), p)
				Y_new = Yp + hp[:, None]*_combine(K, DP5_B, 6)
				K[6] = self._f(tp + hp, Y_new, p)
				sk = self.atol + self.rtol*np.maximum(np.abs(Yp), np.abs(Y_new))
				errv = hp[:, None] * _combine(K, DP5_E, 7)
				err = np.sqrt(np.sum((errv/sk)**2, axis=-1) / n)
			else:
				co = self.co
				for s in range(1, 12):
					K[s] = self._f(tp + co.C[s]*hp, Yp + hp[:, None]*_combine(K, co.A[s], s), p)
				Y_new = Yp + hp[:, None]*_combine(K, co.B, 12)
				sk = self.atol + self.rtol*np.maximum(np.abs(Yp), np.abs(Y_new))
				e5 = _combine(K, co.E5, 12) / sk
				e3 = _combine(K, co.E3, 12) / sk
				s5 = np.sum(e5*e5, axis=-1)
				s3 = np.sum(e3*e3, axis=-1)
				deno = s5 + 0.01*s3
				deno = np.where(deno <= 0, 1.0, deno)
				err = np.abs(hp) * s5 * np.sqrt(1.0/(n*deno))
			with np.errstate(divide="ignore"):
				fac11 = err ** self.expo1
				fac = fac11 / facold[p]**self.beta
			fac = np.maximum(facc2, np.minimum(facc1, fac/self.safety))
			hnew = hp / fac
			ok = err <= 1
			# accepted
			a = p[ok]
			if len(a):
				facold[a] = np.maximum(err[ok], 1e-4)
				n_acc[a] += 1
				if not dp5:
					K[12][ok] = self._f(tp[ok] + hp[ok], Y_new[ok], a)
				check = (n_acc[a] % 1000 == 0) | (iasti[a] > 0)
				if check.any():
					c = np.flatnonzero(ok)[check]          # positions in p
					w = a[check]
					row = S - 2                            # the last proper stage: its input is y + h Σ a(row, j) K_j
					A_row = DP5_A[row] if dp5 else self.co.A[row]
					stage_input = Yp[c] + hp[c, None]*_combine(K[:, c], A_row, row)
					stnum = np.sum((K[S-1][c] - K[row][c])**2, axis=-1)
					stden = np.sum((Y_new[c] - stage_input)**2, axis=-1)
					with np.errstate(divide="ignore", invalid="ignore"):
						hlamb[w] = np.where(stden > 0, np.abs(hp[c])*np.sqrt(stnum/stden), hlamb[w])
					stiff = hlamb[w] > (3.25 if dp5 else 6.1)
					nonsti[w] = np.where(stiff, 0, nonsti[w] + 1)
					iasti[w] = np.where(stiff, iasti[w] + 1, np.where(nonsti[w] == 6, 0, iasti[w]))
				# IDID = −4 (scipy.integrate.ode: "problem is probably stiff (interrupted)"): the solver returns before it
				# commits the step that raised the 15th suspicion — t and y stay where they were
				gave_up = iasti[a] == 15
				status[a[gave_up]] = 3
				active[a[gave_up]] = False
				keep = ~gave_up
				a, ok_keep = a[keep], np.flatnonzero(ok)[keep]
				K1[a] = K[S-1][ok_keep]
				Y[a] = Y_new[ok_keep]
				t[a] = np.where(last[ok_keep], T, tp[ok_keep] + hp[ok_keep])
				hn = np.minimum(np.abs(hnew[ok_keep]), hmax[a])
				hn = np.where(reject[a], np.minimum(hn, np.abs(hp[ok_keep])), hn)
				reject[a] = False
				h[a] = hn
				active[a[last[ok_keep]]] = False
				if getattr(self, "trace", None) is not None:
					self.trace.append((t[a].copy(), hp[ok_keep].copy(), err[ok_keep].copy()))
			r = p[~ok]
			if len(r):
				if dp5 or not self.verbatim_dop853_reject:
					h[r] = hp[~ok] / np.minimum(facc1, fac11[~ok]/self.safety)
				else:
					h[r] = hp[~ok] * self.dfactor
				reject[r] = True
				n_rej[r] += (n_acc[r] >= 1)
				if getattr(self, "trace_rejected", None) is not None:
					self.trace_rejected.append((tp[~ok].copy(), hp[~ok].copy(), err[~ok].copy()))
		return {"y": Y, "t": t, "status": status, "n_accepted": n_acc, "n_rejected": n_rej,
				"n_attempts": nstep, "nfev": self.nfev.copy()}
