"""
TEST INFRASTRUCTURE — CPU oracle, not product code.

The three sampling semantics with which the reference drives SciPy, restated (the
reference file jitcode/integrator_tools.py cannot travel to the GPU box and must not be
copied; where /root/reference exists, tests/test_oracle_pin.py checks these restatements
against the real file, call by call):

  * OdeBackend ........ integrator_tools.py:130-151 (ODE_wrapper): scipy.integrate.ode,
                        i.e. Hairer's dopri5 / dop853; every integrate(T) is a fresh
                        solver start from (t, y) that lands exactly on T; failure →
                        UnsuccessfulIntegration; T < t → ValueError; T == t → current state.
  * IvpDense .......... :39-108 (IVP_wrapper): one persistent solve_ivp stepper with
                        t_bound = ∞; integrate(T) steps until the solver time is ≥ T and
                        returns the dense output of the last step at T.
  * IvpLand ........... :110-128 (IVP_wrapper_no_interpolation): moves t_bound to T,
                        un-finishes the stepper and steps until it lands on T.
"""

import numpy as np
from scipy.integrate import ode
from scipy.integrate._ivp.ivp import METHODS as IVP_METHODS


class UnsuccessfulIntegration(Exception):
	pass

ODE_NAMES = ("dopri5", "dop853")
IVP_NAMES = ("RK45", "DOP853", "RK23")


class OdeBackend:
	def __init__(self, name, f, nsteps=10**6, **params):
		assert name in ODE_NAMES
		self.solver = ode(f)
		# jitcode.set_integrator passes nsteps=10**6 by default (_jitcode.py:560,602-606)
		self.solver.set_integrator(name, nsteps=nsteps, **params)

	def set_initial_value(self, y0, t0=0.0):
		self.solver.set_initial_value(np.array(y0, dtype=float), t0)

	@property
	def t(self):
		return self.solver.t

	@property
	def _y(self):
		return self.solver._y

	def integrate(self, T):
		if T > self.solver.t:
			result = self.solver.integrate(T)
			if not self.solver.successful():
				raise UnsuccessfulIntegration
			return result
		if T == self.solver.t:
			return self.solver._y
		raise ValueError("Target time smaller than current time. Cannot integrate backwards in time")


class IvpDense:
	def __init__(self, name, f, **params):
		self.cls = IVP_METHODS[name]
		self.kwargs = {"fun": f, "t_bound": np.inf, "vectorized": False, **params}
		self.stepper = None
		self.t = None
		self._y = None

	def set_initial_value(self, y0, t0=0.0):
		self.t, self._y = t0, np.array(y0, dtype=float)
		self.stepper = self.cls(t0=t0, y0=self._y, **self.kwargs)

	def _advance(self):
		self.stepper.step()
		if self.stepper.status == "failed":
			raise UnsuccessfulIntegration

	def integrate(self, T):
		while self.stepper.t < T or self.stepper.t_old is None:
			self._advance()
		self._y = self.stepper.dense_output()(T)
		self.t = T
		return self._y

	@property
	def nfev(self):
		return self.stepper.nfev


class IvpLand(IvpDense):
	def integrate(self, T):
		if self.stepper.t > T:
			raise ValueError("Target time smaller than current time. Cannot integrate backwards in time")
		if self.stepper.t < T:
			self.stepper.t_bound = T
			self.stepper.status = "running"
			while self.stepper.status == "running":
				self.stepper.step()
			self._y, self.t = self.stepper.y, T
			if self.stepper.status == "failed":
				raise UnsuccessfulIntegration
		return self._y


def make_backend(name, f, interpolate=True, **params):
	"""the choice jitcode.set_integrator makes (_jitcode.py:599-617)."""
	if name in ODE_NAMES:
		return OdeBackend(name, f, **params)
	if name in IVP_NAMES:
		params.pop("nsteps", None)
		return (IvpDense if interpolate else IvpLand)(name, f, **params)
	raise ValueError(f"The oracle covers explicit Runge–Kutta methods only, not {name!r}.")
