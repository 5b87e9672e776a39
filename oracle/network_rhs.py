"""
TEST INFRASTRUCTURE — CPU oracle, not product code.

Right-hand side of the reference's Kuramoto example (examples/kuramoto_network.py:19-26) for networks too
large to be written out term by term:

    dy_i/dt = ω_i + c/(n−1) · Σ_{j : A[j,i]} sin(y_j − y_i)

The reference generates one sine term of straight-line C per edge (`generate_f_C`, _jitcode.py:186-249); for
n = 1000 at 20 % density those are 2·10⁵ terms / 8.4 MB of C, which gcc needs > 10 minutes for at -O1 and does
not finish at the reference's -Ofast.  This module evaluates THE SAME SUM, edge by edge in the same order
(j ascending), as a loop over the adjacency lists — compiled C with the reference's flag family, called from
SciPy's integrators through the same wrappers (oracle/wrappers.py) as every other oracle run.  It does NOT use
the sin/cos factorisation of the CUDA engine (csrc/jb_dense.cu), so it checks that factorisation.
Pinned against the written-out form (oracle/c_rhs.py, itself pinned to the reference's container) at n = 40 in
tests/test_oracle_pin.py.

The handle offers what oracle/ensemble.py needs: `.f(t, y)`, `.initialise(*pars)`, picklable.
"""

import ctypes
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np

from .c_rhs import BUILD_DIR, COMPILE_ARGS, _host_tag

SOURCE = r"""
#include <math.h>
/* dy_i = omega_i + scale * sum_{k in [rowptr_i, rowptr_{i+1})} sin(y[col_k] - y_i)  —  kuramoto_network.py:19-26 */
void network_f(int n, const int *rowptr, const int *col, const double *omega, double scale, const double *y, double *dy)
{
	for (int i = 0; i < n; ++i) {
		const double yi = y[i];
		double sum = 0.0;
		for (int k = rowptr[i]; k < rowptr[i+1]; ++k)
			sum += sin(y[col[k]] - yi);
		dy[i] = omega[i] + scale * sum;
	}
}
"""


def _library():
	key = hashlib.sha1((SOURCE + " ".join(COMPILE_ARGS) + _host_tag()).encode()).hexdigest()[:16]
	BUILD_DIR.mkdir(parents=True, exist_ok=True)
	target = BUILD_DIR / f"jbo_network_{key}.so"
	if not target.exists():
		source = BUILD_DIR / f"jbo_network_{key}.c"
		source.write_text(SOURCE)
		tmp = BUILD_DIR / f"jbo_network_{key}.{os.getpid()}.tmp.so"
		subprocess.run(["gcc", "-shared", "-fPIC", *COMPILE_ARGS, str(source), "-o", str(tmp), "-lm"], check=True, capture_output=True)
		os.replace(tmp, target)
	return target


class NetworkRhsHandle:
	"""A: (n, n) adjacency with A[j, i] ≠ 0 meaning "j acts on i" (as in the example); omega: (n,); c: coupling."""

	def __init__(self, A, omega, c):
		A = np.asarray(A)
		self.n = n = len(omega)
		self.n_par = 0
		incoming = [np.flatnonzero(A[:, i]) for i in range(n)]
		self.rowptr = np.concatenate([[0], np.cumsum([len(js) for js in incoming])]).astype(np.int32)
		self.col = np.concatenate(incoming).astype(np.int32) if n else np.zeros(0, np.int32)
		self.omega = np.ascontiguousarray(omega, dtype=float)
		self.scale = float(c) / (n - 1)
		self.path = str(_library())
		self._fn = None

	def __getstate__(self):
		state = dict(self.__dict__)
		state["_fn"] = None
		return state

	def _bind(self):
		lib = ctypes.CDLL(self.path)
		lib.network_f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
		lib.network_f.restype = None
		fn, n, scale = lib.network_f, self.n, self.scale
		rowptr, col, omega = self.rowptr.ctypes.data, self.col.ctypes.data, self.omega.ctypes.data

		def f(t, y):
			y = np.ascontiguousarray(y, dtype=float)
			dy = np.empty(n)
			fn(n, rowptr, col, omega, scale, y.ctypes.data, dy.ctypes.data)
			return dy
		return f

	@property
	def f(self):
		if self._fn is None:
			self._fn = self._bind()
		return self._fn

	def f_many(self, t, Y):
		return np.stack([self.f(0.0, row) for row in np.atleast_2d(Y)])

	def initialise(self, *pars):
		pass
