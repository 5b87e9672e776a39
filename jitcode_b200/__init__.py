"""
jitcode_b200 — a B200-native ensemble ODE engine behind the JiTCODE API
(public names as in the reference's jitcode/__init__.py:1-9 and sympy_symbols.py).
"""
from ._integrator import UnsuccessfulIntegration
from ._jitcode import jitcode, jitcode_lyap, jitcode_restricted_lyap, jitcode_transversal_lyap, test
from ._symbolic import t, y

__all__ = ["UnsuccessfulIntegration", "jitcode", "jitcode_lyap", "jitcode_restricted_lyap", "jitcode_transversal_lyap", "t", "test", "y"]
