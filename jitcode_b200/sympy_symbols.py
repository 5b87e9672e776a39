"""
`y` and `t` under the name the reference offers them for SymPy users (jitcode/sympy_symbols.py:1-8).
Here SymPy is the only flavour, so these are the package's own symbols.
"""
from ._symbolic import t, y

__all__ = ["t", "y"]
