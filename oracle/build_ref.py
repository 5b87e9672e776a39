"""
TEST INFRASTRUCTURE — CPU oracle, not product code.

Builds the right-hand side of a system inside the REFERENCE'S OWN extension template,
`/root/reference/jitcode/jitced_template.c`, rendered with Jinja2 exactly as
`jitcxde._render_template` is called at _jitcode.py:397-406 and compiled with gcc from
where it lies.  Nothing of the reference is copied into the repository: the rendered
translation unit lives in a temporary directory and only the shared object is kept, in
`oracle/_ref/` (git-ignored, travels to the GPU box like our own built files).

The generated include files have the names the template's `#include`s fix
(jitced_template.c:58,62,65,105,110,113): `f.c`, `f_definitions.c`, `general_helpers.c`,
`general_helpers_definitions.c`; their bodies are the very statements the port
(rhs_module.c.in) compiles, so port and reference container can be compared bit for bit.

Usable only where /root/reference exists (this container).  `python -m oracle.build_ref`
builds the containers for the benchmark systems of workloads.py.
"""

import hashlib
import json
import os
import tempfile
from pathlib import Path

from . import c_rhs

REFERENCE_TEMPLATE = Path("/root/reference/jitcode/jitced_template.c")
REF_DIR = c_rhs.ORACLE_DIR / "_ref"


def reference_available():
	return REFERENCE_TEMPLATE.exists()


def ref_module_name(f_lines, helper_lines, par_names):
	key = hashlib.sha1("\n".join(f_lines + helper_lines + par_names).encode()).hexdigest()[:16]
	return f"jbref_{key}"


def build_ref_rhs(f_sym, helpers=(), control_pars=(), n=None, label=None):
	"""returns an RhsHandle on the reference container (building it if /root/reference is there),
	or None when neither the built file nor the reference exists."""
	f_lines, helper_lines, n, par_names = c_rhs.print_statements(f_sym, helpers, control_pars, n)
	module = ref_module_name(f_lines, helper_lines, par_names)
	target = REF_DIR / f"{module}.so"
	if not target.exists():
		if not reference_available():
			return None
		import jinja2
		REF_DIR.mkdir(parents=True, exist_ok=True)
		with tempfile.TemporaryDirectory() as tmp:
			tmp = Path(tmp)
			rendered = jinja2.Template(REFERENCE_TEMPLATE.read_text()).render(
				module_name=module,
				n=n,
				has_Jacobian=False,
				number_of_f_helpers=0,
				number_of_jac_helpers=0,
				number_of_general_helpers=len(helper_lines),
				sparse_jac=None,
				control_pars=par_names,
				callbacks=[],
			)
			(tmp / f"{module}.c").write_text(rendered)
			(tmp / "f.c").write_text("\n".join(f_lines) + "\n")
			(tmp / "f_definitions.c").write_text("")
			(tmp / "general_helpers.c").write_text("\n".join(helper_lines) + "\n")
			(tmp / "general_helpers_definitions.c").write_text("")
			out = tmp / "out.so"
			c_rhs.compile_extension(tmp / f"{module}.c", out, extra_includes=[tmp], portable=True)
			os.replace(out, target)
		if label:
			index = REF_DIR / "index.json"
			known = json.loads(index.read_text()) if index.exists() else {}
			known[label] = module
			index.write_text(json.dumps(known, indent=1, sort_keys=True))
	return c_rhs.RhsHandle(target, module, n, len(par_names))


def build_benchmark_systems():
	import workloads as W
	built = {}
	for label, system in {
		"lotka_volterra": W.lotka_volterra(),
		"lotka_volterra_gamma": W.lotka_volterra(gamma_as_parameter=True),
		"fhn": W.fhn(),
		"roessler_100": W.roessler_network(100),
	}.items():
		handle = build_ref_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"], label=label)
		built[label] = handle.path if handle else None
	return built


if __name__ == "__main__":
	for label, path in build_benchmark_systems().items():
		print(label, "->", path)
