"""
TEST INFRASTRUCTURE — CPU oracle, not product code.

Builds the right-hand side of a system inside the REFERENCE'S OWN extension template,
`/root/reference/jitcode/jitced_template.c`, rendered with Jinja2 exactly as
`jitcxde._render_template` is called at _jitcode.py:397-406 and compiled with gcc from
where it lies.  Nothing of the reference is copied into the repository's history: the recipe's
outputs — the shared object (portable flags, `-march=x86-64-v3`) and the rendered translation
unit it was compiled from — live only in `oracle/_ref/` (git-ignored, travels to the GPU box like
our own built files).  The rendered unit is kept so that the machine that TIMES the reference path
can recompile it with the reference's own flags (`-march=native -mtune=native`, the family of
jitcxde_common's DEFAULT_COMPILE_ARGS): `native_handle`.

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
import shutil
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
			# the rendered unit, for native_handle() on the machine that runs the benchmark
			kept = REF_DIR / "rendered" / module
			kept.mkdir(parents=True, exist_ok=True)
			for name in (f"{module}.c", "f.c", "f_definitions.c", "general_helpers.c", "general_helpers_definitions.c"):
				shutil.copyfile(tmp / name, kept / name)
		if label:
			index = REF_DIR / "index.json"
			known = json.loads(index.read_text()) if index.exists() else {}
			known[label] = module
			index.write_text(json.dumps(known, indent=1, sort_keys=True))
	return c_rhs.RhsHandle(target, module, n, len(par_names))


def native_handle(handle):
	"""the same reference container recompiled ON THIS HOST with the reference's flags (-march=native -mtune=native),
	from the rendered unit build_ref_rhs kept; returns (handle, how).  Falls back to the prebuilt portable file when
	gcc or the rendered unit is missing."""
	module = handle.module_name
	rendered = REF_DIR / "rendered" / module
	if shutil.which("gcc") is None or not (rendered / f"{module}.c").exists():
		return handle, "prebuilt with -march=x86-64-v3 (no gcc or no rendered unit on this host)"
	target_dir = REF_DIR / f"native_{c_rhs._host_tag()}"
	target = target_dir / f"{module}.so"
	if not target.exists():
		target_dir.mkdir(parents=True, exist_ok=True)
		tmp = target_dir / f"{module}.{os.getpid()}.tmp.so"
		try:
			c_rhs.compile_extension(rendered / f"{module}.c", tmp, extra_includes=[rendered], portable=False)
			os.replace(tmp, target)
		except Exception as error:      # a host without Python headers: keep the portable build
			return handle, f"prebuilt with -march=x86-64-v3 (native rebuild failed: {type(error).__name__})"
	return c_rhs.RhsHandle(target, module, handle.n, handle.n_par), "compiled on this host with -Ofast -march=native -mtune=native"


def build_benchmark_systems():
	import workloads as W
	built = {}
	fhn_lyap, fhn_helpers = c_rhs.lyapunov_system(W.fhn()["f_sym"], 4, W.fhn()["helpers"], 4)
	for label, system in {
		"lotka_volterra": W.lotka_volterra(),
		"lotka_volterra_gamma": W.lotka_volterra(gamma_as_parameter=True),
		"fhn": W.fhn(),
		"fhn_lyap4": {"f_sym": fhn_lyap, "helpers": fhn_helpers, "control_pars": [], "n": len(fhn_lyap)},
		"roessler_100": W.roessler_network(100),
		"roessler_100_random": W.roessler_network(100, topology="random"),
		"roessler_100_er": W.roessler_network(100, topology="erdos_renyi"),
	}.items():
		handle = build_ref_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"], label=label)
		built[label] = handle.path if handle else None
	return built


if __name__ == "__main__":
	for label, path in build_benchmark_systems().items():
		print(label, "->", path)
