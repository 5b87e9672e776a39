"""
CUDA code generator: the B200 counterpart of generate_f_C / generate_helpers_C
(_jitcode.py:186-249, 332-360) and of the container jitced_template.c.

The reference prints one `set_dy(i, <expr>)` statement per component into f.c and wraps it in a
CPython extension; here the same statements (`dy.set(i, <expr>);`, reading `y(i)`, `t`,
`par.v[k]` and helper temporaries) become the body of `jb_generated::Sys::rhs`, a device
function template that the hand-written Runge–Kutta kernels of csrc/jb_rk.cuh inline into their
stage loop (register storage) or call out of line (global storage).  The vector types behind
`y` and `dy` decide where the state lives; the printed text is the same for both — the role the
accessor macros play in jitced_template.c:38-55.
"""

import os
import struct

import sympy
from sympy.core.function import AppliedUndef
from sympy.printing.c import C99CodePrinter


JB_DP5, JB_DOP853, JB_RK23 = 0, 1, 2
JB_DRV_ODE, JB_DRV_IVP_DENSE, JB_DRV_IVP_LAND = 0, 1, 2
JB_STORE_REGISTERS, JB_STORE_GLOBAL, JB_STORE_WARP, JB_STORE_SHARED, JB_STORE_QUAD = 0, 1, 2, 3, 5

METHOD_NAMES = {JB_DP5: "JB_DP5", JB_DOP853: "JB_DOP853", JB_RK23: "JB_RK23"}
DRIVER_NAMES = {JB_DRV_ODE: "JB_DRV_ODE", JB_DRV_IVP_DENSE: "JB_DRV_IVP_DENSE", JB_DRV_IVP_LAND: "JB_DRV_IVP_LAND"}
STORAGE_NAMES = {JB_STORE_REGISTERS: "JB_STORE_REGISTERS", JB_STORE_GLOBAL: "JB_STORE_GLOBAL", JB_STORE_WARP: "JB_STORE_WARP", JB_STORE_SHARED: "JB_STORE_SHARED", JB_STORE_QUAD: "JB_STORE_QUAD"}


class CudaPrinter(C99CodePrinter):
	"""FP64 CUDA C++: small integer powers become products (no pow() call on the FP64 pipe),
	everything else is the C99 text the reference's generated code would contain."""

	MAX_PRODUCT_POWER = 8

	def __init__(self, literals=None):
		super().__init__({"allow_unknown_functions": True})
		# numbers that do not fit an instruction's 32-bit immediate (the upper half of a double) are read from the constant
		# bank as JB_LIT[k] if the caller keeps a table: as a literal each costs two UMOV per use — 16 % of the instructions
		# of the FitzHugh–Nagumo Lyapunov kernel were UMOV (profiles/r2_fhn_quad_summary.txt)
		self.literals = literals

	def _constant(self, value, text):
		if self.literals is None or os.environ.get("JITCODE_B200_INLINE_LITERALS"):
			return text
		if struct.unpack("<Q", struct.pack("<d", value))[0] & 0xffffffff == 0:
			return text
		if value not in self.literals:
			self.literals.append(value)
		return f"JB_LIT[{self.literals.index(value)}]"

	def _print_Float(self, expr):
		return self._constant(float(expr), super()._print_Float(expr))

	def _print_Rational(self, expr):
		return self._constant(float(expr), super()._print_Rational(expr))

	def _print_Pow(self, expr):
		base, exponent = expr.base, expr.exp
		if exponent.is_Integer and 2 <= abs(int(exponent)) <= self.MAX_PRODUCT_POWER:
			text = self._print(base)
			if not (base.is_Symbol or isinstance(base, AppliedUndef) or (base.is_Number and base >= 0)):
				text = f"({text})"
			product = "*".join([text] * abs(int(exponent)))
			return product if exponent > 0 else f"1.0/({product})"
		return super()._print_Pow(expr)

	def _print_Integer(self, expr):
		# keep every literal a double: no integer arithmetic, no integer division
		return f"{int(expr)}.0"

	def _print_AppliedUndef(self, expr):
		# y(3): the index stays an integer literal
		return f"{expr.func.__name__}({', '.join(str(int(arg)) if arg.is_Integer else self._print(arg) for arg in expr.args)})"

	def _print_Function(self, expr):
		if isinstance(expr, AppliedUndef):
			return self._print_AppliedUndef(expr)
		return super()._print_Function(expr)


def _repeated_function_calls(expressions):
	"""transcendental calls (sin(t), exp(y(0)) …) that occur in more than one place, innermost first."""
	counts = {}
	for expr in expressions:
		for atom in expr.atoms(sympy.Function):
			if not isinstance(atom, AppliedUndef):
				counts[atom] = counts.get(atom, 0) + expr.count(atom)
	repeated = [atom for atom, count in counts.items() if count > 1]
	repeated.sort(key=lambda atom: (sympy.count_ops(atom), sympy.default_sort_key(atom)))
	return repeated


class Statements(list):
	"""lines of code, and the table of numbers they refer to as JB_LIT[k]"""
	def __init__(self, *args):
		super().__init__(*args)
		self.literals = []


def rhs_statements(f, helpers, control_pars, do_cse=False):
	"""
	f            : list of expressions in y(i), t, helper symbols, control parameters
	helpers      : sorted list of (symbol, definition)
	control_pars : list of symbols
	returns the lines of the body of Sys::rhs.
	"""
	subs = {par: sympy.Symbol(f"par.v[{k}]") for k, par in enumerate(control_pars)}
	for k, (symbol, _) in enumerate(helpers):
		subs[symbol] = sympy.Symbol(f"jb_helper_{k}")
	helper_defs = [definition.xreplace(subs) for _, definition in helpers]
	f = [entry.xreplace(subs) for entry in f]

	lines = Statements()
	printer = CudaPrinter(lines.literals)

	# shared temporaries: repeated function calls always; full common-subexpression elimination on request
	everything = helper_defs + f
	temporaries = []
	if do_cse:
		replacements, everything = sympy.cse(everything, symbols=(sympy.Symbol(f"jb_cse_{k}") for k in range(10**9)), order="none")
		temporaries = list(replacements)
	else:
		for k, call in enumerate(_repeated_function_calls(everything)):
			symbol = sympy.Symbol(f"jb_call_{k}")
			everything = [expr.xreplace({call: symbol}) for expr in everything]
			temporaries = [(s, d.xreplace({call: symbol})) for s, d in temporaries]
			temporaries.append((symbol, call))
	helper_defs, f = everything[:len(helpers)], everything[len(helpers):]

	# temporaries may use helpers and vice versa: emit each as soon as its inputs exist
	helper_symbols = {sympy.Symbol(f"jb_helper_{k}"): k for k in range(len(helpers))}
	emitted = set()
	pending_tmp = list(temporaries)
	def flush_temporaries():
		progress = True
		while progress:
			progress = False
			for item in list(pending_tmp):
				needs = {s for s in item[1].free_symbols if s in helper_symbols or s.name.startswith(("jb_cse_", "jb_call_"))}
				if needs <= emitted:
					lines.append(f"const double {item[0].name} = {printer.doprint(item[1])};")
					emitted.add(item[0])
					pending_tmp.remove(item)
					progress = True
	flush_temporaries()
	for k, definition in enumerate(helper_defs):
		lines.append(f"const double jb_helper_{k} = {printer.doprint(definition)};")
		emitted.add(sympy.Symbol(f"jb_helper_{k}"))
		flush_temporaries()
	assert not pending_tmp, "unresolvable temporaries"
	for i, entry in enumerate(f):
		lines.append(f"dy.set({i}, {printer.doprint(entry)});")
	return lines


def _vectorise_row_length(n):
	from ._vectorise import row_length
	return row_length(n)


def _c_array(values, per_line=16):
	values = list(values) or ["0"]
	return ",\n\t".join(", ".join(values[k:k+per_line]) for k in range(0, len(values), per_line))


def module_source(statements, n, n_params, n_basic, n_lyap, tangent_indices, method, driver, storage, threads, tables=None, component_of=None, kregs=0, min_blocks=1, team_warps=1, group=1):
	"""the complete translation unit of one module (what _render_template produces from
	jitced_template.c in the reference, _jitcode.py:397-406).

	statements : body of the right-hand side — of Sys::rhs (thread storage: `y(i)` / `dy.set(i, …)` on whatever
		vector type the stepper uses) or of Sys::rhs_warp (warp storage: the vectorised form of _vectorise.py,
		with `tables` = its _vectorise.Tables; quad storage: the system with ONE tangent vector, 2·n_basic statements).
	group : quad storage: lanes per trajectory (power of two ≥ n_lyap)."""
	tangent_indices = list(tangent_indices or [])
	if tangent_indices:
		table = ", ".join(str(int(k)) for k in tangent_indices)
		tangent = (
			f"\tstatic __device__ __forceinline__ int tangent_index(int k)\n"
			f"\t{{\n\t\tconstexpr int table[{len(tangent_indices)}] = {{ {table} }};\n\t\treturn table[k];\n\t}}\n"
		)
	else:
		tangent = "\tstatic __device__ __forceinline__ int tangent_index(int) { return 0; }\n"

	body = "\n".join("\t\t" + line for line in statements)
	warp = storage == JB_STORE_WARP
	tf = [repr(float(v)) for v in tables.f64] if warp else []
	tu = [str(int(v)) for v in tables.u] if warp else []
	literals = [repr(float(v)) for v in getattr(statements, "literals", [])]
	tu_type = "unsigned short" if (not tu or max(max(int(v) for v in tu), n) < 65536) else "unsigned int"
	components = [str(int(c)) for c in (component_of if warp else range(n))]
	n_row = _vectorise_row_length(n) if warp else n
	if warp:
		function = f"""	template <class VI, class VO>
	static __device__ __forceinline__ void rhs_warp(const double t, const VI &y, VO &dy, const Par &par, const int lane, const void *jb_tables, double *jb_scratch)
	{{
		const double *const JB_TF = static_cast<const double *>(jb_tables);
		const TU *const JB_TU = reinterpret_cast<const TU *>(JB_TF + TF_COUNT);
{body}
		(void) t; (void) par; (void) JB_TF; (void) JB_TU; (void) jb_scratch;
	}}"""
	else:
		function = f"""	template <class VI, class VO>
	static __device__ __forceinline__ void rhs(const double t, const VI &y, VO &dy, const Par &par)
	{{
{body}
		(void) t; (void) par;
	}}"""
	return f"""// generated by jitcode_b200._codegen — one ODE system fused into the ensemble Runge–Kutta kernels
#include "jb_rk.cuh"

namespace jb_generated {{

// tables of the vectorised right-hand side (warp storage); staged in shared memory by the kernels
__device__ const double jb_tf[{max(1, len(tf))}] = {{
	{_c_array(tf, 8)}
}};
__device__ const {tu_type} jb_tu[{max(1, len(tu))}] = {{
	{_c_array(tu, 24)}
}};
// numbers of the right-hand side that do not fit an instruction's immediate: operands from the constant bank
__constant__ double jb_lit[{max(1, len(literals))}] = {{
	{_c_array(literals, 4)}
}};
#define JB_LIT jb_generated::jb_lit
// component stored at each position of the kernel's vectors (identity unless the vectoriser regrouped them)
__device__ const {tu_type} jb_component[{n}] = {{
	{_c_array(components, 24)}
}};

struct Sys {{
	static constexpr int N = {n};      // dimension of the integrated system
	static constexpr int NP = {n_params};     // control parameters per trajectory
	static constexpr int NB = {n_basic};     // dimension of the underlying system
	static constexpr int NL = {n_lyap};     // tangent vectors
	static constexpr int NT = {len(tangent_indices)};     // transversal tangent coordinates
	static constexpr int TF_COUNT = {len(tf)}, TU_COUNT = {len(tu)};
	static constexpr int NROW = {n_row};   // shared-memory row length (warp storage): positions N … NROW-1 hold 0
	static constexpr int TEAM_WARPS = {int(team_warps)};   // warps that share one trajectory (warp storage)
	static constexpr int GROUP = {int(group)};   // lanes that share one trajectory (quad storage)
	static constexpr int THREADS = {int(threads)};   // threads per block (= row stride of shared storage)
	typedef {tu_type} TU;
	static __device__ __forceinline__ const TU *components(const void *tables)
	{{
		return reinterpret_cast<const TU *>(static_cast<const double *>(tables) + TF_COUNT) + TU_COUNT;
	}}
	static __device__ __forceinline__ const TU *positions(const void *tables)      // component → position
	{{
		return components(tables) + N;
	}}
	struct Par {{ double v[NP > 0 ? NP : 1]; }};
{tangent}
{function}
}};

}}  // namespace jb_generated

#define JB_MODULE_METHOD {METHOD_NAMES[method]}
#define JB_MODULE_DRIVER {DRIVER_NAMES[driver]}
#define JB_MODULE_STORAGE {STORAGE_NAMES[storage]}
#define JB_MODULE_THREADS {int(threads)}
#define JB_MODULE_KREGS {int(kregs)}
#define JB_MODULE_MIN_BLOCKS {int(min_blocks)}
#include "jb_module.cuh"
"""
