"""
Test infrastructure: the vectorised right-hand side that `_vectorise.warp_statements` prints for the GPU, compiled
for the HOST with g++ and run lane after lane, so that the printed code itself (component order, tables, windows,
masks) is checked on a machine without a GPU.  Team reductions are emulated in rounds: every round runs all
lanes, collects their contributions to each reduction and hands the totals to the next round.
"""

import ctypes
import subprocess

import numpy as np

from jitcode_b200 import _vectorise

SOURCE = r"""
#include <math.h>
#include <vector>
static std::vector<double> totals, partial;
static int call_index;
namespace jb {{
static inline double flip_sign(double v, unsigned word) {{ return (word & 0x80000000u) ? -v : v; }}
// a shuffle between lanes that run one after the other: the value the source lane offered in the previous round
static std::vector<std::vector<double>> offered, offered_before;
static int shuffle_index, current_lane_index;
static inline double shfl(double v, unsigned source)
{{
	const int c = shuffle_index++;
	if (c >= (int) offered.size()) {{ offered.resize(c + 1, std::vector<double>(256, NAN)); offered_before.resize(c + 1, std::vector<double>(256, NAN)); }}
	offered[c][current_lane_index] = v;
	return offered_before[c][source];
}}
template <int W> struct Team {{
	static double sum(double v, double *)
	{{
		const int c = call_index++;
		if (c >= (int) partial.size()) {{ partial.resize(c + 1, 0.0); totals.resize(c + 1, 0.0); }}
		partial[c] += v;
		return totals[c];
	}}
	template <int K> static double butterfly(double v, double *scratch) {{ return K == 0 ? sum(v, scratch) : v; }}
}};
}}
struct Vec {{
	double *p;
	double operator()(int i) const {{ return p[i]; }}
	double at(unsigned bytes) const {{ return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(p) + bytes); }}
	void set(int i, double x) const {{ p[i] = x; }}
}};
struct Par {{ double v[{n_par}]; }};
typedef {tu_type} TU;
alignas(8) static const double jb_tf[] = {{ {tf} }};
alignas(8) static const TU jb_tu[] = {{ {tu} }};
static void rhs_lane(const double t, const Vec &y, Vec &dy, const Par &par, const int lane, double *jb_scratch)
{{
	const double *const JB_TF = jb_tf;
	const TU *const JB_TU = jb_tu;
{body}
	(void) t; (void) par; (void) JB_TF; (void) JB_TU; (void) jb_scratch;
}}
extern "C" void evaluate(double t, double *row, double *out, const double *pars, int lanes, int rounds)
{{
	Par par;
	for (int k = 0; k < {n_par}; ++k) par.v[k] = pars[k];
	Vec y{{row}}, dy{{out}};
	for (int round = 0; round < rounds; ++round) {{
		partial.assign(partial.size(), 0.0);
		for (int lane = 0; lane < lanes; ++lane) {{
			call_index = 0;
			jb::shuffle_index = 0;
			jb::current_lane_index = lane;
			rhs_lane(t, y, dy, par, lane, nullptr);
		}}
		totals = partial;
		jb::offered_before = jb::offered;
	}}
}}
"""


def evaluate_printed(lines, tables, component_of, n_par, t, state, pars, lanes, rounds, workdir):
	"""f(t, state) computed by the printed code; `rounds` = 1 + depth of the chain of reductions."""
	n = len(component_of)
	wide = bool(tables.u) and max(max(tables.u), n) >= 65536
	source = SOURCE.format(
		n_par=max(1, n_par), tu_type="unsigned int" if wide else "unsigned short",
		tf=", ".join(repr(float(v)) for v in tables.f64) or "0.0", tu=", ".join(str(int(v)) for v in tables.u) or "0",
		body="\n".join("\t" + line for line in lines),
	)
	import hashlib
	tag = hashlib.sha1(source.encode()).hexdigest()[:12]      # dlopen caches by path: one file per source
	path = workdir / f"lanes_{tag}.cpp"
	path.write_text(source)
	library = workdir / f"lanes_{tag}.so"
	subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", str(library), str(path)], check=True, capture_output=True)
	lib = ctypes.CDLL(str(library))
	row = np.zeros(_vectorise.row_length(n))
	row[:n] = np.asarray(state, dtype=float)[component_of]
	out = np.full_like(row, np.nan)
	parameters = np.asarray(list(pars) + [0.0], dtype=float)
	as_pointer = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
	if any("jb::shfl" in line for line in lines):
		rounds += 1          # (what a lane fetches from another lane is there one round later)
	lib.evaluate(ctypes.c_double(t), as_pointer(row), as_pointer(out), as_pointer(parameters), ctypes.c_int(lanes), ctypes.c_int(rounds))
	result = np.empty(n)
	result[component_of] = out[:n]
	return result
