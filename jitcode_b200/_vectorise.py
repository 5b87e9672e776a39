"""
SIMT-vectorising code generator for large systems ("warp" storage: one trajectory per warp, lanes
across components).

The reference prints one C statement per component (generate_f_C, _jitcode.py:186-249) and, for
large systems, cuts them into chunks (chunk_size, :198-200) that a CPU executes one after the
other.  A warp cannot run 32 different statements at once — but the equations of a network of
identical units differ only in their numbers (ω_i, degree) and in which components they read.
This module finds that structure:

  1. every statement is brought into a canonical tree in which numeric constants and the indices
     of dynamical variables are *slots*, the children of commutative operations are ordered by
     shape, and — in sums — runs of equally shaped terms (neighbour sums Σ_j y(3j), Σ_j sin(y_j−y_i))
     become a *loop* of variable length;
  2. statements of equal shape form a group; per group the slot values go into tables (slots whose
     value is the same for the whole group are printed as literals);
  3. per group ONE code block is printed; lane ℓ evaluates statements ℓ, ℓ+32, … of the group,
     fetching its numbers and indices from the tables (staged in shared memory by the kernel).

Helpers (e.g. a mean field Σ_j y(3j+2)) are evaluated cooperatively: their loops are spread over
the lanes and reduced with shuffles; the scalar rest is computed redundantly by every lane.

What the printer does beyond that, all of it measured on the B200 (profiles/):

  * components are permuted so that the outputs of a group are contiguous (`component_positions`);
  * bare neighbour sums Σ g(y[idx]) use 16-bit pre-scaled byte offsets, two per word, padded per pass
    to compile-time trip counts, and dealt to their slots by shared-memory bank (`_spread_over_banks`);
  * lattice-like sums need no table for the neighbours most statements share ("diagonals");
  * rings whose size splits into lanes × passes are printed in BLOCKED order (`ring_blockings`): a lane runs
    consecutive ring elements and reads their overlapping neighbourhoods once; nearly complete rings are
    summed without masks and the absent neighbours subtracted through the (signed) pair table;
  * groups of equal size share one loop, so that operands common to their equations are loaded once.

`tests/host_lanes.py` compiles the printed code for the host and runs it lane after lane: the printer is
tested on a machine without a GPU.

`evaluate_plan` is a NumPy interpreter of the same plan, used by the CPU tests to check the
vectoriser against direct evaluation of the symbolic input.
"""

import os
from collections import OrderedDict

import numpy as np
import sympy
from sympy.core.function import AppliedUndef

from ._codegen import CudaPrinter, _repeated_function_calls

LOOP_MIN = 3        # a run of at least this many equally shaped terms in a sum becomes a loop
FUSE_MAX = 4        # groups of equal size printed into one loop
SPREAD_OVER_BANKS = True
SUBTRACT_MISSING = True
SIGNED_PAIRS = True


# ------------------------------------------------------------------ canonical trees

class Node:
	__slots__ = ("kind", "value", "children", "key", "has_loop")

	def __init__(self, kind, value=None, children=()):
		self.kind, self.value, self.children = kind, value, list(children)
		self.has_loop = kind == "loop" or any(child.has_loop for child in self.children)
		if kind == "const":
			self.key = "C"
		elif kind == "y":
			self.key = "Y"
		elif kind == "sym":
			self.key = "S:" + value.name
		elif kind == "loop":
			self.key = "Loop[" + self.children[0].key + "]"
		elif kind == "pow":
			self.key = f"Pow({self.children[0].key};{sympy.srepr(value)})"
		else:   # op
			self.key = f"{value.__name__}(" + ",".join(child.key for child in self.children) + ")"


def _is_dynvar(expr):
	return isinstance(expr, AppliedUndef) and expr.func.__name__ == "y"


def _first(node, kind):
	if node.kind == kind:
		return node.value
	for child in node.children:
		found = _first(child, kind)
		if found is not None:
			return found
	return None


def _role(node, own):
	"""tie-break between equally shaped children of a commutative operation, chosen so that the same ROLE lands in
	the same slot in every statement of a group: the distance of the first dynamical variable to the statement's
	own component (a·y(i+1) + ω·y(i) vs ω'·y(j) + a·y(j+1)), then the first constant."""
	index, constant = _first(node, "y"), _first(node, "const")
	return (0 if index is None else 1, 0 if index is None else (index - own if own is not None else index), 0.0 if constant is None else constant)


def analyse(expr, allow_loops=True, own=None, force=()):
	"""canonical tree of a SymPy expression (`own`: index of the component whose equation this is; `force`: keys of
	terms that form a loop however few of them there are — see make_plan)."""
	if expr.is_Number:
		return Node("const", float(expr))
	if _is_dynvar(expr):
		return Node("y", int(expr.args[0]))
	if expr.is_Symbol:
		return Node("sym", expr)
	if expr.is_Pow and expr.exp.is_Number:
		return Node("pow", expr.exp, [analyse(expr.base, allow_loops, own, force)])
	if expr.is_Add or expr.is_Mul:
		children = sorted((analyse(arg, allow_loops, own, force) for arg in expr.args), key=lambda node: (node.key, _role(node, own)))
		if expr.is_Add and allow_loops:
			runs = OrderedDict()
			for child in children:
				runs.setdefault(child.key, []).append(child)
			children = []
			for key, members in runs.items():
				varying = "C" in key or "Y" in key
				if (len(members) >= LOOP_MIN or key in force) and varying and not any(member.has_loop for member in members):
					children.append(Node("loop", None, members))
				else:
					children.extend(members)
			children.sort(key=lambda node: (node.key, _role(node, own) if node.kind != "loop" else ()))
		return Node("op", expr.func, children)
	if isinstance(expr, sympy.Function) or isinstance(expr, AppliedUndef):
		return Node("op", expr.func, [analyse(arg, allow_loops, own, force) for arg in expr.args])
	raise NotImplementedError(f"cannot vectorise {expr!r}")


def slots(node, consts, indices, loops):
	"""collects slot values in canonical order; loops get their own term-local slots."""
	if node.kind == "const":
		consts.append(node.value)
	elif node.kind == "y":
		indices.append(node.value)
	elif node.kind == "loop":
		terms = []
		for member in node.children:
			tc, ti = [], []
			slots(member, tc, ti, None)
			terms.append((tc, ti))
		loops.append(terms)
	else:
		for child in node.children:
			slots(child, consts, indices, loops)


class SlotNames:
	"""hands out placeholder symbols while a template expression is rebuilt from a canonical tree."""
	def __init__(self, prefix):
		self.prefix = prefix
		self.n_const = self.n_index = self.n_loop = 0

	def const(self):
		self.n_const += 1
		return sympy.Symbol(f"{self.prefix}c{self.n_const-1}")

	def index(self):
		self.n_index += 1
		return sympy.Symbol(f"{self.prefix}i{self.n_index-1}", integer=True)

	def loop(self):
		self.n_loop += 1
		return sympy.Symbol(f"{self.prefix}l{self.n_loop-1}")


Y = sympy.Function("y")

def template(node, names, loop_templates):
	"""SymPy expression with placeholders for the slots; loop bodies are appended to loop_templates."""
	if node.kind == "const":
		return names.const()
	if node.kind == "y":
		return Y(names.index())
	if node.kind == "sym":
		return node.value
	if node.kind == "loop":
		symbol = names.loop()
		term_names = SlotNames(f"{symbol.name}_t")
		loop_templates.append((symbol, template(node.children[0], term_names, None), term_names))
		return symbol
	if node.kind == "pow":
		return sympy.Pow(template(node.children[0], names, loop_templates), node.value)
	return node.value(*[template(child, names, loop_templates) for child in node.children])


# ------------------------------------------------------------------ the plan

class Loop:
	def __init__(self, symbol, body, names):
		self.symbol, self.body = symbol, body
		self.n_const, self.n_index = names.n_const, names.n_index
		self.prefix = names.prefix
		self.start = [0]                               # prefix sums over the statements of the group
		self.consts = [[] for _ in range(self.n_const)]    # per slot: one value per term
		self.indices = [[] for _ in range(self.n_index)]


class Group:
	"""statements of one shape."""
	def __init__(self, key, node, prefix):
		self.key = key
		names = SlotNames(prefix)
		loop_templates = []
		self.body = template(node, names, loop_templates)
		self.prefix = prefix
		self.n_const, self.n_index = names.n_const, names.n_index
		self.loops = [Loop(*item) for item in loop_templates]
		self.outs = []
		self.consts = [[] for _ in range(self.n_const)]
		self.indices = [[] for _ in range(self.n_index)]

	def add(self, out, node):
		consts, indices, loops = [], [], []
		slots(node, consts, indices, loops)
		assert len(consts) == self.n_const and len(indices) == self.n_index and len(loops) == len(self.loops)
		self.outs.append(out)
		for slot, value in zip(self.consts, consts):
			slot.append(value)
		for slot, value in zip(self.indices, indices):
			slot.append(value)
		for loop, terms in zip(self.loops, loops):
			for tc, ti in terms:
				assert len(tc) == loop.n_const and len(ti) == loop.n_index
				for slot, value in zip(loop.consts, tc):
					slot.append(value)
				for slot, value in zip(loop.indices, ti):
					slot.append(value)
			loop.start.append(loop.start[-1] + len(terms))

	@property
	def size(self):
		return len(self.outs)

	def reordered(self, order):
		"""the same group with its statements in another order (place m holds what was statement order[m])."""
		other = object.__new__(Group)
		other.key, other.body, other.prefix = self.key, self.body, self.prefix
		other.n_const, other.n_index = self.n_const, self.n_index
		other.outs = [self.outs[m] for m in order]
		other.consts = [[slot[m] for m in order] for slot in self.consts]
		other.indices = [[slot[m] for m in order] for slot in self.indices]
		other.loops = []
		for loop in self.loops:
			moved = object.__new__(Loop)
			moved.symbol, moved.body, moved.prefix = loop.symbol, loop.body, loop.prefix
			moved.n_const, moved.n_index = loop.n_const, loop.n_index
			terms = [k for m in order for k in range(loop.start[m], loop.start[m+1])]
			moved.start = [0]
			for m in order:
				moved.start.append(moved.start[-1] + loop.start[m+1] - loop.start[m])
			moved.consts = [[slot[k] for k in terms] for slot in loop.consts]
			moved.indices = [[slot[k] for k in terms] for slot in loop.indices]
			other.loops.append(moved)
		return other


class Plan:
	"""everything the code printer and the NumPy interpreter need."""
	def __init__(self, n, n_params):
		self.n, self.n_params = n, n_params
		self.temporaries = []       # (symbol, expression): repeated function calls, uniform over the warp
		self.helpers = []           # (symbol, Group of size 1)
		self.groups = []

	@property
	def lane_passes(self):
		return sum((group.size + 31) // 32 for group in self.groups)

	@property
	def efficiency(self):
		"""fraction of lane slots doing useful work."""
		return self.n / (32.0 * max(1, self.lane_passes))


def make_plan(f, helpers, control_pars):
	"""f: list of expressions; helpers: sorted (symbol, definition) pairs; control_pars: symbols."""
	subs = {par: sympy.Symbol(f"par.v[{k}]") for k, par in enumerate(control_pars)}
	for k, (symbol, _) in enumerate(helpers):
		subs[symbol] = sympy.Symbol(f"jb_helper_{k}")
	helper_defs = [definition.xreplace(subs) for _, definition in helpers]
	f = [entry.xreplace(subs) for entry in f]
	everything = helper_defs + f
	plan = Plan(len(f), len(control_pars))
	# shared temporaries are evaluated by EVERY lane, which pays off for sin(t) but not for sin(y(3) − y(24)):
	# calls that depend on the state stay inside their statements
	uniform_calls = [call for call in _repeated_function_calls(everything) if not any(_is_dynvar(atom) for atom in call.atoms(AppliedUndef))]
	for k, call in enumerate(uniform_calls):
		symbol = sympy.Symbol(f"jb_call_{k}")
		everything = [expr.xreplace({call: symbol}) for expr in everything]
		plan.temporaries = [(s, d.xreplace({call: symbol})) for s, d in plan.temporaries]
		plan.temporaries.append((symbol, call))
	helper_defs, f = everything[:len(helpers)], everything[len(helpers):]
	for k, definition in enumerate(helper_defs):
		node = analyse(sympy.sympify(definition))
		group = Group(node.key, node, f"jb_h{k}_")
		group.add(-1, node)
		plan.helpers.append((sympy.Symbol(f"jb_helper_{k}"), group))
	nodes = [analyse(sympy.sympify(entry), own=i) for i, entry in enumerate(f)]
	# Stragglers: a node of a network with only one or two neighbours has no "run of equally shaped terms", so its equation
	# would be a group of its own (and keep its siblings' groups from being fused).  Where reading its few terms as a
	# (short) loop gives it the shape of a large group, it joins that group.
	census = {}
	for node in nodes:
		census[node.key] = census.get(node.key, 0) + 1
	large = {key for key, count in census.items() if count >= 8}
	def loop_terms(node, found):
		if node.kind == "loop":
			found.add(node.children[0].key)
		for child in node.children:
			loop_terms(child, found)
		return found
	force = set()
	for node in nodes:
		if node.key in large:
			loop_terms(node, force)
	if force:
		for i, node in enumerate(nodes):
			if node.key not in large:
				again = analyse(sympy.sympify(f[i]), own=i, force=force)
				if again.key in large:
					nodes[i] = again
	groups = OrderedDict()
	for i, node in enumerate(nodes):
		if node.key not in groups:
			groups[node.key] = Group(node.key, node, f"jb_g{len(groups)}_")
		groups[node.key].add(i, node)
	plan.groups = list(groups.values())
	return plan


# ------------------------------------------------------------------ NumPy interpreter (test infrastructure for the plan)

def evaluate_plan(plan, t, state, pars=()):
	"""dY = f(t, state) computed by walking the plan the way the generated kernel does."""
	state = np.asarray(state, dtype=float)
	scalars = {sympy.Symbol("t", real=True): float(t)}
	for k, value in enumerate(pars):
		scalars[sympy.Symbol(f"par.v[{k}]")] = float(value)

	def run(body, names, consts, indices, loops):
		values = dict(scalars)
		for c, value in enumerate(consts):
			values[sympy.Symbol(f"{names}c{c}")] = value
		expr = body
		for i, index in enumerate(indices):
			expr = expr.xreplace({Y(sympy.Symbol(f"{names}i{i}", integer=True)): sympy.Float(state[index])})
		for symbol, total in loops:
			values[symbol] = total
		return float(expr.xreplace({k: sympy.Float(v) for k, v in values.items()}))

	def run_group(group, m):
		loops = []
		for loop in group.loops:
			total = 0.0
			for k in range(loop.start[m], loop.start[m+1]):
				total += run(loop.body, loop.prefix, [slot[k] for slot in loop.consts], [slot[k] for slot in loop.indices], [])
			loops.append((loop.symbol, total))
		return run(group.body, group.prefix, [slot[m] for slot in group.consts], [slot[m] for slot in group.indices], loops)

	pending = list(plan.temporaries)
	def flush():
		progress = True
		while progress:
			progress = False
			for symbol, definition in list(pending):
				expr = definition
				for atom in definition.atoms(AppliedUndef):
					if _is_dynvar(atom):
						expr = expr.xreplace({atom: sympy.Float(state[int(atom.args[0])])})
				expr = expr.xreplace({k: sympy.Float(v) for k, v in scalars.items()})
				if not expr.free_symbols:
					scalars[symbol] = float(expr)
					pending.remove((symbol, definition))
					progress = True
	flush()
	for symbol, group in plan.helpers:
		scalars[symbol] = run_group(group, 0)
		flush()
	out = np.empty(plan.n)
	for group in plan.groups:
		for m, i in enumerate(group.outs):
			out[i] = run_group(group, m)
	return out


# ------------------------------------------------------------------ printing

class Tables:
	"""two flat tables (doubles, unsigned integers) that the kernel stages in shared memory."""
	def __init__(self):
		self.f64, self.u = [], []
		self.helper_lanes = None      # set while a batch is printed whose idle lanes help with the gathers (_print_loop)
		self.single_warp = True       # one warp per trajectory (the helpers' shuffles are warp-wide)

	def add_f64(self, values):
		offset = len(self.f64)
		self.f64.extend(float(v) for v in values)
		return offset

	def shared_f64(self, values):
		"""(sign, offset) of a table that holds these numbers or their negatives (ω_i of one equation, −ω_i of another):
		equal addresses let the compiler load them once."""
		values = [float(v) for v in values]
		for sign in ("", "-"):
			wanted = values if not sign else [-v for v in values]
			for offset in range(0, len(self.f64) - len(values) + 1):
				if self.f64[offset] == wanted[0] and self.f64[offset:offset + len(values)] == wanted:
					return sign, offset
		return "", self.add_f64(values)

	def add_u(self, values, align=1):
		while len(self.u) % align:
			self.u.append(0)
		offset = len(self.u)
		self.u.extend(int(v) for v in values)
		return offset

	@property
	def wide(self):
		return bool(self.u) and max(self.u) >= 65536

	@property
	def nbytes(self):
		return 8 * len(self.f64) + ((4 if self.wide else 2) * len(self.u) + 7) // 8 * 8 + 16


def _uniform(values):
	return len(values) > 0 and all(v == values[0] for v in values)


def component_positions(plan, groups=None):
	"""Where component i lives inside the kernel's vectors: the outputs of every group are made contiguous
	(group after group, statement after statement), so that a group's lanes write — and, for neighbour
	couplings, mostly read — consecutive shared-memory words.  The permutation costs nothing: every index
	the right-hand side uses comes from the tables printed here; only loading and storing the caller's
	state goes through it (Sys::component_of)."""
	position = {}
	for group in (plan.groups if groups is None else groups):
		for out in group.outs:
			position[out] = len(position)
	assert sorted(position) == list(range(plan.n)), "every component needs exactly one equation"
	return position


# Blocked rings.  With statements dealt to the lanes round-robin (lane ℓ runs statements ℓ, ℓ+32, …), a lane that
# sums its ring neighbours m±1 … m±D reads 2D words per statement and never reads a word twice.  If instead every
# lane runs a BLOCK of consecutive ring elements (lane ℓ: elements ℓ·Q … ℓ·Q+Q−1, one per pass), the neighbourhoods
# of its Q statements overlap: it reads the 2D+Q words of the block's surroundings once, keeps them in registers and
# serves all its passes from there — the shared-memory traffic of the gather falls by 2D·Q/(2D+Q).  The words are
# stored pass-major (place q·L + ℓ for element ℓ·Q + q, L lanes in use), so that at every step of the window the
# lanes of a warp still read consecutive words.

def _ring_split(size, team):
	"""(lanes, passes) with lanes·passes = size, lanes ≤ team and at most a quarter more passes than the round-robin
	order needs; None if the ring cannot be cut that way."""
	least = -(-size // team)
	for passes in range(max(2, least), least + max(1, least // 4) + 1):
		if size % passes == 0 and 16 <= size // passes <= team:      # at least half a warp at work
			return size // passes, passes
	return None


def _blocked_order(size, lanes, passes):
	"""place q·lanes + ℓ holds ring element ℓ·passes + q."""
	return [(place % lanes) * passes + place // lanes for place in range(size)]


def _bare_gather(loop):
	return loop.n_index == 1 and all(_uniform(values) for values in loop.consts) and _paddable(loop)


def ring_blockings(plan, team):
	"""{group size: (lanes, passes)} for the sizes of groups that sum over ring neighbours of their own kind and gain
	from the blocked order; every group of such a size is printed in that order."""
	owner = {out: (g, m) for g, group in enumerate(plan.groups) for m, out in enumerate(group.outs)}
	found = {}
	for group in plan.groups:
		split = _ring_split(group.size, team)
		if split is None or group.size in found or min(group.outs) < 0:
			continue
		lanes, passes = split
		for loop in group.loops:
			if not _bare_gather(loop):
				continue
			targets = {owner[c][0] for c in loop.indices[0]}
			if len(targets) != 1 or plan.groups[next(iter(targets))].size != group.size:
				continue
			size, popularity = group.size, {}
			for m in range(size):
				for k in range(loop.start[m], loop.start[m+1]):
					d = (owner[loop.indices[0][k]][1] - m) % size
					popularity[d] = popularity.get(d, 0) + 1
			common = sorted(d for d, count in popularity.items() if count >= 0.6 * size)[:32]
			if len(common) < 4:
				continue
			signed = [d if d <= size // 2 else d - size for d in common]
			window = {divmod(q + d, passes) for q in range(passes) for d in signed}
			carries = [carry for carry, _ in window]
			if max(carries) - min(carries) < lanes and 4 * len(window) <= 3 * len(common) * (-(-size // team)):
				found[size] = split
	return found


def _paddable(loop):
	"""can a loop be padded to uniform length with terms that contribute exactly zero (varying constants 0,
	indices pointing to the zero slot)?"""
	body = loop.body
	for c, values in enumerate(loop.consts):
		symbol = sympy.Symbol(f"{loop.prefix}c{c}")
		body = body.xreplace({symbol: sympy.Float(values[0]) if _uniform(values) else sympy.Integer(0)})
	body = body.replace(lambda e: _is_dynvar(e), lambda e: sympy.Integer(0))
	try:
		return body == 0 or sympy.simplify(body) == 0
	except Exception:
		return False


HELPER_LANES = not os.environ.get("JITCODE_B200_NO_HELPER_LANES")
SPLIT_REDUCTIONS = not os.environ.get("JITCODE_B200_NO_SPLIT_REDUCTIONS")


def team_is_warp(team):
	"""a single warp whose idle lanes can shadow working lanes of the second half-warp"""
	return 24 <= team < 32


def _helper_plan(lists, team, n_pass, size):
	"""per pass (cap, owners): every statement keeps `cap` (even) gather slots, the statements in `owners` hand the rest to
	one idle lane each; None if no pass gets narrower that way.  The smallest cap is taken for which the idle lanes
	(but one, whose empty row serves everybody else) suffice and the rest of a statement fits a helper's row."""
	if size != team * n_pass:
		return None
	helpers = min(32 - team, 8) - 1
	if helpers < 1:
		return None
	plan, gain = [], 0
	for p in range(n_pass):
		counts = {m: len(lists[m]) for m in range(team * p, team * p + team)}
		widest = max(counts.values())
		widest += widest % 2
		cap = widest
		for candidate in range(2, widest, 2):
			owners = [m for m, count in counts.items() if count > candidate]
			if len(owners) <= helpers and all(counts[m] - candidate <= candidate for m in owners):
				cap = candidate
				break
		plan.append((max(cap, 2), [m for m, count in counts.items() if count > cap]))
		gain += widest - cap
	return plan if gain > 0 else None


def _ell(values, start, size, width, pad):
	"""CSR → ELL: entry (k, m) at k*size + m, padded with `pad`."""
	out = [pad] * (width * size)
	for m in range(size):
		for k, value in enumerate(values[start[m]:start[m+1]]):
			out[k*size + m] = value
	return out


# Positions N … N + ZERO_WORDS − 1 of every shared-memory row hold 0.0: where a padded gather table has no entry, the
# lane reads one of them.  One such word per bank, because the padding reads take part in the bank arithmetic like any
# other: with a single zero word every real gather that fell into ITS bank cost an extra wavefront (ncu on the Rössler
# network: 18 of the 32 gathers per evaluation, 6 % of all shared-memory wavefronts of the kernel).
ZERO_WORDS = 16


def row_length(n):
	"""doubles per shared-memory row of warp storage: the n components and the zero words, even (16-byte rows)"""
	return (n + ZERO_WORDS + 1) // 2 * 2


def _spread_over_banks(entries, width, zero_slot=None, physical=lambda lane: lane):
	"""The terms of a sum may run in any order, and padding may sit anywhere: every statement's gathers are dealt
	to the `width` slots of its pass so that the lanes which execute a slot together (half a warp per 64-bit
	shared-memory wavefront) hit as few equal banks with different words as possible.  entries[lane]: positions.
	With zero_slot (the first of ZERO_WORDS zero words, one per bank) the padding entries are resolved as well: to the
	zero word of a bank that no other word of the same slot and half-warp occupies.  physical: the lane of the warp
	that executes entry `lane` (the blocked ring order rotates the lanes, see _print_loop)."""
	load = {}       # (slot, half-warp, bank) → words already read there
	placed = []
	for lane, wanted in enumerate(entries):
		row, free = [None] * width, list(range(width))
		for where in wanted:
			def cost(k):
				others = load.get((k, (physical(lane) % 32) // 16, where % 16), ())
				return 0 if where in others else len(others)
			best = min(free, key=lambda k: (cost(k), k))
			free.remove(best)
			row[best] = where
			load.setdefault((best, (physical(lane) % 32) // 16, where % 16), set()).add(where)
		placed.append(row)
	if zero_slot is not None:
		for lane, row in enumerate(placed):
			half = (physical(lane) % 32) // 16
			for k, where in enumerate(row):
				if where is None:
					# the emptiest bank of this slot and half-warp (a zero word already read there is the same word: no cost)
					def crowd(bank):
						zero = zero_slot + (bank - zero_slot) % 16
						return len(load.get((k, half, bank), set()) - {zero})
					bank = min(range(16), key=lambda b: (crowd(b), b))
					zero = zero_slot + (bank - zero_slot) % 16
					row[k] = zero
					load.setdefault((k, half, bank), set()).add(zero)
	return placed


def _print_loop(printer, loop, group, position, tables, lines, zero_slot, team, blocked=None, prologue=None, base_of=None):
	"""one variable-length sum of a group, in ELL form: all lanes run the same (unrolled) number of terms.
	team: lanes that share the group's statements (lane ℓ runs statements ℓ, ℓ+team, …); blocked: passes per lane if
	the group is printed in the blocked ring order (then team = lanes in use, see _ring_split); prologue: lines that
	go in front of the loop over the passes."""
	size = group.size
	lengths = [loop.start[m+1] - loop.start[m] for m in range(size)]
	width = max(lengths)
	name = loop.symbol.name
	# terms ordered by distance from the statement's own component: lanes m, m+1, … then read neighbouring words
	order = []
	for m in range(size):
		terms = list(range(loop.start[m], loop.start[m+1]))
		if loop.n_index and group.outs[m] >= 0:
			own = position[group.outs[m]]
			terms.sort(key=lambda k: (position[loop.indices[0][k]] - own) % len(position))
		order.extend(terms)
	consts = [[slot[k] for k in order] for slot in loop.consts]
	indices = [[position[slot[k]] for k in order] for slot in loop.indices]
	padded = _paddable(loop)

	if padded and loop.n_index == 1 and not any(not _uniform(values) for values in consts) and 8 * zero_slot < 65536 and not tables.wide:
		# bare gather Σ g(y[idx]).  Shared-memory bandwidth is what this kernel runs out of, so the table is
		# as small as it gets: 16-bit byte offsets, two per 32-bit word (one wavefront per 64 gathers), and every
		# pass of 32 statements is padded only to ITS longest statement.
		n_pass = (size + team - 1) // team
		per_statement = [indices[0][loop.start[m]:loop.start[m+1]] for m in range(size)]

		# Diagonals: in lattice-like networks most statements read the SAME relative neighbours (m+1, m-1, m+2 …
		# within a ring of positions lo … lo+L-1).  Such a neighbour needs no table at all — lane m reads position
		# lo + (m + d) mod L, consecutive lanes read consecutive words (no bank conflict), and where no lane of a
		# pass wraps around the ring the address is the lane's own base plus a compile-time offset.  A bit mask per
		# statement says which diagonals it really has.  Everything else stays in the padded pair table.
		diagonals, lo, ring = [], 0, 0
		gathered = [p for terms in per_statement for p in terms]
		distance = None
		if gathered and min(group.outs) >= 0:
			if blocked:
				# ring coordinates instead of places: place q·team + ℓ of the gathered group holds element ℓ·blocked + q
				spans = {base_of(p) for p in gathered}
				if len(spans) == 1 and next(iter(spans))[1] == size:
					lo, ring = next(iter(spans))[0], size
					element = lambda place: (place % team) * blocked + place // team
					distance = lambda m, p: (element(p - lo) - element(m)) % ring
			else:
				lo, ring = min(gathered), max(gathered) - min(gathered) + 1
				distance = lambda m, p: (p - lo - m) % ring
		if distance:
			popularity = {}
			for m, terms in enumerate(per_statement):
				for p in terms:
					d = distance(m, p)
					popularity[d] = popularity.get(d, 0) + 1
			diagonals = sorted(d for d, count in popularity.items() if count >= 0.6 * size)[:32]
			if len(diagonals) < 4 or ring > 32767:
				diagonals = []
		masks = []
		if diagonals:
			where = {d: k for k, d in enumerate(diagonals)}
			remainder = []
			for m, terms in enumerate(per_statement):
				mask, rest, seen = 0, [], set()
				for p in terms:
					d = distance(m, p)
					if d in where and d not in seen:
						mask |= 1 << where[d]
						seen.add(d)
					else:
						rest.append(p)
				masks.append(mask)
				remainder.append(rest)
			per_statement = remainder
		# Nearly complete rings (blocked order): testing a mask bit for every neighbour of every statement costs four
		# instructions per term.  Where few neighbours are absent, every statement instead sums ALL the ring
		# neighbours (no masks; what all passes of a lane have in common is added once) and the absent ones are read
		# through a second pair table and subtracted.
		missing = None
		if diagonals and blocked and SUBTRACT_MISSING:
			full = (1 << len(diagonals)) - 1
			absent = sum(bin(full ^ mask).count("1") for mask in masks)
			if 0 < absent <= 0.25 * len(diagonals) * size:
				place = lambda e: (e % blocked) * team + e // blocked
				missing = [[lo + place((element(m) + d) % ring) for k, d in enumerate(diagonals) if not mask >> k & 1] for m, mask in enumerate(masks)]
				masks = [full] * size

		# one table for both kinds where the offsets leave a bit for the sign: padded to the longest statement of a pass
		# once, not once per kind
		# Rotated lanes (blocked order): lane p of the warp does the work of ring block (p + R) mod team.  A window word
		# at relative block c is then read by lanes 0 … 15 — one 64-bit wavefront — at blocks R+c … R+15+c: consecutive
		# words, no wrap-around inside the wavefront, as long as −c ≤ R ≤ team − 16 − c for every c.  Unrotated, the
		# wrap put words 16 apart (= the same bank) into one wavefront for every negative c (ncu: 10 of the 26 window
		# loads of the Rössler network took 3 wavefronts instead of 2).
		rotation = 0
		if diagonals and blocked:
			window_carries = [divmod(q + (d if d <= ring // 2 else d - ring), blocked)[0] for q in range(blocked) for d in diagonals]
			rotation = max(0, min(-min(window_carries), team - 16 - max(window_carries)))
			if rotation < -min(window_carries):
				rotation = 0        # the ring is too short for a wavefront without wrap-around either way
		who = "jb_lane" if blocked else "lane"
		physical = (lambda lane: (lane - rotation) % team) if blocked else (lambda lane: lane)
		signed_pairs = bool(missing) and 8 * zero_slot < 32768 and SIGNED_PAIRS
		negative = [set(terms) for terms in missing] if signed_pairs else [set()] * size
		if signed_pairs:
			per_statement = [list(plus) + list(minus) for plus, minus in zip(per_statement, missing)]
			missing = None
		# Idle lanes as helpers (blocked order of a single warp that leaves lanes idle): a pass is padded to its longest
		# statement, but few statements are long (Rössler benchmark: 3.7 entries per statement on average, 8 at most).  The
		# lanes beyond `team` run the same gather loop on rows of their own that hold what did NOT fit `cap` slots of an
		# over-long statement (one such statement per helper lane and pass); its owner fetches the helper's sum with one
		# shuffle.  For the warp these gathers are free: an LDS.64 of 32 lanes takes the two wavefronts that one of 25 takes.
		helper_plan = _helper_plan(per_statement, team, n_pass, size) if (signed_pairs and HELPER_LANES and blocked and tables.single_warp and team_is_warp(team) and 8 * (zero_slot + ZERO_WORDS) <= 4096) else None
		if helper_plan:
			tables.helper_lanes = team
		if blocked and prologue is not None and not any("const int jb_lane" in line for line in prologue):
			# (helper lanes shadow a working lane of their own half-warp in everything but the gathers: the same addresses,
			# hence no additional shared-memory wavefronts, and nothing out of range)
			source = f"(lane >= {team} ? lane - {team - 16} : lane)" if helper_plan else "lane"
			shifted = f"{source} + {rotation}"
			prologue.insert(0, f"\tconst int jb_lane = {f'{shifted} >= {team} ? {shifted} - {team} : {shifted}' if rotation else source};")

		def pair_table(lists):
			"""padded table of 16-bit byte offsets, two per 32-bit word, pass after pass → (offset, bases, pairs per pass)"""
			words, bases, widths = [], [], []
			for p in range(n_pass):
				members = range(team*p, min(size, team*p + team))
				pass_width = max(len(lists[m]) for m in members)
				pass_width += pass_width % 2
				bases.append(len(words) // 2)      # in 32-bit words
				widths.append(pass_width // 2)
				placed = _spread_over_banks([lists[m] for m in members], pass_width, zero_slot, physical) if SPREAD_OVER_BANKS else [list(lists[m]) + [None] * (pass_width - len(lists[m])) for m in members]
				for kp in range(pass_width // 2):
					for lane in range(team):
						for k in (2*kp, 2*kp + 1):
							where = placed[lane][k] if lane < len(placed) else None
							words.append(8 * (zero_slot if where is None else where) | (0x8000 if where in negative[min(team*p + lane, size - 1)] else 0))
			return (tables.add_u(words, align=2), bases, widths) if words else None

		def helper_table(lists):
			"""the same table with 32 rows per pass, indexed by the lane of the warp: rows of the working lanes hold the first
			`cap` entries of their statement and, in bits 12-14 of the first entry, which lane beyond `team` has summed the
			rest (the last one holds no entries at all: its sum is the 0.0 of statements that fit) → (offset, bases, widths)"""
			words, bases, widths = [], [], []
			spare = 32 - team
			for p in range(n_pass):
				cap, owners = helper_plan[p]
				rows, owner_of_row, code = [[] for _ in range(32)], [None] * 32, [spare - 1] * 32
				for lane in range(team):
					m = team * p + (lane + rotation) % team       # the statement lane `lane` of the warp works on
					rows[lane], owner_of_row[lane] = lists[m][:cap], m
					if m in owners:
						helper = owners.index(m)
						rows[team + helper], owner_of_row[team + helper], code[lane] = lists[m][cap:], m, helper
				bases.append(len(words) // 2)
				widths.append(cap // 2)
				placed = _spread_over_banks(rows, cap, zero_slot) if SPREAD_OVER_BANKS else [row + [None] * (cap - len(row)) for row in rows]
				for kp in range(cap // 2):
					for lane in range(32):
						for k in (2*kp, 2*kp + 1):
							where = placed[lane][k]
							entry = 8 * (zero_slot if where is None else where)
							if owner_of_row[lane] is not None and where in negative[owner_of_row[lane]]:
								entry |= 0x8000
							if k == 0 and lane < team:
								entry |= code[lane] << 12
							words.append(entry)
			return tables.add_u(words, align=2), bases, widths

		def print_helped_pairs(table, first, second):
			offset, bases, widths = table
			lines.append("\t{")
			lines.append(f"\t\tconstexpr int pass_base[{n_pass}] = {{ {', '.join(map(str, bases))} }}, pass_pairs[{n_pass}] = {{ {', '.join(map(str, widths))} }};")
			lines.append(f"\t\tconst unsigned *const pairs = reinterpret_cast<const unsigned *>(JB_TU + {offset}) + pass_base[pass] + lane;")
			lines.append("\t\tconst unsigned jb_first = pairs[0];")
			lines.append("#pragma unroll")
			lines.append("\t\tfor (int k = 0; k < pass_pairs[pass]; ++k) {")
			lines.append("\t\t\tconst unsigned pair = k ? pairs[k * 32] : jb_first;")
			lines.append(f"\t\t\t{first} += jb::flip_sign({term.replace(reads, 'y.at(pair & 0x0fffu)')}, pair << 16);")
			lines.append(f"\t\t\t{second} += jb::flip_sign({term.replace(reads, 'y.at((pair >> 16) & 0x0fffu)')}, pair);")
			lines.append("\t\t}")
			lines.append("\t\t// what did not fit the lane's own slots was summed by one of the idle lanes")
			lines.append(f"\t\t{first} += {second};")
			lines.append(f"\t\t{second} = jb::shfl({first}, {team} + ((jb_first >> 12) & 7u));")
			lines.append("\t}")

		def print_pairs(table, first, second, sign):
			offset, bases, widths = table
			lines.append("\t{")
			lines.append(f"\t\tconstexpr int pass_base[{n_pass}] = {{ {', '.join(map(str, bases))} }}, pass_pairs[{n_pass}] = {{ {', '.join(map(str, widths))} }};")
			lines.append(f"\t\tconst unsigned *const pairs = reinterpret_cast<const unsigned *>(JB_TU + {offset}) + pass_base[pass] + {who};")
			lines.append("#pragma unroll")
			lines.append("\t\tfor (int k = 0; k < pass_pairs[pass]; ++k) {")
			lines.append(f"\t\t\tconst unsigned pair = pairs[k * {team}];")
			if signed_pairs:
				lines.append(f"\t\t\t{first} += jb::flip_sign({term.replace(reads, 'y.at(pair & 0x7fffu)')}, pair << 16);")
				lines.append(f"\t\t\t{second} += jb::flip_sign({term.replace(reads, 'y.at((pair >> 16) & 0x7fffu)')}, pair);")
			else:
				lines.append(f"\t\t\t{first} {sign}= {term.replace(reads, 'y.at(pair & 0xffffu)')};")
				lines.append(f"\t\t\t{second} {sign}= {term.replace(reads, 'y.at(pair >> 16)')};")
			lines.append("\t\t}")
			lines.append("\t}")

		extra_table = (helper_table if helper_plan else pair_table)(per_statement)
		missing_table = pair_table(missing) if missing else None
		for c, values in enumerate(consts):
			lines.append(f"\tconst double {loop.prefix}c{c} = {printer.doprint(sympy.Float(values[0]))};")
		term = printer.doprint(loop.body)
		reads = f"y({loop.prefix}i0)"
		assert reads in term
		lines.append(f"\tdouble {name}_a = 0.0, {name}_b = 0.0, {name}_c = 0.0, {name}_d = 0.0;")
		if diagonals:
			full = (1 << len(diagonals)) - 1
			all_present = all(mask == full for mask in masks)
			lines.append("\t{")
			if not all_present:
				# the mask is split into two 16-bit table entries
				low = tables.add_u([mask & 0xffff for mask in masks])
				high = tables.add_u([mask >> 16 for mask in masks])
				if not blocked:
					lines.append(f"\t\tconst unsigned present = JB_TU[{low} + m] | ((unsigned) JB_TU[{high} + m] << 16);")
			if blocked:
				# the window: every word any pass of this lane needs is read ONCE, in front of the passes, and added to the
				# sums of all the passes that have it as a neighbour (registers: two partial sums per pass, no window array)
				signed = [d if d <= ring // 2 else d - ring for d in diagonals]
				tag = lambda number: f"{'m' if number < 0 else 'p'}{abs(number)}"
				for c, values in enumerate(consts):
					prologue.append(f"\tconst double {loop.prefix}c{c} = {printer.doprint(sympy.Float(values[0]))};")
				users = {}
				for q in range(blocked):
					for k, d in enumerate(signed):
						users.setdefault(divmod(q + d, blocked), []).append((q, k))
				for carry in sorted({carry for carry, _ in users}):
					shifted = f"jb_lane {'-' if carry < 0 else '+'} {abs(carry)}"
					wrapped = f"{shifted} < 0 ? {shifted} + {team} : {shifted}" if carry < 0 else f"{shifted} >= {team} ? {shifted} - {team} : {shifted}"
					prologue.append(f"\tconst int {name}_l{tag(carry)} = {wrapped if carry else 'jb_lane'};")
				for q in range(blocked):
					prologue.append(f"\tdouble {name}_s{q}a = 0.0, {name}_s{q}b = 0.0, {name}_s{q}c = 0.0, {name}_s{q}d = 0.0;")
					if not all_present:
						prologue.append(f"\tconst unsigned {name}_present{q} = JB_TU[{low} + jb_lane + {q * team}] | ((unsigned) JB_TU[{high} + jb_lane + {q * team}] << 16);")
				shared_by_all = [key for key, wanted in users.items() if all_present and blocked > 1 and len(wanted) == blocked]
				if shared_by_all:
					prologue.append(f"\tdouble {name}_core_a = 0.0, {name}_core_b = 0.0, {name}_core_c = 0.0, {name}_core_d = 0.0;")
				for count, ((carry, source), wanted) in enumerate(sorted(users.items())):
					prologue.append("\t{")
					prologue.append(f"\t\tconst double {name}_w = {term.replace(reads, f'y({lo + source*team} + {name}_l{tag(carry)})')};")
					if (carry, source) in shared_by_all:
						prologue.append(f"\t\t{name}_core_{'abcd'[count % 4]} += {name}_w;")       # a neighbour of every pass of this lane
					else:
						for q, k in wanted:
							# a select on the VALUE keeps the selects off the dependent chain of additions
							value = f"{name}_w" if all_present else f"(({name}_present{q} & {1 << k}u) ? {name}_w : 0.0)"
							prologue.append(f"\t\t{name}_s{q}{('ab' if all_present else 'abcd')[count % (2 if all_present else 4)]} += {value};")
					prologue.append("\t}")
				if shared_by_all:
					prologue.append(f"\tconst double {name}_core = ({name}_core_a + {name}_core_b) + ({name}_core_c + {name}_core_d);")
				lines.append("\t\tswitch (pass) {      // a constant once the passes are unrolled")
				for q in range(blocked):
					lines.append(f"\t\tcase {q}: {name}_a = {name}_s{q}a + {name}_s{q}c; {name}_b = {name}_s{q}b + {name}_s{q}d{f' + {name}_core' if shared_by_all else ''}; break;")
				lines.append("\t\t}")
			for k, d in enumerate([] if blocked else diagonals):
				# first case: no thread of this pass wraps around the ring; second: all of them do (pass is a constant after unrolling)
				index = (f"(({team} * pass + {team - 1} + {d} < {ring}) ? m + {d} : (({team} * pass + {d} >= {ring}) ? m + {d} - {ring} : "
						f"(m + {d} >= {ring} ? m + {d} - {ring} : m + {d})))")
				value = term.replace(reads, f"y({lo} + {index})")
				accumulator = f"{name}_{'abcd'[k % 4]}"
				if all_present:
					lines.append(f"\t\t{accumulator} += {value};")
				else:
					lines.append(f"\t\tif (present & {1 << k}u) {accumulator} += {value};")
			lines.append("\t}")
		if extra_table and helper_plan:
			print_helped_pairs(extra_table, f"{name}_c", f"{name}_d")
		elif extra_table:
			print_pairs(extra_table, f"{name}_c", f"{name}_d", "+")
		if missing_table:
			# absent ring neighbours were summed with all the others: taken out again
			print_pairs(missing_table, f"{name}_c", f"{name}_d", "-")
		lines.append(f"\tconst double {name} = ({name}_a + {name}_b) + ({name}_c + {name}_d);")
		return

	# general terms: one ELL table per varying slot; padding terms vanish, or are skipped by a per-lane count
	lines.append(f"\tdouble {name} = 0.0;")
	lines.append("\t{")
	if not padded:
		lines.append(f"\t\tconst int count = JB_TU[{tables.add_u(lengths)} + m];")
	lines.append("#pragma unroll 4")
	lines.append(f"\t\tfor (int k = 0; k < {width}; ++k) {{")
	indent = "\t\t\t"
	if not padded:
		lines.append("\t\t\tif (k < count) {")
		indent = "\t\t\t\t"
	for c, values in enumerate(consts):
		if _uniform(values):
			lines.append(f"{indent}const double {loop.prefix}c{c} = {printer.doprint(sympy.Float(values[0]))};")
		else:
			offset = tables.add_f64(_ell(values, loop.start, size, width, 0.0))
			lines.append(f"{indent}const double {loop.prefix}c{c} = JB_TF[{offset} + k * {size} + m];")
	for j, values in enumerate(indices):
		offset = tables.add_u(_ell(values, loop.start, size, width, zero_slot))
		lines.append(f"{indent}const int {loop.prefix}i{j} = JB_TU[{offset} + k * {size} + m];")
	lines.append(f"{indent}{name} += {printer.doprint(loop.body)};")
	if not padded:
		lines.append("\t\t\t}")
	lines.append("\t\t}")
	lines.append("\t}")


def warp_statements(plan, team_warps=1, blocked_rings=True, fuse_groups=True):
	"""returns (lines of the body of Sys::rhs_warp, Tables, component_of) where component_of[q] is the
	component stored at position q of the kernel's vectors; position n is a slot that always holds 0.
	team_warps: how many warps share one trajectory (`lane` in the printed code is the rank within that team)."""
	team = 32 * team_warps
	printer = CudaPrinter()
	tables = Tables()
	tables.single_warp = team_warps == 1
	lines = []
	blockings = ring_blockings(plan, team) if blocked_rings else {}
	groups = [group.reordered(_blocked_order(group.size, *blockings[group.size])) if group.size in blockings else group for group in plan.groups]
	# Round-robin groups of several passes: a pass is padded to ITS longest sum, so the statements are dealt to the passes
	# by length — the longest sums share the first pass, the shortest the last (Erdős–Rényi graph of mean degree 20:
	# 32 + 23 + 20 + 14 padded gather slots per lane instead of 4 × ≈ 30).  All groups of one size take the same order,
	# so that the equations of one unit stay at the same place of their groups (and are fused as before).
	orders = {}
	for group in groups:
		if group.size in blockings or group.size in orders or group.size <= team or not group.loops or min(group.outs) < 0:
			continue
		weight = [sum(loop.start[m+1] - loop.start[m] for loop in group.loops) for m in range(group.size)]
		order = sorted(range(group.size), key=lambda m: (-weight[m], m))
		if order != list(range(group.size)):
			orders[group.size] = order
	groups = [group.reordered(orders[group.size]) if group.size in orders else group for group in groups]
	position = component_positions(plan, groups)
	zero_slot = plan.n
	spans = [(position[group.outs[0]], group.size) for group in groups]
	base_of = lambda p: next(span for span in spans if span[0] <= p < span[0] + span[1])

	def located(expr):
		"""y(i) with a literal index → y(position of i)."""
		return expr.replace(lambda e: _is_dynvar(e) and e.args[0].is_Integer, lambda e: Y(sympy.Integer(position[int(e.args[0])])))

	emitted = set()
	pending = list(plan.temporaries)
	helper_symbols = {symbol for symbol, _ in plan.helpers}
	def flush():
		progress = True
		while progress:
			progress = False
			for item in list(pending):
				needs = {s for s in item[1].free_symbols if s in helper_symbols or s.name.startswith("jb_call_")}
				if needs <= emitted:
					lines.append(f"const double {item[0].name} = {printer.doprint(located(item[1]))};")
					emitted.add(item[0])
					pending.remove(item)
					progress = True
	flush()

	def slot_preamble(prefix, consts, indices, where, indent):
		for c, values in enumerate(consts):
			if _uniform(values):
				lines.append(f"{indent}const double {prefix}c{c} = {printer.doprint(sympy.Float(values[0]))};")
			else:
				sign, offset = tables.shared_f64(values)
				lines.append(f"{indent}const double {prefix}c{c} = {sign}JB_TF[{offset} + {where}];")
		for j, values in enumerate(indices):
			values = [position[v] for v in values]
			if _uniform(values):
				lines.append(f"{indent}const int {prefix}i{j} = {values[0]};")
			elif where != "0" and _uniform([v - m for m, v in enumerate(values)]):
				lines.append(f"{indent}const int {prefix}i{j} = {values[0]} + {where};")      # consecutive positions: no table
			else:
				lines.append(f"{indent}const int {prefix}i{j} = JB_TU[{tables.add_u(values)} + {where}];")

	# helpers: loops spread over the lanes and reduced, the rest redundantly on every lane
	# Split reduction (one helper with one sum, a single warp per trajectory): the five shuffle steps of the sum over the
	# warp are a chain of dependent shuffles and additions — 9 % of the stall samples of the Rössler network's kernel for
	# 4.5 % of its instructions.  Its steps are handed to the first batch, which spreads them over the loads and additions
	# of its register window (every shuffle sits in a basic block of its own, so the printed order is the executed order).
	split_steps, split_tail = [], []
	can_split = SPLIT_REDUCTIONS and team_warps == 1 and len(plan.helpers) == 1 and len(plan.helpers[0][1].loops) == 1 and not pending
	for symbol, group in plan.helpers:
		if can_split:
			loop = group.loops[0]
			lines.append(f"double {loop.symbol.name}_v;")
		lines.append("{")
		for loop in group.loops:
			count = loop.start[1]
			if _bare_gather(loop):
				# a plain sum may run in any order: by position, so that lanes read consecutive words (and need no table)
				by_position = sorted(range(count), key=lambda k: position[loop.indices[0][k]])
				loop = group.reordered([0]).loops[group.loops.index(loop)]
				loop.indices = [[slot[k] for k in by_position] for slot in loop.indices]
				loop.consts = [[slot[k] for k in by_position] for slot in loop.consts]
			lines.append(f"\tdouble {loop.symbol.name}_acc = 0.0;")
			lines.append("#pragma unroll" if count <= 16 * team else "#pragma unroll 4")      # independent loads in flight, not one dependent load per trip
			lines.append(f"\tfor (int k = lane; k < {count}; k += {team}) {{")
			slot_preamble(loop.prefix, loop.consts, loop.indices, "k", "\t\t")
			lines.append(f"\t\t{loop.symbol.name}_acc += {printer.doprint(loop.body)};")
			lines.append("\t}")
			if can_split:
				lines.append(f"\t{loop.symbol.name}_v = {loop.symbol.name}_acc;")
				lines.append("}")
				split_steps = [f"{loop.symbol.name}_v = jb::Team<1>::butterfly<{k}>({loop.symbol.name}_v, jb_scratch);" for k in range(5)]
				saved, lines = lines, ["{", f"\tconst double {loop.symbol.name} = {loop.symbol.name}_v;"]
			else:
				lines.append(f"\tconst double {loop.symbol.name} = jb::Team<{team_warps}>::sum({loop.symbol.name}_acc, jb_scratch);")
		slot_preamble(group.prefix, group.consts, group.indices, "0", "\t")
		lines.append(f"\t{symbol.name} = {printer.doprint(group.body)};")
		lines.append("}")
		if can_split:
			split_tail, lines = lines, saved
		emitted.add(symbol)
		flush()
	assert not pending, "unresolvable temporaries"

	# groups of one size run in ONE loop over the lanes' statements (at most FUSE_MAX of them): what several of them
	# read — the oscillator's own x, y, z in each of its three equations — is then loaded once
	batches = OrderedDict()
	for g, group in enumerate(groups):
		shape, part = (group.size, blockings.get(group.size, (team, None))), 0
		while len(batches.get((shape, part), ())) >= (FUSE_MAX if fuse_groups else 1):
			part += 1
		batches.setdefault((shape, part), []).append((g, group))
	for ((size, (lanes, blocked)), _), members in batches.items():
		n_pass = (size + lanes - 1) // lanes
		header, prologue, stores = [], [], []
		tables.helper_lanes = None
		if blocked:
			header.append(f"if (lane < {lanes}) {{")
		if any(group.loops for _, group in members) and n_pass <= 16:
			# passes unrolled: every pass knows its own trip counts at compile time
			header.append("#pragma unroll")
			header.append(f"for (int pass = 0; pass < {n_pass}; ++pass) {{")
			header.append(f"\tconst int m = {'jb_lane' if blocked else 'lane'} + {lanes} * pass;")
			if not blocked:
				header.append(f"\tif (m >= {size}) break;")
		else:
			header.append(f"for (int m = {'jb_lane' if blocked else 'lane'}, pass = 0; m < {size}; m += {lanes}, ++pass) {{")
			header.append("\t(void) pass;")
		body, lines = lines, []
		for g, group in members:
			base = position[group.outs[0]]
			body.append(f"// group {g}: {group.size} statement(s), components at positions {base} … {base + group.size - 1}"
					+ (f"; blocked ring order, {lanes} lanes × {blocked} consecutive elements" if blocked else ""))
			for loop in group.loops:
				_print_loop(printer, loop, group, position, tables, lines, zero_slot, lanes, blocked, prologue, base_of)
			slot_preamble(group.prefix, group.consts, group.indices, "m", "\t")
			lines.append(f"\tconst double {group.prefix}value = {printer.doprint(group.body)};")
			stores.append(f"\tdy.set({base} + m, {group.prefix}value);")
		if blocked and tables.helper_lanes:
			# the idle lanes take part (gathers, shuffles); they shadow working lanes everywhere else and store nothing
			header[0] = "{"
			stores = [f"\tif (lane < {lanes}) " + store.strip() for store in stores]
		lines.extend(stores)
		lines.append("}")
		if blocked:
			lines.append("}")
		if blocked and not any("const int jb_lane" in line for line in prologue):
			prologue.insert(0, "\tconst int jb_lane = lane;")
		if split_steps:
			# the pending reduction: its steps between the blocks of this batch's window if all lanes of the warp get there
			# (no `if (lane < …)` around the batch) and the window is long enough, otherwise in one piece in front of it
			closing = [k for k, line in enumerate(prologue) if line == "\t}"]
			if blocked and header[0] == "{" and len(closing) >= 2 * len(split_steps):
				for step in reversed(range(len(split_steps))):
					where = closing[(step + 1) * len(closing) // (len(split_steps) + 1) - 1] + 1
					prologue.insert(where, "\t" + split_steps[step])
				prologue.extend("\t" + line for line in split_tail)
			else:
				body.extend(split_steps + split_tail)
			split_steps, split_tail = [], []
		lines = body + (header[:1] + prologue + header[1:] if blocked else header) + lines
	lines.extend(split_steps + split_tail)      # (no batch at all)
	helper_declarations = [f"double {symbol.name};" for symbol, _ in plan.helpers]
	component_of = [None] * plan.n
	for component, where in position.items():
		component_of[where] = component
	return helper_declarations + lines, tables, component_of
