"""
Bookkeeping for transversal Lyapunov exponents: the `GroupHandler` that jitcode_transversal_lyap
inherits from (jitcxde_common, not in the reference tree; behaviour derived from its call sites
_jitcode.py:782,803-852,869-873,887-892,915 — see SURVEY.md App. A).

The n variables are partitioned into groups that are identical on the synchronisation manifold.
The transformed system keeps, per group and in the order of `groups`, one *main* slot (the
synchronised state) followed by len(group)-1 *tangent* slots holding the differences of the
tangent-vector components of consecutive group members.
"""


class GroupHandler:
	def __init__(self, groups):
		self.groups = [sorted(int(i) for i in group) for group in groups]
		members = [i for group in self.groups for i in group]
		if sorted(members) != list(range(len(members))):
			raise ValueError("groups must be a partition of the indices 0 … n-1.")
		self.n_members = len(members)
		self.main_indices = []
		self.tangent_indices = []
		self._group_of = {}
		slot = 0
		for g, group in enumerate(self.groups):
			self.main_indices.append(slot)
			self.tangent_indices.extend(range(slot+1, slot+len(group)))
			slot += len(group)
			for member in group:
				self._group_of[member] = g

	def map_to_main(self, i):
		"""slot that holds the synchronised state of variable i's group."""
		return self.main_indices[self._group_of[i]]

	def iterate(self, sequence):
		"""walks the slots of the transformed system: the group index (int) at a main slot, the pair of
		entries of `sequence` (indexed by original variable) to be differenced at a tangent slot."""
		sequence = list(sequence)
		for g, group in enumerate(self.groups):
			yield g
			for a, b in zip(group[:-1], group[1:]):
				yield (sequence[a], sequence[b])

	def back_transform(self, z_vector):
		"""tangent-vector components of the original variables expressed through the difference coordinates
		stored in the tangent slots of `z_vector`, under the constraint that they sum to zero within each
		group (no component along the synchronisation manifold)."""
		result = [None] * self.n_members
		for g, group in enumerate(self.groups):
			k = len(group)
			base = self.main_indices[g]
			differences = [z_vector[base+1+j] for j in range(k-1)]   # d_j = δ(group[j]) − δ(group[j+1])
			first = sum(((k-1-j) * d for j, d in enumerate(differences)), 0) / k if k > 1 else 0
			value = first
			result[group[0]] = value
			for j in range(1, k):
				value = value - differences[j-1]
				result[group[j]] = value
		return result

	def extract_main(self, f_list):
		"""the representative equation per main slot (first member of each group)."""
		return {self.main_indices[g]: f_list[group[0]] for g, group in enumerate(self.groups)}
