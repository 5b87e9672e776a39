"""
Multi-GPU plumbing: one process per GPU (torchrun), the ensemble sharded over ranks.

Trajectories are independent, so the data path needs no collective at all: every
rank integrates its contiguous shard of the batch axis with its own kernel launches.  Collectives
appear only where the caller wants a global view afterwards:
  * gather_states ......... all_gather of the (B_g, n) result shards (NCCL over NVLink on GPUs, gloo on CPUs)
  * reduce_lyapunov ....... all_reduce(SUM) of [Σ λ, Σ λ², count] → ensemble mean and standard error
The reference has no counterpart (single process, single thread).
"""

import torch
import torch.distributed as dist


def world():
	"""(rank, world_size) of the default process group, (0, 1) without one."""
	if dist.is_available() and dist.is_initialized():
		return dist.get_rank(), dist.get_world_size()
	return 0, 1


def shard_bounds(B, world_size=None, rank=None):
	"""[begin, end) of this rank's contiguous shard of B trajectories; sizes differ by at most one."""
	if world_size is None or rank is None:
		rank, world_size = world()
	base, extra = divmod(B, world_size)
	begin = rank * base + min(rank, extra)
	return begin, begin + base + (1 if rank < extra else 0)


def shard(array, world_size=None, rank=None):
	"""this rank's rows of a (B, …) array or tensor."""
	begin, end = shard_bounds(len(array), world_size, rank)
	return array[begin:end]


def gather_states(local, B=None):
	"""(B, n) tensor of all ranks' shards, in rank order, on every rank.  `local` is this rank's (B_g, n) tensor
	(CUDA with the nccl backend, CPU with gloo); B_g may differ by one between ranks."""
	rank, size = world()
	if size == 1:
		return local
	counts = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
	all_counts = [torch.zeros_like(counts) for _ in range(size)]
	dist.all_gather(all_counts, counts)
	all_counts = [int(c.item()) for c in all_counts]
	width = max(all_counts)
	padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
	padded[:local.shape[0]] = local
	pieces = [torch.empty_like(padded) for _ in range(size)]
	dist.all_gather(pieces, padded)
	result = torch.cat([piece[:count] for piece, count in zip(pieces, all_counts)], dim=0)
	if B is not None and result.shape[0] != B:
		raise ValueError(f"gathered {result.shape[0]} trajectories, expected {B}")
	return result


class GatheredEnsemble:
	"""
	One ensemble of B trajectories sharded over the ranks of the default process group, with the result of every
	`integrate` call gathered on all ranks: each rank integrates its contiguous shard, its
	integrate kernel's result is written straight into the rank's slot of the gather buffer (no staging copy), and one
	in-place `all_gather_into_tensor` (ncclAllGather over NVLink) of (B/G)·n doubles per rank completes the global
	state.  Shards must be even (B divisible by the world size) for the in-place collective; otherwise `gather_states`
	is used on the shard.

	ODE : a jitcode object with its integrator set; Y0 / pars : the WHOLE ensemble ((B, n) and (B, p), host or device —
	every rank passes the same arrays and keeps only its shard).
	"""

	def __init__(self, ODE, Y0, pars=None, time=0.0):
		self.ODE = ODE
		self.rank, self.size = world()
		self.B = len(Y0)
		self.begin, self.end = shard_bounds(self.B, self.size, self.rank)
		self.even = self.B % self.size == 0
		if pars is not None and pars.shape[1]:
			ODE.set_parameters(pars[self.begin:self.end])
		ODE.set_initial_value(Y0[self.begin:self.end], time)
		device = ODE.integrator.device
		self.gathered = torch.empty((self.B, ODE.n), dtype=torch.float64, device=device)

	def integrate(self, T):
		"""advance all shards to T; returns the (B, n) CUDA tensor of the whole ensemble (valid on every rank)."""
		slot = self.gathered[self.begin:self.end]
		if self.even or self.size == 1:
			self.ODE.integrator.integrate(T, into=slot)
			if self.size > 1:
				dist.all_gather_into_tensor(self.gathered, slot)
		else:
			self.ODE.integrator.integrate(T)
			self.gathered.copy_(gather_states(self.ODE.integrator.y_device, self.B))
		return self.gathered


def reduce_lyapunov(local_exponents):
	"""ensemble mean and standard error of the mean of local Lyapunov exponents.

	local_exponents : (samples, B_g, n_lyap) tensor of this rank (what jitcode_lyap.integrate_many returns);
	each trajectory's time average counts as one sample of the ensemble (reference tests/test_jitcode.py:151-162
	takes the standard error over time; over an ensemble the trajectories are the independent samples)."""
	per_trajectory = local_exponents.mean(dim=0)                   # (B_g, n_lyap)
	stats = torch.stack([per_trajectory.sum(dim=0), (per_trajectory**2).sum(dim=0),
			torch.full_like(per_trajectory[0], per_trajectory.shape[0])])
	if world()[1] > 1:
		dist.all_reduce(stats, op=dist.ReduceOp.SUM)
	total, total_sq, count = stats
	mean = total / count
	variance = (total_sq / count - mean**2).clamp_min(0) * count / (count - 1).clamp_min(1)
	return mean, (variance / count).sqrt()
