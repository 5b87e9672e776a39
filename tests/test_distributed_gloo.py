"""
The multi-GPU host logic on CPU: two processes, gloo backend (the N>1 path of bench.py and of
jitcode_b200._distributed; NCCL runs the same calls on the GPU box).
"""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jitcode_b200 import _distributed as D


def _free_port():
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


def _worker(rank, world_size, port, B, queue):
	os.environ["MASTER_ADDR"] = "127.0.0.1"
	os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world_size)
	try:
		rng = np.random.default_rng(0)
		states = torch.as_tensor(rng.random((B, 3)))
		exponents = torch.as_tensor(rng.normal(size=(7, B, 2)))
		begin, end = D.shard_bounds(B)
		mine = D.shard(states)
		assert mine.shape[0] == end - begin
		gathered = D.gather_states(mine * 2.0, B)          # stand-in for "integrate my shard"
		mean, sem = D.reduce_lyapunov(exponents[:, begin:end])
		queue.put((rank, begin, end, gathered.numpy(), mean.numpy(), sem.numpy()))
	finally:
		dist.destroy_process_group()


def test_shard_bounds_partition():
	for B in (1, 7, 8, 1000003):
		for size in (1, 2, 3, 8):
			bounds = [D.shard_bounds(B, size, r) for r in range(size)]
			assert bounds[0][0] == 0 and bounds[-1][1] == B
			assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
			sizes = [e - b for b, e in bounds]
			assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo():
	B, world_size = 11, 2          # odd: the shards differ in size
	port = _free_port()
	ctx = mp.get_context("spawn")
	queue = ctx.Queue()
	procs = [ctx.Process(target=_worker, args=(r, world_size, port, B, queue)) for r in range(world_size)]
	for p in procs:
		p.start()
	results = [queue.get(timeout=120) for _ in procs]
	for p in procs:
		p.join(timeout=60)
		assert p.exitcode == 0
	rng = np.random.default_rng(0)
	states = rng.random((B, 3))
	exponents = rng.normal(size=(7, B, 2))
	per_trajectory = exponents.mean(axis=0)
	covered = np.zeros(B, dtype=int)
	for rank, begin, end, gathered, mean, sem in results:
		covered[begin:end] += 1
		np.testing.assert_allclose(gathered, 2.0 * states)             # every rank sees the whole ensemble, in order
		np.testing.assert_allclose(mean, per_trajectory.mean(axis=0), rtol=1e-12)
		np.testing.assert_allclose(sem, per_trajectory.std(axis=0, ddof=1) / np.sqrt(B), rtol=1e-10)
	assert np.all(covered == 1)


def test_single_process_is_identity():
	x = torch.arange(6.0).reshape(3, 2)
	assert D.gather_states(x) is x
	assert D.shard_bounds(5) == (0, 5)


class _StubIntegrator:
	"""stands where CudaEnsembleIntegrator stands in GatheredEnsemble: 'integrating' doubles the state"""
	device = torch.device("cpu")

	def __init__(self):
		self.rows = None

	def integrate(self, T, into=None):
		result = self.rows * 2.0 + T
		if into is not None:
			into.copy_(result)
			result = into
		self.rows = result
		return result

	@property
	def y_device(self):
		return self.rows


class _StubODE:
	n = 3

	def __init__(self):
		self.integrator = _StubIntegrator()
		self.pars = None

	def set_parameters(self, pars):
		self.pars = pars

	def set_initial_value(self, rows, time=0.0):
		self.integrator.rows = torch.as_tensor(rows).clone()


def _gather_worker(rank, world_size, port, B, queue):
	os.environ["MASTER_ADDR"] = "127.0.0.1"
	os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world_size)
	try:
		states = torch.as_tensor(np.random.default_rng(1).random((B, 3)))
		pars = torch.arange(float(B)).reshape(B, 1)
		ensemble = D.GatheredEnsemble(_StubODE(), states, pars)
		result = ensemble.integrate(0.5).clone()
		queue.put((rank, ensemble.begin, ensemble.end, ensemble.ODE.pars.numpy(), result.numpy()))
	finally:
		dist.destroy_process_group()


def test_gathered_ensemble_two_ranks_gloo():
	"""the strong-scaling path of bench.py: every rank integrates its shard into its slot of the gather buffer, one
	in-place all-gather completes it (even shards) or gather_states does (uneven shards)"""
	for B in (12, 11):
		port = _free_port()
		ctx = mp.get_context("spawn")
		queue = ctx.Queue()
		procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, B, queue)) for r in range(2)]
		for p in procs:
			p.start()
		results = [queue.get(timeout=120) for _ in procs]
		for p in procs:
			p.join(timeout=60)
			assert p.exitcode == 0
		states = np.random.default_rng(1).random((B, 3))
		for rank, begin, end, pars, result in results:
			np.testing.assert_allclose(pars[:, 0], np.arange(begin, end))
			np.testing.assert_allclose(result, 2.0 * states + 0.5)
