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
