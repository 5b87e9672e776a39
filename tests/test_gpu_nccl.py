"""
The multi-GPU path on hardware: two ranks over NCCL integrate the two halves of one ensemble;
every rank's integrate kernel writes its result into its slot of the gather buffer, one in-place all-gather
(ncclAllGather) completes it; Lyapunov averages are reduced with all_reduce.  The gathered result equals what ONE GPU
computes for the whole ensemble, bit for bit (trajectories are independent of their position in the batch).
Needs two GPUs (`gpurun --gpus 2`); skipped on a one-GPU box.
"""

import os
import socket

import numpy as np
import pytest
import torch

import workloads as W

pytestmark = pytest.mark.gpu


def _free_port():
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


def _systems():
	return {
		"lv": (W.lotka_volterra(gamma_as_parameter=True), W.lotka_volterra_ensemble(3000, seed=3, gamma_as_parameter=True), 15.0),
		"roessler": (W.roessler_network(100), W.roessler_ensemble(600, seed=4), 2.0),
	}


def _make(system):
	from jitcode_b200 import jitcode
	ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)
	ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12)
	return ODE


def _rank(rank, world_size, port, queue):
	import torch.distributed as dist
	from jitcode_b200 import _distributed as D, jitcode_lyap
	os.environ["MASTER_ADDR"] = "127.0.0.1"
	os.environ["MASTER_PORT"] = str(port)
	torch.cuda.set_device(rank)
	dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
	try:
		results = {}
		for name, (system, (Y0, pars), T) in _systems().items():
			ensemble = D.GatheredEnsemble(_make(system), Y0, pars)
			gathered = ensemble.integrate(T)
			results[name] = gathered.cpu().numpy()
			results[name + "_again"] = ensemble.integrate(1.5 * T).cpu().numpy()      # a second call continues every shard
		# Lyapunov averages: every rank's share of the trajectories, reduced over the ranks
		B = 64
		begin, end = D.shard_bounds(B)
		ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=2, verbose=False, seed=7)
		ODE.set_integrator("dopri5")
		state = np.tile(W.FHN_Y0, (B, 1)) + 1e-3 * np.random.default_rng(5).normal(size=(B, 4))
		tangents = np.random.default_rng(6).normal(size=(B, 2, 4))
		tangents /= np.linalg.norm(tangents, axis=2, keepdims=True)
		full = np.hstack([state, tangents.reshape(B, 8)])
		from jitcode_b200 import jitcode
		jitcode.set_initial_value(ODE, torch.as_tensor(full[begin:end]).cuda(), 0.0)
		states, lyaps = ODE.integrate_many([10.0 * k for k in range(1, 11)])
		mean, sem = D.reduce_lyapunov(lyaps)
		results["lyap_mean"], results["lyap_sem"] = mean.cpu().numpy(), sem.cpu().numpy()
		queue.put((rank, results))
	finally:
		dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_nccl_equal_one_gpu():
	import torch.multiprocessing as mp
	from jitcode_b200 import jitcode, jitcode_lyap
	port = _free_port()
	ctx = mp.get_context("spawn")
	queue = ctx.Queue()
	procs = [ctx.Process(target=_rank, args=(r, 2, port, queue)) for r in range(2)]
	for p in procs:
		p.start()
	results = dict(queue.get(timeout=600) for _ in procs)
	for p in procs:
		p.join(timeout=120)
		assert p.exitcode == 0
	torch.cuda.set_device(0)
	for name, (system, (Y0, pars), T) in _systems().items():
		ODE = _make(system)
		ODE.set_parameters(pars)
		ODE.set_initial_value(Y0, 0.0)
		single = ODE.integrate(T)
		again = ODE.integrate(1.5 * T)
		for rank in (0, 1):
			assert np.array_equal(results[rank][name], single), (name, rank)
			assert np.array_equal(results[rank][name + "_again"], again), (name, rank)
	# the reduced Lyapunov averages equal those of the whole ensemble on one GPU
	B = 64
	ODE = jitcode_lyap(W.fhn()["f_sym"], n_lyap=2, verbose=False, seed=7)
	ODE.set_integrator("dopri5")
	state = np.tile(W.FHN_Y0, (B, 1)) + 1e-3 * np.random.default_rng(5).normal(size=(B, 4))
	tangents = np.random.default_rng(6).normal(size=(B, 2, 4))
	tangents /= np.linalg.norm(tangents, axis=2, keepdims=True)
	jitcode.set_initial_value(ODE, np.hstack([state, tangents.reshape(B, 8)]), 0.0)
	states, lyaps = ODE.integrate_many([10.0 * k for k in range(1, 11)])
	per_trajectory = lyaps.mean(axis=0)
	for rank in (0, 1):
		np.testing.assert_allclose(results[rank]["lyap_mean"], per_trajectory.mean(axis=0), rtol=1e-12, atol=1e-15)
		np.testing.assert_allclose(results[rank]["lyap_sem"], per_trajectory.std(axis=0, ddof=1) / np.sqrt(B), rtol=1e-9)
