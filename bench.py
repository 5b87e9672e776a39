#!/usr/bin/env python
"""
bench.py — ensemble Dormand–Prince trajectory-steps/s (BASELINE.json's metric) for every BASELINE config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference] [--only-headline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N … bench.py --gpus N …

A *step* of the benchmark is one pass of the hot path over one batch: (re)initialise the ensemble and integrate
every trajectory from t = 0 through the workload's sample times.  A *trajectory-step* is one Runge–Kutta step attempt
(accepted or rejected) of one trajectory; they are counted on the device (per-trajectory counters of the kernel).

Headline (the top-level keys of the ONE JSON line rank 0 prints): BASELINE.json configs[2], the workload the north
star's ≥ 100× target is quoted on — small-world network of 100 Rössler oscillators (n = 300), 1 Mi trajectories per GPU
with per-trajectory coupling strength and initial state, dopri5 (Dormand–Prince 5(4) under Hairer's controller, what
examples/SW_of_Roesslers.py:107 uses), rtol 1e-6, atol 1e-12, t = 0 → 10.

`configs` (N = 1; the same record for every other BASELINE config, SURVEY.md §8d C2–C5, and for two networks on which
the lattice-specific paths of the code generator do not apply):
    lv               1 Mi Lotka–Volterra initial conditions, dopri5, t = 0 → 1000
    lv_RK45          the same under SciPy's solve_ivp controller with dense output ("RK45", BASELINE's wording)
    fhn_lyap         100k FitzHugh–Nagumo pairs with 4 tangent vectors, QR every Δt = 10, 20 calls in one launch
    kuramoto         1000 oscillators, dense adjacency, 4096 trajectories — as examples/kuramoto_network.py:29,34:
                     dop853, atol 1e-6, rtol 0, samples at t = 1 … 50
    kuramoto_dopri5  the same under dopri5
    roessler_random  config 3's network rewired with probability 1 (no lattice left), 256 Ki trajectories
    roessler_er      Erdős–Rényi network of the same mean degree, 256 Ki trajectories
    roessler_t50     config 3 sampled every 10 time units up to t = 50 (the example's loop) in one launch, 256 Ki trajectories
`strong` (N > 1): 1 Mi trajectories IN TOTAL sharded over the ranks, every rank's result written by its integrate kernel's
unpack straight into its slot of the all-gather buffer, one in-place ncclAllGather inside the timed region.

Keys of every record
  value        device-timed (CUDA events) with the ensemble already resident in HBM
  e2e          through the public API with pinned HOST buffers, copies inside the timed region
  roofline     the integrate kernel against the limit its design has: "fp64" (state on chip: useful flops against the
               DFMA peak measured in this run), "tensor" (dense engine: against the measured DMMA peak), "hbm" (global
               storage: 280·n bytes per trajectory-step against MEASURED_PEAKS.json); `traffic` = DRAM bytes per launch
               from the committed ncu capture of that workload (profiles/traffic.json)
  cpu_baseline the reference CPU path (the reference's own extension template compiled ON THIS HOST with its own flags,
               under scipy.integrate, one process per core) on a bounded sample of the same ensemble
  parity       ≥ 256 trajectories (dense engine: 32) of the timed result against oracle/ on the same inputs
`--impl reference` times only the CPU path of the headline workload (or of --workload).
"""

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

import workloads as W  # noqa: E402

METRIC = "ensemble RK45 trajectory-steps/s"
UNIT = "trajectory-steps/s"
RTOL, ATOL = 1e-6, 1e-12
HEADLINE = "roessler"
CONFIGS = ("lv", "lv_RK45", "fhn_lyap", "kuramoto", "kuramoto_dopri5", "roessler_random", "roessler_er", "roessler_t50")
STRONG = ("roessler", "lv")
DEFAULT_TRAJECTORIES = {"roessler": 1 << 20, "lv": 1 << 20, "lv_RK45": 1 << 20, "fhn_lyap": 100_000, "kuramoto": 4096,
		"kuramoto_dopri5": 4096, "roessler_random": 1 << 18, "roessler_er": 1 << 18, "roessler_t50": 1 << 18}
CPU_SAMPLE_PER_CORE = {"roessler": 2048, "lv": 1024, "lv_RK45": 48, "fhn_lyap": 2048, "kuramoto": 2, "kuramoto_dopri5": 2,
		"roessler_random": 1024, "roessler_er": 1024, "roessler_t50": 256}


def workload(name, B, seed_offset=0):
	"""the synthetic input of one config: system, seeded ensemble, sample times, integrator, tolerances"""
	if name.startswith("roessler"):
		topology = {"roessler": "small_world", "roessler_random": "random", "roessler_er": "erdos_renyi", "roessler_t50": "small_world"}[name]
		system = W.roessler_network(100, topology=topology)
		Y0, pars = W.roessler_ensemble(B, seed=43 + seed_offset)
		label = {"small_world": "SW_of_Roesslers N=100 (n=300), per-trajectory k and Y0",
				"random": "Roessler network N=100 (n=300), ring rewired with p=1 (random graph, degree 20)",
				"erdos_renyi": "Roessler network N=100 (n=300), Erdos-Renyi graph of mean degree 20"}[topology]
		if name == "roessler_t50":
			# the example's sampling (outputs every 10 time units) as ONE launch; chaotic (λ ≈ 0.08): two implementations
			# of the same method agree to 1e-3 of the state's scale up to t = 50
			return {"kind": "generic", "system": system, "Y0": Y0, "pars": pars, "T": 50.0, "times": [10.0, 20.0, 30.0, 40.0, 50.0], "integrator": "dopri5",
					"rtol": RTOL, "atol": ATOL, "label": label + ", samples t=10..50 in one launch", "parity": {"n": 256, "rule": "scale", "bound": 1e-3}}
		# chaotic: agreement to 1e-6 of the state's scale at the horizon T = 10
		return {"kind": "generic", "system": system, "Y0": Y0, "pars": pars, "T": 10.0, "times": [10.0], "integrator": "dopri5",
				"rtol": RTOL, "atol": ATOL, "label": label, "parity": {"n": 256, "rule": "scale", "bound": 1e-6}}
	if name in ("lv", "lv_RK45"):
		system = W.lotka_volterra()
		Y0, pars = W.lotka_volterra_ensemble(B, seed=42 + seed_offset)
		return {"kind": "generic", "system": system, "Y0": Y0, "pars": pars, "T": 1000.0, "times": [1000.0],
				"integrator": "dopri5" if name == "lv" else "RK45", "rtol": RTOL, "atol": ATOL,
				"label": "Lotka-Volterra ensemble (n=2), random initial conditions",
				# neutrally stable orbits: the phase error grows linearly with t; 2500 steps of two implementations of
				# the same method stay within the 10·(atol + rtol·|y|) of the north star
				"parity": {"n": 256, "rule": "tolerance", "bound": 10.0, "horizon": 100.0}}
	if name == "fhn_lyap":
		system = dict(W.fhn(), lyap=4)
		return {"kind": "lyap", "system": system, "Y0": W.fhn_lyap_ensemble(B, seed=42 + seed_offset), "pars": np.empty((B, 0)), "T": 200.0,
				"times": [10.0 * (i + 1) for i in range(20)], "integrator": "dopri5", "rtol": RTOL, "atol": ATOL,
				"label": "double_fhn_lyapunov: FHN pair + 4 tangent vectors (n=20), QR every dt=10, 20 calls",
				"parity": {"n": 256, "rule": "lyap"}}
	if name in ("kuramoto", "kuramoto_dopri5"):
		# the symbolic form has 2·10^5 sine terms (SymPy needs minutes to build it), so the system enters through its
		# data — the same kernels the recognised symbolic form ends up in (tests/test_gpu_dense.py)
		data = W.kuramoto_data(1000)
		n = data["n"]
		Wm = data["c"] / (n - 1) * data["A"].T.astype(float)
		np.fill_diagonal(Wm, 0.0)
		example = name == "kuramoto"
		return {"kind": "dense", "system": {"n": n, "sine_coupling": (Wm, data["omega"]), "control_pars": [], "data": data},
				"Y0": W.kuramoto_ensemble(B, n, seed=44 + seed_offset), "pars": np.empty((B, 0)), "T": 50.0,
				"times": [float(T) for T in range(1, 51)], "integrator": "dop853" if example else "dopri5",
				"rtol": 0.0 if example else RTOL, "atol": 1e-6 if example else 1e-9,
				"label": "kuramoto_network n=1000 dense adjacency, samples t=1..50",
				"parity": {"n": 32, "rule": "tolerance", "bound": 10.0}}
	raise SystemExit(f"unknown workload {name}")


def rk_flops_per_component(integrator):
	"""FP64 operations per component and step attempt outside the right-hand side: stage combinations (2 per non-zero
	coefficient + 2 for y + h·Σ), error estimate and scale (72 for the 5th-order pair)."""
	if integrator in ("dopri5", "RK45"):
		return 72.0
	from scipy.integrate._ivp import dop853_coefficients as co
	nonzero = int(np.count_nonzero(co.A[:12])) + int(np.count_nonzero(co.B))
	return 2.0 * nonzero + 2.0 * 12 + 2.0 * (np.count_nonzero(co.E5) + np.count_nonzero(co.E3)) + 14.0


# ------------------------------------------------------------------ the reference CPU path (oracle)

def cpu_reference_handle(wl):
	"""(handle, kind, how): the reference's own extension template around the system's right-hand side, compiled on this
	host with the reference's flags (oracle/build_ref.native_handle) when oracle/_ref holds it, else the oracle's port;
	for the dense network the loop-form port of oracle/network_rhs.py (the written-out form cannot be compiled)."""
	from oracle import build_ref, c_rhs
	system = wl["system"]
	if wl["kind"] == "dense":
		from oracle import network_rhs
		data = system["data"]
		return (network_rhs.NetworkRhsHandle(data["A"], data["omega"], data["c"]), "port",
				"loop-form C right-hand side (the reference's written-out form: 2e5 sine terms, 8.4 MB of C, does not compile in reasonable time), -Ofast -march=native")
	if system.get("lyap"):
		full, helpers = c_rhs.lyapunov_system(system["f_sym"], system["lyap"], system["helpers"], system["n"])
		system = {"f_sym": full, "helpers": helpers, "control_pars": [], "n": len(full)}
	try:
		handle = build_ref.build_ref_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"])
		if handle is not None:
			handle, how = build_ref.native_handle(handle)
			handle.f  # noqa: B018  (import check: the file was built for this Python)
			return handle, "reference", "jitced_template.c " + how
	except Exception:
		pass
	return c_rhs.build_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"]), "port", "oracle/rhs_module.c.in, -Ofast -march=native"


def cpu_run(wl, handle, n_sample, integrator=None):
	from oracle import ensemble
	Y0, pars = wl["Y0"][:n_sample], wl["pars"][:n_sample]
	lyap = None
	if wl["system"].get("lyap"):
		# jitcode_lyap: random unit tangent vectors appended to the state, QR after every call (_jitcode.py:723-775)
		k, n = wl["system"]["lyap"], wl["system"]["n"]
		tangents = np.random.default_rng(1).normal(size=(len(Y0), k, n))
		tangents /= np.linalg.norm(tangents, axis=2, keepdims=True)
		Y0 = np.hstack([Y0, tangents.reshape(len(Y0), k * n)])
		lyap = (n, k)
	result = ensemble.run_ensemble(handle, integrator or wl["integrator"], Y0, wl["times"], pars=pars, lyap=lyap, rtol=wl["rtol"], atol=wl["atol"])
	return result["attempts"], result["seconds"], result["cores"]


def backend_name(integrator):
	return "scipy.integrate.ode" if integrator in ("dopri5", "dop853") else "scipy.integrate.solve_ivp steppers (IVP_wrapper)"


def cpu_baseline(wl, per_core, integrator=None):
	from oracle import ensemble
	handle, kind, how = cpu_reference_handle(wl)
	integrator = integrator or wl["integrator"]
	cores = ensemble.available_cores()
	n_sample = min(len(wl["Y0"]), per_core * cores)
	cpu_run(wl, handle, min(n_sample, 2 * cores), integrator)     # warm the pool and the page cache
	attempts, seconds, cores = cpu_run(wl, handle, n_sample, integrator)
	return {"value": attempts / seconds, "unit": UNIT, "cores": cores, "kind": kind, "integrator": integrator,
			"sample": f"{n_sample} trajectories of the same ensemble, {integrator} via {backend_name(integrator)}, t=0..{wl['T']:g}, {seconds:.1f} s; {how}"}


def run_reference(args):
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return None
	from oracle import ensemble
	cores = ensemble.available_cores()
	per_core = args.cpu_sample_per_core or CPU_SAMPLE_PER_CORE[args.workload]
	wl = workload(args.workload, per_core * cores)
	if args.integrator:
		wl["integrator"] = args.integrator
	handle, kind, how = cpu_reference_handle(wl)
	for _ in range(args.warmup):
		cpu_run(wl, handle, min(len(wl["Y0"]), 4 * cores))
	attempts = seconds = 0.0
	for _ in range(args.steps):
		a, s, cores = cpu_run(wl, handle, len(wl["Y0"]))
		attempts += a
		seconds += s
	value = attempts / seconds
	return {
		"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
		"warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps, "higher_is_better": True, "scaling": "weak",
		"vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": wl["label"], "integrator": wl["integrator"], "rtol": wl["rtol"], "atol": wl["atol"], "t_end": wl["T"],
				"trajectories_per_step": len(wl["Y0"])},
		"cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
				"sample": f"{len(wl['Y0'])} trajectories per step, one process per core, {wl['integrator']} via {backend_name(wl['integrator'])}; {how}"},
		"e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}


# ------------------------------------------------------------------ clocks

CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
		"clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

class ClockSampler:
	def __init__(self, device_index):
		self.device_index = device_index
		self.process = None

	def __enter__(self):
		try:
			self.process = subprocess.Popen(
				["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={CLOCK_QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
				stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
		except OSError:
			self.process = None
		return self

	def __exit__(self, *exc):
		self.samples = []
		if self.process is None:
			return
		self.process.terminate()
		try:
			out, _ = self.process.communicate(timeout=5)
		except subprocess.TimeoutExpired:
			self.process.kill()
			out, _ = self.process.communicate()
		for line in out.splitlines():
			parts = [p.strip() for p in line.split(",")]
			if len(parts) >= 9:
				try:
					self.samples.append({"sm": float(parts[1]), "max": float(parts[2]), "power": float(parts[3]),
							"hw_slowdown": parts[5], "hw_thermal": parts[6], "sw_thermal": parts[7], "sw_power_cap": parts[8]})
				except ValueError:
					pass

	def summary(self):
		if not getattr(self, "samples", None):
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
		busy = [s for s in self.samples if s["power"] > 200] or self.samples
		reasons = [key for key in ("hw_slowdown", "hw_thermal", "sw_thermal", "sw_power_cap")
				if any(s[key].lower().startswith("active") for s in self.samples)]
		return {"sm_mhz": statistics.median(s["sm"] for s in busy), "sm_max_mhz": max(s["max"] for s in self.samples),
				"reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------ our arm

class Bench:
	"""one process of the benchmark (one GPU); `record(name, …)` measures one workload"""

	def __init__(self, args):
		import torch
		import torch.distributed as dist
		self.torch, self.dist, self.args = torch, dist, args
		self.world = int(os.environ.get("WORLD_SIZE", "1"))
		self.rank = int(os.environ.get("RANK", "0"))
		self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
		if not torch.cuda.is_available():
			raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the CPU arm).")
		torch.cuda.set_device(self.local_rank)
		self.device = torch.device("cuda", self.local_rank)
		if self.world > 1:
			dist.init_process_group("nccl", device_id=self.device)

	def close(self):
		if self.world > 1:
			self.dist.destroy_process_group()

	def barrier(self):
		if self.world > 1:
			self.dist.barrier()
		self.torch.cuda.synchronize()

	def make_ode(self, wl):
		from jitcode_b200 import jitcode, jitcode_lyap
		args, system = self.args, wl["system"]
		if wl["kind"] == "lyap":
			ODE = jitcode_lyap(system["f_sym"], n_lyap=system["lyap"], verbose=False, seed=1)
			ODE.set_integrator(wl["integrator"], rtol=wl["rtol"], atol=wl["atol"], **({"storage": args.storage} if args.storage else {}))
		elif wl["kind"] == "dense":
			ODE = jitcode.from_sine_coupling(*system["sine_coupling"])
			ODE.set_integrator(wl["integrator"], rtol=wl["rtol"], atol=wl["atol"])
		else:
			ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)
			ODE.set_integrator(wl["integrator"], rtol=wl["rtol"], atol=wl["atol"], storage=args.storage, kregs=args.kregs)
		return ODE

	# ---- parity of the timed result against the oracle

	def parity(self, wl, ODE, result, lyap_initial=None, lyaps=None):
		"""result: device tensor — (B, n) final states, or (n_times, B, n) samples (dense engine) — of the last timed
		device step; compares the first trajectories with the oracle on the same inputs."""
		from oracle import ensemble
		spec = wl["parity"]
		k = min(spec["n"], len(wl["Y0"]))
		handle, kind, _ = cpu_reference_handle(wl)
		record = {"trajectories": k, "oracle": f"oracle/ ({kind}) under {backend_name(wl['integrator'])}"}
		if spec["rule"] == "lyap":
			n, nl = wl["system"]["n"], wl["system"]["lyap"]
			calls = 5                                       # t ≤ 50: the pair is weakly chaotic (λ₁ ≈ 0.007)
			ref = ensemble.run_ensemble(handle, wl["integrator"], lyap_initial[:k].cpu().numpy(), wl["times"][:calls], lyap=(n, nl), rtol=wl["rtol"], atol=wl["atol"])
			ours = lyaps[:calls, :k].cpu().numpy()
			dt = np.diff(np.r_[0.0, wl["times"][:calls]])[:, None, None]
			# A tangent vector that shrinks by more than ~e^-20 relative to the leading one between two renormalisations is
			# lost to round-off in ANY implementation (on this ensemble the two trailing exponents are ≈ -3.8: e^-38 per
			# call; the oracle's own QR and a Gram–Schmidt of the same vectors differ by 0.5 there) — compared are the
			# exponents the arithmetic determines
			determined = ref["lyaps"] * dt > -20.0
			deviation = np.abs(ours - ref["lyaps"])
			allowed = 1e-6 + 1e-4 * np.abs(ref["lyaps"])
			record.update({"compared": f"local Lyapunov exponents of the first {calls} calls (t <= {wl['times'][calls-1]:g}) with lambda*dt > -20: {int(determined.sum())} of {determined.size}",
					"tolerance": "1e-6 + 1e-4*|lambda|", "max_ratio": float((deviation / allowed)[determined].max()), "max_abs": float(deviation[determined].max()),
					"max_abs_undetermined": float(deviation[~determined].max()) if (~determined).any() else 0.0})
			record["ok"] = bool(record["max_ratio"] <= 1.0 and determined.sum() >= determined.size // 4)
			return record
		times = wl["times"]
		run = lambda Y0, sample_times, integrator=wl["integrator"], **tol: ensemble.run_ensemble(      # noqa: E731
				handle, integrator, Y0, sample_times, pars=wl["pars"][:len(Y0)], **({"rtol": wl["rtol"], "atol": wl["atol"]} | tol))["y"]
		ref = run(wl["Y0"][:k], times)
		ours = result[..., :k, :].cpu().numpy()
		if ours.ndim == 2:
			ours, ref = ours[None], ref[-1:]
		deviation = np.abs(ours - ref)
		if spec["rule"] == "scale":
			allowed = spec["bound"] * np.abs(ref).max(axis=-1, keepdims=True)
			record["tolerance"] = f"{spec['bound']:g} * max|y| per trajectory (chaotic; horizon t = {wl['T']:g})"
		else:
			allowed = spec["bound"] * (wl["atol"] + wl["rtol"] * np.abs(ref))
			record["tolerance"] = f"{spec['bound']:g} * (atol + rtol*|y|)"
		record.update({"compared": f"state at t = {', '.join(f'{T:g}' for T in times[-3:])}{' (and all earlier samples)' if len(times) > 3 and ours.shape[0] > 1 else ''}",
				"max_ratio": float((deviation / allowed).max()), "max_abs": float(deviation.max())})
		record["ok"] = bool(record["max_ratio"] <= 1.0)
		if spec.get("horizon") and not record["ok"]:
			# Beyond the horizon the SEQUENCE OF STEP SIZES of the reference's own integrator depends on round-off (measured:
			# SciPy's dopri5 and its step-for-step restatement oracle/rk_numpy.py, identical for 300 time units, drift apart
			# in their step times from t ≈ 400 on); every such sequence is a valid discretisation whose distance from the
			# true solution is the global error, thousands of times the local tolerance after 2500 steps.  So: (1) the same
			# trajectories, same kernel, to the horizon — within the tolerance of the north star; (2) the timed result —
			# as close to a tight-tolerance solution as the oracle is.
			horizon = spec["horizon"]
			ODE.set_initial_value(wl["Y0"][:k], 0.0)
			short = np.asarray(ODE.integrate(horizon))
			short_ref = run(wl["Y0"][:k], [horizon])[0]
			ratio = float((np.abs(short - short_ref) / (spec["bound"] * (wl["atol"] + wl["rtol"] * np.abs(short_ref)))).max())
			truth = run(wl["Y0"][:k], times, "dop853", rtol=1e-13, atol=1e-15)[-1]
			scale = wl["atol"] + wl["rtol"] * np.abs(truth)
			error_ours, error_oracle = float((np.abs(ours[-1] - truth) / scale).max()), float((np.abs(ref[-1] - truth) / scale).max())
			record.update({"timed_result_vs_oracle": {"max_ratio": record["max_ratio"], "max_abs": record["max_abs"]},
					"horizon": {"t": horizon, "max_ratio": ratio, "tolerance": record["tolerance"]},
					"global_error_in_units_of_atol_plus_rtol_y": {"ours": error_ours, "oracle": error_oracle, "truth": "dop853, rtol 1e-13"},
					"compared": f"state at t = {horizon:g} against the oracle (step sequences still locked); state at t = {wl['T']:g} against a tight-tolerance solution, next to the oracle's own distance from it",
					"max_ratio": ratio})
			record["ok"] = bool(ratio <= 1.0 and error_ours <= 1.25 * error_oracle)
		return record

	# ---- one workload

	def record(self, name, steps, warmup, B, cpu=True, parity=True, integrator=None, min_seconds=0.0):
		"""min_seconds: the `configs` records time at least that long (more steps than asked for if a step is short):
		nvidia-smi delivers its first clock sample a few hundred milliseconds after it was started"""
		torch = self.torch
		args, world, rank, device = self.args, self.world, self.rank, self.device
		wl = workload(name, B, seed_offset=1000 * rank)      # every rank integrates its own shard
		if integrator:
			wl["integrator"] = integrator
		ODE = self.make_ode(wl)
		solver = ODE.integrator
		module = solver.module
		kind = wl["kind"]
		n = ODE.n
		times = wl["times"]

		Y0_host = torch.as_tensor(wl["Y0"]).pin_memory()
		pars_host = torch.as_tensor(wl["pars"]).pin_memory() if wl["pars"].shape[1] else None
		Y0_dev = Y0_host.to(device)
		pars_dev = pars_host.to(device) if pars_host is not None else None
		out_host = torch.empty_like(Y0_host).pin_memory() if kind == "generic" else None
		keep = {}

		def device_step():
			if pars_dev is not None:
				ODE.set_parameters(pars_dev)
			ODE.set_initial_value(Y0_dev, 0.0)
			if kind == "lyap":
				keep["initial"] = solver.y_device
				states, exponents = ODE.integrate_many(times, record_states=False)     # the calls of the example's loop as one launch
				keep["lyaps"] = exponents
				return states
			if kind == "dense" or len(times) > 1:
				return ODE.integrate_many(times)
			return ODE.integrate(wl["T"])

		def host_step():
			"""the same through host buffers: what a user of the reference's API with NumPy arrays gets"""
			if kind == "lyap":
				ODE.set_initial_value(Y0_host, 0.0)
				states, exponents = ODE.integrate_many(times, record_states=False)     # the result of such a run: the exponents
				return exponents
			if kind == "dense":
				ODE.set_initial_value(Y0_host, 0.0)
				return ODE.integrate_many(times)                                       # every sample, as the example stores them
			if pars_host is not None:
				ODE.set_parameters(pars_host)
			if len(times) > 1:
				ODE.set_initial_value(Y0_host, 0.0)
				return ODE.integrate_many(times)                                       # every sample back to the host
			if hasattr(solver, "integrate_streamed") and B >= 65536:
				# the public call for host-resident ensembles: uploads and downloads overlap the kernels chunk by chunk
				return ODE.integrate_streamed(Y0_host, wl["T"], out_host, 0.0, chunks=args.chunks)
			ODE.set_initial_value(Y0_host, 0.0)
			return ODE.integrate(wl["T"], out=out_host)

		# ---- device-resident arm
		for _ in range(warmup):
			device_step()
		self.barrier()
		if min_seconds > 0.0:
			probe_start, probe_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			probe_start.record()
			device_step()
			probe_end.record()
			torch.cuda.synchronize()
			wanted = torch.tensor([min_seconds / max(probe_start.elapsed_time(probe_end) * 1e-3, 1e-4)], device=device)
			if world > 1:
				self.dist.all_reduce(wanted, op=self.dist.ReduceOp.MAX)      # every rank times the same number of steps
			steps = max(steps, min(400, int(math.ceil(wanted.item()))))
		solver.reset_step_counts()
		solver.kernel_events = []
		launches_before = module.launch_count()
		start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		with ClockSampler(self.local_rank) as clocks:
			self.barrier()
			start.record()
			for _ in range(steps):
				result = device_step()
			end.record()
			self.barrier()
		seconds = start.elapsed_time(end) * 1e-3
		launches = module.launch_count() - launches_before
		accepted, rejected = solver.step_counts()
		attempts = accepted + rejected
		kernel_ms = sum(a.elapsed_time(b) for a, b in solver.kernel_events) / steps     # integrate kernels of one step
		solver.kernel_events = None
		rounds = getattr(solver, "rounds", None)
		parity_record = None
		if parity and rank == 0:
			try:
				parity_record = self.parity(wl, ODE, result, keep.get("initial"), keep.get("lyaps"))
			except Exception as error:      # the line must still be printed; a parity check that could not run is not "ok"
				parity_record = {"ok": False, "error": f"{type(error).__name__}: {error}"}
		del result

		# ---- end-to-end arm: pinned host buffers in and out through the public API
		# (the warm-up also lets torch's caching allocator for page-locked memory settle; it keeps its result until the next
		# one exists, like the timed loop and like a caller does, so that both blocks the steady state needs get allocated
		# here: page-locking 1.6 GB of samples of config 5 takes a second)
		returned = None
		for _ in range(warmup):
			returned = host_step()
		self.barrier()
		solver.reset_step_counts()
		t0 = time.perf_counter()
		for _ in range(steps):
			t_step = time.perf_counter()
			returned = host_step()
			if os.environ.get("JITCODE_B200_BENCH_DEBUG"):
				print(f"[{name}] end-to-end step: {1e3 * (time.perf_counter() - t_step):.2f} ms", file=sys.stderr)
		self.barrier()
		e2e_seconds = time.perf_counter() - t0
		e2e_attempts = sum(solver.step_counts())
		d2h_bytes = int(returned.numel() if hasattr(returned, "numel") else returned.size) * 8
		del returned

		if world > 1:
			stats = torch.tensor([seconds, e2e_seconds], dtype=torch.float64, device=device)
			self.dist.all_reduce(stats, op=self.dist.ReduceOp.MAX)
			seconds, e2e_seconds = stats.tolist()
			counts = torch.tensor([attempts, e2e_attempts, accepted, launches], dtype=torch.float64, device=device)
			self.dist.all_reduce(counts, op=self.dist.ReduceOp.SUM)
			attempts, e2e_attempts, accepted, launches = counts.tolist()
		if rank != 0:
			return None

		from jitcode_b200 import _peaks
		import sympy
		own_attempts = (attempts / world) / steps              # per launch (per step of one rank)
		kernel_s = kernel_ms * 1e-3
		storage_name = {0: "registers", 1: "global", 2: "warp", 3: "shared", 4: "dense", 5: "quad"}.get(module.info.storage, str(module.info.storage))
		stages = 12 if wl["integrator"] in ("dop853", "DOP853") else 6
		if kind == "dense":
			rhs_flops = 4.0 * n * n + 30.0 * n                  # the two GEMM halves + sincos and the combination, per evaluation
		else:
			rhs_flops = float(sum(sympy.count_ops(entry) for entry in ODE._f_list))
		flops_per_step = rk_flops_per_component(wl["integrator"]) * n + stages * rhs_flops
		fp64_peak = _peaks.fp64_fma_tflops(device)
		fp64_achieved = flops_per_step * own_attempts / kernel_s / 1e12
		peaks_path = ROOT / "MEASURED_PEAKS.json"
		if peaks_path.exists():
			hbm_peak, hbm_source = json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (burst copy)"
		else:
			hbm_peak, hbm_source = 6650.0, "fallback of B200_PROFILING.md"
		hbm_achieved = 280.0 * n * own_attempts / kernel_s / 1e9     # 35·n·8 B per trajectory-step if the stages streamed through HBM
		traffic, traffic_note = None, None
		traffic_file = ROOT / "profiles" / "traffic.json"
		if traffic_file.exists():
			# dram__bytes_read.sum + dram__bytes_write.sum from the committed ncu --set full capture of this workload
			entry = json.loads(traffic_file.read_text()).get(f"{name}:{storage_name}")
			if entry and "per_trajectory_step" in entry:
				traffic = entry["per_trajectory_step"] * own_attempts
			elif entry and "per_launch" in entry:
				traffic = entry["per_launch"] * B / entry["trajectories"]
			elif entry and "per_gemm_launch" in entry:
				traffic, traffic_note = entry["per_gemm_launch"], "per launch of the GEMM kernel (one right-hand-side evaluation of the ensemble)"
		hbm_record = {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "peak_source": hbm_source,
				"algorithmic_bytes_per_trajectory_step": 280 * n}
		fp64_record = {"achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
				"flops_per_trajectory_step": flops_per_step, "peak_source": "DFMA microbenchmark measured in this run (jitcode_b200/_peaks.py; MEASURED_PEAKS.json has no FP64 figure)"}
		if storage_name == "global":
			roofline = dict(hbm_record, bound="hbm", fp64=fp64_record)
		elif kind == "dense":
			dmma_peak = _peaks.fp64_dmma_tflops(device)
			roofline = {"bound": "tensor", "achieved": fp64_achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": fp64_achieved / dmma_peak,
					"flops_per_trajectory_step": flops_per_step,
					"peak_source": "DMMA.8x8x4 microbenchmark measured in this run (jitcode_b200/_peaks.py); DFMA peak of the same run: %.1f TFLOP/s" % fp64_peak}
		else:
			# state and stages never leave the chip between the sample times: the FP64 pipe is the roof; what an
			# HBM-streaming design would have needed is kept as a second figure ("report both")
			roofline = dict(fp64_record, bound="fp64", hbm_streaming_equivalent=hbm_record)
		if traffic_note:
			roofline["traffic_note"] = traffic_note
		roofline.update({"traffic": traffic, "kernel_ms": kernel_ms,
				"kernel": "jb_dense_integrate: dgemm_dmma_pipelined_kernel + element-wise stage kernels" if kind == "dense" else "jb_integrate_kernel"})
		line = {
			"metric": METRIC, "value": attempts / seconds, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
			"ms_per_step": 1e3 * seconds / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
			"dtype": "f64", "data": "synthetic",
			"config": {"workload": wl["label"], "integrator": wl["integrator"], "rtol": wl["rtol"], "atol": wl["atol"], "t_end": wl["T"],
					"samples_per_step": len(times), "trajectories_per_gpu": B, "n": n, "storage": storage_name, "threads_per_block": module.info.threads_per_block,
					"parallelism": f"trajectory shards x{world}, no data-path collective",
					"l2": f"state per GPU {B*n*8/1e6:.0f} MB vs 126 MB L2 (no flush needed)" if B*n*8 > 2.5e8 else "ensemble re-packed from a fresh copy every step; state on chip during the launch",
					"accepted_fraction": accepted / max(1.0, attempts), "attempts_per_trajectory": attempts / world / steps / B},
			"e2e": {"value": e2e_attempts / e2e_seconds, "unit": UNIT,
					"h2d_bytes_per_step": int(Y0_host.numel() * 8 + (pars_host.numel() * 8 if pars_host is not None else 0)),
					"d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * e2e_seconds / steps},
			"gpu_launches": int(launches),
			"clocks": clocks.summary(),
			"roofline": roofline,
		}
		if rounds is not None:
			line["config"]["lock_step_rounds_per_step"] = rounds / max(1, steps)
		if parity_record is not None:
			line["parity"] = parity_record
		if cpu and world == 1 and not args.no_cpu_baseline:
			per_core = args.cpu_sample_per_core or CPU_SAMPLE_PER_CORE[name]
			line["cpu_baseline"] = cpu_baseline(wl, per_core)
			if name == HEADLINE and not integrator:
				# BASELINE.json says "RK45": the same pair under solve_ivp's controller is the slower CPU path (SciPy's
				# stepper is Python); dopri5 above is the honest denominator — both are reported
				line["cpu_baseline_RK45"] = cpu_baseline(wl, max(8, per_core // 16), "RK45")
		return line

	# ---- strong scaling with the result gathered over NCCL (N > 1)

	def strong(self, name, steps, warmup, total):
		"""`total` trajectories sharded over the ranks; every rank's integrate kernel writes its result into its slot of
		the all-gather buffer, one in-place ncclAllGather completes the (total, n) state on every rank; rank 0 copies it
		to the host.  Timed end to end per step: H2D of the shard, integrate, all-gather, D2H."""
		torch, dist = self.torch, self.dist
		from jitcode_b200 import _distributed as D
		wl = workload(name, total)                          # the same ensemble on every rank
		ODE = self.make_ode(wl)
		solver = ODE.integrator
		begin, end = D.shard_bounds(total, self.world, self.rank)
		Y0_host = torch.as_tensor(wl["Y0"][begin:end]).pin_memory()
		pars_host = torch.as_tensor(wl["pars"][begin:end]).pin_memory() if wl["pars"].shape[1] else None
		gathered = torch.empty((total, ODE.n), dtype=torch.float64, device=self.device)
		out_host = torch.empty((total, ODE.n), dtype=torch.float64).pin_memory() if self.rank == 0 else None
		slot = gathered[begin:end]
		even = total % self.world == 0
		collective_events = []

		def step(timed):
			if pars_host is not None:
				ODE.set_parameters(pars_host)
			ODE.set_initial_value(Y0_host, 0.0)
			solver.integrate(wl["T"], into=slot if even else None)
			a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			a.record()
			if even:
				dist.all_gather_into_tensor(gathered, slot)
			else:
				gathered.copy_(D.gather_states(solver.y_device, total))
			b.record()
			if timed:
				collective_events.append((a, b))
			if out_host is not None:
				out_host.copy_(gathered, non_blocking=True)
			torch.cuda.current_stream().synchronize()

		for _ in range(warmup):
			step(False)
		self.barrier()
		solver.reset_step_counts()
		t0 = time.perf_counter()
		for _ in range(steps):
			step(True)
		self.barrier()
		seconds = time.perf_counter() - t0
		attempts = sum(solver.step_counts())
		collective_ms = sum(a.elapsed_time(b) for a, b in collective_events) / steps
		stats = torch.tensor([seconds, collective_ms], dtype=torch.float64, device=self.device)
		dist.all_reduce(stats, op=dist.ReduceOp.MAX)
		# the rank that arrives last at the collective measures the transfer alone; the others also wait for it
		fastest = torch.tensor([collective_ms], dtype=torch.float64, device=self.device)
		dist.all_reduce(fastest, op=dist.ReduceOp.MIN)
		counts = torch.tensor([attempts], dtype=torch.float64, device=self.device)
		dist.all_reduce(counts, op=dist.ReduceOp.SUM)
		seconds, collective_ms = stats.tolist()
		transfer_ms = fastest.item()
		if self.rank != 0:
			return None
		return {"value": counts.item() / seconds, "unit": UNIT, "scaling": "strong", "n_gpus": self.world, "steps": steps, "warmup": warmup,
				"ms_per_step": 1e3 * seconds / steps, "total_trajectories": total, "workload": wl["label"], "integrator": wl["integrator"],
				"timed_region": "per step: H2D of the rank's shard (pinned), integrate, all-gather on NCCL, D2H of the gathered ensemble on rank 0; wall clock, max over ranks",
				"collective": {"name": "ncclAllGather (torch.distributed.all_gather_into_tensor, in place: the integrate kernel's unpack writes into the rank's slot)" if even else "all_gather of padded shards",
						"doubles_per_rank": (end - begin) * ODE.n, "ms_per_step": transfer_ms, "share_of_step": transfer_ms / (1e3 * seconds / steps),
						"ms_per_step_including_the_wait_for_the_slowest_rank": collective_ms},
				"h2d_bytes_per_step_per_rank": int(Y0_host.numel() * 8 + (pars_host.numel() * 8 if pars_host is not None else 0)),
				"d2h_bytes_per_step_rank0": int(total * ODE.n * 8)}


def run_ours(args):
	bench = Bench(args)
	torch = bench.torch
	name = args.workload
	B = args.trajectories or DEFAULT_TRAJECTORIES[name]
	line = bench.record(name, args.steps, args.warmup, B, integrator=args.integrator)
	extras = name == HEADLINE and not args.only_headline and not args.integrator and not args.storage
	if extras and bench.world == 1:
		configs = {}
		for other in CONFIGS:
			torch.cuda.empty_cache()
			try:
				configs[other] = bench.record(other, args.config_steps, 3, DEFAULT_TRAJECTORIES[other], min_seconds=2.0)
			except Exception as error:      # one failing config must not cost the headline line
				configs[other] = {"error": f"{type(error).__name__}: {error}"}
		line["configs"] = configs
	if extras and bench.world > 1:
		strong = {}
		for other in STRONG:
			torch.cuda.empty_cache()
			try:
				record = bench.strong(other, args.config_steps, 3, 1 << 20)
			except Exception as error:
				record = {"error": f"{type(error).__name__}: {error}"}
			if bench.rank == 0:
				strong[other] = record
		if bench.rank == 0:
			line["strong"] = strong
	bench.close()
	return line


def keep_stdout_for_the_result():
	"""Libraries print to stdout now and then (NCCL announces its version there); the contract is ONE JSON line.
	Everything written to file descriptor 1 from here on goes to stderr; the returned writer is the real stdout."""
	sys.stdout.flush()
	real = os.fdopen(os.dup(1), "w")
	os.dup2(2, 1)
	sys.stdout = os.fdopen(os.dup(2), "w")
	return real


def main():
	parser = argparse.ArgumentParser()
	parser.add_argument("--gpus", type=int, default=1)
	parser.add_argument("--steps", type=int, default=5)
	parser.add_argument("--warmup", type=int, default=3)
	parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
	parser.add_argument("--workload", default=HEADLINE, choices=[HEADLINE, *CONFIGS])
	parser.add_argument("--trajectories", type=int, default=None, help="per GPU (default: the config's own size)")
	parser.add_argument("--integrator", default=None, choices=[None, "dopri5", "dop853", "RK45", "DOP853", "RK23"],
			help="instead of the workload's own (dopri5, as in the reference's examples); RK45 = the same pair under solve_ivp's controller")
	parser.add_argument("--storage", default=None, choices=[None, "registers", "global", "warp", "shared", "quad"])
	parser.add_argument("--kregs", type=int, default=None)
	parser.add_argument("--chunks", type=int, default=None, help="pieces of the host-resident ensemble in the end-to-end arm (default: one per 256 MB of state, at most 8)")
	parser.add_argument("--cpu-sample-per-core", type=int, default=None)
	parser.add_argument("--no-cpu-baseline", action="store_true")
	parser.add_argument("--only-headline", action="store_true", help="skip the `configs` (N = 1) / `strong` (N > 1) records")
	parser.add_argument("--config-steps", type=int, default=5, help="timed steps of each of the `configs` / `strong` records")
	args = parser.parse_args()
	args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
	result = keep_stdout_for_the_result()
	line = run_reference(args) if args.impl == "reference" else run_ours(args)
	if line is not None:
		result.write(json.dumps(line) + "\n")
		result.flush()


if __name__ == "__main__":
	main()
