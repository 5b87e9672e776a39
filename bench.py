#!/usr/bin/env python
"""
bench.py — ensemble Dormand–Prince RK45 trajectory-steps/s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload roessler|lv] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N … bench.py --gpus N …

A *step* of the benchmark is one pass of the hot path over one batch: (re)initialise the ensemble and
integrate every trajectory from t = 0 to the workload's end time with one `integrate` call.  A
*trajectory-step* is one Runge–Kutta step attempt (accepted or rejected) of one trajectory; they are
counted on the device (per-trajectory counters of the kernel).

Default workload (the one the north star's ≥100× target is quoted on, BASELINE.json configs[2]):
small-world network of 100 Rössler oscillators (n = 300), 1 Mi trajectories per GPU with per-trajectory
coupling strength and initial state, dopri5 semantics (Dormand–Prince 5(4), Hairer's controller — what
examples/SW_of_Roesslers.py:107 uses), rtol 1e-6, atol 1e-12, t = 0 → 10.  `--workload lv` runs configs[1]
(1 Mi Lotka–Volterra initial conditions, t = 0 → 100).

Output: ONE JSON line on rank 0 (keys as the driver's contract demands).
  value        device-timed (CUDA events) with the ensemble already resident in HBM
  e2e          through jitcode.set_initial_value / integrate with pinned HOST buffers, copies inside the timed region
  roofline     the integrate kernel against the HBM roofline of MEASURED_PEAKS.json (algorithmic bytes 280·n per
               trajectory-step)
  cpu_baseline the reference CPU path (compiled C right-hand side under scipy.integrate.ode, one process per core)
               on a bounded sample of the same ensemble
`--impl reference` times only that CPU path.
"""

import argparse
import json
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


def workload(name, B, seed_offset=0):
	if name == "roessler":
		system = W.roessler_network(100)
		Y0, pars = W.roessler_ensemble(B, seed=43 + seed_offset)
		return {"system": system, "Y0": Y0, "pars": pars, "T": 10.0, "integrator": "dopri5",
				"label": "SW_of_Roesslers N=100 (n=300), per-trajectory k and Y0"}
	if name == "lv":
		system = W.lotka_volterra()
		Y0, pars = W.lotka_volterra_ensemble(B, seed=42 + seed_offset)
		return {"system": system, "Y0": Y0, "pars": pars, "T": 100.0, "integrator": "dopri5",
				"label": "Lotka-Volterra ensemble (n=2), random initial conditions"}
	if name == "fhn_lyap":
		# config 4: FitzHugh–Nagumo pair with all 4 Lyapunov exponents (n = 20), re-orthonormalised every Δt = 10
		system = dict(W.fhn(), lyap=4)
		return {"system": system, "Y0": W.fhn_lyap_ensemble(B, seed=42 + seed_offset), "pars": np.empty((B, 0)), "T": 200.0,
				"times": [10.0 * (i + 1) for i in range(20)], "integrator": "dopri5",
				"label": "double_fhn_lyapunov: FHN pair + 4 tangent vectors (n=20), QR every dt=10, 20 calls"}
	if name == "kuramoto":
		# config 5: 1000 oscillators, dense adjacency; the symbolic form has 2·10^5 sine terms (SymPy needs minutes to
		# build it), so the system enters through its data — same kernels as the recognised symbolic form
		data = W.kuramoto_data(1000)
		n = data["n"]
		Wm = data["c"] / (n - 1) * data["A"].T.astype(float)
		np.fill_diagonal(Wm, 0.0)
		return {"system": {"n": n, "sine_coupling": (Wm, data["omega"]), "control_pars": []}, "Y0": W.kuramoto_ensemble(B, n, seed=44 + seed_offset),
				"pars": np.empty((B, 0)), "T": 5.0, "integrator": "dopri5", "label": "kuramoto_network n=1000 dense adjacency"}
	raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------ the reference CPU path (oracle)

def cpu_reference_handle(system, label):
	"""oracle/_ref (the reference's own jitced_template.c, compiled in the build container) when it is there,
	else the oracle's port of it."""
	from oracle import build_ref, c_rhs
	handle = None
	if system.get("lyap"):
		full, helpers = c_rhs.lyapunov_system(system["f_sym"], system["lyap"], system["helpers"], system["n"])
		system = {"f_sym": full, "helpers": helpers, "control_pars": [], "n": len(full)}
	try:
		handle = build_ref.build_ref_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"])
		if handle is not None:
			handle.f  # noqa: B018  (import check: the file was built for this Python)
			return handle, "reference"
	except Exception:
		handle = None
	return c_rhs.build_rhs(system["f_sym"], system["helpers"], system["control_pars"], system["n"]), "port"


def cpu_run(wl, handle, n_sample):
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
	result = ensemble.run_ensemble(handle, wl["integrator"], Y0, wl.get("times", [wl["T"]]), pars=pars, lyap=lyap, rtol=RTOL, atol=ATOL)
	return result["attempts"], result["seconds"], result["cores"]


def cpu_baseline(wl, per_core):
	handle, kind = cpu_reference_handle(wl["system"], wl["label"])
	from oracle import ensemble
	cores = ensemble.available_cores()
	n_sample = min(len(wl["Y0"]), per_core * cores)
	cpu_run(wl, handle, min(n_sample, 2 * cores))     # warm the pool and the page cache
	attempts, seconds, cores = cpu_run(wl, handle, n_sample)
	return {"value": attempts / seconds, "unit": UNIT, "cores": cores, "kind": kind,
			"sample": f"{n_sample} trajectories of the same ensemble, {wl['integrator']} via {'scipy.integrate.ode' if wl['integrator'] in ('dopri5', 'dop853') else 'scipy.integrate.solve_ivp steppers'}, t=0..{wl['T']:g}, {seconds:.1f} s"}


def run_reference(args):
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return None
	if args.workload == "kuramoto":
		return {"impl": "reference", "unavailable": "the reference's written-out form of the dense 1000-oscillator network (2e5 sine terms of generated C) is not built by the default run"}
	per_core = args.cpu_sample_per_core
	from oracle import ensemble
	cores = ensemble.available_cores()
	wl = workload(args.workload, per_core * cores)
	if args.integrator:
		wl["integrator"] = args.integrator
	handle, kind = cpu_reference_handle(wl["system"], wl["label"])
	for _ in range(args.warmup):
		cpu_run(wl, handle, min(len(wl["Y0"]), 4 * cores))
	attempts = seconds = 0.0
	for _ in range(args.steps):
		a, s, cores = cpu_run(wl, handle, len(wl["Y0"]))
		attempts += a
		seconds += s
	value = attempts / seconds
	line = {
		"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
		"warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps, "higher_is_better": True, "scaling": "weak",
		"vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": wl["label"], "integrator": wl["integrator"], "rtol": RTOL, "atol": ATOL, "t_end": wl["T"],
				"trajectories_per_step": len(wl["Y0"])},
		"cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
				"sample": f"{len(wl['Y0'])} trajectories per step, one process per core"},
		"e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}
	return line


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
				["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={CLOCK_QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
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

def run_ours(args):
	import torch
	import torch.distributed as dist
	from jitcode_b200 import jitcode

	world = int(os.environ.get("WORLD_SIZE", "1"))
	rank = int(os.environ.get("RANK", "0"))
	local_rank = int(os.environ.get("LOCAL_RANK", "0"))
	if not torch.cuda.is_available():
		raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the CPU arm).")
	torch.cuda.set_device(local_rank)
	device = torch.device("cuda", local_rank)
	if world > 1:
		dist.init_process_group("nccl", device_id=device)

	B = args.trajectories
	wl = workload(args.workload, B, seed_offset=1000 * rank)      # every rank integrates its own shard
	if args.integrator:
		wl["integrator"] = args.integrator
	system = wl["system"]
	n = system["n"]
	dense = "sine_coupling" in system
	lyap = system.get("lyap")
	if lyap:
		from jitcode_b200 import jitcode_lyap
		ODE = jitcode_lyap(system["f_sym"], n_lyap=lyap, verbose=False, seed=1)
		ODE.set_integrator(wl["integrator"], rtol=RTOL, atol=ATOL)
		n = ODE.n
	elif dense:
		ODE = jitcode.from_sine_coupling(*system["sine_coupling"])
		ODE.set_integrator(wl["integrator"], rtol=RTOL, atol=1e-9)
	else:
		ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=n, control_pars=system["control_pars"], verbose=False)
		ODE.set_integrator(wl["integrator"], rtol=RTOL, atol=ATOL, storage=args.storage, kregs=args.kregs)
	integrator = ODE.integrator
	module = integrator.module

	Y0_host = torch.as_tensor(wl["Y0"]).pin_memory()
	pars_host = torch.as_tensor(wl["pars"]).pin_memory() if wl["pars"].shape[1] else None
	out_host = torch.empty_like(Y0_host).pin_memory()
	Y0_dev = Y0_host.to(device)
	pars_dev = pars_host.to(device) if pars_host is not None else None

	def barrier():
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	def device_step():
		if pars_dev is not None:
			ODE.set_parameters(pars_dev)
		ODE.set_initial_value(Y0_dev, 0.0)
		if lyap:
			return ODE.integrate_many(wl["times"], record_states=False)       # the 20 calls of the example's loop as one launch
		return ODE.integrate(wl["T"])

	def host_step():
		if lyap:
			ODE.set_initial_value(Y0_host, 0.0)
			states, exponents = ODE.integrate_many(wl["times"], record_states=False)     # the result of such a run: the exponents
			return exponents
		if pars_host is not None:
			ODE.set_parameters(pars_host)
		if hasattr(integrator, "integrate_streamed") and B >= 65536:
			# the public call for host-resident ensembles: uploads and downloads overlap the kernels chunk by chunk
			return ODE.integrate_streamed(Y0_host, wl["T"], out_host, 0.0, chunks=args.chunks)
		ODE.set_initial_value(Y0_host, 0.0)
		return ODE.integrate(wl["T"], out=out_host)

	# ---- device-resident arm
	for _ in range(args.warmup):
		device_step()
	barrier()
	integrator.reset_step_counts()
	integrator.kernel_events = []
	launches_before = module.launch_count()
	start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	with ClockSampler(local_rank) as clocks:
		barrier()
		start.record()
		for _ in range(args.steps):
			device_step()
		end.record()
		barrier()
	seconds = start.elapsed_time(end) * 1e-3
	launches = module.launch_count() - launches_before
	accepted, rejected = integrator.step_counts()
	attempts = accepted + rejected
	kernel_ms = [a.elapsed_time(b) for a, b in integrator.kernel_events]
	integrator.kernel_events = None

	# ---- end-to-end arm: pinned host buffers in and out through the public API
	for _ in range(min(args.warmup, 2)):
		host_step()
	barrier()
	integrator.reset_step_counts()
	t0 = time.perf_counter()
	for _ in range(args.steps):
		host_step()
	barrier()
	e2e_seconds = time.perf_counter() - t0
	e2e_attempts = sum(integrator.step_counts())

	if world > 1:
		stats = torch.tensor([seconds, e2e_seconds], dtype=torch.float64, device=device)
		dist.all_reduce(stats, op=dist.ReduceOp.MAX)
		seconds, e2e_seconds = stats.tolist()
		counts = torch.tensor([attempts, e2e_attempts, accepted, launches], dtype=torch.float64, device=device)
		dist.all_reduce(counts, op=dist.ReduceOp.SUM)
		attempts, e2e_attempts, accepted, launches = counts.tolist()

	if rank == 0:
		peaks_path = ROOT / "MEASURED_PEAKS.json"
		if peaks_path.exists():
			peak, peak_source = json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (burst copy)"
		else:
			peak, peak_source = 6650.0, "fallback of B200_PROFILING.md"
		own_attempts = (attempts / world) / args.steps      # per launch
		bytes_per_launch = 280.0 * n * own_attempts           # 35·n·8 B per trajectory-step
		kernel_s = statistics.mean(kernel_ms) * 1e-3
		achieved = bytes_per_launch / kernel_s / 1e9
		storage_name = {0: "registers", 1: "global", 2: "warp", 3: "shared", 4: "dense"}.get(module.info.storage)
		traffic_file = ROOT / "profiles" / "traffic.json"
		traffic = None
		if traffic_file.exists():
			# dram__bytes_read.sum + dram__bytes_write.sum per trajectory-step from the committed ncu --set full capture
			per_step = json.loads(traffic_file.read_text()).get(f"{args.workload}:{storage_name}")
			traffic = None if per_step is None else per_step * own_attempts
		# the state never leaves the chip between samples in register / warp storage: FP64 is the second yardstick
		import sympy
		from jitcode_b200 import _peaks
		rhs_flops = float(sum(sympy.count_ops(entry) for entry in ODE._f_list)) if not dense else 4.0 * n * n
		if lyap:
			line_extra = {"d2h_bytes_per_step": int(B * (system["n"] + len(wl["times"]) * lyap) * 8)}
		flops_per_step = 72.0 * n + 6.0 * rhs_flops
		fp64_peak = _peaks.fp64_fma_tflops(device)
		fp64_achieved = flops_per_step * own_attempts / kernel_s / 1e12
		line = {
			"metric": METRIC, "value": attempts / seconds, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
			"ms_per_step": 1e3 * seconds / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
			"dtype": "f64", "data": "synthetic",
			"config": {"workload": wl["label"], "integrator": wl["integrator"], "rtol": RTOL, "atol": ATOL, "t_end": wl["T"],
					"trajectories_per_gpu": B, "n": n, "storage": storage_name, "threads_per_block": module.info.threads_per_block,
					"parallelism": f"trajectory shards x{world}, no data-path collective",
					"l2": f"state per GPU {B*n*8/1e6:.0f} MB vs 126 MB L2 (no flush needed)" if B*n*8 > 2.5e8 else "ensemble re-packed from a fresh copy every step",
					"accepted_fraction": accepted / max(1.0, attempts)},
			"e2e": {"value": e2e_attempts / e2e_seconds, "unit": UNIT,
					"h2d_bytes_per_step": int(Y0_host.numel() * 8 + (pars_host.numel() * 8 if pars_host is not None else 0)),
					"d2h_bytes_per_step": int(out_host.numel() * 8), "ms_per_step": 1e3 * e2e_seconds / args.steps},
			"gpu_launches": int(launches),
			"clocks": clocks.summary(),
			"roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
					"kernel": "jb_integrate_kernel", "kernel_ms": statistics.mean(kernel_ms), "peak_source": peak_source,
					"algorithmic_bytes_per_trajectory_step": 280 * n,
					"fp64": {"achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
							"flops_per_trajectory_step": flops_per_step, "peak_source": "DFMA microbenchmark measured in this run (jitcode_b200/_peaks.py)"}},
		}
		if lyap:
			line["e2e"]["d2h_bytes_per_step"] = line_extra["d2h_bytes_per_step"]
		if dense:
			# the contraction is the kernel: FP64 tensor cores against the measured FP64 peak
			line["roofline"].update({"bound": "tensor", "achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
					"peak_source": "FP64 FMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)", "kernel": "dgemm_dmma_kernel + element-wise stages (jb_dense_integrate)"})
			line["config"]["lock_step_rounds_per_step"] = integrator.rounds / max(1, args.steps)
		if world == 1 and not args.no_cpu_baseline and not dense:
			line["cpu_baseline"] = cpu_baseline(wl, args.cpu_sample_per_core)
	else:
		line = None
	if world > 1:
		dist.destroy_process_group()
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
	parser.add_argument("--workload", default="roessler", choices=["roessler", "lv", "fhn_lyap", "kuramoto"])
	parser.add_argument("--trajectories", type=int, default=1 << 20, help="per GPU")
	parser.add_argument("--integrator", default=None, choices=[None, "dopri5", "dop853", "RK45", "DOP853", "RK23"],
			help="instead of the workload's own (dopri5, as in the reference's examples); RK45 = the same pair under solve_ivp's controller")
	parser.add_argument("--storage", default=None, choices=[None, "registers", "global", "warp"])
	parser.add_argument("--kregs", type=int, default=None)
	parser.add_argument("--chunks", type=int, default=None, help="pieces of the host-resident ensemble in the end-to-end arm (default: one per 256 MB of state, at most 8)")
	parser.add_argument("--cpu-sample-per-core", type=int, default=None)
	parser.add_argument("--no-cpu-baseline", action="store_true")
	args = parser.parse_args()
	if args.workload == "kuramoto" and args.trajectories == 1 << 20:
		args.trajectories = 4096
	if args.workload == "fhn_lyap" and args.trajectories == 1 << 20:
		args.trajectories = 100_000
	args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
	if args.cpu_sample_per_core is None:
		args.cpu_sample_per_core = {"roessler": 2048, "lv": 16384, "fhn_lyap": 4096}.get(args.workload, 256)
		if args.integrator in ("RK45", "DOP853", "RK23"):
			args.cpu_sample_per_core = max(64, args.cpu_sample_per_core // 8)     # SciPy's solve_ivp steppers are Python: ≈ 10× slower per step
	result = keep_stdout_for_the_result()
	line = run_reference(args) if args.impl == "reference" else run_ours(args)
	if line is not None:
		result.write(json.dumps(line) + "\n")
		result.flush()


if __name__ == "__main__":
	main()
