"""A/B of compile-time variants of one workload's integrate kernel (development aid):
    python scripts/ab_variants.py roessler 262144 -- "" "-DJB_NO_STIFFNESS" …
    python scripts/ab_variants.py fhn_lyap 100000 threads=64,min_blocks=4 threads=128,min_blocks=3 …
every argument after the batch size is one variant: extra nvcc flags, or key=value pairs handed to set_integrator.
Switches the kernels understand: -DJB_NO_STIFFNESS (Hairer's stiffness test compiled out), -DJB_LITERAL_COEF (Runge–Kutta
coefficients as 64-bit literals instead of constant-bank operands), -DJB_REFILL_LANES=k (lanes that change trajectories together).
Switches of the generators (environment): JITCODE_B200_NO_HELPER_LANES (idle lanes of a blocked ring do not help with the
gathers), JITCODE_B200_NO_SPLIT_REDUCTIONS (the mean field's warp sum in one piece), JITCODE_B200_INLINE_LITERALS (numbers of
the right-hand side as literals instead of constant-bank operands)."""
import sys
import time
import torch
sys.path.insert(0, ".")
import bench

name, B = sys.argv[1], int(sys.argv[2])
variants = [v for v in sys.argv[3:] if v != "--"] or [""]
wl = bench.workload(name, B)
for variant in variants:
	flags, options = [], {}
	for token in variant.replace(",", " ").split():
		if "=" in token and not token.startswith("-"):
			key, value = token.split("=")
			options[key] = int(value) if value.lstrip("-").isdigit() else value
		else:
			flags.append(token)
	from jitcode_b200 import jitcode, jitcode_lyap
	system = wl["system"]
	if wl["kind"] == "lyap":
		ODE = jitcode_lyap(system["f_sym"], n_lyap=system["lyap"], verbose=False, seed=1)
	else:
		ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)
	ODE._compile_options["extra_compile_args"] = tuple(flags)
	try:
		ODE.set_integrator(wl["integrator"], rtol=wl["rtol"], atol=wl["atol"], **options)
	except Exception as error:
		print(f"{variant!r}: {type(error).__name__}: {str(error)[:300]}")
		continue
	Y0 = torch.as_tensor(wl["Y0"]).cuda()
	pars = torch.as_tensor(wl["pars"]).cuda() if wl["pars"].shape[1] else None
	best = None
	for rep in range(4):
		if pars is not None:
			ODE.set_parameters(pars)
		ODE.set_initial_value(Y0, 0.0)
		ODE.integrator.reset_step_counts()
		ODE.integrator.kernel_events = []
		if wl["kind"] == "lyap":
			ODE.integrate_many(wl["times"], record_states=False)
		else:
			ODE.integrate(wl["T"])
		torch.cuda.synchronize()
		ms = sum(a.elapsed_time(b) for a, b in ODE.integrator.kernel_events)
		best = ms if best is None or ms < best else best
	attempts = sum(ODE.step_counts())
	info = ODE.integrator.module.info
	print(f"{variant!r:50s} storage {info.storage} threads {info.threads_per_block}: best kernel {best:8.3f} ms  {attempts / best * 1e3:.4e} steps/s", flush=True)
