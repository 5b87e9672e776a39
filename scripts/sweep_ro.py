"""sweep of warp-storage configurations on the Rössler network (development aid)."""
import sys, itertools
import torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = 10.0
N = next((int(a[2:]) for a in sys.argv[2:] if a.startswith("N=")), 100)
system = W.roessler_network(N)
Y0, pars = W.roessler_ensemble(B, N=N)
Yd, Pd = (torch.as_tensor(Y0).cuda(), torch.as_tensor(pars).cuda()) if torch.cuda.is_available() else (None, None)
configs = [("global", None, None)] + [("warp", k, th) for k, th in [(0, None), (2, None), (4, None), (5, None), (6, None), (7, None), (7, 192), (7, 128), (6, 256), (5, 256), (4, 256), (0, 256)]]
if len(sys.argv) > 2:
	# kregs:threads[:team_warps]  (0 = default); "N=…" changes the network size
	configs = []
	for a in sys.argv[2:]:
		if a.startswith("N=") or a.startswith("blocked=") or a.startswith("fuse="):
			continue
		parts = [int(x) for x in a.split(":")]
		configs.append(("warp", parts[0] if parts[0] >= 0 else None, parts[1] or None, (parts[2] or None) if len(parts) > 2 else None))
configs = [c if len(c) == 4 else (*c, None) for c in configs]
ref = None
blocked = next((bool(int(a[8:])) for a in sys.argv[2:] if a.startswith("blocked=")), True)
fuse = next((bool(int(a[5:])) for a in sys.argv[2:] if a.startswith("fuse=")), True)
compile_only = not torch.cuda.is_available()
for storage, kregs, threads, team in configs:
	ODE = jitcode(system["f_sym"], helpers=system["helpers"], n=system["n"], control_pars=system["control_pars"], verbose=False)
	ODE.blocked_rings, ODE.fuse_groups = blocked, fuse
	if compile_only:
		ODE.compile_cuda(storage=storage, kregs=kregs, threads=threads, team_warps=team); continue
	try:
		ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12, storage=storage, kregs=kregs, threads=threads, team_warps=team)
	except Exception as error:
		print(storage, kregs, threads, "FAILED", str(error)[-300:]); continue
	ODE.set_parameters(Pd)
	best = 1e9
	try:
		for rep in range(3):
			ODE.set_initial_value(Yd, 0.0)
			ODE.integrator.reset_step_counts()
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			e0.record(); out = ODE.integrate(T); e1.record(); torch.cuda.synchronize()
			best = min(best, e0.elapsed_time(e1))
	except Exception as error:
		print(storage, kregs, threads, "FAILED", str(error)[-200:]); continue
	acc, rej = ODE.step_counts()
	if ref is None: ref = out
	info = ODE.integrator.module.info
	print(f"blocked={int(blocked)} fuse={int(fuse)} {storage:6s} team={team} kregs={kregs} threads={info.threads_per_block:4d} smem={info.shared_bytes:6d} {best:8.2f} ms {(acc+rej)/best*1e3:.3e} steps/s  maxdiff {float((out-ref).abs().max()):.2e}", flush=True)
