"""how much of a warp idles because its 32 trajectories need different numbers of step attempts (development aid)."""
import sys
import torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
B = 1 << 20
system = W.lotka_volterra()
Y0, _ = W.lotka_volterra_ensemble(B)
ODE = jitcode(system["f_sym"], n=2, verbose=False)
ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-12)
ODE.set_initial_value(torch.as_tensor(Y0).cuda(), 0.0)
ODE.integrate(100.0)
it = ODE.integrator
attempts = (it.n_accepted[:B] + it.n_rejected[:B]).double().view(-1, 32)
print("mean attempts", attempts.mean().item(), "mean of per-warp max", attempts.max(dim=1).values.mean().item(), "lane efficiency", (attempts.mean() / attempts.max(dim=1).values.mean()).item())
order = torch.argsort(attempts.view(-1))
sorted_attempts = attempts.view(-1)[order].view(-1, 32)
print("if sorted by cost: efficiency", (sorted_attempts.mean() / sorted_attempts.max(dim=1).values.mean()).item())
