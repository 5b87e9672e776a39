"""would pre-sorting the ensemble help the lock-step rounds of the dense engine? (development aid)"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
n, B, T = 1000, 4096, 5.0
data = W.kuramoto_data(n)
Wm = data["c"] / (n - 1) * data["A"].T.astype(float); np.fill_diagonal(Wm, 0.0)
ODE = jitcode.from_sine_coupling(Wm, data["omega"])
ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-9)
Y0 = torch.as_tensor(W.kuramoto_ensemble(B, n)).cuda()
ODE.set_initial_value(Y0, 0.0)
ODE.integrator.reset_step_counts()
ODE.integrate(T)
integ = ODE.integrator
attempts = (integ.n_accepted[:B] + integ.n_rejected[:B]).double().cpu().numpy()
print("attempts: min", attempts.min(), "mean", attempts.mean(), "max", attempts.max(), "rounds", integ.rounds, "slot use", attempts.sum() / (integ.rounds * B))
f0 = ODE.f(0.0, Y0)
proxies = {
	"|f(y0)|": f0.norm(dim=1).cpu().numpy(),
	"max|f(y0)|": f0.abs().max(dim=1).values.cpu().numpy(),
	"order parameter": torch.complex(torch.cos(Y0), torch.sin(Y0)).mean(dim=1).abs().cpu().numpy(),
}
def tile_cost(order):
	a = attempts[order]
	pad = (-len(a)) % 64
	a = np.concatenate([a, np.zeros(pad)]).reshape(-1, 64)
	return (a.max(axis=1) * 64).sum()
print("useful/executed unsorted:", attempts.sum() / tile_cost(np.arange(B)), " ideal (sorted by the true cost):", attempts.sum() / tile_cost(np.argsort(attempts)))
from scipy.stats import spearmanr
for name, proxy in proxies.items():
	print(f"{name:18s} spearman {spearmanr(proxy, attempts)[0]:+.3f}  useful/executed if sorted by it: {attempts.sum() / tile_cost(np.argsort(proxy)):.3f}")
