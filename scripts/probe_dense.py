"""throughput of the dense Kuramoto engine (development aid)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import workloads as W
from jitcode_b200 import jitcode
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
data = W.kuramoto_data(n)
Wm = data["c"] / (n - 1) * data["A"].T.astype(float); np.fill_diagonal(Wm, 0.0)
ODE = jitcode.from_sine_coupling(Wm, data["omega"])
ODE.set_integrator("dopri5", rtol=1e-6, atol=1e-9)
Y0 = torch.as_tensor(W.kuramoto_ensemble(B, n)).cuda()
for rep in range(2):
	ODE.set_initial_value(Y0, 0.0)
	ODE.integrator.reset_step_counts()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record(); ODE.integrate(T); e1.record(); torch.cuda.synchronize()
	ms = e0.elapsed_time(e1)
	acc, rej = ODE.step_counts(); rounds = ODE.integrator.rounds
	flops = rounds * 7 * 2.0 * n * n * 2 * ODE.integrator.ld      # what the GEMMs executed
	useful = (acc + rej) * 6 * 4.0 * n * n
	print(f"n={n} B={B} T={T} {ms:.1f} ms rounds={rounds} attempts={acc+rej} (slot use {(acc+rej)/(rounds*B):.2f}) {(acc+rej)/ms*1e3:.3e} steps/s  GEMM executed {flops/ms/1e9:.1f} TFLOP/s, useful {useful/ms/1e9:.1f} TFLOP/s")
