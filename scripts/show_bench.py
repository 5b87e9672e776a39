"""prints the records of a bench.py line (development aid): python scripts/show_bench.py file.json"""
import json, sys
line = json.loads(open(sys.argv[1]).read())
def show(name, r):
	if "error" in r:
		print(f"== {name}: ERROR {r['error']}"); return
	ro = r["roofline"]
	cpu = r.get("cpu_baseline", {})
	print(f"== {name}: value {r['value']:.4g}  e2e {r['e2e']['value']:.4g}  ms/step {r['ms_per_step']:.1f} (kernel {ro['kernel_ms']:.1f}, e2e {r['e2e']['ms_per_step']:.1f})  launches {r['gpu_launches']}  storage {r['config']['storage']}")
	print(f"   roofline {ro['bound']}: {ro['achieved']:.3g} / {ro['peak']:.3g} {ro['unit']} = {ro['frac']:.3f}  traffic {ro['traffic']}")
	p = r.get("parity", {})
	print(f"   parity ok={p.get('ok')} max_ratio={p.get('max_ratio')} | {p.get('compared', p.get('error'))}")
	if cpu:
		print(f"   cpu {cpu['value']:.4g} on {cpu['cores']} cores ({cpu['kind']}) -> e2e ratio {r['e2e']['value']/cpu['value']:.1f}x, device {r['value']/cpu['value']:.1f}x")
	if "cpu_baseline_RK45" in r:
		print(f"   cpu RK45 {r['cpu_baseline_RK45']['value']:.4g}")
	print(f"   clocks {r['clocks']}")
show("headline", line)
for k, v in line.get("configs", {}).items():
	show(k, v)
for k, v in line.get("strong", {}).items():
	print(f"== strong {k}:", json.dumps(v)[:600])
