"""Prints the headline numbers and the per-kernel table of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.0f}  launches {d['gpu_launches']}  n_gpus {d['n_gpus']}")
r = d['roofline']
print(f"dominant {r['kernel']}: {r['achieved']:.0f} GB/s = {r['frac']:.3f} of peak, avg {r['avg_ms']:.3f} ms, share {r['share_of_step']:.3f}")
steps = d['steps']
for k, v in r['kernels'].items():
    gbs = v['GB/s']
    print(f"  {k:18s} {v['ms_total']/steps:8.3f} ms/step  {v['launches']/steps:5.1f} launches/step  {'' if not gbs else f'{gbs:7.0f} GB/s'}")
print("vcycles", (d.get('details') or d['config']).get('vcycles_per_solve'))
