"""Development probe (run under gpurun): integer-pipe microbenchmarks and an MSM stage breakdown."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import numpy as np

import dvpari
from oracle import oracle as O

ctx = dvpari.Context(0)
res = {}
for name, op, it in [("gf_mul", 0, 200), ("gf_sqr", 1, 1000), ("fr_mul", 2, 400)]:
    v = ctx.microbench(op, it)
    res[name + "_per_s"] = v
    print(f"{name}: {v:.3e} ops/s", flush=True)
G = O.generator()
sizes = [int(a) for a in sys.argv[1:]] or [12, 16, 18, 20]
for lg in sizes:
    n = 1 << lg
    pts = O.chain_points(n, O.pt_mul(G, 0x1234567), O.pt_mul(G, 0x7654321))
    t0 = time.time()
    ctx.srs_load(0, O.encode_batch(pts))
    t_load = time.time() - t0
    sc = dvpari.random_fr_mont(n, 77 + lg)
    d = ctx.dev_alloc(n * 32)
    ctx.dev_upload(d, sc)
    ctx.set("timing", 1)
    for cbits in ([0] if lg < 20 else [0, 13, 14, 16]):
        ctx.set("msm_window_bits", cbits)
        best = None
        for rep in range(3):
            t0 = time.time()
            out = ctx.multi_scalar_mul_device(d, n, 0)
            dt = time.time() - t0
            best = dt if best is None else min(best, dt)
        st = ctx.msm_stats()
        print(f"n=2^{lg} c={st['window_bits']} W={st['windows']} rounds={st['rounds_main']}/{st['rounds_a']}/{st['rounds_b']} "
              f"launches={st['launches']} wall={best*1e3:.2f} ms  sort={st['ms_recode_sort']:.2f} acc={st['ms_accumulate']:.2f} "
              f"red={st['ms_reduce']:.2f} tail={st['ms_tail']:.2f}  -> {n/best:.3e} pts/s  (srs_load {t_load:.2f}s)", flush=True)
        res[f"msm_2^{lg}_c{st['window_bits']}"] = {"wall_ms": best * 1e3, **st}
    ctx.dev_free(d)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
