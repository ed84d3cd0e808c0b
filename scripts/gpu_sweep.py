"""Development probe: knob sweep of the MSM at one size: python scripts/gpu_sweep.py lg knob=v1,v2 knob2=..."""
import itertools, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
lg = int(sys.argv[1])
knobs = [(a.split("=")[0], [int(v) for v in a.split("=")[1].split(",")]) for a in sys.argv[2:]]
n = 1 << lg
ctx.srs_random(0, n, 5)
d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
ref = ctx.multi_scalar_mul_device(d, n, 0)
for combo in itertools.product(*[v for _, v in knobs]):
    for (name, _), v in zip(knobs, combo):
        ctx.set(name, v)
    best = 1e9
    for rep in range(8):
        t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
    assert out == ref
    st = ctx.msm_stats()
    print(f"n=2^{lg} " + " ".join(f"{k}={v}" for (k, _), v in zip(knobs, combo)) + f": {best*1e3:.3f} ms {n/best:.3e} pts/s launches={st['launches']} c={st['window_bits']} W={st['windows']} rounds={st['rounds_main']},{st['rounds_a']},{st['rounds_b']}", flush=True)
