import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [16, 18, 20, 22]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    for mx in (3000, 5000, 7000, 12000, 20000, 40000, 70000, 140000, 300000):
        ctx.set("ld_tree_max", mx)
        best = 1e9
        for rep in range(8):
            t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
        assert out == ref
        st = ctx.msm_stats()
        print(f"n=2^{lg} ld_tree_max={mx}: {best*1e3:.3f} ms {n/best:.3e} pts/s rounds={st['rounds_main']},{st['rounds_a']}", flush=True)
    ctx.dev_free(d); ctx.srs_free(0)
