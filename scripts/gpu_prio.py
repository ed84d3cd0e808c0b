import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [18, 20, 22]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    for prio in (0, 1):
        for bmax in (64, 16, 4):
            for lanes in (2, 3, 4):
                ctx.set("prio_split", prio); ctx.set("pass_b_max", bmax); ctx.set("msm_lanes", lanes)
                best = 1e9
                for rep in range(5):
                    t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
                assert out == ref
                print(f"n=2^{lg} prio={prio} bmax={bmax} lanes={lanes}: {best*1e3:.2f} ms {n/best:.3e} pts/s", flush=True)
    ctx.dev_free(d); ctx.srs_free(0)
