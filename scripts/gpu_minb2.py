"""Development probe (run under gpurun): MSM time against the occupancy of the pass-2 kernel (separate-launch path)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [20, 22]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    sc = dvpari.random_fr_mont(n, 6)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, sc)
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    for acc in (1, 0):
        ctx.set("use_accumulate", acc)
        for minb in (2, 1, 3) if acc == 0 else (2,):
            ctx.set("pass2_minb", minb)
            best = 1e9
            for rep in range(4):
                t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
            assert out == ref
            st = ctx.msm_stats()
            print(f"n=2^{lg} use_accumulate={acc} pass2_minb={minb}: {best*1e3:.2f} ms {n/best:.3e} pts/s launches={st['launches']}", flush=True)
    ctx.dev_free(d)
