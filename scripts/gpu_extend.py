"""Development probe (run under gpurun): ECFFT extend of 3 polynomials, ms and IMAD.WIDE issue fraction."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [20, 22]:
    n = 1 << lg
    dom = dvpari.Domain(ctx, lg + 1)
    d = ctx.dev_alloc(3 * n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(3 * n, 5))
    dom.extend_device(d, 3)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); dom.extend_device(d, 3); best = min(best, time.perf_counter() - t0)
    wides = 3 * (n // 2) * 2 * lg * 339
    print(f"extend 3 x 2^{lg}: {best*1e3:.3f} ms  {wides/best:.3e} IMAD.WIDE/s = {wides/best/9.2e12:.3f} of peak")
    ctx.dev_free(d); dom.close()
