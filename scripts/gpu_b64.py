"""Development probe: MSM time with / without 64 additions per thread in the large rounds."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [20, 22, 24]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    for b64 in (0, 1 << 23, 1 << 22, 1 << 21):
        ctx.set("b64_min", b64)
        best = 1e9
        for rep in range(4):
            t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
        assert out == ref
        print(f"n=2^{lg} b64_min={b64}: {best*1e3:.2f} ms {n/best:.3e} pts/s", flush=True)
    ctx.set("b64_min", 0)
    ctx.dev_free(d); ctx.srs_free(0)
