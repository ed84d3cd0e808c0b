"""Persistent kernel (use_accumulate = 2) against the separate-launch path (the default above 3 * 2^20 points) at large sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [22, 23, 24]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    for rnd in range(2):
        for ua in (1, 2):
            ctx.set("use_accumulate", ua)
            best = 1e9
            for rep in range(4):
                t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
            assert out == ref
            st = ctx.msm_stats()
            print(f"n=2^{lg} use_accumulate={ua}: {best*1e3:.2f} ms {n/best:.3e} pts/s launches={st['launches']} rounds={st['rounds_main']},{st['rounds_a']} c={st['window_bits']}", flush=True)
    ctx.set("use_accumulate", 1)
    ctx.dev_free(d); ctx.srs_free(0)
