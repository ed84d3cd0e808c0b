"""Development probe: wall time of the MSM against the number of lanes (staggered by stream priority)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [16, 18, 20, 22]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    sc = dvpari.random_fr_mont(n, 6)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, sc)
    ref = None
    for lanes in (1, 2, 3, 4):
        ctx.set("msm_lanes", lanes)
        best = 1e9
        for rep in range(8):
            t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
        ref = ref or out
        assert out == ref
        st = ctx.msm_stats()
        print(f"n=2^{lg} lanes={lanes} c={st['window_bits']} W={st['windows']}: {best*1e3:.3f} ms  {n/best:.3e} pts/s  launches={st['launches']}", flush=True)
    ctx.set("msm_lanes", 0)
    ctx.dev_free(d)
