"""Development probe (run under gpurun): separate-launch path with the fused per-round kernel against pass1 / binv / pass2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [22]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    sc = dvpari.random_fr_mont(n, 6)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, sc)
    ctx.set("fused_rounds", 0)
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    for acc in (1, 0):
        ctx.set("use_accumulate", acc)
        for fused in (0, 1):
            ctx.set("fused_rounds", fused)
            for lanes in (0, 1, 2, 3):
                ctx.set("msm_lanes", lanes)
                best = 1e9
                for rep in range(4):
                    t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
                assert out == ref
                st = ctx.msm_stats()
                print(f"n=2^{lg} use_accumulate={acc} fused={fused} lanes={lanes}: {best*1e3:.2f} ms {n/best:.3e} pts/s launches={st['launches']}", flush=True)
    ctx.dev_free(d)
