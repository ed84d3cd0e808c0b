"""Development probe (run under gpurun): per-category kernel time of one MSM on the separate-launch path, by pass-2 occupancy."""
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
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    ctx.set("use_accumulate", 0)
    for minb in (2, 1):
        ctx.set("pass2_minb", minb)
        for lanes in (2, 1):
            ctx.set("msm_lanes", lanes)
            best = 1e9
            for rep in range(3):
                t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
            assert out == ref
            print(f"n=2^{lg} minb={minb} lanes={lanes}: {best*1e3:.2f} ms", flush=True)
        ctx.set("msm_lanes", 0)
        ctx.set("msm_profile", 1); ctx.set("timing", 1)
        for rep in range(2):
            out = ctx.multi_scalar_mul_device(d, n, 0)
        assert out == ref
        pr = ctx.msm_profile(); st = ctx.msm_stats()
        ctx.set("msm_profile", 0); ctx.set("timing", 0)
        tot = sum(v[0] for v in pr.values())
        print(f"  minb={minb} profile (1 lane): " + "  ".join(f"{k} {v[0]:.3f}ms/{v[1]}" for k, v in pr.items()) + f"  sum {tot:.3f} ms")
        print(f"  stages: sort {st['ms_recode_sort']:.3f} accumulate {st['ms_accumulate']:.3f} reduce {st['ms_reduce']:.3f} tail {st['ms_tail']:.3f}; pass2 round0 {st['ms_pass2_round0']:.3f} ms for {st['adds_round0']} adds")
    ctx.dev_free(d)
