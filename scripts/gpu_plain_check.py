"""Plain layout (msm_tables = 0) at 2^lg points: preplan on/off, ld_tree_max default / 65536; and the timing-mode stage split."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
ctx.srs_random(0, n, 5)
d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
for tables in (1, 0):
    ctx.set("msm_tables", tables)
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    for pre in (1, 0):
        for ldm in (0, 65536):
            ctx.set("msm_preplan", pre); ctx.set("ld_tree_max", ldm)
            best = 1e9
            for rep in range(6):
                t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
            assert out == ref
            st = ctx.msm_stats()
            print(f"tables={tables} preplan={pre} ld_tree_max={ldm}: {best*1e3:.3f} ms launches={st['launches']} rounds={st['rounds_main']},{st['rounds_a']} c={st['window_bits']} W={st['windows']}", flush=True)
    ctx.set("msm_preplan", 1); ctx.set("ld_tree_max", 0)
    ctx.set("timing", 1); ctx.set("msm_lanes", 1)
    for rep in range(3):
        ctx.multi_scalar_mul_device(d, n, 0)
        st = ctx.msm_stats()
        print(f"tables={tables} timing: sort {st['ms_recode_sort']:.3f} acc {st['ms_accumulate']:.3f} reduce {st['ms_reduce']:.3f} tail {st['ms_tail']:.3f} device {st['ms_device']:.3f}", flush=True)
    ctx.set("timing", 0); ctx.set("msm_lanes", 0)
