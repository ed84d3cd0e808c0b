"""Development probe: Gantt rows (lane, kernel category, start, end) of one MSM with CUDA events around every launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = dvpari.Context(0)
n = 1 << lg
ctx.srs_random(0, n, 5)
d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
for kv in filter(None, os.environ.get("DVP_KNOBS", "").split(",")):  # e.g. DVP_KNOBS=msm_tables=0,msm_preplan=0
    ctx.set(kv.split("=")[0], int(kv.split("=")[1]))
ctx.set("msm_lanes", lanes); ctx.set("msm_profile", 1); ctx.set("timing", 1)
for _ in range(3):
    ctx.multi_scalar_mul_device(d, n, 0)
st = ctx.msm_stats()
print(f"# n=2^{lg} lanes={lanes} device {st['ms_device']:.3f} ms sort {st['ms_recode_sort']:.3f}")
for lane, cat, t0, t1 in sorted(ctx.msm_timeline(), key=lambda r: r[2]):
    print(f"{lane} {cat:12s} {t0:8.3f} {t1:8.3f}  {1e3*(t1-t0):8.1f} us")
