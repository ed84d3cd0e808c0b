"""Development probe (run under gpurun): raw integer-pipe issue rates and field-op rates vs occupancy."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
res = {}
names = ["IMAD.WIDE", "LOP3", "IMAD", "IMAD.WIDE:LOP3 1:2", "SHF", "IADD"]
for mode in range(6):
    for bps in (1, 2, 4, 8):
        v = ctx.pipebench(mode, 2000, bps)
        res[f"{names[mode]}@{bps*8}warps"] = v
        print(f"{names[mode]:20s} {bps*8:3d} warps/SM: {v:.3e} thread-instr/s = {v/148/32:.3e} warp-instr/s/SM", flush=True)
for prim, nm, it in [(0, "gf_mul", 100), (1, "gf_sqr", 500), (2, "fr_mul", 200)]:
    for minb in (1, 2, 3):
        v = ctx.microbench(prim + 10 * (minb - 1), it)
        res[f"{nm}@minb{minb}"] = v
        print(f"{nm} minBlocks={minb}: {v:.3e} ops/s", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "pipes.json"), "w"), indent=1)
