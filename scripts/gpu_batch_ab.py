"""A/B of the batched MSM call on one GPU: one call per MSM, batch without / with the sort of the next MSM running ahead.
Usage (on a GPU box): python scripts/gpu_batch_ab.py [lg ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari  # noqa: E402
import torch  # noqa: E402


def main():
    lgs = [int(a) for a in sys.argv[1:]] or [20]
    ctx = dvpari.Context(0)
    batch = 8
    for lg in lgs:
        n = 1 << lg
        ctx.srs_random(0, n, 0xD5A10002)
        pinned = torch.empty((batch, n, 4), dtype=torch.int64).pin_memory()
        sc = pinned.numpy().view(np.uint64)
        d = []
        for b in range(batch):
            sc[b] = dvpari.random_fr_mont(n, 77 + b)
            d.append(ctx.dev_alloc(n * 32))
            ctx.dev_upload(d[b], sc[b])
        host = [sc[b] for b in range(batch)]
        ref = [ctx.multi_scalar_mul_device(d[b], n, 0) for b in range(batch)]

        def t(fn, reps=8):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = fn()
            torch.cuda.synchronize()
            assert out == ref
            return 1e3 * (time.perf_counter() - t0) / (reps * batch)

        for rnd in range(2):
            for pre in (0, 1):
                ctx.set("msm_preplan", pre)
                row = {}
                row["single dev"] = t(lambda: [ctx.multi_scalar_mul_device(d[b], n, 0) for b in range(batch)])
                row["single e2e"] = t(lambda: [ctx.multi_scalar_mul(sc[b], 0) for b in range(batch)])
                for ahead in (0, 1):
                    ctx.set("msm_sort_ahead", ahead)
                    row[f"batch dev ahead={ahead}"] = t(lambda: ctx.multi_scalar_mul_batch(d, 0, on_device=True, n=n))
                    row[f"batch e2e ahead={ahead}"] = t(lambda: ctx.multi_scalar_mul_batch(host, 0))
                print(f"2^{lg} preplan={pre} ms per MSM: " + ", ".join(f"{k} {v:.3f}" for k, v in row.items()), flush=True)
        for p in d:
            ctx.dev_free(p)
    ctx.close()


if __name__ == "__main__":
    main()
