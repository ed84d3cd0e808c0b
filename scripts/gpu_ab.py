"""Development probe: the same MSM sizes with two builds of the library (DVP_LIB) and knob settings, interleaved so that
both see the same box and clocks: python scripts/gpu_ab.py libA.so libB.so lg[,lg..] [knob=v1,v2 ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs, lgs, knobs = sys.argv[1:3], [int(x) for x in sys.argv[3].split(",")], sys.argv[4:]
for lg in lgs:
    for rep in range(2):
        for lib in libs:
            env = dict(os.environ, DVP_LIB=os.path.join(ROOT, "dv-pari_b200", lib))
            # knobs an old build does not know are skipped for it
            ks = [k for k in knobs if not (lib.endswith("_old.so") and k.split("=")[0] in ("even_waves",))]
            out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_sweep.py"), str(lg)] + ks, env=env,
                                 capture_output=True, text=True)
            for line in out.stdout.splitlines():
                print(f"[{lib}] {line}", flush=True)
            if out.returncode:
                print(f"[{lib}] FAILED rc={out.returncode}\n{out.stderr[-2000:]}", flush=True)
