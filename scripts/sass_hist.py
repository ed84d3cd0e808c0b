"""Static SASS opcode histogram of the hot kernels (cuobjdump on the built objects, no GPU needed):
    python scripts/sass_hist.py > profiles/<round>_sass_histogram.txt
Per function: instruction total, the integer-port classes (ALU pipe: LOP3 SHF PRMT IADD3 ISETP SEL ...; FMA-heavy pipe:
IMAD.WIDE / IMAD.HI; FMA-lite: other IMAD), memory, control, and the full opcode table."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"msm.o": ("k_accumulate", "k_pass2<16, 2>", "k_pass1<16>", "gf_mul_call", "gf_inv_tab", "k_binv_coop", "k_round_warp", "k_ld_tree<false, false>"),
        "prover.o": ("k_extend_level<3, 1, true, true>", "k_extend_level<3, 2, true, true>", "k_extend_level<3, 0, true, false>", "k_extend_fused",
                     "k_r1cs_sides", "k_extend_down_sel", "k_extend_up_sel")}
def classify(op):
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"): return "fma_heavy (IMAD.WIDE/HI)"
    if op.startswith("IMAD") or op.startswith("IMUL"): return "fma_lite (IMAD)"
    if re.match(r"(LOP3|SHF|PRMT|IADD3|IADD|ISETP|SEL|VIADD|LEA|MOV|VIMNMX|PLOP3|ICMP|POPC|FLO|BREV|SGXT|BMSK|R2P|P2R|LOP|IABS|UIADD|ULOP|USHF|UMOV|ULEA|UISETP|USEL|UPRMT)", op): return "alu"
    if re.match(r"(LDG|STG|LDS|STS|LDL|STL|LDC|LD\.|ST\.|ATOM|RED|LDGSTS|ULDC|LDSM|CCTL|MEMBAR|ERRBAR)", op): return "memory"
    if re.match(r"(BRA|BSSY|BSYNC|CALL|RET|EXIT|BAR|WARPSYNC|NANOSLEEP|YIELD|BREAK|BRX|JMP|KILL|NOP|DEPBAR|BMOV|CS2R|S2R|S2UR|VOTE|SHFL|REDUX|MATCH|ELECT)", op): return "control/warp"
    return "other"
for obj, names in WANT.items():
    path = os.path.join(ROOT, "dv-pari_b200", "csrc", obj)
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            funcs[cur][m.group(1)] += 1
    for mangled, c in funcs.items():
        short = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip().split("(")[0]
        if not any(short.endswith(n) or short.endswith(n.replace("dvp::", "")) for n in names):
            continue
        cls = collections.Counter()
        for op, v in c.items():
            cls[classify(op)] += v
        tot = sum(c.values())
        print(f"== {short}   [{obj}]   {tot} instructions")
        print("   classes: " + ", ".join(f"{k} {v} ({100.0*v/tot:.1f} %)" for k, v in cls.most_common()))
        base = collections.Counter()
        for op, v in c.items():
            base[op.split(".")[0] + (".WIDE" if ".WIDE" in op else "")] += v
        print("   opcodes: " + ", ".join(f"{k} {v}" for k, v in base.most_common()))
        print()
