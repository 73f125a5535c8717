"""profiles/r02_sass_*.txt: SASS of the hot-path kernels from the built library (cuobjdump -sass), one file per kernel, plus a
summary of the memory instructions in the SpMV inner loop (what proves 128-bit / coalesced accesses; there are no tensor-core
or TMA instructions on this path by design -- DESIGN.md section 4)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "highs_b200", "libb200pdlp.so")
OUT = os.path.join(ROOT, "profiles")
WANT = {
    "K1_primal_step": "primal_step_kernel",
    "K2_spmv_dual": "spmv_sell_kernelINS_13DualEpilogueTILi1",
    "K3_spmv_primal": "spmv_sell_kernelINS_14PrimalEpilogue",
    "K4_step_rule": "step_rule_kernelE",
    "C2_spmv_check_rows": "spmv_sell_kernelINS_17CheckRowEpilogueTILb1",
    "C3_spmv_check_cols": "spmv_sell_kernelINS_16CheckColEpilogue",
    "C4_check_decide": "check_decide_kernelE",
}
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
summary = ["# r02: SASS of the hot-path kernels (cuobjdump -sass highs_b200/libb200pdlp.so, sm_100a)\n",
           "| kernel | instructions | LDG.E.64 | LDG.E.128 | LDG (32-bit) | STG | DFMA | DMUL | DADD | SHFL | BAR | ATOM/RED |", "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for tag, pat in WANT.items():
    body = next((p for p in parts[1:] if p.split("\n", 1)[0].find(pat) >= 0), None)
    if body is None:
        print("missing", tag, file=sys.stderr)
        continue
    name = body.split("\n", 1)[0]
    if tag.startswith("K"):   # full listings of the four pass kernels; the check kernels appear in the summary only
        with open(os.path.join(OUT, f"r02_sass_{tag}.txt"), "w") as f:
            f.write("Function : " + body)
    ops = collections.Counter()
    n = 0
    for line in body.splitlines():
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            n += 1
            ops[m.group(1)] += 1
    cnt = lambda pref: sum(v for k, v in ops.items() if k.startswith(pref))
    ldg64 = sum(v for k, v in ops.items() if k.startswith("LDG") and ".64" in k)
    ldg128 = sum(v for k, v in ops.items() if k.startswith("LDG") and ".128" in k)
    ldg32 = cnt("LDG") - ldg64 - ldg128
    summary.append(f"| {tag} `{name[:60]}` | {n} | {ldg64} | {ldg128} | {ldg32} | {cnt('STG')} | {cnt('DFMA')} | {cnt('DMUL')} | {cnt('DADD')} | "
                   f"{cnt('SHFL')} | {cnt('BAR')} | {cnt('ATOM') + cnt('RED')} |")
summary.append("\nNotes: compiled with -fmad=false, so products and sums stay separate (DMUL + DADD, no DFMA) -- the reference CPU path's "
               "rounding; the SpMV body loads `col` as 32-bit and `val` / gathers as 64-bit per lane, warp-coalesced (k-major, lane-minor "
               "layout: one 128-byte line of `col` and two of `val` per k step); K1 moves pairs with 128-bit accesses.  No UTC*MMA / "
               "UTMALDG: there is no dense contraction and no dense tile to stage on this path.")
open(os.path.join(OUT, "r02_sass_summary.md"), "w").write("\n".join(summary) + "\n")
print("\n".join(summary))
