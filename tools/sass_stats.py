#!/usr/bin/env python3
"""Static SASS statistics of the main kernels (cuobjdump on the objects under build/obj[_opt]): registers, stack and the
instruction mix that matters on this path.  No GPU needed.   python tools/sass_stats.py [build/obj] > table.md"""
import re
import subprocess
import sys
from collections import Counter

OBJ = sys.argv[1] if len(sys.argv) > 1 else "build/obj"
KERNELS = [("inst_bn254_g1.o", "k_msm_accumulateI"), ("inst_bn254_g1.o", "k_msm_accumulate52I"),
           ("inst_bn254_g1.o", "k_msm_ba_levelI"), ("inst_bn254_g1.o", "k_msm_combineI"),
           ("inst_bn254_g2.o", "k_msm_accumulateI"), ("inst_bn254_g2.o", "k_msm_ba_levelI"),
           ("inst_bls12_381_g1.o", "k_msm_accumulateI"), ("inst_bls12_381_g1.o", "k_msm_accumulate52I"),
           ("inst_bw6_761_g1.o", "k_msm_accumulateI"),
           ("inst_bn254_fr.o", "k_ntt_passI"), ("inst_bn254_fr.o", "k_plonk_constraintsI"), ("inst_bn254_fr.o", "k_h_pointwiseI")]
MIX = ("IMAD.WIDE", "IMAD", "IADD3", "DFMA", "DADD", "LOP3", "SHF", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "CALL")


def main():
    print(f"| object ({OBJ}) | kernel | regs | stack B | instr | " + " | ".join(MIX) + " |")
    print("|---|---|---|---|---|" + "---|" * len(MIX))
    cache = {}
    for obj, kern in KERNELS:
        path = f"{OBJ}/{obj}"
        if path not in cache:
            try:
                sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
                res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True, check=True).stdout
            except Exception as e:  # object not built
                print(f"| {obj} | {kern} | n/a ({e.__class__.__name__}) |")
                continue
            cache[path] = (sass, res)
        sass, res = cache[path]
        # split per function
        blocks = re.split(r"\n\s*Function : ", sass)
        for b in blocks[1:]:
            name = b.split("\n", 1)[0].strip()
            if kern not in name:
                continue
            ops = Counter()
            total = 0
            for ln in b.splitlines():
                m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
                if not m:
                    continue
                total += 1
                op = m.group(1)
                for key in MIX:
                    if op == key or op.startswith(key + "."):
                        # IMAD.WIDE counted separately from plain IMAD
                        if key == "IMAD" and op.startswith("IMAD.WIDE"):
                            continue
                        ops[key] += 1
                        break
            rm = re.search(re.escape(name) + r":\s*\n\s*(REG:\d+ STACK:\d+)", res)
            regs, stack = ("?", "?")
            if rm:
                regs, stack = re.findall(r"\d+", rm.group(1))[:2]
            short = re.sub(r"^_ZN5gb200\d+", "", name)[:46]
            print(f"| {obj} | {short} | {regs} | {stack} | {total} | " + " | ".join(str(ops[k]) for k in MIX) + " |")


if __name__ == "__main__":
    main()
