"""Static look at one kernel's SASS (no GPU needed): instruction mix of every loop (backward branch) of a function.

usage: python tools/sass_loops.py <lib.so> <substring of the mangled kernel name> [min_instructions]
Prints, innermost-first by size, each backward-branch region with its counts of FP64 / shared / global / local
(spill) / integer-move instructions -- the numbers behind "FP64 share of the issue slots" in DESIGN.md."""
import re
import subprocess
import sys
from collections import Counter


def function_sass(lib, name):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.splitlines()
    start = [i for i, l in enumerate(out) if "Function :" in l]
    for k, i in enumerate(start):
        if name in out[i]:
            end = start[k + 1] if k + 1 < len(start) else len(out)
            return out[i:end]
    raise SystemExit(f"no function matching {name}")


def main():
    lib, name = sys.argv[1], sys.argv[2]
    min_ins = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    ins = []
    for l in function_sass(lib, name):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, text) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    print(f"{name}: {len(ins)} instructions, {len(loops)} backward branches")
    for s, e in sorted(loops, key=lambda t: t[1] - t[0]):
        n = e - s + 1
        if n < min_ins:
            continue
        c = Counter()
        for _, text in ins[s:e + 1]:
            op = text.split()[0] if not text.startswith("@") else text.split()[1]
            base = op.split(".")[0]
            c[base] += 1
        fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        print(f"  loop [{ins[s][0]:#x}..{ins[e][0]:#x}] {n} instr: FP64 {fp64} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']}) "
              f"LDS {c['LDS']} STS {c['STS']} LDG {c['LDG']} STG {c['STG']} LDL {c['LDL']} STL {c['STL']} "
              f"MOV {c['MOV']} IMAD {c['IMAD']} other {n - fp64 - c['LDS'] - c['STS'] - c['LDG'] - c['STG'] - c['LDL'] - c['STL'] - c['MOV'] - c['IMAD']}")
        rest = Counter({k: v for k, v in c.items() if k not in ("DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "LDL", "STL", "MOV", "IMAD")})
        print("     other:", dict(rest.most_common(12)))


if __name__ == "__main__":
    main()
