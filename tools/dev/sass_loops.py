"""Instruction counts of the innermost loops of a kernel in an object file (cuobjdump -sass):
    python tools/dev/sass_loops.py gaussian_splatting_b200/_build/gsr_render.o k_render_bwdILb1
Prints every backward branch (loop) with its body length and an opcode histogram of the largest loops."""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s+Function : ", txt)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for line in f.splitlines():
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        addr_idx = {a: i for i, (a, _) in enumerate(ins)}
        print(name, "total", len(ins))
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_idx:
                    loops.append((i - addr_idx[tgt] + 1, tgt, a))
        for n, tgt, a in sorted(loops, reverse=True)[:8]:
            body = ins[addr_idx[tgt]:addr_idx[a] + 1]
            hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
            print(f"  loop {tgt:#06x}..{a:#06x}: {n} instr  " + " ".join(f"{k}:{v}" for k, v in hist.most_common(14)))


if __name__ == "__main__":
    main()
