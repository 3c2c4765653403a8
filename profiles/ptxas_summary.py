"""Turns `nvcc ... -Xptxas -v` output (one file per .cu) into one line per kernel: registers, spills, stack, shared memory.

    for f in lab_mean nn_topk greedy similarity anneal render; do
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xptxas -v -c csrc/$f.cu -o /tmp/$f.o 2> /tmp/ptx/$f.txt
    done
    python profiles/ptxas_summary.py /tmp/ptx/*.txt > profiles/r02_ptxas_summary.txt
(the Makefile's flags; no GPU needed)."""
import re
import subprocess
import sys


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(files):
    rows = []
    for f in files:
        text = open(f).read()
        unit = f.rsplit("/", 1)[-1].replace(".txt", ".cu")
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(.*?)(?=ptxas info    : Compiling entry|\Z)", text, flags=re.S):
            name, body = m.group(1), m.group(2)
            fp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", body)
            used = re.search(r"Used (\d+) registers", body)
            smem = re.search(r"(\d+) bytes smem", body)
            rows.append((unit, name, int(used.group(1)), int(fp.group(1)), int(fp.group(2)), int(fp.group(3)),
                         int(smem.group(1)) if smem else 0))
    names = demangle([r[1] for r in rows])
    print(f"{'file':14s} {'regs':>4s} {'stack':>6s} {'spill st/ld':>12s} {'static smem':>11s}  kernel")
    for unit, name, regs, stack, sst, sld, smem in rows:
        short = re.sub(r"\(anonymous namespace\)::", "", names.get(name, name))
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\(.*$", "", short)[:110]
        print(f"{unit:14s} {regs:4d} {stack:6d} {sst:5d}/{sld:<6d} {smem:11d}  {short}")
    spilled = [r for r in rows if r[4] or r[5]]
    print(f"\n{len(rows)} kernels; {len(spilled)} with register spills" + (": " + ", ".join(sorted({r[0] for r in spilled})) if spilled else ""))


if __name__ == "__main__":
    main(sys.argv[1:])
