#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what the library's kernels are made of (cuobjdump -sass of the
built libmosaic_b200.so): TMA bulk copies (UBLKCP / UTMALDG), mbarriers (SYNCS), warp reductions (REDUX / CREDUX),
packed FP32 (FADD2 / FMUL2 / FFMA2), integer dot products (IDP), MUFU, tensor-core MMA (UTC*MMA / HMMA — none expected:
inner dimension 3 is not a contraction).   python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mosaic-art-rust_b200", "libmosaic_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS"), ("REDUX", r"\bC?REDUX"),
                                ("F*2 (packed f32)", r"\b(FADD2|FMUL2|FFMA2)\b"), ("IDP", r"\bIDP"), ("MUFU", r"\bMUFU"),
                                ("ATOMS/ATOMG", r"\bATOM[SG]"), ("tensor MMA", r"\b(UTC\w*MMA|HMMA|IMMA|QGMMA)"), ("STG.128/LDG.128", r"\b(STG|LDG)\.E\.128")])
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, name, total = collections.OrderedDict(), None, 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"mosaic::\(anonymous namespace\)::", "", name)
        kern[name] = collections.Counter()
        continue
    if name is None or not re.search(r"/\*[0-9a-f]{4}\*/", line):
        continue
    kern[name]["instructions"] += 1
    for k, p in pats.items():
        if re.search(p, line):
            kern[name][k] += 1
print(f"libmosaic_b200.so: {len(kern)} kernels, architectures {arch}")
cols = ["instructions"] + list(pats)
print("kernel".ljust(78) + " ".join(c[:12].rjust(12) for c in cols))
tot = collections.Counter()
for k, c in kern.items():
    print(k[:77].ljust(78) + " ".join(str(c.get(col, 0)).rjust(12) for col in cols))
    tot.update(c)
print("TOTAL".ljust(78) + " ".join(str(tot.get(col, 0)).rjust(12) for col in cols))
