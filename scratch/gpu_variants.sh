#!/bin/bash
# A/B: the same run with differently built libraries (scratch/lib_*.so)
cp mosaic-art-rust_b200/libmosaic_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp scratch/lib_$v.so mosaic-art-rust_b200/libmosaic_b200.so
  echo "== $v"; python scratch/chain_prof.py C3 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[-300:]); continue
    print({k: d.get(k) for k in ('ms','n_fast','n_slow','cyc_slow','cyc_scan','cyc_listen','cyc_total_both_warps','cyc_per_slow')})"
done
cp /tmp/lib_orig.so mosaic-art-rust_b200/libmosaic_b200.so
