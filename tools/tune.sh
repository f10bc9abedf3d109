#!/bin/bash
# scratch: sweep the gather-warp split on one box
for gw in 16 14 12 8; do
  for lse in auto exact; do
    for wl in c2 c3; do
      echo -n "gw=$gw lse=$lse $wl: "
      RNNT_B200_GATHER_WARPS=$gw python bench.py --workload $wl --lse $lse --steps 200 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1),'us', round(d['roofline']['frac'],3))"
    done
  done
done
