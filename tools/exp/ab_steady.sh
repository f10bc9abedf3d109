#!/bin/bash
mkdir -p gpurun_out/ab
for v in "RNNT_B200_STEADY=0" "RNNT_B200_STEADY=1"; do
  for m in exact fast; do
  env $v RNNT_B200_PIPELINE=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ab/l.csv python tools/one_call.py c4 $m > /dev/null 2>&1
  python - "$v $m" <<'PY'
import csv, sys
rows=[r for r in csv.reader(open("gpurun_out/ab/l.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
seq=[(r[ki][:40], float(r[vi].replace(",",""))/1e3) for r in rows[1:] if "rnnt" in r[ki]]
print(sys.argv[1], " | ".join("%s %.0f" % (k.split("(")[0][-16:], v) for k, v in seq[-3:]))
PY
  done
done
