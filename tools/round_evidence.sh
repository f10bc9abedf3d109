#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): tests on every kernel path, smoke, bench lines for all
# workloads (+ the reference arm with REF=1), ncu launch lists and full captures (c4 with C4=1), fused-kernel
# phase timelines.  Outputs under gpurun_out/ev/ ; summarised into profiles/ by tools/ncu_summary.py afterwards.
set -u
O=gpurun_out/ev; mkdir -p $O
{
echo "== pytest -m gpu (default paths)";            timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
echo "== pytest -m gpu, RNNT_B200_PATH=general";    RNNT_B200_PATH=general timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -2
echo "== pytest parity, LDG gather + STG fill";     RNNT_B200_GATHER=ldg RNNT_B200_FILL=stg timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cabi.py -q -m gpu 2>&1 | tail -2
echo "== pytest parity, RNNT_B200_LSE=fast";        RNNT_B200_LSE=fast timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -2
echo "== smoke";                                    python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
} > $O/tests.log 2>&1
python bench.py --steps 100 --warmup 10 > $O/bench_c2.json 2> $O/bench_c2.err
if [ "${REF:-0}" = 1 ]; then python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_c2_reference.json 2> $O/bench_c2_reference.err; fi
for w in c3 c4 c5mb; do python bench.py --workload $w --steps 20 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; done
python tools/fused_timeline.py c2 $O/timeline_c2.json > $O/timeline_c2.log 2>&1
python tools/fused_timeline.py c3 $O/timeline_c3.json > $O/timeline_c3.log 2>&1
python tools/compact_time.py > $O/compact_time.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -f -o $O/fused_c2 python tools/one_call.py c2 exact > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -f -o $O/fused_c3 python tools/one_call.py c3 exact > /dev/null 2>&1
if [ "${C4:-0}" = 1 ]; then
RNNT_B200_PIPELINE=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_bench_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
RNNT_B200_PIPELINE=0 ncu --set full --clock-control none --import-source on -k regex:"k_gather|k_wavefront|k_expand" -s 3 -c 3 -f -o $O/general_c4 python tools/one_call.py c4 exact > /dev/null 2>&1
fi
cat $O/tests.log; cat $O/compact_time.log; for f in $O/bench_c2.json $O/bench_c3.json $O/bench_c4.json $O/bench_c5mb.json; do python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['roofline']['frac'], (d.get('lse_fast') or {}).get('ms_per_step'), d['e2e']['value'])"; done
