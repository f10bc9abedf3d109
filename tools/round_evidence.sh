#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): tests on every kernel path, smoke, bench lines for all
# workloads (+ the reference arm with REF=1), ncu launch lists and full captures (c4 with C4=1, k_fused at c2 with
# FUSED=1), fused-kernel phase timelines, compute-sanitizer.  QUICK=1 skips the forced-path test passes and synccheck.  Outputs under gpurun_out/ev/ ; copied / summarised into profiles/ afterwards
# (tools/ncu_summary.py, tools/collect_profiles.py).
set -u
O=gpurun_out/ev; mkdir -p $O
TLOG=tests.log; [ "${QUICK:-0}" = 1 ] && TLOG=tests_quick.log
{
echo "== pytest -m gpu (default paths)";            timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
if [ "${QUICK:-0}" != 1 ]; then
echo "== pytest -m gpu, RNNT_B200_PATH=general";    RNNT_B200_PATH=general timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py::test_python_api_one_launch_and_upstream_scaling 2>&1 | tail -2
echo "== pytest parity, LDG gather + STG fill";     RNNT_B200_GATHER=ldg RNNT_B200_FILL=stg timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cabi.py tests/test_gpu_bf16.py -q -m gpu 2>&1 | tail -2
echo "== pytest parity, one row buffer per gather warp"; RNNT_B200_ROW_BUFS=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16.py -q -m gpu 2>&1 | tail -2
fi
echo "== pytest parity, RNNT_B200_LSE=fast";        RNNT_B200_LSE=fast timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -2
echo "== smoke";                                    python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
[ "${QUICK:-0}" = 1 ] || { echo "== synccheck"; }
[ "${QUICK:-0}" = 1 ] ||                            timeout 300 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden_dense or dense_vs_oracle" 2>&1 | tail -3
} > $O/$TLOG 2>&1
python bench.py --steps 100 --warmup 10 > $O/bench_c2.json 2> $O/bench_c2.err
if [ "${REF:-0}" = 1 ]; then python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_c2_reference.json 2> $O/bench_c2_reference.err; fi
for w in c2g c3 c3d c4 c4d c5mb c2b c5mbb c2l c5mbl; do timeout 400 python bench.py --workload $w --steps 20 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; done
python tools/fused_timeline.py c2 $O/timeline_c2.json > $O/timeline_c2.log 2>&1
python tools/fused_timeline.py c3 $O/timeline_c3.json > $O/timeline_c3.log 2>&1
# launch lists: the API step as the bench runs it (eager, so that ncu sees every launch), cold-cache + serialised: shares only
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > /dev/null 2>&1
if [ "${FUSED:-0}" = 1 ]; then
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -f -o $O/fused_c2 python tools/one_call.py c2 exact > /dev/null 2>&1
fi
if [ "${C4:-0}" = 1 ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_bench_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_bench_c4d.csv python bench.py --workload c4d --steps 2 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gather|k_wavefront|k_expand" -s 3 -c 3 -f -o $O/general_c4 python tools/one_call.py c4 exact > /dev/null 2>&1
fi
cat $O/$TLOG; for f in $O/bench_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],4), round(d['roofline']['frac'],3) if 'roofline' in d else None, round(d['e2e']['value']))
except Exception as e: print('$f', 'FAILED', e)"; done
