mkdir -p gpurun_out/split
echo "== parity with RNNT_B200_SPLIT=1"
RNNT_B200_SPLIT=1 timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_holes.py tests/test_gpu_fullsize.py tests/test_gpu_bf16.py tests/test_gpu_cabi.py -q -m gpu -x 2>&1 | tail -4
for v in 0 1; do
  RNNT_B200_SPLIT=$v timeout -s KILL 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/split/c2_$v.json 2>gpurun_out/split/c2_$v.err
  for w in c2g c2b; do RNNT_B200_SPLIT=$v timeout -s KILL 200 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/split/${w}_$v.json 2>gpurun_out/split/${w}_$v.err; done
done
RNNT_B200_SPLIT=1 timeout -s KILL 120 python tools/fused_timeline.py c2 gpurun_out/split/timeline_c2_split.json > gpurun_out/split/timeline.log 2>&1
for f in gpurun_out/split/c2*.json; do python -c "
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('RES $f', round(d['ms_per_step'],4), round(d['roofline']['frac'],3), d.get('operator_ms_per_step'))
except Exception as e: print('RES $f FAILED', e)"; done
python -c "
import json
d=json.load(open('gpurun_out/split/timeline_c2_split.json')); print('TL', d['exact']['median_us'])"
