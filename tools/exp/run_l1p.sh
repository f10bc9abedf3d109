mkdir -p gpurun_out/l1p
timeout 300 python -m pytest tests/test_gpu_holes.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/l1p/c2.json 2>gpurun_out/l1p/c2.err
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/l1p/c2_again.json 2>gpurun_out/l1p/c2.err
for f in gpurun_out/l1p/c2*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('RES $f', round(d['ms_per_step'],4), round(d['roofline']['frac'],3), d.get('operator_ms_per_step'))"; done
