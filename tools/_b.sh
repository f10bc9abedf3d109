O=gpurun_out/ev
f() { python -c "
import sys,json
for line in sys.stdin:
    m,_,j=line.partition(' '); d=json.loads(j)['median_us']; print('$1', m, 'gather',d['gather_done'],'d96',d['a_d96'],'d160',d['a_d160'],'ns/step',round((d['a_d160']-d['a_d96'])/64*1000),'alpha_done',d['alpha_done'],'end',d['end'])
"; }
p() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['ms_per_step'], d['roofline']['frac'], d['lse_fast']['ms_per_step'])"; }
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -2
RNNT_B200_GATHER=ldg timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -1
python tools/fused_timeline.py c2 $O/timeline_c2c.json | f base
python bench.py --steps 100 --warmup 10 | tee $O/bench_c2c.json | p c2
python bench.py --workload c3 --steps 20 --warmup 5 | tee $O/bench_c3c.json | p c3
