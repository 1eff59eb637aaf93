# throughput of the BASELINE configs and the widening cases, one B200 (profiles/configs_bench_r02.jsonl)
python tools/configs_bench.py 2>/dev/null | grep "^{"
python tools/wm_probe.py 10000 50
python tools/wm_probe.py 200000 20
for d in 60 64 72 96; do python tools/dense_probe.py $d 29600 16; done
python tools/c5_probe.py 20000 8
python bench.py --dense --ntraj 148000 --steps 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print(json.dumps({'config':'rotated AS d=60 (bench.py --dense)','value':j['value'],'e2e':j['e2e']['value'],'kernel_frac':r['frac'],'whole_step_frac':r['whole_step']['frac'],'kernel_ms':r['whole_step']['kernel_ms']}))"
