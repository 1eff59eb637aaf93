# throughput against the dimension: dense harmonic molecule (dense pipeline / k_hk_small) and the AS model (structured pipeline)
for d in 12 16 17 20 24 32 40 48 56 60; do
  n=$((148000 * 60 * 60 / (d * d) / 10)); if [ $n -gt 400000 ]; then n=400000; fi
  python tools/dense_probe.py $d $n 16 | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); d=$d; dr=d-6
F=16*d**3+8*dr*d*d+8*dr*dr*d+8/3*dr**3
print(json.dumps({'model':'dense harmonic','d':d,'ntraj':j['ntraj'],'traj_steps_per_s':j['traj_steps_per_s'],'algorithmic_tflops':F*j['traj_steps_per_s']*1e-12,'kernel':j['kernel']}))"
done
for d in 8 16 17 24 32 40 48 56 64; do
  n=$((148000 * 60 * 60 / (d * d) / 10 * 4)); if [ $n -gt 1000000 ]; then n=1000000; fi
  python bench.py --dim $d --ntraj $n --steps 20 --warmup 3 --no-cpu-baseline --no-dense-legs 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); d=$d
F=16*d**3+8/3*d**3
print(json.dumps({'model':'AS (separable)','d':d,'ntraj':$n,'traj_steps_per_s':j['value'],'e2e':j['e2e']['value'],'algorithmic_tflops':F*j['value']*1e-12,'kernel':j['config']['kernel']}))"
done
