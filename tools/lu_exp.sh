# LU kernel crossover experiment of round 2: rotated AS models (d' = d), k_lu_warp against k_lu_mma at 2 / 3 / 4 CTAs
# per SM; SC_LU_MMA_MIN = rank up to which k_lu_warp is used, SC_LU_MMA_CTAS needs the experiment hook of that commit
run() { python bench.py --dense --dim $1 --ntraj 148000 --steps 16 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('$2', $1, '%.3e' % j['value'], 'lu %.1f' % r['whole_step']['kernel_ms']['lu'])"; }
for d in 26 32; do for c in 2 3 4; do SC_LU_MMA_MIN=16 SC_LU_MMA_CTAS=$c run $d mma_ctas$c; done; done
for d in 20 24; do run $d warp; for c in 3 4; do SC_LU_MMA_MIN=16 SC_LU_MMA_CTAS=$c run $d mma_ctas$c; done; done
