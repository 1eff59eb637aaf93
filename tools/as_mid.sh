for d in 17 20 24 28 32; do for m in 33 17; do n=$((148000 * 60 * 60 / (d * d) / 10 * 4)); if [ $n -gt 1000000 ]; then n=1000000; fi
SC_CHUNK_DMIN=$m python bench.py --dim $d --ntraj $n --steps 20 --warmup 3 --no-cpu-baseline --no-dense-legs 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($d, 'dmin=$m', '%.3e' % j['value'], j['config']['kernel'], j['roofline']['whole_step']['kernel_ms'])"; done; done
