# 1/2/4/8-GPU bench lines of the round, back to back on one box (launched as the driver does); NS = list of GPU counts
NS=${NS:-"1 2 4 8"}
for n in $NS; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 > gpurun_out/scale_r02_n1.json 2> gpurun_out/scale_r02_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n > gpurun_out/scale_r02_n$n.json 2> gpurun_out/scale_r02_n$n.err
  fi
done
for n in $NS; do python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/scale_r02_n$n.json") if l.startswith("{")][-1])
    print($n, j["steps"], j["value"], j["e2e"]["value"], j["roofline"]["whole_step"]["frac"])
except Exception as e:
    print($n, "FAILED", e)
PY
done
