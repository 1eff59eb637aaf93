# ncu evidence of round 2 (one GPU): launch list of the bench command + full captures of the dense-pipeline kernels.
# The reports are condensed on the box (tools/ncu_summary.py, tools/ncu_lines.py) and deleted: gpurun_out/ is capped at 64 MiB.
cap() {  # name kernel-regex skip command...
  name=$1; kern=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -o gpurun_out/prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$name.ncu-rep > gpurun_out/ncu_r02_$name.txt 2>&1
  python tools/ncu_lines.py gpurun_out/prof_$name.ncu-rep $kern 25 >> gpurun_out/ncu_r02_$name.txt 2>&1
  rm -f gpurun_out/prof_$name.ncu-rep
  tail -n 1 gpurun_out/ncu_$name.log
}
python tools/dense_probe.py 60 29600 16 > gpurun_out/dp_plain.log 2>&1 || exit 1
cap stream_harm k_rk4_stream 1 python tools/dense_probe.py 60 29600 16
cap rmult k_rmult 1 python tools/dense_probe.py 60 29600 16
python bench.py --dense --ntraj 29600 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/dense_plain.log 2>&1 || exit 1
cap stream_rot k_rk4_stream 1 python bench.py --dense --ntraj 29600 --steps 16 --warmup 3 --no-cpu-baseline
python bench.py --ntraj 44400 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r02.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_a.csv python bench.py --ntraj 44400 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r02_l.log 2>&1
cap wcols_k20 k_rk4_wcols 2 python bench.py --ntraj 44400 --steps 20 --warmup 3 --no-cpu-baseline --no-dense-legs
cap lu_mma_k20 k_lu_mma 2 python bench.py --ntraj 44400 --steps 20 --warmup 3 --no-cpu-baseline --no-dense-legs
ls -la gpurun_out
