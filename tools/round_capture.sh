# full GPU check of the round: parity tests, bench line, ncu launch list, ncu --set full of the two dominant kernels
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01h_tests.log 2>&1; tail -2 gpurun_out/r01h_tests.log
python bench.py > gpurun_out/bench_r01_h.json 2> gpurun_out/bench_r01_h.err || tail -5 gpurun_out/bench_r01_h.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_h_ref.json 2>> gpurun_out/bench_r01_h.err
python bench.py --ntraj 44400 --steps 10 --warmup 1 --no-cpu-baseline > gpurun_out/plain_h.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_h.csv python bench.py --ntraj 44400 --steps 10 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rk4_wcols -s 2 -c 1 -o gpurun_out/prof_wcols -f python bench.py --ntraj 44400 --steps 10 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fh1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lu_mma -s 2 -c 1 -o gpurun_out/prof_lumma_app -f python bench.py --ntraj 44400 --steps 10 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fh2.log 2>&1
tail -n 2 gpurun_out/ncu_fh1.log; tail -n 2 gpurun_out/ncu_fh2.log
