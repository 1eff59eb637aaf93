# final evidence of round 2 on one B200: GPU parity suite, smoke, driver-shaped bench (both arms), launch list of the bench command
python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_tests.log 2>&1; tail -2 gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -3 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err || tail -5 gpurun_out/bench_r02_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_final_ref.json 2>> gpurun_out/bench_r02_final.err
python bench.py --ntraj 44400 --steps 20 --warmup 3 --no-cpu-baseline --dense-ntraj 14800 > gpurun_out/plain_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --ntraj 44400 --steps 20 --warmup 3 --no-cpu-baseline --dense-ntraj 14800 > gpurun_out/ncu_final_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rk4_stream -s 1 -c 1 -o gpurun_out/prof_final_stream -f python tools/dense_probe.py 60 29600 16 > gpurun_out/ncu_final_s.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_final_stream.ncu-rep > gpurun_out/ncu_r02_final_k_rk4_stream.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_final_stream.ncu-rep k_rk4_stream 16 >> gpurun_out/ncu_r02_final_k_rk4_stream.txt 2>&1
rm -f gpurun_out/prof_final_stream.ncu-rep
