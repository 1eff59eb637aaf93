# end-of-round evidence refresh on one B200: full GPU suite, smoke, bench (both arms), throughput of the other configs
python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_tests.log 2>&1; tail -2 gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -3 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err || tail -5 gpurun_out/bench_r02_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_final_ref.json 2>> gpurun_out/bench_r02_final.err
bash tools/round2_configs.sh > gpurun_out/configs_bench_r02.jsonl 2> gpurun_out/configs_bench_r02.err
python tools/wm_big_probe.py 24 4000 4 | tail -1 >> gpurun_out/configs_bench_r02.jsonl
python tools/wm_big_probe.py 60 2368 2 | tail -1 >> gpurun_out/configs_bench_r02.jsonl
bash tools/dim_sweep.sh > gpurun_out/dim_sweep_r02.jsonl 2> gpurun_out/dim_sweep_r02.err
