python tools/wm_probe.py 200000 20
ncu --set full --clock-control none --import-source on -k regex:k_wm_fused -s 1 -c 1 -o gpurun_out/prof_wm -f python tools/wm_probe.py 200000 20 > gpurun_out/ncu_wm.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_wm.ncu-rep > gpurun_out/ncu_r02_wm_fused.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_wm.ncu-rep k_wm_fused 25 >> gpurun_out/ncu_r02_wm_fused.txt 2>&1
rm -f gpurun_out/prof_wm.ncu-rep
