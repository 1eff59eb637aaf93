for w in c1 c3; do
python tools/small_probe.py $w
ncu --set full --clock-control none --import-source on -k regex:k_hk_small -s 1 -c 1 -o gpurun_out/prof_small_$w -f python tools/small_probe.py $w > gpurun_out/ncu_small_$w.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_small_$w.ncu-rep > gpurun_out/ncu_r02_k_hk_small_$w.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_small_$w.ncu-rep k_hk_small 30 >> gpurun_out/ncu_r02_k_hk_small_$w.txt 2>&1
rm -f gpurun_out/prof_small_$w.ncu-rep
done
