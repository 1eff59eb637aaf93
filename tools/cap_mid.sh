D=${D:-24}
python tools/dense_probe.py $D 92500 16
ncu --set full --clock-control none --import-source on -k regex:k_rk4_stream -s 1 -c 1 -o gpurun_out/prof_mid -f python tools/dense_probe.py $D 92500 16 > gpurun_out/ncu_mid.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_mid.ncu-rep > gpurun_out/ncu_r02_k_rk4_stream_d$D.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_mid.ncu-rep k_rk4_stream 25 >> gpurun_out/ncu_r02_k_rk4_stream_d$D.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lu_warp -s 1 -c 1 -o gpurun_out/prof_mid2 -f python tools/dense_probe.py $D 92500 16 > gpurun_out/ncu_mid2.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_mid2.ncu-rep > gpurun_out/ncu_r02_k_lu_warp_d$D.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_mid2.ncu-rep k_lu_warp 20 >> gpurun_out/ncu_r02_k_lu_warp_d$D.txt 2>&1
rm -f gpurun_out/prof_mid.ncu-rep gpurun_out/prof_mid2.ncu-rep
