ncu --set full --clock-control none --import-source on -k regex:k_lu_mma -s 4 -c 1 -o gpurun_out/prof_lu1 -f ./tools/lu_mma_bench 14208 60 > gpurun_out/ncu_lu1.log 2>&1
ncu -i gpurun_out/prof_lu1.ncu-rep --page source --csv > gpurun_out/lu1_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_lu1.ncu-rep > gpurun_out/ncu_lu_occ1.txt 2>&1
rm -f gpurun_out/prof_lu1.ncu-rep
head -c 3000000 gpurun_out/lu1_source.csv > gpurun_out/lu1_source_head.csv; rm gpurun_out/lu1_source.csv
