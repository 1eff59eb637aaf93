export SC_GDML_V1=1; python tools/c5_probe.py 20000 8; unset SC_GDML_V1
python tools/c5_probe.py 20000 8
ncu --set full --clock-control none --import-source on -k regex:k_gdml_eval2 -s 4 -c 1 -o gpurun_out/prof_gdml2 -f python tools/c5_probe.py 20000 4 > gpurun_out/ncu_gdml2.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_gdml2.ncu-rep > gpurun_out/ncu_r02_gdml2.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_gdml2.ncu-rep k_gdml_eval2 30 >> gpurun_out/ncu_r02_gdml2.txt 2>&1
rm -f gpurun_out/prof_gdml2.ncu-rep
