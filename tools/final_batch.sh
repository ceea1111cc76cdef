set -x
python bench.py > gpurun_out/final_cfg2.json 2> gpurun_out/final_cfg2.err
python bench.py --workload cfg5 > gpurun_out/final_cfg5.json 2> gpurun_out/final_cfg5.err
python tools/step_breakdown.py cfg2 8 > gpurun_out/r2_step_breakdown_cfg2.txt 2>&1
python tools/conv_one.py 8 256 256 320 320 3 sb > gpurun_out/conv_plain.txt 2>&1
python tools/conv_one.py 8 256 256 320 320 3 sbl >> gpurun_out/conv_plain.txt 2>&1
python tools/conv_one.py 8 256 256 320 320 3 srl >> gpurun_out/conv_plain.txt 2>&1
python tools/conv_one.py 8 256 256 320 320 3 srLl >> gpurun_out/conv_plain.txt 2>&1
python tools/gn_one.py 8 256 256 320 >> gpurun_out/conv_plain.txt 2>&1
python tools/gn_one.py 8 256 256 320 lo >> gpurun_out/conv_plain.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 6 -c 1 -o gpurun_out/conv_lo_prof -f python tools/conv_one.py 8 256 256 320 320 3 sbl > gpurun_out/conv_lo_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_apply_kernel -s 4 -c 1 -o gpurun_out/gn_lo_prof -f python tools/gn_one.py 8 256 256 320 lo > gpurun_out/gn_lo_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1200 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-pass --no-gpu-eager-reference --no-other-workloads > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/conv_plain.txt
