python tools/step_breakdown.py cfg2 8 > gpurun_out/breakdown_cfg2.txt 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-pass > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-pass > gpurun_out/ncu6.log 2>&1
python tools/conv_one.py 8 256 256 320 320 > gpurun_out/conv_one2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 8 -c 1 -o gpurun_out/conv_r1v3 python tools/conv_one.py 8 256 256 320 320 > gpurun_out/ncu7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 3 -c 1 -o gpurun_out/gn_r1 python tools/microbench.py gn > gpurun_out/ncu8.log 2>&1
tail -3 gpurun_out/breakdown_cfg2.txt
