python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-pass > gpurun_out/plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-pass > gpurun_out/ncu11.log 2>&1
python tools/conv_one.py 8 256 256 320 320 3 srb > gpurun_out/conv_one3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 8 -c 1 -o gpurun_out/conv_r1v4 python tools/conv_one.py 8 256 256 320 320 3 srb > gpurun_out/ncu12.log 2>&1
python tools/gn_one.py 8 256 256 320 > gpurun_out/gn_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 4 -c 1 -o gpurun_out/gn_r1b python tools/gn_one.py 8 256 256 320 > gpurun_out/ncu13.log 2>&1
cat gpurun_out/conv_one3.log gpurun_out/gn_one.log
