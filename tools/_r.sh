ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 8 -c 1 -o gpurun_out/conv_epi python tools/conv_one.py 8 256 256 8 320 3 s > gpurun_out/ncu10.log 2>&1
tail -1 gpurun_out/ncu10.log
