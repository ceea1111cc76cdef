for f in "" s; do python tools/conv_one.py 8 256 256 8 320 3 "$f"; done
B200DIFF_SKIP_EPI=1 python tools/conv_one.py 8 256 256 8 320 3 ""
B200DIFF_NO_VEC32=1 python tools/conv_one.py 8 256 256 8 320 3 ""
python tools/conv_one.py 8 256 256 8 256 3 ""
python tools/conv_one.py 8 256 256 8 64 3 ""
