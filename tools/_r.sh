python tools/gn_one.py 8 8 8 960
python tools/gn_one.py 8 16 16 960
python tools/gn_one.py 8 32 32 640
python tools/gn_one.py 8 64 64 640
python tools/gn_one.py 8 128 128 320
