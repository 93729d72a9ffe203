python tools/probe_iter.py 2048x1024x1024 20
python tools/probe_iter.py 512x512x512 20
