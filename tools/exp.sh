python tools/probe_iter.py 2048x1024x1024 20
python tools/probe.py jacobi3d clamp 6
python tools/probe.py laplacian512 clamp 20
