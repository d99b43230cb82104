timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multiproc.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
python tools/bh_phases.py 4194304 3 0.5 8
for g in 8 4 2; do host/_build/nbody_run --ic galaxy --n 4194304 --dims 3 --eps 0.01 --dt 0.001 --steps 1000 --algo bh --theta 0.5 --near-leaves on --gpus $g 2>&1 | tail -2; done
