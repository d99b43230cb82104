timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multiproc.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
