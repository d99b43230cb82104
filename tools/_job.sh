timeout 900 python -m pytest tests/test_gpu_bh.py -m gpu -q -x -p no:cacheprovider -k "variants" 2>&1 | tail -12
