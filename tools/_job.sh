timeout 600 python -m pytest tests/test_gpu_bh.py -m gpu -q -x -p no:cacheprovider -k "tree_matches or golden or octree_matches" 2>&1 | tail -2
python tools/bench_refscene.py 25000 2>&1 | sed -n 2p | cut -c1-330
