timeout 900 python -m pytest tests/test_gpu_bh.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for v in 1 0 1 0; do NBODY_BH_COUNT_SCAN=$v python tools/bench_refscene.py 25000 2>&1 | sed -n 2p | cut -c1-330; done
for v in 1 0; do NBODY_BH_COUNT_SCAN=$v python tools/bh_phases.py 1000000 2 1.0 | cut -c1-120; NBODY_BH_COUNT_SCAN=$v python tools/bh_phases.py 4194304 3 0.5 | cut -c1-120; done
