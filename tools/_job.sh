NBODY_BH_TRACE=1 python tools/bh_profile.py 25000 0 1.0 2 2>&1 | grep "top of" | tail -1
timeout 600 python -m pytest tests/test_gpu_bh.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
python tools/bench_refscene.py 25000 2>&1 | sed -n 2p | cut -c1-330
NBODY_BH_CTA_CLIMB=0 python tools/bench_refscene.py 25000 2>&1 | sed -n 2p | cut -c1-330
python tools/bench_refscene.py 32000 2>&1 | sed -n 2p | cut -c1-330
NBODY_BH_CTA_CLIMB=0 python tools/bench_refscene.py 32000 2>&1 | sed -n 2p | cut -c1-330
