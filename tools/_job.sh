for v in 1 0 1 0; do NBODY_BH_FUSE_INSERT=$v python tools/bench_refscene.py 25000 2>&1 | sed -n 2p | cut -c1-330; done
