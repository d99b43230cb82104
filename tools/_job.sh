timeout 900 python -m pytest tests/test_collide.py tests/test_reference_scene.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for sw in 0 150 75 37.5 18.75 9.375 0 37.5; do echo "== strip $sw"; NBODY_COL_STRIP=$sw python tools/bench_refscene.py 25000 2>&1 | sed -n 2p | cut -c1-330; done
