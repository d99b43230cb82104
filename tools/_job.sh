timeout 300 build/sort_check > gpurun_out/r2b_sort_check_coop.log 2>&1; echo "sort_check coop rc=$?"; grep -E "FAILED|ALL OK|MISMATCH|error" gpurun_out/r2b_sort_check_coop.log | head
NBODY_SORT_COOP=0 timeout 300 build/sort_check > gpurun_out/r2b_sort_check_nocoop.log 2>&1; echo "sort_check nocoop rc=$?"; grep -E "FAILED|ALL OK|MISMATCH|error" gpurun_out/r2b_sort_check_nocoop.log | head
timeout 900 python -m pytest tests/test_gpu_bh.py tests/test_collide.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
python tools/bench_refscene.py 25000 2>&1 | sed -n 2p | cut -c1-330
python tools/bench_refscene.py 100000 2>&1 | sed -n 2p | cut -c1-330
python tools/bh_phases.py 1000000 2 1.0 | cut -c1-120;  python tools/bh_phases.py 4194304 3 0.5 | cut -c1-120
