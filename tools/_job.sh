timeout 300 build/sort_check > gpurun_out/r2b_sort_check_coop.log 2>&1; echo "sort_check coop rc=$?"; grep -E "FAILED|ALL OK|MISMATCH|error" gpurun_out/r2b_sort_check_coop.log | head
NBODY_SORT_COOP=0 timeout 300 build/sort_check > gpurun_out/r2b_sort_check_nocoop.log 2>&1; echo "sort_check nocoop rc=$?"; grep -E "FAILED|ALL OK|MISMATCH|error" gpurun_out/r2b_sort_check_nocoop.log | head
timeout 600 python -m pytest tests/test_gpu_bh.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
NBODY_BH_LOCAL=0 timeout 600 python -m pytest tests/test_gpu_bh.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for v in "" "NBODY_BH_LOCAL=0" "NBODY_SORT_LAZY=0" "NBODY_BH_LOCAL=0 NBODY_SORT_LAZY=0"; do echo "== $v"; env $v python tools/bench_refscene.py 25000 2>&1 | sed -n 2p; done
for v in "" "NBODY_BH_LOCAL=0" "NBODY_SORT_LAZY=0" "NBODY_SORT_COOP=0" "NBODY_BH_LOCAL=0 NBODY_SORT_LAZY=0 NBODY_SORT_COOP=0"; do echo "== $v"; env $v python tools/bh_phases.py 1000000 2 1.0; env $v python tools/bh_phases.py 4194304 3 0.5; done
