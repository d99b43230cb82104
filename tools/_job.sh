ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/r2b_launches_refscene_warm.csv python tools/bh_profile.py 25000 0 1.0 6 > gpurun_out/ncu1.log 2>&1
python tools/bench_refscene.py 25000 2>&1 | sed -n 2p | cut -c1-330
