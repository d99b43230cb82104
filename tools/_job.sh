ncu --set full --clock-control none -k regex:"bh_walk_direct|bh_emit_local|os_sort_all|bh_climb" --launch-skip 8 --launch-count 4 -f -o gpurun_out/r2c_bh1m_kernels python tools/bh_profile.py 1000000 2 1.0 4 > gpurun_out/ncu_1m.log 2>&1
tail -2 gpurun_out/ncu_1m.log | cut -c1-150
