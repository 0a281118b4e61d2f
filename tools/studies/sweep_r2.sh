for t in 4 6 8 16 24; do python tools/profile_align.py cfg2 1.0 target_pts_per_cell=$t > gpurun_out/sw_t$t.log 2>&1; done
for m in 0.02 0.1 0.2; do python tools/profile_align.py cfg2 1.0 lazy_margin=$m > gpurun_out/sw_lm$m.log 2>&1; done
for r in 0.5 1.0; do python tools/profile_align.py cfg2 1.0 skip_reach_cells=$r > gpurun_out/sw_reach$r.log 2>&1; done
grep -H "^align:" gpurun_out/sw_*.log
