#!/bin/bash
# round 2, GPU call F: what does emitting the candidate stream cost the seed scan?  ncu counters, EMIT vs plain
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"
for m in 12 13; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:k_seed_scan -c 2 --csv --log-file gpurun_out/r02f_scan_mode$m.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 --scan-mode $m > /dev/null 2>&1; echo "ncu mode $m rc=$?"
  python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02f_scan_mode$m.csv")) if len(r) > 10 and r[0].isdigit()]
last = {}
for r in rows: last[r[-3]] = (r[-1], r[-2])
for k, v in last.items(): print("mode $m  %-75s %s %s" % (k, v[0], v[1]))
PY
done
