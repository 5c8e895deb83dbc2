#!/bin/bash
# Regenerate the committed profile summaries from one launch list and one `ncu --set full` capture of
#   python bench.py --steps 1 --warmup 1 --no-cpu-baseline
# usage: profiles/make_profiles.sh <launches.csv> <step.ncu-rep>
set -e
cd "$(dirname "$0")/.."
L=$1; R=$2
cp "$L" profiles/r01_launches_bench_steps1.csv
python profiles/launch_summary.py profiles/r01_launches_bench_steps1.csv 1 > profiles/r01_launch_shares.txt
ncu -i "$R" --page raw --csv 2>/dev/null > /tmp/_raw.csv
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/_raw.csv')))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size']
with open('profiles/r01_ncu_raw_selected.csv', 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[idx[c]] for c in cols])
    for r in rows[2:]:
        if 'at::' not in r[idx['Kernel Name']]:
            w.writerow([r[idx[c]] for c in cols])
PY
(for k in "ap_kernel" "walk_kernel<1" "coop_radix_kernel" "post_kernel" "match_kernel"; do
    python profiles/ncu_summary.py "$R" "$k" 0 2>/dev/null | grep -v "^total inst\|^ *[0-9]* inst"; done) > profiles/r01_ncu_kernel_summaries.txt
python profiles/ncu_source_lines.py "$R" ap_kernel 22 2>/dev/null | head -24 > profiles/r01_ncu_ap_hot_lines.txt
python profiles/ncu_source_lines.py "$R" walk_kernel 22 "(bool)1" 2>/dev/null | grep -v "total inst 0 \|total inst 75360" > profiles/r01_ncu_walk_hot_lines.txt
python profiles/ncu_source_lines.py "$R" coop_radix 22 2>/dev/null | head -50 > profiles/r01_ncu_sort_hot_lines.txt
python profiles/ncu_source_lines.py "$R" post_kernel 16 2>/dev/null | head -40 > profiles/r01_ncu_post_hot_lines.txt
