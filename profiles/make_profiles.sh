#!/bin/bash
# Regenerate the committed profile summaries of a round from one launch list and one `ncu --set full` capture of
#   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph
# usage: profiles/make_profiles.sh <launches.csv> <step.ncu-rep> [tag = r02]
set -e
cd "$(dirname "$0")/.."
L=$1; R=$2; T=${3:-r02}
cp "$L" profiles/${T}_launches_bench_steps1.csv
python profiles/launch_list.py profiles/${T}_launches_bench_steps1.csv > profiles/${T}_launch_shares.txt
ncu -i "$R" --page raw --csv 2>/dev/null > /tmp/_raw.csv
python - "$T" <<'PY'
import csv, sys
tag = sys.argv[1]
rows = list(csv.reader(open('/tmp/_raw.csv')))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size']
with open(f'profiles/{tag}_ncu_raw_selected.csv', 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[idx[c]] for c in cols])
    for r in rows[2:]:
        if 'at::' not in r[idx['Kernel Name']]:
            w.writerow([r[idx[c]] for c in cols])
PY
(for k in "ap_kernel" "walk2_kernel" "walk_kernel" "bucket_partition" "bucket_local_kernel" "coop_radix_kernel" "post_kernel" "match_kernel" "prep_kernel"; do
    python profiles/ncu_summary.py "$R" "$k" 0 2>/dev/null | grep -v "^total inst\|^ *[0-9]* inst" | awk '/^----/{n++} n<=1'; done) > profiles/${T}_ncu_kernel_summaries.txt
python profiles/ncu_source_lines.py "$R" ap_kernel 30 2>/dev/null | head -32 > profiles/${T}_ncu_ap_hot_lines.txt
python profiles/ncu_source_lines.py "$R" walk2_kernel 30 2>/dev/null | head -32 > profiles/${T}_ncu_walk_hot_lines.txt
python profiles/ncu_source_lines.py "$R" post_kernel 16 2>/dev/null | head -40 > profiles/${T}_ncu_post_hot_lines.txt
python profiles/ncu_source_lines.py "$R" bucket_partition 16 2>/dev/null | head -40 > profiles/${T}_ncu_sort_hot_lines.txt
