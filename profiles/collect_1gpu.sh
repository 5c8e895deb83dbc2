#!/bin/bash
# One-GPU evidence of a round, in one gpurun call (run from the repository root on the GPU box):
#   profiles/collect_1gpu.sh [tag = r02]
# 1. ncu counter pass per workload -> profiles/r02_kernel_counters.json (bench.py's roofline reads the instruction counts)
# 2. the bench line of every BASELINE config, with the CPU baseline and the parity block, and the reference arm
# 3. launch list and one `ncu --set full` capture of a coco5000 step
# Everything lands in gpurun_out/<tag>_*; copy what is to be kept into profiles/ (profiles/make_profiles.sh).
T=${1:-r02}
O=gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for wl in coco5000 voc4952 smoke500 coco5000_ori coco5000_dcsb; do
    ncu --metrics $M --clock-control none --csv --log-file $O/${T}_ctr_$wl.csv python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --no-graph > $O/${T}_ctr_$wl.log 2>&1
done
python profiles/ncu_counters.py coco5000 $O/${T}_ctr_coco5000.csv voc4952 $O/${T}_ctr_voc4952.csv smoke500 $O/${T}_ctr_smoke500.csv coco5000_ori $O/${T}_ctr_coco5000_ori.csv coco5000_dcsb $O/${T}_ctr_coco5000_dcsb.csv > $O/${T}_ctr_summary.txt 2>&1
cp profiles/r02_kernel_counters.json $O/${T}_kernel_counters.json
for wl in coco5000 smoke500 voc4952 coco5000_ori coco5000_dcsb; do
    python bench.py --workload $wl --steps 20 --warmup 5 > $O/${T}_bench_${wl}_1gpu.json 2> $O/${T}_bench_${wl}_1gpu.err
    python profiles/sumbench.py $O/${T}_bench_${wl}_1gpu.json
done
python bench.py --workload sweep50k --steps 3 --warmup 3 > $O/${T}_bench_sweep50k_1gpu.json 2> $O/${T}_bench_sweep50k_1gpu.err
python profiles/sumbench.py $O/${T}_bench_sweep50k_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_coco5000_reference_arm.json 2> $O/${T}_bench_coco5000_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $O/${T}_launches.log 2>&1
ncu --set full --import-source on --clock-control none -c 28 -o $O/${T}_step python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph > $O/${T}_step.log 2>&1
ls -la $O/${T}_*
