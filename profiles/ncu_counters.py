"""Per-launch counters of the engine's kernels from an ncu CSV log of one bench step:

    ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,\
dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ctr_<workload>.csv \
        python bench.py --workload <workload> --steps 1 --warmup 1 --no-cpu-baseline
    python profiles/ncu_counters.py <workload> gpurun_out/ctr_<workload>.csv [more pairs ...]

Merges {workload: {kernel: {inst, lanes_per_inst, dram_bytes, ncu_us, launches_in_log}}} into
profiles/r02_kernel_counters.json, which bench.py's roofline block reads (instruction count of one launch of the
dominant kernel; the duration it is divided by is measured live with CUDA events).  For a kernel launched several times
in the log the LARGEST launch (by instructions) is taken — that is the timed step's launch; smaller ones belong to
parity checks on sub-ranges.
"""
import csv
import json
import os
import sys

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_kernel_counters.json")
KERNELS = ["walk_kernel", "ap_kernel", "match_kernel", "coop_radix_kernel", "post_kernel", "prep_kernel", "ens_sample_kernel",
           "finalize_kernel", "layout_kernel", "bucket_partition_kernel", "bucket_local_kernel", "bucket_local_cta_kernel"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1.0, "us": 1.0, "nsecond": 1e-3, "ns": 1e-3,
         "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}


def parse(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        key = (r[ix["ID"]], r[ix["Kernel Name"]])
        try:
            val = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        launches.setdefault(key, {})[r[ix["Metric Name"]]] = val * SCALE.get(r[ix["Metric Unit"]], 1.0)
    return launches


def summarise(launches):
    out = {}
    for k in KERNELS:
        mine = [(m, name) for (_, name), m in launches.items() if k in name and "membership" not in name]
        if k == "walk_kernel":      # the detection walk (two batches per warp, or DETS = true), not the label walk
            mine = [(m, name) for (_, name), m in launches.items()
                    if "walk2_kernel" in name or "walk_kernel<1" in name or "walk_kernel<(bool)1" in name]
        if not mine:
            continue
        m, name = max(mine, key=lambda t: t[0].get("smsp__inst_executed.sum", 0.0))
        out[k] = {"inst": m.get("smsp__inst_executed.sum"),
                  "lanes_per_inst": m.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                  "dram_bytes": m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0),
                  "ncu_us": m.get("gpu__time_duration.sum"), "launches_in_log": len(mine), "kernel": name[:80]}
    return out


def main(argv):
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for wl, path in zip(argv[0::2], argv[1::2]):
        data[wl] = summarise(parse(path))
        print(wl, json.dumps(data[wl], indent=1))
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
