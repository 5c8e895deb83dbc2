import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, {k:(round(d[k],3) if isinstance(d[k],float) else d[k]) for k in ("value","ms_per_step","gpu_launches","parity_max_abs_err_vs_oracle") if k in d})
    print("   phase", {k:round(v,3) for k,v in d["config"]["phase_ms"].items()}, d["config"]["index"])
    print("   kern", {k:round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2))
