import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","gpu_launches","n_gpus")}, "frac %.4f kernel_ms %.4f call_ms %.4f"%(l["roofline"]["frac"], l["roofline"]["kernel_ms"], l["roofline"]["call_ms"]))
print("e2e", l["e2e"]); print("clocks", l["clocks"])
for e in l["extra_workloads"]: print({k:e.get(k) for k in ("workload","value","ms_per_step","call_ms","kernel_ms","ctas_per_sample","kernel_config","grid","relation_table_ms","error") if e.get(k) is not None})
