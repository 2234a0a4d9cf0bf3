python bench.py --extras --steps 10 --warmup 3 > gpurun_out/bench_s3_extras.json 2> gpurun_out/bench_s3_extras.err; tail -c 600 gpurun_out/bench_s3_extras.err; python - <<PYEOF
import json
d=json.loads(open("gpurun_out/bench_s3_extras.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
for k,v in d["extra"]["widened"].items():
    if isinstance(v,dict): print(k, round(v["gpu_ms"],3), round(v["gpu_mcells_s"]), v["cpu_mcells_s"] and round(v["cpu_mcells_s"],1))
PYEOF
