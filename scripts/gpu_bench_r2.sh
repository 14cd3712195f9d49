#!/bin/bash
# round 2 bench: our arm (with sweep + cfg3) and the reference arm, the way the driver runs them
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench.json").read().strip().splitlines()[-1])
print("value %.0f tok/s, e2e %.0f, ms/step %.1f, frac %.3f, launch_ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["launch_ms"]))
for s in d["sweep"] or []:
    print(s if "error" in s else "%-10s mode %d: %.0f tok/s, %.1f us/step, launch %.1f ms, prefill %.2f ms (%d rows), mean KV %.0f, frac %.3f" % (
        s["workload"], s["decode_mode"], s["tokens_per_s"], s["us_per_decode_step"], s["launch_ms"], s["prefill_ms"], s["prefill_rows"], s["mean_kv_positions"], s["frac"]))
print("cfg3:", json.dumps(d["cfg3"])[:900])
print("cpu_baseline:", json.dumps(d["cpu_baseline"])[:600])
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 --cpu-budget 60 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_ref.json').read().strip().splitlines()[-1]); print('reference arm: %.1f tok/s, ms/step %.0f, kind %s, cores %d' % (d['value'], d['ms_per_step'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])); print(d['cpu_baseline']['sample'])"
