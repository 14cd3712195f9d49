#!/bin/bash
# per-kernel device times of ONE large-batch decode step on the tcgen05 graph path (mode 3, plain stream launches = mode 2 has the same
# kernels; here mode 3 itself under ncu: graph kernel nodes are profiled individually)
mkdir -p gpurun_out
python scripts/profile_step.py --batch 256 --steps 3 --mode 1 --tc 1 --tcmin 64 > gpurun_out/m3_plain.log 2>&1; tail -1 gpurun_out/m3_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"k_gemm_tc|k_phase|k_gather_x0" -c 400 --csv \
    --log-file gpurun_out/launches_mode3_b256.csv python scripts/profile_step.py --batch 256 --steps 3 --mode 1 --tc 1 --tcmin 64 > gpurun_out/m3_ncu.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_mode3_b256.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki][:60]; v = float(r[vi].replace(",", "")); 
    if r[ui] == "ns": v /= 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in agg.items(): print("%-62s n=%4d total %9.1f us  mean %7.2f us" % (k, a[0], a[1], a[1] / a[0]))
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
PY
