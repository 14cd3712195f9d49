"""Turns gpurun_out ncu artefacts into the tracked summaries under profiles/."""
import collections, csv, io, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = []
# (a) launch list of the bench command
fn = os.path.join(ROOT, "gpurun_out", "launches_bench.csv")
if os.path.exists(fn):
    lines = [l for l in open(fn) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        n = r["Kernel Name"].split("(")[0].replace("void ", "")
        agg.setdefault(n, []).append(float(r["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out.append(f"## Launch list (ncu --metrics gpu__time_duration.sum --clock-control none) of `python bench.py --steps 1 --warmup 1 --no-cpu`\n")
    out.append(f"{len(rows)} launches of this library's kernels (weight-packing kernels at engine creation excluded by the kernel-name filter); "
               f"per-launch times are cold-cache and serialised: compare SHARES.\n")
    out.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in agg.items():
        out.append(f"| `{k}` | {len(v)} | {sum(v)/1e6:.3f} | {100*sum(v)/tot:.1f}% |")
    out.append("")
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches_bench.csv"), "w") as f:
        f.write("kernel,launches,total_ns\n")
        for k, v in agg.items():
            f.write(f"{k},{len(v)},{sum(v):.0f}\n")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic"]
import json
for rep, title in (("prof_decode_b32", "k_decode_cluster: the bench workload's decode launch (B=32, 1000 steps in ONE launch of 7 clusters x 16 CTAs, mean KV 743) — `ncu --set full --clock-control none`"),
                   ("prof_gemm_b32", "k_gemm_tc<128> (tcgen05/TMEM + TMA prefill projections), B=32 (7755 rows) — `ncu --set full`"),
                   ("prof_gemm_tcp_b32", "k_gemm_tcp<256> (round 2: persistent 128 x 256-tile tcgen05 GEMM), four consecutive projections of the bench prefill (launches 40-43 of the call: linear2 of layer 9, then QKV, out_proj, linear1 of layer 10) at B=32 (7755 rows) — `ncu --set full`"),
                   ("prof_pattn_b32", "k_prefill_attn_tc (prefix-LM flash attention on mma.sync), B=32 — `ncu --set full`"),
                   ("prof_wide_b1", "k_decode_wide (round 2, mode 6: 144 CTAs, TMA weight ring, L2 {value, tag} hand-offs), batch 1, 300 steps in one launch — `ncu --set full`")):
    path = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out.append(f"## {title}\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        out.append(f"### `{name}` (id {r[0]})\n\n| metric | value | unit |\n|---|---:|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"| {w} | {r[i]} | {units[i]} |")
        out.append("")
        if rep == "prof_decode_b32":
            def val(nm):
                i = hdr.index(nm); x = float(r[i].replace(",", "")); u = units[i].lower()
                return x * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1, "tbyte": 1e12}.get(u, 1)
            tr = {"kernel": name, "workload": "cfg2_b32", "decode_steps": 1000, "decode_mode": 4 if "cluster" in name else 1,
                  "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                  "note": "one ncu --set full capture of the whole 1000-step persistent launch (profiler-time duration is not a bench number)"}
            json.dump(tr, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:6000])
