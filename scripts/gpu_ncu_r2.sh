#!/bin/bash
# round-2 ncu evidence for profiles/: (a) per-launch device times of the bench command (sweep / cfg3 / cpu legs off), (b) --set full of the decode launch
mkdir -p gpurun_out
KREG='regex:k_decode_cluster|k_decode_persistent|k_decode_wide|k_gemm_tc|k_prefill_attn|k_prefill_attn_tc|k_ln_rows|k_phase|k_embed_rows|k_bert|k_init_session|k_finalize|k_rows_stats|k_admit'
python bench.py --steps 1 --warmup 1 --no-cpu --no-sweep --no-cfg3 > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KREG" -c 1200 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-sweep --no-cfg3 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python scripts/profile_step.py --steps 1000 --tc 1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode_cluster -c 1 -o gpurun_out/prof_decode_b32 \
    python scripts/profile_step.py --steps 1000 --tc 1 > gpurun_out/ncu_decode.log 2>&1
echo "decode full rc=$?"; tail -2 gpurun_out/ncu_decode.log
python scripts/profile_step.py --batch 1 --steps 300 --mode 6 --tc 1 > gpurun_out/prof_plain_wide.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode_wide -c 1 -o gpurun_out/prof_wide_b1 \
    python scripts/profile_step.py --batch 1 --steps 300 --mode 6 --tc 1 > gpurun_out/ncu_wide.log 2>&1
echo "wide full rc=$?"
ls -la gpurun_out/*.ncu-rep
