#!/bin/bash
# round-2 (second half) ncu evidence: (a) per-launch device times of the bench command with the persistent prefill GEMM,
# (b) --set full of the four projections of one prefill layer at the bench row count (7,755 rows)
mkdir -p gpurun_out
KREG='regex:k_decode_cluster|k_decode_persistent|k_decode_wide|k_gemm_tc|k_prefill_attn|k_prefill_attn_tc|k_ln_rows|k_phase|k_embed_rows|k_bert|k_init_session|k_finalize|k_rows_stats|k_admit'
python bench.py --steps 1 --warmup 1 --no-cpu --no-sweep --no-cfg3 > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KREG" -c 1200 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-sweep --no-cfg3 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python scripts/profile_step.py --steps 3 --tc 1 > gpurun_out/prof_plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tcp --launch-skip 40 -c 4 -f -o gpurun_out/prof_gemm_tcp_b32 \
    python scripts/profile_step.py --steps 3 --tc 1 > gpurun_out/ncu_gemm_tcp_b32.log 2>&1
echo "gemm full rc=$?"; tail -2 gpurun_out/ncu_gemm_tcp_b32.log
ls -la gpurun_out/prof_gemm_tcp_b32.ncu-rep
