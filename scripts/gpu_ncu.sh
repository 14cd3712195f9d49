#!/bin/bash
# ncu evidence for profiles/: (a) per-launch device times of the bench command, (b) --set full of the top kernels.
mkdir -p gpurun_out
KREG='regex:k_decode_cluster|k_decode_persistent|k_gemm_tc|k_prefill_attn|k_prefill_attn_tc|k_ln_rows|k_phase|k_embed_rows|k_bert|k_init_session|k_finalize|k_rows_stats'
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KREG" -c 1200 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# the bench workload's decode launch itself: B=32, 1000 steps, one launch
python scripts/profile_step.py --steps 1000 --tc 1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode_cluster -c 1 -o gpurun_out/prof_decode_b32 \
    python scripts/profile_step.py --steps 1000 --tc 1 > gpurun_out/ncu_decode.log 2>&1
echo "decode full rc=$?"; tail -2 gpurun_out/ncu_decode.log
ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc" -s 4 -c 4 -o gpurun_out/prof_gemm_b32 \
    python scripts/profile_step.py --steps 3 --tc 1 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_prefill_attn_tc" -s 2 -c 2 -o gpurun_out/prof_pattn_b32 \
    python scripts/profile_step.py --steps 3 --tc 1 > gpurun_out/ncu_pattn.log 2>&1
echo "prefill full rc=$?"
ls -la gpurun_out/*.ncu-rep
