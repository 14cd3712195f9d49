#!/bin/bash
# ncu evidence for the decode kernel only (launch list of the bench command + --set full of the decode launch)
mkdir -p gpurun_out
KREG='regex:k_decode_cluster|k_decode_persistent|k_gemm_tc|k_prefill_attn|k_prefill_attn_tc|k_ln_rows|k_phase|k_embed_rows|k_bert|k_init_session|k_finalize|k_rows_stats'
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KREG" -c 1200 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; grep -c k_decode_cluster gpurun_out/launches_bench.csv
python scripts/profile_step.py --steps 1000 --tc 1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode_cluster -c 1 -o gpurun_out/prof_decode_b32 \
    python scripts/profile_step.py --steps 1000 --tc 1 > gpurun_out/ncu_decode.log 2>&1
echo "decode full rc=$?"; tail -3 gpurun_out/ncu_decode.log; cat gpurun_out/prof_plain.log | tail -1
