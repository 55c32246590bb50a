#!/bin/bash
# Round-2 profiling pass (one B200, under gpurun): launch list of the bench command + ncu --set full captures of the
# dominant kernels.  Outputs land in gpurun_out/; scripts/summarize_launches.py and scripts/ncu_summary.py turn them into
# the tables committed under profiles/.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02b.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --min-seconds 0 > gpurun_out/bench_under_ncu_r02b.json 2> gpurun_out/bench_under_ncu_r02b.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 988 -c 494 --csv --log-file gpurun_out/launches_r02b_lane1.csv \
    python scripts/profile_step.py 3 > gpurun_out/profile_step_r02b.log 2>&1
# one encoder layer's four GEMMs (step 2, layer ~25), the fused attention + FSMN kernel, LayerNorm, the front-end
PFASR_GEMM_POLICY=throughput ncu --set full --clock-control none --import-source on -k regex:pf_gemm -s 280 -c 4 -f -o gpurun_out/prof_r02b_gemm \
    python scripts/profile_step.py 2 > gpurun_out/ncu_gemm_r02b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pf_sanm_attention_tc -s 70 -c 1 -f -o gpurun_out/prof_r02b_attn \
    python scripts/profile_step.py 2 > gpurun_out/ncu_attn_r02b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pf_frontend_fbank|pf_layernorm|pf_logsoftmax" -s 4 -c 4 -f -o gpurun_out/prof_r02b_small \
    python scripts/profile_step.py 2 > gpurun_out/ncu_small_r02b.log 2>&1
ls -la gpurun_out/*.ncu-rep
