#!/bin/bash
# Profiling pass of one round on the GPU box (run through gpurun from the repo root); profiles/make_summary.py turns the
# artefacts it leaves in gpurun_out/ into the tracked files under profiles/.   usage: bash profiles/collect.sh r01
R=${1:-r01}
O=gpurun_out
mkdir -p $O
# 1. the bench line the profiles belong to (not under a profiler), plus the default invocation with the CPU baseline
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_final.json 2> $O/bench_final.err || exit 1
python bench.py > $O/bench_default.json 2> $O/bench_default.err
# 2. launch list of the bench command (durations are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tb_ -c 3000 --csv --log-file $O/${R}_launches_hg38.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${R}_launches_bench.log 2>&1
# 3. one full-set capture per kernel, chr50mb workload (a: everything but the fit rounds, b: fit rounds, c: MINPACK-exact fit rounds)
ncu --set full --clock-control none --import-source on -k 'regex:^(void )?tb_(pileup_kernel|k2_scatter|k2_weights|k2_hist_smem|k2_fold|score_dwm2|correct_kernel|footprint_kernel)' -c 10 \
    -f -o $O/${R}_full_a python bench.py --workload chr50mb --steps 1 --warmup 0 --no-cpu-baseline > $O/${R}_full_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tb_fitg_round_kernel -c 2 \
    -f -o $O/${R}_full_b python bench.py --workload chr50mb --steps 1 --warmup 0 --no-cpu-baseline > $O/${R}_full_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tb_fitg_round_kernel -c 2 \
    -f -o $O/${R}_full_c python bench.py --workload chr50mb --steps 1 --warmup 0 --no-cpu-baseline --lm-exact > $O/${R}_full_c.log 2>&1
# 4. the dominant kernel at the bench's own size (DRAM traffic per launch for the bench line's roofline.traffic)
ncu --set full --clock-control none --import-source on -k regex:tb_score_dwm2 -c 1 \
    -f -o $O/${R}_full_hg38 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/${R}_full_hg38.log 2>&1
ls -la $O/${R}_full_*.ncu-rep $O/${R}_launches_hg38.csv
