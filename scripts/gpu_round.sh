#!/bin/bash
# smoke + bench (both modes) + ncu launch list + ncu full captures; outputs under gpurun_out/
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_faithful.json 2> gpurun_out/bench_faithful.err; echo "bench exit $?"; cat gpurun_out/bench_faithful.json; tail -n 3 gpurun_out/bench_faithful.err
timeout 300 python bench.py --steps 6 --warmup 3 --dedup 1 --cpu-baseline 0 > gpurun_out/bench_dedup.json 2> gpurun_out/bench_dedup.err; cat gpurun_out/bench_dedup.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --also-19 0 --also-dedup 0 > gpurun_out/ncu_launch_run.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dualnet_tc -s 3 -c 2 -o gpurun_out/prof_dualnet -f python bench.py --games 1024 --steps 1 --warmup 3 --cpu-baseline 0 --also-19 0 --also-dedup 0 > gpurun_out/ncu_dualnet_run.log 2>&1; echo "ncu dualnet exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_planes|k_backup|k_descend_sh|k_move_end|k_root_begin" -s 10 -c 8 -o gpurun_out/prof_search -f python bench.py --games 4096 --steps 1 --warmup 3 --cpu-baseline 0 --also-19 0 --also-dedup 0 > gpurun_out/ncu_search_run.log 2>&1; echo "ncu search exit $?"
ls -la gpurun_out
