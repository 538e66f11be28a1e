#!/bin/bash
# round 2: smoke, bench (both arms), ncu launch lists and full captures (summarised on the box); outputs under gpurun_out/
# usage: gpu_round2.sh [tests]   -- with "tests" the full -m gpu suite runs first
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
if [ "$1" = "tests" ]; then
  S=$(date +%s); timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest -m gpu exit $? in $(( $(date +%s) - S )) s"; tail -n 3 gpurun_out/r02_pytest_gpu.log
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r02_smoke.log
fi
S=$(date +%s); timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench exit $? in $(( $(date +%s) - S )) s"; tail -n 2 gpurun_out/r02_bench_1gpu.err
if [ "$2" = "ref" ]; then
  S=$(date +%s); timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "bench reference exit $? in $(( $(date +%s) - S )) s"
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --cpu-baseline 0 --extras 0 --also-dedup 0 > gpurun_out/r02_ncu_launch_run.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_c5.csv python scripts/one_config.py c5 3 > /dev/null 2>&1; echo "ncu c5 launches exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c4.csv python scripts/one_config.py c4 1 > /dev/null 2>&1; echo "ncu c4 launches exit $?"
timeout 900 ncu --set full --clock-control none -k regex:k_dualnet_tc -s 3 -c 2 -o gpurun_out/r02_prof_dualnet_tc -f python scripts/one_config.py c2 1 > gpurun_out/r02_ncu_dualnet_run.log 2>&1; echo "ncu dualnet exit $?"
timeout 900 ncu --set full --clock-control none -k regex:k_dualnet_tc -s 3 -c 1 -o gpurun_out/r02_prof_dualnet_tc_19x19 -f python scripts/one_config.py sh19 1 > gpurun_out/r02_ncu_dualnet19_run.log 2>&1; echo "ncu dualnet19 exit $?"
timeout 900 ncu --set full --clock-control none -k "regex:k_planes|k_backup|k_descend_sh|k_move_end|k_root_begin" -s 10 -c 8 -o gpurun_out/r02_prof_search_kernels -f python scripts/one_config.py c2 1 > gpurun_out/r02_ncu_search_run.log 2>&1; echo "ncu search exit $?"
timeout 900 ncu --set full --clock-control none -k "regex:k_wave_puct_blk|k_expand_leaves_blk|k_backup_blk|k_backup_priors_blk" -s 4 -c 8 -o gpurun_out/r02_prof_puct_block -f python scripts/one_config.py c5 1 > gpurun_out/r02_ncu_puct_run.log 2>&1; echo "ncu puct exit $?"
python scripts/ncu_box_summary.py r02
du -sh gpurun_out
