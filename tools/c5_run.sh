#!/bin/bash
# usage (on the GPU box): tools/c5_run.sh <tag>  -- multi-query test, C5 probe (32 queries), kernel time from an ncu launch list
tag=$1
timeout 300 python -m pytest tests/test_gpu_screen.py -x -q -m gpu -k "multi_query" > gpurun_out/t_multi_$tag.log 2>&1; tail -2 gpurun_out/t_multi_$tag.log
timeout 120 python tools/c5_probe.py 32 > gpurun_out/c5_probe_$tag.log 2>&1; tail -1 gpurun_out/c5_probe_$tag.log
C5_REP=1 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:score_screen_multi -c 4 --csv --log-file gpurun_out/c5_k_$tag.csv python tools/c5_probe.py 16 > /dev/null 2>&1
tail -2 gpurun_out/c5_k_$tag.csv | awk -F'","' '{print $(NF-2), $NF}'
