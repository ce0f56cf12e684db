#!/bin/bash
# the record of the final build: ncu of the bench kernel (full workload) and of the primary-ray kernel (config 4),
# launch list, all five configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ag_plain.json 2> gpurun_out/ag_plain.err || { echo 'bench failed'; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ag_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ag_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_lane2 -s 3 -c 1 -f -o gpurun_out/ag_lane2_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ag_ncu_lane2.log 2>&1
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ag_plain_c4.json 2> gpurun_out/ag_plain_c4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_primary -s 3 -c 1 -f -o gpurun_out/ag_primary_c4 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ag_ncu_c4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/ag_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ag_ncu_launch_c4.log 2>&1
timeout 1500 python tools/report_configs.py > gpurun_out/ag_configs.log 2>&1
tail -n 5 gpurun_out/ag_configs.log
ls -la gpurun_out/ag_*
