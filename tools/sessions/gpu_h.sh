#!/bin/bash
# round-2 GPU session H: the record -- ncu of the default kernel on the full workload, ncu of the TMA kernel on
# config 5, launch list, all five configs, the non-headline bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/h_plain.json 2> gpurun_out/h_plain.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/h_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_lane2 -s 3 -c 1 -f -o gpurun_out/h_lane2_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/h_ncu_lane2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -f -o gpurun_out/h_smemtab_c5 python bench.py --workload c5 --spp 4 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/h_ncu_c5.log 2>&1
for w in c1 c2 c4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/h_bench_$w.json 2> gpurun_out/h_bench_$w.err
done
timeout 900 python bench.py --workload c5 --steps 2 --warmup 3 --no-extras > gpurun_out/h_bench_c5.json 2> gpurun_out/h_bench_c5.err
timeout 1500 python tools/report_configs.py > gpurun_out/h_configs.log 2>&1
