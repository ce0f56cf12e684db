#!/bin/bash
# ncu --set full of the wavefront kernel on a 5-sphere and on the cover scene (reduced spp)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for c in default_1920x1080x16 cover_1920x1080x16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 2 -c 1 -f -o gpurun_out/b_wave_$c python tools/quick_perf.py $c 3 > gpurun_out/b_ncu_$c.log 2>&1
done
