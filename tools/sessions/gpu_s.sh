#!/bin/bash
# one path per lane: the round-1 kernel (render_kernel<true>) against the same loop built on path_step()
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python - > gpurun_out/s_parity.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import oracle_lib as O
import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render
S, CAM = R.scenes, R.camera
ok = True
for world, cam, spp, depth, flags, unit in ((S.cover_hittables(7), CAM.main_camera(160, 90, **S.COVER_CAMERA), 8, 50, O.FLAGS_MAIN, 0),
                                            (S.realm_hittables(), CAM.realm_camera(200), 32, 50, O.FLAGS_REALM, 7),
                                            (S.main_hittables(), CAM.main_camera(200), 70, 50, O.FLAGS_MAIN, 70),
                                            (S.i_hittables(), CAM.i_camera(200), 16, 50, O.FLAGS_I, 0),
                                            ([], CAM.realm_camera(32), 4, 50, O.FLAGS_REALM, 0)):
    soa = S.to_soa(world)
    u = unit if unit else spp
    lo, ro, so = O.render(soa, cam, spp, depth, seed=3, flags=flags, threads=8, samples_per_unit=u)
    lg, rg, sg = render.render(soa, cam, spp, depth, seed=3, flags=flags | (1 << 22), samples_per_unit=u)
    good = bool(np.array_equal(lo, lg) and np.array_equal(ro, rg) and so.segments == sg["segments"])
    ok = ok and good
    print(len(world), spp, good, flush=True)
print("PARITY", ok)
PY
tail -n 2 gpurun_out/s_parity.log
for k in lane lane1p lane lane1p; do RTCLJ_QP_KERNEL=$k timeout 200 python tools/quick_perf.py 2>/dev/null | grep -o '"case": "[^"]*"\|"ms": [0-9.]*' | paste - - | tr '\n' ' '; echo; done
for k in lane lane1p; do timeout 300 python bench.py --kernel $k --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$k', round(d['value']/1e9,4), round(d['ms_per_step'],2), round(d['roofline']['frac'],4))"; done
for w in c1 c2 c4; do for k in lane lane1p; do timeout 300 python bench.py --workload $w --kernel $k --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w $k', round(d['value']/1e9,4), round(d['ms_per_step'],3))"; done; done
