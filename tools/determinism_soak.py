"""Twelve renders of the bench scene at 40 spp must hash identically: the atomic work queue hands units to
lanes in a different order every launch, and nothing may depend on it."""
import hashlib, sys
sys.path.insert(0, '.')
import raytracing_clj_b200 as R
from raytracing_clj_b200 import render, _abi
S, CAM = R.scenes, R.camera
world, cam = S.cover_hittables(7), CAM.main_camera(1920, 1080, **S.COVER_CAMERA)
hs = set()
for i in range(12):
    lin, rgb, st = render.render(world, cam, 40, 50, seed=1)
    hs.add((hashlib.sha256(lin.tobytes()).hexdigest(), hashlib.sha256(rgb.tobytes()).hexdigest(), st["segments"]))
print(len(hs), "distinct results in 12 renders of 1920x1080x40spp;", list(hs)[0][0][:16], list(hs)[0][2])
