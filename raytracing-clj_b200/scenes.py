"""Hittable lists: the reference's literals and the synthetic scenes BASELINE.json
names.  A scene is a list of body maps in hittable-list ORDER (the order is part of
the semantics: the first body wins an exact tie)."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

from . import hittable, material

_MASK = (1 << 64) - 1


class SplitMix64:
    """Tiny seeded generator for SCENE LAYOUT only (never for rendering).  Restated
    natively in csrc/host.cpp so both sides build identical scenes."""

    def __init__(self, seed: int):
        self.s = seed & _MASK

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)

    def uniform(self, lo: float = 0.0, hi: float = 1.0) -> float:
        u = (self.next_u64() >> 11) * (1.0 / 9007199254740992.0)
        return lo + (hi - lo) * u


def body(geom: dict, mat: dict) -> dict:
    """`(merge (hittable/sphere ...) (material/... ...))`, raytracing.clj:65-66."""
    return {**geom, **mat}


def main_hittables() -> List[dict]:
    """raytracing.clj:63-78: ground, center, left, bubble, right."""
    return [
        body(hittable.sphere((0.0, -100.5, -1.0), 100.0), material.lambertian((0.8, 0.8, 0.0))),
        body(hittable.sphere((0.0, 0.0, -1.2), 0.5), material.lambertian((0.1, 0.2, 0.5))),
        body(hittable.sphere((-1.0, 0.0, -1.0), 0.5), material.dielectric(1.5)),
        body(hittable.sphere((-1.0, 0.0, -1.0), 0.4), material.dielectric(1.00 / 1.5)),
        body(hittable.sphere((1.0, 0.0, -1.0), 0.5), material.metal((0.8, 0.6, 0.2), 1.0)),
    ]


def realm_hittables() -> List[dict]:
    """realm/raytracing.clj:292-301: center, ground, left, bubble, right."""
    m = main_hittables()
    return [m[1], m[0], m[2], m[3], m[4]]


def i_hittables() -> List[dict]:
    """experimental/raytracing_i.clj:118-124: two spheres, shaded by their normal
    (the material is never consulted in that variant)."""
    return [
        body(hittable.sphere((0.0, 0.0, -1.0), 0.5), material.lambertian((0.5, 0.5, 0.5))),
        body(hittable.sphere((0.0, -100.5, -1.0), 100.0), material.lambertian((0.5, 0.5, 0.5))),
    ]


def _random_field(seed: int, lo: int, hi: int) -> List[dict]:
    """The book's random-sphere field (RTIOW final scene; the reference follows the
    book, raytracing.clj:15) over the integer grid [lo,hi)^2.  SURVEY.md Appendix D."""
    rng = SplitMix64(seed)
    out = [body(hittable.sphere((0.0, -1000.0, 0.0), 1000.0), material.lambertian((0.5, 0.5, 0.5)))]
    for a in range(lo, hi):
        for b in range(lo, hi):
            choose = rng.uniform()
            cx = a + 0.9 * rng.uniform()
            cz = b + 0.9 * rng.uniform()
            # every candidate consumes the same number of draws, kept or not
            d = [rng.uniform() for _ in range(7)]
            dx, dz = cx - 4.0, cz - 0.0
            if math.sqrt(dx * dx + dz * dz) <= 0.9:
                continue
            geom = hittable.sphere((cx, 0.2, cz), 0.2)
            if choose < 0.8:
                mat = material.lambertian((d[0] * d[1], d[2] * d[3], d[4] * d[5]))
            elif choose < 0.95:
                mat = material.metal((0.5 + 0.5 * d[0], 0.5 + 0.5 * d[1], 0.5 + 0.5 * d[2]), 0.5 * d[3])
            else:
                mat = material.dielectric(1.5)
            out.append(body(geom, mat))
    out.append(body(hittable.sphere((0.0, 1.0, 0.0), 1.0), material.dielectric(1.5)))
    out.append(body(hittable.sphere((-4.0, 1.0, 0.0), 1.0), material.lambertian((0.4, 0.2, 0.1))))
    out.append(body(hittable.sphere((4.0, 1.0, 0.0), 1.0), material.metal((0.7, 0.6, 0.5), 0.0)))
    return out


def cover_hittables(seed: int = 7) -> List[dict]:
    """BASELINE.json config 3: ~488 spheres on the grid [-11,11)^2."""
    return _random_field(seed, -11, 11)


def field_hittables(seed: int = 7, half: int = 50) -> List[dict]:
    """BASELINE.json config 5: ~10 000 spheres on [-50,50)^2, brute force, no BVH."""
    return _random_field(seed, -half, half)


COVER_CAMERA = dict(vfov=20.0, look_from=(13.0, 2.0, 3.0), look_at=(0.0, 0.0, 0.0),
                    vup=(0.0, 1.0, 0.0), defocus_angle=0.6, focus_dist=10.0)
FIELD_CAMERA = dict(vfov=30.0, look_from=(52.0, 14.0, 12.0), look_at=(0.0, 0.0, 0.0),
                    vup=(0.0, 1.0, 0.0), defocus_angle=0.6, focus_dist=40.0)


def to_soa(bodies: List[dict]) -> Tuple[np.ndarray, ...]:
    """Body maps -> the structure of arrays that crosses the C ABI (rtclj_scene)."""
    n = len(bodies)
    center = np.array([b[hittable.CENTER] for b in bodies], dtype=np.float64).reshape(n, 3)
    radius = np.array([b[hittable.RADIUS] for b in bodies], dtype=np.float64).reshape(n)
    kind = np.array([b[material.KIND] for b in bodies], dtype=np.int32).reshape(n)
    albedo = np.array([b[material.ALBEDO] for b in bodies], dtype=np.float64).reshape(n, 3)
    fuzz = np.array([b[material.FUZZ] for b in bodies], dtype=np.float64).reshape(n)
    ior = np.array([b[material.IOR] for b in bodies], dtype=np.float64).reshape(n)
    return tuple(np.ascontiguousarray(a) for a in (center, radius, kind, albedo, fuzz, ior))
