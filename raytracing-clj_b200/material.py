"""`material/lambertian`, `material/metal`, `material/dielectric` with the
reference's names and arities (src/material.clj:13, :21, :34); they return maps that
are merged into a body, as raytracing.clj:63-78 does with `merge`.  The parameters are
recorded as data instead of being closed over."""
from __future__ import annotations

from typing import Sequence

KIND = "material/kind"
ALBEDO = "material/albedo"
FUZZ = "material/fuzz"
IOR = "material/refraction-index"

LAMBERTIAN, METAL, DIELECTRIC = 0, 1, 2  # == RTCLJ_LAMBERTIAN/METAL/DIELECTRIC


def lambertian(albedo: Sequence[float]) -> dict:
    r, g, b = (float(c) for c in albedo)
    return {KIND: LAMBERTIAN, ALBEDO: (r, g, b), FUZZ: 0.0, IOR: 1.0}


def metal(albedo: Sequence[float], fuzz: float) -> dict:
    r, g, b = (float(c) for c in albedo)
    return {KIND: METAL, ALBEDO: (r, g, b), FUZZ: float(fuzz), IOR: 1.0}


def dielectric(refraction_index: float) -> dict:
    return {KIND: DIELECTRIC, ALBEDO: (1.0, 1.0, 1.0), FUZZ: 0.0, IOR: float(refraction_index)}
