"""`hittable/sphere` with the reference's name and arity (src/hittable.clj:7).

In the reference the centre and radius live only inside the `::hit-fn` closure
(hittable.clj:7-9); a drop-in host needs them as data, so this constructor records
them in the body map ("Trap")."""
from __future__ import annotations

from typing import Sequence

CENTER = "hittable/center"
RADIUS = "hittable/radius"


def sphere(center: Sequence[float], radius: float) -> dict:
    cx, cy, cz = (float(c) for c in center)
    return {CENTER: (cx, cy, cz), RADIUS: float(radius)}
