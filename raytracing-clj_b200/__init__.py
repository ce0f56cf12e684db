"""raytracing-clj_b200: a B200-native (sm_100a) backend for the per-pixel render loop of
keychera/raytracing-clj, behind a C ABI (include/rtclj_b200.h).

Host-side modules mirror the reference's namespaces for this path:
  hittable.sphere / material.lambertian|metal|dielectric   (src/hittable.clj, src/material.clj)
  camera.main_camera / realm_camera / i_camera             (the -main let-blocks)
  scenes.*                                                 (the hittable lists)
  render.*                                                 (the render loop -> C ABI -> CUDA)
There is no CPU fallback: render.* raises if the CUDA library or a GPU is missing."""
from . import camera, hittable, material, scenes  # noqa: F401

__all__ = ["camera", "hittable", "material", "scenes", "render", "main"]


def __getattr__(name):  # render / main load the CUDA library lazily (and loudly)
    if name in ("render", "main", "_abi"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
