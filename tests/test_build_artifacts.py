"""Properties of the built sm_100a code that the design depends on and that a compiler or source change
can silently break (checked on the CPU box with cuobjdump)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "raytracing-clj_b200", "librtclj_b200.so")
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", line):
            funcs[name].append(line)
    return funcs


def kernels(funcs, key):
    """Every instantiation whose mangled name contains `key` (the strict-order variants are separate kernels)."""
    names = [n for n in funcs if key in n]
    assert names, key
    return [funcs[n] for n in names]


def test_only_sm_100a_code_is_shipped():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], check=True, capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("key", ["render_wave_kernel", "render_kernelILb1", "render_lane2_kernel", "render_split_kernel"])
def test_small_scene_cull_takes_its_table_through_uniform_registers(sass, key):
    """<= 512 spheres: the cull table is a kernel parameter and must reach FFMA2 as UNIFORM operands
    (LDCU.64 UR, c[0x0][UR+..] -> FFMA2 R, R.F32, UR.F32x2, ..).  ptxas silently falls back to per-lane LDC
    when it cannot prove the warp converged (e.g. after a spin-wait without __syncwarp), which costs ~12 %."""
    for body in kernels(sass, key):
        uniform = [l for l in body if re.search(r"LDCU\.64 UR\d+, c\[0x0\]\[UR\d+", l)]
        per_lane = [l for l in body if re.search(r"LDC\.64 R\d+, c\[0x0\]\[R\d+", l)]
        assert len(uniform) >= 32 and not per_lane, (len(uniform), len(per_lane))
        assert sum("FFMA2" in l and ".F32x2" in l and " UR" in l for l in body) >= 48


def test_large_scene_table_is_staged_by_tma(sass):
    for body in kernels(sass, "render_kernelILb0"):
        assert any("UBLKCP" in l for l in body) and any("SYNCS" in l for l in body)


def test_no_module_global_constant_table():
    """The round-1 __constant__ cull table (shared by every context of a device) is gone."""
    out = subprocess.run(["cuobjdump", "-elf", LIB], check=True, capture_output=True, text=True).stdout
    assert "g_ctab" not in out
    assert not re.search(r"\.nv\.constant3", out)


def test_wave_kernel_fits_the_instruction_cache_budget(sass):
    """59 KB of SASS stalled 5.5 warps per issue on instruction fetch (profiles/r2_wave_v1_*): keep the
    kernel's resident body under 32 KB and the whole function under 48 KB."""
    body = kernels(sass, "render_wave_kernel")[0]
    first_exit = next(i for i, l in enumerate(body) if re.search(r"\bEXIT\b", l))
    assert len(body) * 16 <= 48 * 1024, len(body) * 16
    assert first_exit * 16 <= 32 * 1024, first_exit * 16


def test_lane_kernels_fit_the_instruction_cache(sass):
    """The chunked-mode instantiations of the two lane kernels: at 35 KB the two-paths-per-lane kernel stalled
    0.4 warps per issue on instruction fetch, at 33 KB 0.24 (ncu, profiles/r2_render_lane2_*)."""
    for key in ("render_kernelILb1ELb0ELb0", "render_kernelILb1ELb0ELb1", "render_lane2_kernelILb0"):
        body = kernels(sass, key)[0]
        assert len(body) * 16 <= 34 * 1024, (key, len(body) * 16)



def test_primary_ray_kernel_reads_no_memory_in_its_loop(sass):
    """render_primary_kernel: the sphere constants come from the kernel parameters, the sums
    live in registers -- no global / shared / local loads anywhere, no spills, and a body that fits the
    instruction cache."""
    # (the scan is unrolled over six sphere slots; a scene executes only the slots it has)
    for key, budget in (("render_primary_kernelILb0", 40), ("render_primary_kernelILb1", 44)):  # without / with the defocus disk
        body = kernels(sass, key)[0]
        text = "\n".join(body)
        assert not re.search(r"\b(LDG|LDS|LDL|STL|STS)\b", text), key
        assert len(re.findall(r"\bSTG\b", text)) == 3          # one unit sum: three doubles
        assert re.search(r"\bDFMA\b", text) and re.search(r"c\[0x0\]\[", text)
        assert len(body) * 16 <= budget * 1024, (key, len(body) * 16)
