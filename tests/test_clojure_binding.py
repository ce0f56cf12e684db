"""The Clojure/Panama shim (integration/clojure/src/rtclj/native.clj) cannot run here (no JVM), so it is
pinned to the C header instead: every struct size and byte offset written in its `layouts` table -- and in
the copy INTEGRATION.md prints -- must equal ctypes' sizeof / offsetof for the structs of
include/rtclj_b200.h, every function it binds must be declared by the header with the argument count the
shim's FunctionDescriptor has, and the flag constants must equal the header's."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLJ = os.path.join(ROOT, "integration", "clojure", "src", "rtclj", "native.clj")
from raytracing_clj_b200 import _abi  # noqa: E402

STRUCTS = {"rtclj_scene": _abi.Scene, "rtclj_camera": _abi.Camera, "rtclj_params": _abi.Params,
           "rtclj_stats": _abi.Stats}


def parse_layouts(text):
    block = text[text.index("BEGIN-LAYOUTS"):text.index("END-LAYOUTS")]
    out = {}
    for m in re.finditer(r":(rtclj_\w+)\s*\{:size\s+(\d+)\s*:fields\s*\{([^}]*)\}", block):
        fields = {k: int(v) for k, v in re.findall(r":(\w+)\s+(\d+)", m.group(3))}
        out[m.group(1)] = (int(m.group(2)), fields)
    return out


@pytest.mark.parametrize("path", [CLJ, os.path.join(ROOT, "INTEGRATION.md")])
def test_struct_layouts_match_the_header(path):
    layouts = parse_layouts(open(path).read())
    assert set(layouts) == set(STRUCTS), "every struct of the header is described"
    for name, (size, fields) in layouts.items():
        st = STRUCTS[name]
        assert size == C.sizeof(st), name
        real = {f: getattr(st, f).offset for f, _ in st._fields_ if not f.startswith("_")}
        assert fields == real, f"{name}: shim {fields} != header {real}"


def test_ctypes_structs_match_the_header_text():
    """...and the ctypes structs themselves follow the header: same field names, in order."""
    hdr = open(os.path.join(ROOT, "include", "rtclj_b200.h")).read()
    for name, st in STRUCTS.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [re.search(r"(\w+)(\[\d+\])?$", d.strip()).group(1) for d in body.split(";") if d.strip()]
        assert names == [f for f, _ in st._fields_], name


def test_bound_functions_exist_with_the_declared_arity():
    text = open(CLJ).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rtclj_b200.h")).read(), flags=re.S)
    bound = re.findall(r'\(downcall "(\w+)"', text)
    assert {"rtclj_render", "rtclj_render_multi", "rtclj_render_multi_ppm", "rtclj_host_alloc", "rtclj_host_free", "rtclj_last_error",
            "rtclj_encode_ppm_p3_gpu"} <= set(bound)
    for sym in bound:
        decl = re.search(r"\b%s\s*\(([^)]*)\)\s*;" % sym, hdr)
        assert decl, f"{sym} is not declared in the header"
        args = [a for a in decl.group(1).split(",") if a.strip() and a.strip() != "void"]
        assert len(args) == len(_abi.SYMBOLS[sym][1]), sym
        # the shim's descriptor: the balanced form that starts at (downcall "sym"
        start = text.index('(downcall "%s"' % sym)
        depth, end = 0, start
        for end in range(start, len(text)):
            depth += {"(": 1, ")": -1}.get(text[end], 0)
            if depth == 0:
                break
        form = text[start:end + 1]
        if "repeat 6" in form:
            n = 6
        elif "make-array MemoryLayout 0" in form:
            n = 0
        else:
            n = len(re.findall(r"ValueLayout/(?:ADDRESS|JAVA_INT|JAVA_LONG)", form.split("fd-int", 1)[1]))
        assert n == len(args), f"{sym}: shim passes {n} arguments, header declares {len(args)}"


def test_flag_constants():
    text = open(CLJ).read()
    assert "(def flags-main  (bit-or 1 2 4 8))" in text and _abi.FLAGS_MAIN == 15
    assert "(def flags-realm 0)" in text and _abi.FLAGS_REALM == 0
    assert "(def flags-i     (bit-or 16 32))" in text and _abi.FLAGS_I == 48
