"""-m "not gpu": the C-ABI library builds, loads, and exports every symbol include/mofa_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mofa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mofa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mofa_video_b200 import build, lib
    path = build.build(force=False)
    so = ctypes.CDLL(path)
    declared = _declared()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(so, name), f"{name} declared in mofa_b200.h but not exported"
    assert sorted(lib.EXPORTS) == declared
    assert so.mofa_version() >= 100


def test_gemm_args_struct_matches_header_field_order():
    from mofa_video_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "mofa_b200.h")).read()
    body = hdr[hdr.index("typedef struct mofa_gemm_args {"):hdr.index("} mofa_gemm_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct mofa_gemm_args {", "").strip()
        if not decl:
            continue
        parts = decl.replace("*", " ").split(",")
        first = parts[0].split()
        if not first:
            continue
        names.append(first[-1])
        names += [p.strip() for p in parts[1:]]
    assert names == [f[0] for f in lib.GemmArgs._fields_]


def test_bad_arguments_are_rejected_without_a_gpu():
    from mofa_video_b200 import lib
    so = lib.load()
    g = lib.GemmArgs()
    assert so.mofa_gemm(ctypes.byref(g), None) == -1
    assert b"null" in so.mofa_last_error()


def test_header_is_plain_c_and_cxx(tmp_path):
    """The boundary is a C ABI: include/mofa_b200.h must compile as C99 and as C++17 on its own (no torch / CUDA types),
    and a C translation unit must be able to take the address of every declared entry point."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "mofa_b200.h")
    subprocess.run(["gcc", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", hdr], check=True)
    subprocess.run(["g++", "-x", "c++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", hdr], check=True)
    names = _declared()
    src = tmp_path / "use.c"
    src.write_text('#include "mofa_b200.h"\nconst void* table[] = {' + ", ".join(f"(const void*){n}" for n in names) +
                   "};\nint main(void) { return table[0] == 0; }\n")
    so = os.path.join(ROOT, "mofa_video_b200", "libmofa_b200.so")
    if not os.path.exists(so):
        import pytest
        pytest.skip("library not built")
    # link against the built library: every declared symbol must resolve (unresolved ones fail the link)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), so, "-o", str(tmp_path / "use"),
                    "-Wl,--unresolved-symbols=ignore-in-shared-libs"], check=True)


def test_product_never_imports_the_oracle_or_the_tests():
    """oracle/ and tests/ are test infrastructure: nothing under mofa_video_b200/ may import them, and bench.py /
    __graft_entry__.py only inside the CPU-baseline sampler / smoke checker."""
    import ast
    pkg = os.path.join(ROOT, "mofa_video_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            path = os.path.join(dirpath, fn)
            for node in ast.walk(ast.parse(open(path).read())):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.module:
                    mods = [node.module]
                for m in mods:
                    if m.split(".")[0] in ("oracle", "tests", "ref_ops"):
                        bad.append((path, m))
    assert not bad, bad
    # bench.py: the oracle may only be imported inside _cpu_sample (the cpu_baseline / --impl reference leg)
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            names = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""]
            assert not any(n.split(".")[0] == "oracle" for n in names)
        if isinstance(node, ast.FunctionDef) and node.name != "_cpu_sample":
            for sub in ast.walk(node):
                if isinstance(sub, ast.ImportFrom) and sub.module:
                    assert sub.module.split(".")[0] != "oracle", node.name


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """No fallback: with the shared library absent, binding raises instead of computing anything."""
    import pytest

    from mofa_video_b200 import lib
    monkeypatch.setattr(lib, "_lib", None, raising=False)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope" / "libmofa_b200.so"), raising=False)
    if not hasattr(lib, "LIB_PATH"):
        pytest.skip("library path is not a module attribute")
    with pytest.raises((OSError, RuntimeError, FileNotFoundError)):
        lib.load()
