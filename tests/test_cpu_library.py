"""not gpu: the C-ABI library loads and exports exactly what include/kiez_b200.h declares
(no compute calls -- there is no GPU here), and the argument validation of the entry
points that can be exercised without a device."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kiez_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kb2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from kiez_b200 import _lib

    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(_lib.lib, name), f"{name} declared in the header but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\bT (kb2_[a-z0-9_]+)", out)))
    assert exported == names, "exported symbols and header declarations differ"


def test_ctypes_signatures_cover_header():
    from kiez_b200 import _lib

    declared = set(_declared()) - {"kb2_last_error"}
    assert declared == set(_lib.SIGNATURES)
    # argument counts agree with the header prototypes
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        proto = re.search(rf"\b{name}\s*\(([^;]*?)\)\s*;", text, flags=re.S).group(1).strip()
        n_args = 0 if proto in ("", "void") else proto.count(",") + 1
        assert n_args == len(args), f"{name}: header has {n_args} args, ctypes has {len(args)}"


def test_pure_host_entry_points():
    from kiez_b200 import _lib

    lib = _lib.lib
    assert lib.kb2_version() == 4
    assert lib.kb2_max_candidates() == 128
    assert [lib.kb2_padded_dim(d) for d in (1, 32, 33, 50, 256)] == [32, 32, 64, 64, 256]
    # plenty of query tiles: never split; few query tiles: split to fill 148 SMs
    assert lib.kb2_suggest_splits(1_000_000, 1_000_000, 16, 148) == 1
    s = lib.kb2_suggest_splits(15_000, 15_000, 56, 148)
    assert 2 <= s <= 14 and s * 56 <= 2048
    assert lib.kb2_suggest_splits(100, 50, 16, 148) == 1


def test_argument_validation_needs_no_device():
    """Bad arguments are rejected before anything is launched."""
    from kiez_b200 import _lib

    with pytest.raises(RuntimeError, match="dpad"):
        _lib.call("kb2_knn_candidates", 0, None, None, 10, None, None, None, 10, 33, 16, 1, None,
                  None, None)
    with pytest.raises(RuntimeError, match="cap"):
        _lib.call("kb2_knn_candidates", 0, None, None, 10, None, None, None, 10, 32, 500, 1, None,
                  None, None)
    with pytest.raises(RuntimeError, match="k=0"):
        _lib.call("kb2_topk_rows", None, None, 4, 8, 1, 0, 0, None, None, None)
    with pytest.raises(RuntimeError, match="elem_size"):
        _lib.call("kb2_refine_topk", None, 1, 4, None, 1, 4, 4, 2, None, None, None, 8, 0, 0, 0, 0,
                  1, None, None, None)
    assert isinstance(_lib.lib.kb2_last_error(), bytes)


def test_sass_is_blackwell_native():
    """The search kernel must be tcgen05 + TMA (UTC*MMA / LDTM / UTMALDG in SASS)."""
    from kiez_b200 import _lib

    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass
    assert "LDTM" in sass and "UTMALDG" in sass
    assert "HMMA.16816" not in sass          # no legacy mma.sync path


def test_integration_stub_calls_match_the_header_arity():
    """The reference-side binding documented in INTEGRATION.md section 2 passes as many
    arguments to each entry point as include/kiez_b200.h declares (the GPU suite executes it)."""
    from kiez_b200 import _lib

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = text[text.index("## 2. The minimal stub"):]
    code = re.search(r"```python\n(.*?)```", section, re.S).group(1)
    compile(code, "INTEGRATION.md#stub", "exec")
    for name in ("kb2_prepare_rows", "kb2_knn_candidates", "kb2_refine_topk"):
        call = code[code.index("_lib." + name + "("):]
        depth, n_args = 0, 1
        for ch in call[call.index("(") + 1:]:
            if ch in "([":
                depth += 1
            elif ch in ")]":
                if depth == 0:
                    break
                depth -= 1
            elif ch == "," and depth == 0:
                n_args += 1
        assert n_args == len(_lib.SIGNATURES[name]), name


def test_every_entry_point_cites_the_reference():
    """include/kiez_b200.h: the comment block in front of each prototype (group) names the
    reference code it replaces (file.py:lines), as the boundary contract asks."""
    text = open(HEADER).read()
    pattern = re.compile(
        r"(/\*(?:.|\n)*?\*/)\s*((?:(?:int|const char \*)\s*kb2_[a-z0-9_]+\([^;]*;\s*)+)")
    seen = set()
    for comment, protos in pattern.findall(text):
        names = re.findall(r"kb2_[a-z0-9_]+(?=\()", protos)
        seen.update(names)
        assert re.search(r"[a-z_]+\.py:\d", comment), f"{names}: no reference citation"
    assert seen >= set(_declared()) - {"kb2_version", "kb2_last_error"}
