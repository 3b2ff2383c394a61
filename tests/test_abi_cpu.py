"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
keeps the reference's error behaviour for argument errors, and FAILS LOUDLY without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import coffeedb_b200 as cdb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    L = cdb.lib()
    header = open(os.path.join(ROOT, "include", "coffeedb_b200.h")).read()
    declared = set(re.findall(r"\b(cdb_[a-z_]+)\s*\(", header))
    declared -= {"cdb_build_device_"}  # no such thing; keeps the regex honest
    assert declared == set(cdb.EXPORTS), declared ^ set(cdb.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.cdb_version()


def test_no_torch_types_in_signatures():
    header = open(os.path.join(ROOT, "include", "coffeedb_b200.h")).read()
    assert "torch" not in header and "at::" not in header and "#include <cuda" not in header


def test_host_side_argument_errors():
    ix = cdb.StringIndex()
    ix.add(1, b"abc")
    ix.add(2, b"")
    with pytest.raises(RuntimeError, match="has not been built"):
        ix.locate_batch([b"a"])
    with pytest.raises(RuntimeError, match="has not been built"):
        ix.info()
    ix.close()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_build_fails_loudly_without_a_device():
    ix = cdb.StringIndex()
    ix.add(1, b"abc")
    with pytest.raises(RuntimeError, match="no CUDA device available: coffeedb_b200 has no CPU fallback"):
        ix.build()
    ix.close()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_filter_and_numeric_fail_loudly_without_a_device():
    """cdb_filter / cdb_numeric_create have no host path either: the set algebra of src/interface.cpp:79-147 is not
    re-implemented on the CPU behind the ABI."""
    with pytest.raises(RuntimeError, match="no CUDA device available: coffeedb_b200 has no CPU fallback"):
        cdb.NumericIndex(0, [1, 2], [3, 4])
    ix = cdb.StringIndex()
    ix.add(1, b"abc")
    with pytest.raises(RuntimeError, match="no CUDA device available"):
        cdb.filter_batch({"v": ix}, [{"constraints": {"v": "a"}}])
    ix.close()


def test_range_parsing_follows_the_reference():
    # src/utility.h:69-104: "(" / "]" set the id half of the bound pair to INT64_MAX; uint ranges become half-open
    assert cdb.parse_range("[1990, 2000)", 0) == (1990, 0, 2000, 0)
    assert cdb.parse_range("(5,7]", 0) == (5, cdb.INT64_MAX, 7, cdb.INT64_MAX)
    assert cdb.parse_range("[-inf,inf]", 0) == (-(1 << 63), 0, cdb.INT64_MAX, cdb.INT64_MAX)
    lo, _a, hi, _b = cdb.parse_range("[0.5,inf)", 1)
    import numpy as np
    assert np.array([lo, hi], np.int64).view(np.float64).tolist() == [0.5, np.finfo(np.float64).max]
    assert cdb.parse_uint_range("[0,32)") == (0, 32) and cdb.parse_uint_range("(3,3]") == (4, 4)
    with pytest.raises(ValueError):
        cdb.parse_uint_range("[5,2]")
    with pytest.raises(ValueError):
        cdb.parse_range("1..2", 0)


def test_splice_matches_reference_rendering(golden):
    # database.cpp:78-90 — host-side marker splicing, checked against the compiled reference's rendered strings
    import oracle
    from tests import cases
    rend, roff = golden["highlight/rendered"], golden["highlight/rendered_off"]
    for i, (kws, text) in enumerate(cases.highlight_cases()):
        spans = oracle.port.spans(kws, text)
        assert cdb.splice(text, spans, b"<b>", b"</b>") == rend[roff[i]:roff[i + 1]].tobytes()


def test_product_never_imports_the_oracle():
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "coffeedb_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "coffee_oracle" not in src and "libcoffeeref" not in src, f
