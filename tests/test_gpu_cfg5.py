"""BASELINE configs[4] shape at >= 1 GiB on one GPU, against the COMPILED REFERENCE built on this box's host cores:
valid UTF-8 mixing 1-, 2- and 3-byte code points (bytes on both sides of 0x80 -> the note-N1 layout with several
signed-radix levels), documents of up to 64 KB, 64-bit suffix-array elements (bits1 + bits2 > 32).

 * string_index::build() of the unmodified reference (oracle.Ref, src/index.cpp:178-236) builds the same corpus on the
   host; the device array must equal it element by element up to ties between byte-identical suffixes (SURVEY.md note
   N2): cdb_compare_sa reports 0 elements that name different suffixes,
 * cdb_verify_sa accepts the device array under the reference's rules (signed rule for groups > chuck_size) with 0
   inversions, and the signed rule really decided pairs,
 * string_index::query() rows of the reference (src/index.cpp:237-326, its own recurrences on its own array) equal the
   device rows for 2 000 keywords of 4..12 bytes, half sampled from the corpus and half random.

The reference build takes 1-3 minutes of host time; CDB_SKIP_SLOW=1 skips the file."""
import os
import time

import numpy as np
import pytest

import coffeedb_b200 as cdb
import oracle
from tests import corpora

pytestmark = pytest.mark.gpu

TARGET_BYTES = int(os.environ.get("CDB_CFG5_BYTES", 1 << 30))  # a smaller value gives a quick smoke run of this file


@pytest.fixture(scope="module")
def cfg5():
    import torch
    if os.environ.get("CDB_SKIP_SLOW") == "1":
        pytest.skip("CDB_SKIP_SLOW=1")
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref (the compiled reference) is not available")
    if torch.cuda.mem_get_info()[0] < 60 * (1 << 30):
        pytest.skip("needs 60 GB of free device memory")
    text, doc_off, ids, nd, n = corpora.utf8_corpus_on_device(TARGET_BYTES, seed=55)
    ix = cdb.StringIndex(device=0)
    t0 = time.perf_counter()
    ix.build_device(text.data_ptr(), doc_off.data_ptr(), ids.data_ptr(), nd, torch.cuda.current_stream().cuda_stream,
                    keep=(text, doc_off, ids))
    gpu_s = time.perf_counter() - t0
    h_text = text[:n].cpu().numpy()
    h_off = doc_off.cpu().numpy()
    h_ids = ids.cpu().numpy()
    ref = oracle.Ref()
    ref.add_many_borrowed(h_ids, h_text, h_off)
    t0 = time.perf_counter()
    ref.build()
    ref_s = time.perf_counter() - t0
    print(f"\ncfg5-shaped corpus: {nd} docs, {n} bytes; device build {gpu_s:.2f} s ({ix.build_stats()}), "
          f"reference build {ref_s:.1f} s on {oracle.hardware_threads()} host threads")
    yield ix, ref, h_text, h_off, h_ids, n, nd
    ref.close()
    ix.close()


def test_geometry_is_config5(cfg5):
    ix, _ref, _t, h_off, _i, n, nd = cfg5
    inf = ix.info()
    assert n >= TARGET_BYTES and inf["n"] == n and inf["nd"] == nd
    if TARGET_BYTES >= 1 << 30:
        assert inf["width"] == 8  # bits1 + bits2 > 32: 64-bit elements (src/index.cpp:203-208)
    assert int(np.diff(h_off).max()) <= 65536
    assert ix.prefix_directory()["symbols"] == 0  # note-N1 layout: the reference's own recurrences run, no directory


def test_suffix_array_equals_compiled_reference_up_to_ties(cfg5):
    ix, ref, _t, _o, _i, n, _nd = cfg5
    raw, bits, mask, width = ref.export_sa_raw()
    inf = ix.info()
    assert (bits, width, len(raw)) == (inf["bits"], inf["width"], n) and mask == inf["mask"]
    cmp_ = ix.compare_sa(raw)
    assert cmp_["different"] == 0, cmp_
    assert cmp_["identical"] + cmp_["ties"] == n
    print(f"\nidentical elements {cmp_['identical']}, elements inside tie groups in a different order {cmp_['ties']}")


def test_device_verifier_with_signed_levels(cfg5):
    ix, *_ = cfg5
    v = ix.verify_sa()
    assert v["ok"] and v["inversions"] == 0 and v["duplicates"] == 0 and v["invalid"] == 0, v
    assert v["signed_rule_pairs"] > 0, v


def test_query_rows_equal_reference(cfg5):
    ix, ref, h_text, h_off, _ids, n, nd = cfg5
    rng = np.random.default_rng(56)
    pats = []
    for _ in range(1000):  # sampled from the corpus (may start inside a multi-byte sequence, like any byte query)
        m = int(rng.integers(4, 13))
        d = int(rng.integers(0, nd))
        ln = int(h_off[d + 1] - h_off[d])
        if ln < m:
            continue
        o = int(rng.integers(0, ln - m + 1))
        pats.append(h_text[h_off[d] + o: h_off[d] + o + m].tobytes())
    for _ in range(1000):  # random printable ASCII / mixed bytes of the same lengths
        m = int(rng.integers(4, 13))
        pats.append(bytes(rng.integers(0x20, 0xF0, size=m, dtype=np.uint8)) if rng.random() < 0.3
                    else bytes(rng.integers(0x20, 0x7F, size=m, dtype=np.uint8)))
    pats += [b" ", b"e", b"\xc3", b"ab", b"the"]  # long intervals: the large path at this size
    pat, poff = cdb.pack(pats)
    want_off, want = ref.query_batch(pat, poff)
    got_off, got = ix.locate_batch(pat, poff)
    assert np.array_equal(got_off, want_off)
    assert np.array_equal(got, want)
    assert len(want) > 1000
