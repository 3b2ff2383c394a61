"""The C++ adaptor over the C ABI (coffeedb_b200/host/string_index.hpp): compiles against the stand-alone base
and — where the reference tree is present — against the reference's own abstract `index` (src/index.h:9-23),
which is the drop-in arrangement INTEGRATION.md describes."""
import os
import subprocess

import pytest

import coffeedb_b200 as cdb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "test_adaptor.cpp")
OUT = os.path.join(ROOT, "tests", "host", "_build")
REF_INDEX_H = "/root/reference/src/index.h"


def compile_adaptor(name, extra=()):
    os.makedirs(OUT, exist_ok=True)
    cdb.lib()  # fails loudly when the library has not been built
    exe = os.path.join(OUT, name)
    libdir = os.path.dirname(cdb.LIB_PATH)
    cmd = ["g++", "-std=c++20", "-O1", "-w", "-pthread", *extra, SRC, "-o", exe, f"-L{libdir}", "-lcoffeedb_b200",
           f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_adaptor_links_and_fails_loudly_without_device():
    exe = compile_adaptor("test_adaptor")
    r = subprocess.run([exe, "nogpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "adaptor nogpu ok" in r.stdout


@pytest.mark.skipif(not os.path.exists(REF_INDEX_H), reason="reference tree absent (GPU box)")
def test_adaptor_derives_from_reference_index_class():
    exe = compile_adaptor("test_adaptor_refbase", [f'-DCDB_TEST_REFERENCE_BASE="{REF_INDEX_H}"'])
    r = subprocess.run([exe, "nogpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_adaptor_on_device():
    exe = compile_adaptor("test_adaptor")
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adaptor gpu ok" in r.stdout


def test_micro_batcher_host_logic():
    """coffeedb_b200/host/micro_batcher.hpp (the coalescing queue behind cdb_query / string_index::query) with a
    host-only stand-in backend: own rows, batches form, batch-size / in-flight limits, linger, error delivery."""
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "test_batcher")
    src = os.path.join(ROOT, "tests", "host", "test_batcher.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-pthread", src, "-o", exe], check=True,
                   capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "micro_batcher ok" in r.stdout


def test_std_sort_order_restatement_equals_std_sort():
    """coffeedb_b200/host/std_sort_order.hpp (the introsort the device-side filter runs to reproduce the tie order of
    src/interface.cpp:143-146) yields std::sort's permutation on 6 000 sequences, including every length up to 1 100
    with all-equal keys, few-valued keys and median-of-three killers (heap-sort fallback)."""
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "test_sort_order")
    src = os.path.join(ROOT, "tests", "host", "test_sort_order.cpp")
    subprocess.run(["g++", "-std=c++20", "-O2", "-Wall", "-Wextra", src, "-o", exe], check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sort_order ok" in r.stdout
