"""TEST INFRASTRUCTURE — builds two CoffeeDB servers from the reference sources where they lie (/root/reference):

  _build/coffeedb_ref_server     the UNMODIFIED reference (all seven src/*.cpp), the CPU arm of the HTTP-level parity test
  _build/coffeedb_b200_server    the reference with INTEGRATION.md §1 applied to a scratch copy of src/: `class string_index`
                                 (src/index.h:54-86) replaced by `coffeedb_b200::basic_string_index<index>` and its member
                                 definitions (src/index.cpp:75-128, 174-326) dropped; database.cpp / interface.cpp /
                                 server.cpp / main.cpp / command.cpp are compiled UNCHANGED and linked against
                                 libcoffeedb_b200.so

The scratch copy and the binaries live under tests/dropin/_build/ (git-ignored; they travel to the GPU box with the
snapshot).  Nothing of the reference is committed: the edit is made by this script, on anchors it checks first.
Build recipe as in oracle/Makefile: g++ -std=c++20 with the progress-bar shim force-included."""
from __future__ import annotations

import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")
REF = os.environ.get("COFFEEDB_REFERENCE", "/root/reference")
SHIM = os.path.join(ROOT, "oracle", "shim_progress_bar.h")
REF_SERVER = os.path.join(OUT, "coffeedb_ref_server")
B200_SERVER = os.path.join(OUT, "coffeedb_b200_server")
SOURCES = ["main.cpp", "server.cpp", "interface.cpp", "database.cpp", "index.cpp", "command.cpp", "profile.cpp"]


def _match_brace(text: str, open_pos: int) -> int:
    """Index just past the brace that closes the one at open_pos."""
    depth = 0
    for i in range(open_pos, len(text)):
        if text[i] == "{":
            depth += 1
        elif text[i] == "}":
            depth -= 1
            if depth == 0:
                return i + 1
    raise RuntimeError("unbalanced braces")


def _drop_definition(text: str, signature_regex: str) -> str:
    m = re.search(signature_regex, text)
    if not m:
        raise RuntimeError(f"anchor not found: {signature_regex}")
    end = _match_brace(text, text.index("{", m.end() - 1))
    return text[: m.start()] + text[end:]


def apply_integration(src_dir: str) -> None:
    """INTEGRATION.md §1 on a scratch copy of src/."""
    h = os.path.join(src_dir, "index.h")
    t = open(h).read()
    m = re.search(r"class\s+string_index\s*:\s*public\s+index\s*\{", t)
    if not m:
        raise RuntimeError("class string_index not found in index.h")
    end = _match_brace(t, t.index("{", m.start()))
    end = t.index(";", end) + 1
    adaptor = os.path.join(ROOT, "coffeedb_b200", "host", "string_index.hpp")
    t = (t[: m.start()] + f'#include <stdexcept>\n#include <string>\n#include "{adaptor}"\n'
         "using string_index = coffeedb_b200::basic_string_index<index>;  // derives from the reference's own `index`\n" + t[end:])
    open(h, "w").write(t)
    c = os.path.join(src_dir, "index.cpp")
    t = open(c).read()
    t = _drop_definition(t, r"template\s*<\s*typename\s+T\s*>\s*void\s+string_index::parallel_sort\s*\(\s*\)\s*const\s*\{")
    t = _drop_definition(t, r"void\s+string_index::add\s*\([^)]*\)\s*\{")
    t = _drop_definition(t, r"void\s+string_index::build\s*\(\s*\)\s*\{")
    t = _drop_definition(t, r"std::vector<std::pair<int64_t,\s*int64_t>>\s+string_index::query\s*\([^)]*\)\s*const\s*\{")
    if "string_index::" in t:
        raise RuntimeError("a string_index member definition is left in index.cpp")
    open(c, "w").write(t)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + r.stdout[-3000:] + r.stderr[-3000:])


def build(force: bool = False) -> dict | None:
    """Builds both servers when the reference sources are present; returns their paths (or the prebuilt ones), None when
    neither sources nor binaries exist."""
    have_src = os.path.isdir(os.path.join(REF, "src"))
    lib = os.path.join(ROOT, "coffeedb_b200", "libcoffeedb_b200.so")
    fresh = all(os.path.exists(p) for p in (REF_SERVER, B200_SERVER)) and not force
    if fresh and have_src:
        newest_in = max(os.path.getmtime(p) for p in (os.path.abspath(__file__), os.path.join(ROOT, "coffeedb_b200", "host", "string_index.hpp"),
                                                       os.path.join(ROOT, "coffeedb_b200", "host", "micro_batcher.hpp"),
                                                       os.path.join(ROOT, "include", "coffeedb_b200.h")))
        fresh = newest_in < min(os.path.getmtime(REF_SERVER), os.path.getmtime(B200_SERVER))
    if have_src and not fresh:
        if not os.path.exists(lib):
            raise RuntimeError("build libcoffeedb_b200.so first")
        os.makedirs(OUT, exist_ok=True)
        flags = ["-std=c++20", "-O2", "-w", "-DPROGRESS_BAR", "-include", SHIM, f"-I{REF}/package"]
        # the unmodified reference, compiled from where it lies
        _run(["g++", *flags, f"-I{REF}/src", *[os.path.join(REF, "src", s) for s in SOURCES], "-o", REF_SERVER, "-lpthread"])
        # scratch copy with the integration applied
        scratch = os.path.join(OUT, "src")
        shutil.rmtree(scratch, ignore_errors=True)
        shutil.copytree(os.path.join(REF, "src"), scratch)
        apply_integration(scratch)
        libdir = os.path.join(ROOT, "coffeedb_b200")
        _run(["g++", *flags, f"-I{scratch}", *[os.path.join(scratch, s) for s in SOURCES], "-o", B200_SERVER, f"-L{libdir}",
              "-lcoffeedb_b200", "-Wl,-rpath,$ORIGIN/../../../coffeedb_b200", f"-Wl,-rpath,{libdir}", "-lpthread"])
    if all(os.path.exists(p) for p in (REF_SERVER, B200_SERVER)):
        return {"reference": REF_SERVER, "b200": B200_SERVER}
    return None


if __name__ == "__main__":
    print(build(force=True))
