"""End-to-end drop-in proof ("missing" item 5): INTEGRATION.md §1 is applied to a scratch copy of the
reference's src/ (tests/dropin/build_servers.py), the UNMODIFIED database.cpp / interface.cpp / server.cpp / main.cpp are
linked against libcoffeedb_b200.so, the resulting server is started, and the logic of the reference's own black-box tests
runs against it over HTTP — next to the unmodified reference server, whose response bodies must be byte-identical:

  test/test-string.py:25-56       insert / build / query, $correlation of every object == brute-force overlapping count
  test/test-highlight.py:31-59    five OR-ed keywords of one key, returned id set and highlighted text == str.replace chain
  test/test-concurrency.py:44-57  8 threads of random insert / build / query, HTTP 200 only
  examples/example.py + README.md:64-109   the walkthrough's known answers

Sizes are reduced from the reference scripts (5000 x 5000 random letters) so that the file runs in about a minute."""
import json
import os
import random
import socket
import subprocess
import threading
import time

import pytest
import requests

from tests.dropin import build_servers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Server:
    def __init__(self, binary, directory):
        self.port = _free_port()
        os.makedirs(directory, exist_ok=True)
        self.log = open(os.path.join(directory, "server.log"), "w")
        self.proc = subprocess.Popen([binary, f"--port={self.port}", f"--directory={directory}"], stdout=self.log,
                                     stderr=subprocess.STDOUT, cwd=directory)
        self.url = f"http://127.0.0.1:{self.port}/coffeedb"
        deadline = time.time() + 60
        while time.time() < deadline:
            if self.proc.poll() is not None:
                raise RuntimeError(f"server exited with {self.proc.returncode}")
            try:
                socket.create_connection(("127.0.0.1", self.port), timeout=0.5).close()
                return
            except OSError:
                time.sleep(0.1)
        raise RuntimeError("server did not start listening")

    def post(self, obj):
        return requests.post(self.url, json.dumps(obj), timeout=120)

    def send(self, obj):
        r = self.post(obj)
        assert r.status_code == 200, r.text
        return r.text

    def stop(self):
        self.proc.terminate()
        try:
            self.proc.wait(timeout=10)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.log.close()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def binaries():
    b = build_servers.build()
    if b is None:
        pytest.skip("neither /root/reference nor prebuilt tests/dropin/_build servers are present")
    return b


@pytest.fixture()
def pair(binaries, tmp_path):
    """(drop-in server on the GPU, unmodified reference server on the CPU), each on its own empty directory."""
    a = Server(binaries["b200"], str(tmp_path / "b200"))
    b = Server(binaries["reference"], str(tmp_path / "ref"))
    yield a, b
    a.stop()
    b.stop()


def both(pair, obj):
    """The same request to both servers; bodies must be byte-identical."""
    a, b = pair
    ra, rb = a.send(obj), b.send(obj)
    assert ra == rb, (obj, ra[:300], rb[:300])
    return ra


def count(text, sub):  # test/test-string.py:14-19: overlapping occurrences
    return sum(1 for i in range(len(text) - len(sub) + 1) if text[i:i + len(sub)] == sub)


# ---------------------------------------------------------------------------------------------------- without a GPU
@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_dropin_server_without_gpu_fails_loudly_through_the_servers_catch(binaries, tmp_path):
    """No CUDA device: there is no CPU fallback — build() throws std::runtime_error inside the unmodified server, whose
    catch (src/server.cpp:58-62) answers HTTP 500 "[Error] <message>."; inserts (which never touch the index,
    src/database.cpp:283-379) still work."""
    s = Server(binaries["b200"], str(tmp_path / "b200"))
    try:
        assert s.post({"operation": "insert", "data": {"id": 1, "val": "3010103"}}).status_code == 200
        r = s.post({"operation": "build"})
        assert r.status_code == 500
        assert r.text.startswith("[Error] no CUDA device available: coffeedb_b200 has no CPU fallback")
    finally:
        s.stop()


def test_reference_server_known_answers(binaries, tmp_path):
    """The unmodified reference, compiled by the same recipe, reproduces README.md:78-109 (this pins the CPU arm)."""
    s = Server(binaries["reference"], str(tmp_path / "ref"))
    try:
        for i, v in enumerate(["3010103", "301022"]):
            s.send({"operation": "insert", "data": {"id": i, "val": v}})
        s.send({"operation": "build"})
        got = json.loads(s.send({"operation": "query", "constraints": {"val": "010"}, "fields": ["id", "$correlation"]}))
        assert got == [{"$correlation": 2, "id": 0}, {"$correlation": 1, "id": 1}]
    finally:
        s.stop()


# ------------------------------------------------------------------------------------------------------ on the GPU
@pytest.mark.gpu
def test_readme_walkthrough_identical_on_both_servers(pair):
    for i, (v, n) in enumerate([("3010103", 7), ("301022", 6), ("abcabc", 1), ("", 0)]):
        both(pair, {"operation": "insert", "data": {"id": i, "val": v, "num": n, "flag": bool(i & 1), "score": i / 4}})
    both(pair, {"operation": "build"})
    got = json.loads(both(pair, {"operation": "query", "constraints": {"val": "010"}, "fields": ["id", "$correlation"]}))
    assert got == [{"$correlation": 2, "id": 0}, {"$correlation": 1, "id": 1}]  # README.md:91
    hl = json.loads(both(pair, {"operation": "query", "constraints": {"val": "01"}, "fields": ["val"],
                                "highlight": ["<b>", "</b>"]}))
    assert {"val": "3<b>01010</b>3"} not in hl  # touching matches are NOT merged (src/database.cpp:66-76) ...
    assert {"val": "3<b>01</b><b>01</b>03"} in hl  # ... so the markers repeat
    both(pair, {"operation": "query", "constraints": {"val": ["010", "22"], "num": "[6,7]"}, "fields": ["id", "num", "$correlation"]})
    both(pair, {"operation": "query", "constraints": {"val": "0", "$correlation": "[2,3)"}, "fields": ["id", "$correlation"]})
    both(pair, {"operation": "count", "constraints": {"val": "3"}})
    both(pair, {"operation": "query", "constraints": {"val": "abc"}, "span": "[0,1)"})
    a, b = pair
    ra, rb = a.post({"operation": "query", "constraints": {"val": ""}}), b.post({"operation": "query", "constraints": {"val": ""}})
    assert ra.status_code == rb.status_code == 500 and ra.text == rb.text == "[Error] Empty keywords are not allowed."


@pytest.mark.gpu
def test_string_property_over_http(pair):
    """test/test-string.py at 400 x 2000: for every keyword and every object, $correlation == brute-force count."""
    rng = random.Random(20261017)
    vals = ["".join(chr(rng.randint(ord("a"), ord("z"))) for _ in range(2000)) for _ in range(400)]
    for i, v in enumerate(vals):
        both(pair, {"operation": "insert", "data": {"id": i, "val": v}})
    both(pair, {"operation": "build"})
    for _ in range(40):
        kw = "".join(chr(rng.randint(ord("a"), ord("z"))) for _ in range(3))
        body = both(pair, {"operation": "query", "constraints": {"val": kw}, "fields": ["id", "$correlation"]})
        cnt = {int(o["id"]): int(o["$correlation"]) for o in json.loads(body)}
        for i, v in enumerate(vals):
            assert count(v, kw) == cnt.get(i, 0)


@pytest.mark.gpu
def test_highlight_property_over_http(pair):
    """test/test-highlight.py at 300 x 1000: five disjoint 4-letter keywords of one key, OR-ed."""
    rng = random.Random(7)
    vals = ["".join(chr(rng.randint(ord("a"), ord("z"))) for _ in range(1000)) for _ in range(300)]
    for i, v in enumerate(vals):
        both(pair, {"operation": "insert", "data": {"id": i, "val": v}})
    both(pair, {"operation": "build"})
    chars = [chr(ord("a") + i) for i in range(26)]
    for _ in range(25):
        rng.shuffle(chars)
        s = "".join(chars)
        kws = [s[i:i + 4] for i in range(0, 20, 4)]
        body = both(pair, {"operation": "query", "constraints": {"val": kws}, "fields": ["id", "val"], "highlight": ["<b>", "</b>"]})
        result = {int(o["id"]): o["val"] for o in json.loads(body)}
        answer = {}
        for i, text in enumerate(vals):
            t = text
            for k in kws:
                t = t.replace(k, f"<b>{k}</b>")
            if t != text:
                answer[i] = t
        assert result == answer


@pytest.mark.gpu
def test_concurrency_over_http(binaries, tmp_path):
    """test/test-concurrency.py:44-57 at 8 x 48 operations: random insert / build / query from 8 threads, HTTP 200 only
    (a build swaps in a new index while queries run on the old one, src/database.cpp:170-172, 276-281)."""
    s = Server(binaries["b200"], str(tmp_path / "b200"))
    bad = []

    def op(rng):
        k = rng.choice(["insert", "build", "query"])
        if k == "insert":
            r = s.post({"operation": "insert", "data": {"id": rng.randint(1, 10 ** 6),
                                                        "val": "".join(chr(rng.randint(97, 122)) for _ in range(256))}})
        elif k == "build":
            r = s.post({"operation": "build"})
        else:
            chars = [chr(ord("a") + i) for i in range(26)]
            rng.shuffle(chars)
            t = "".join(chars)
            r = s.post({"operation": "query", "constraints": {"val": [t[i:i + 4] for i in range(0, 20, 4)]},
                        "fields": ["id", "val"], "highlight": ["<b>", "</b>"]})
        if r.status_code != 200:
            bad.append((k, r.status_code, r.text[:200]))

    try:
        seed = random.Random(1)
        for _ in range(128):
            s.send({"operation": "insert", "data": {"id": seed.randint(1, 10 ** 6),
                                                    "val": "".join(chr(seed.randint(97, 122)) for _ in range(256))}})
        s.send({"operation": "build"})
        th = [threading.Thread(target=lambda k=k: [op(random.Random(100 + k)) for _ in range(48)]) for k in range(8)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not bad, bad[:5]
    finally:
        s.stop()
