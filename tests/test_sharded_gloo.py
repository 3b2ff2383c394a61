"""Host-side logic of the sharded index (coffeedb_b200/sharded.py) on CPU: world_size 2 and 3 over gloo.

The device engine cannot run here, so every rank gets a TEST-ONLY local engine backed by the oracle; what is under
test is the sharding contract: doc-range split, pattern broadcast, all_gather of row lengths -> global CSR offsets,
all_reduce of occurrence totals, and the rank-ordered row concatenation being exactly string_index::query on the
whole corpus."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import corpora


class OracleShard:
    """TEST-ONLY stand-in for coffeedb_b200.StringIndex with the same host-buffer surface."""

    def __init__(self):
        self.ids, self.text, self.off = [], [], [np.zeros(1, np.int64)]

    def add_many(self, ids, text, doc_off):
        base = self.off[-1][-1]
        self.ids.append(np.asarray(ids, np.int64))
        self.text.append(np.asarray(text, np.uint8)[doc_off[0]:doc_off[-1]])
        self.off.append(np.asarray(doc_off[1:], np.int64) - doc_off[0] + base)

    def build(self):
        import oracle
        self.ids = np.concatenate(self.ids) if self.ids else np.zeros(0, np.int64)
        self.text = np.concatenate(self.text) if self.text else np.zeros(0, np.uint8)
        self.off = np.concatenate(self.off)
        self.sa, self.bits1, _w = oracle.port.build_sa(self.text, self.off)

    def locate_batch(self, pat, pat_off):
        import oracle
        rows = [oracle.port.query(self.text, self.off, self.ids, self.sa, self.bits1, bytes(pat[pat_off[q]:pat_off[q + 1]]))
                for q in range(len(pat_off) - 1)]
        row_off = np.zeros(len(rows) + 1, np.int64)
        row_off[1:] = np.cumsum([len(r) for r in rows])
        return row_off, (np.concatenate(rows) if rows and row_off[-1] else np.zeros((0, 2), np.int64))

    def info(self):
        return {"n": int(len(self.text))}

    def close(self):
        pass


class OracleShardWide(OracleShard):
    """Reports a suffix count beyond 2^31, which keeps the per-batch exchange on 8-byte integers."""

    def info(self):
        return {"n": 1 << 33}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, wide=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from coffeedb_b200.sharded import ShardedStringIndex, shard_range
        text, off, ids = corpora.ragged(901, 50, seed=5, alphabet=b"abc")
        nd = len(ids)
        lo, hi = shard_range(nd, rank, world)
        ix = ShardedStringIndex(device=torch.device("cpu"), index_factory=OracleShardWide if wide else OracleShard)
        ix.add_many(ids[lo:hi], text, off[lo:hi + 1])
        ix.build()
        assert ix.nd_global == nd and ix.doc_base == lo
        assert ix.narrow == (not wide)  # 4-byte exchange whenever every shard has < 2^31 suffixes
        spat, soff = corpora.sampled_patterns(text, off, 60, 1, 6, seed=6)
        pats = [bytes(spat[soff[i]:soff[i + 1]]) for i in range(60)] + [b"zz", b"a"]
        # only the source rank knows the request
        res = ix.locate_batch(pats if rank == 1 % world else None, src=1 % world)
        flat = ix.gather_rows(res, dst=0)
        sa, b1, _w = oracle.port.build_sa(text, off)
        want = [oracle.port.query(text, off, ids, sa, b1, kw) for kw in pats]
        occ = res.occurrences.numpy()
        assert [int(w[:, 1].sum()) if len(w) else 0 for w in want] == occ.tolist()
        assert res.global_row_off.numpy().tolist() == np.concatenate([[0], np.cumsum([len(w) for w in want])]).tolist()
        # this shard's part of every row is the slice of the whole-corpus row at rank_base
        ro = res.row_off.numpy()
        for qi, w in enumerate(want):
            mine = res.pairs.numpy()[ro[qi]:ro[qi + 1]]
            b = int(res.rank_base[qi])
            assert np.array_equal(mine, w[b:b + len(mine)])
        if rank == 0:
            gro, gp = flat
            for qi, w in enumerate(want):
                assert np.array_equal(gp.numpy()[gro[qi]:gro[qi + 1]], w), pats[qi]
        else:
            assert flat is None
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _subgroup_worker(rank, world, port, q):
    """Shards on a SUBSET of the ranks (global ranks 1 and 2 of 3): group-local ranks differ from global ranks, which the
    collectives and the send/recv of gather_rows must translate."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from coffeedb_b200.sharded import ShardedStringIndex, shard_range
        grp = dist.new_group([1, 2])
        if rank in (1, 2):
            text, off, ids = corpora.ragged(400, 40, seed=15, alphabet=b"ab")
            g_rank, g_world = dist.get_rank(grp), 2
            lo, hi = shard_range(len(ids), g_rank, g_world)
            ix = ShardedStringIndex(group=grp, device=torch.device("cpu"), index_factory=OracleShard)
            assert ix.rank == g_rank and ix.world == 2
            ix.add_many(ids[lo:hi], text, off[lo:hi + 1])
            ix.build()
            pats = [b"ab", b"ba", b"aab", b"bbbb", b"zz"]
            res = ix.locate_batch(pats if g_rank == 1 else None, src=1)  # group rank 1 = global rank 2 holds the request
            flat = ix.gather_rows(res, dst=1)
            if g_rank == 1:
                sa, b1, _w = oracle.port.build_sa(text, off)
                gro, gp = flat
                for qi, kw in enumerate(pats):
                    assert np.array_equal(gp.numpy()[gro[qi]:gro[qi + 1]], oracle.port.query(text, off, ids, sa, b1, kw)), kw
            else:
                assert flat is None
        q.put((rank, "ok"))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_index_on_a_subgroup():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_subgroup_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(3)]
    for p in procs:
        p.join(60)
    assert all(msg == "ok" for _r, msg in out), out


@pytest.mark.parametrize("world,wide", [(2, False), (3, False), (2, True)])
def test_sharded_index_over_gloo(world, wide):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, wide)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(msg == "ok" for _r, msg in out), out


def test_shard_ranges_cover_corpus():
    from coffeedb_b200.sharded import shard_range
    for nd in (0, 1, 7, 8, 100, 12345):
        for world in (1, 2, 3, 8):
            r = [shard_range(nd, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nd
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
