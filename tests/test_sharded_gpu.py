"""Sharded index on real devices: world 1 always, world 2 over NCCL when the box has two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

import oracle
from tests import corpora

pytestmark = pytest.mark.gpu


def _check_against_oracle(ix, res, flat, text, off, ids, pats, rank):
    sa, b1, _w = oracle.port.build_sa(text, off)
    want = [oracle.port.query(text, off, ids, sa, b1, kw) for kw in pats]
    assert res.occurrences.cpu().tolist() == [int(w[:, 1].sum()) if len(w) else 0 for w in want]
    ro = res.row_off.cpu().numpy()
    pr = res.pairs.cpu().numpy()
    base = res.rank_base.cpu().numpy()
    for qi, w in enumerate(want):
        mine = pr[ro[qi]:ro[qi + 1]]
        assert np.array_equal(mine, w[base[qi]:base[qi] + len(mine)]), pats[qi]
    if rank == 0:
        gro, gp = flat
        gro, gp = gro.cpu().numpy(), gp.cpu().numpy()
        for qi, w in enumerate(want):
            assert np.array_equal(gp[gro[qi]:gro[qi + 1]], w), pats[qi]


def _corpus():
    text, off, ids = corpora.ragged(3001, 70, seed=91, alphabet=b"abcd")
    spat, soff = corpora.sampled_patterns(text, off, 120, 1, 7, seed=92)
    pats = [bytes(spat[soff[i]:soff[i + 1]]) for i in range(120)] + [b"zz", b"a", b"ab"]
    return text, off, ids, pats


def test_sharded_world1():
    from coffeedb_b200.sharded import ShardedStringIndex
    text, off, ids, pats = _corpus()
    ix = ShardedStringIndex(device=torch.device("cuda", 0))
    ix.add_many(ids, text, off)
    ix.build()
    res = ix.locate_batch(pats)
    flat = ix.gather_rows(res)
    _check_against_oracle(ix, res, flat, text, off, ids, pats, 0)
    ix.close()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from coffeedb_b200.sharded import ShardedStringIndex, shard_range
        text, off, ids, pats = _corpus()
        lo, hi = shard_range(len(ids), rank, world)
        ix = ShardedStringIndex(device=torch.device("cuda", rank))
        ix.add_many(ids[lo:hi], text, off[lo:hi + 1])
        ix.build()
        res = ix.locate_batch(pats if rank == 0 else None, src=0)
        flat = ix.gather_rows(res, dst=0)
        _check_against_oracle(ix, res, flat, text, off, ids, pats, rank)
        q.put((rank, "ok"))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_world2_nccl():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(60)
    assert all(m == "ok" for _r, m in out), out
