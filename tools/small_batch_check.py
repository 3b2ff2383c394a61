"""Quick check of the experimental small-batch path (CDB_SMALL_BATCH) on a GPU box, run from the repo root:
    python tools/small_batch_check.py
Parity of a 256-keyword batch and of single keywords against the general path, and the single-keyword latency of both."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coffeedb_b200 as cdb  # noqa: E402
from tests import corpora  # noqa: E402

text, off, ids = corpora.uniform(20000, 60, seed=81, lo=97, hi=101)
ix = cdb.StringIndex(); ix.add_many(ids, text, off); ix.build()
pat, poff = corpora.sampled_patterns(text, off, 300, 5, 9, seed=82)
pats = [bytes(pat[poff[i]:poff[i + 1]]) for i in range(300)]
want = ix.locate_batch(pats[:256]); w1 = [ix.locate_batch([p]) for p in pats[:20]]
t = time.time(); [ix.locate_batch([p]) for p in pats[:200]]; t0 = (time.time() - t) / 200
os.environ["CDB_SMALL_BATCH"] = "256"
got = ix.locate_batch(pats[:256]); ok = np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
ok1 = all(np.array_equal(ix.locate_batch([p])[1], w[1]) for p, w in zip(pats[:20], w1))
st = cdb.last_locate_stats()["total_ms"]
t = time.time(); [ix.locate_batch([p]) for p in pats[:200]]; t1 = (time.time() - t) / 200
print("small-batch parity", ok, ok1, "general %.1f us, small %.1f us" % (t0 * 1e6, t1 * 1e6), "small path taken:", st == 0.0)
