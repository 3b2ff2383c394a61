#!/bin/bash
# cfg5-shaped test file (skewed lengths: 64-bit elements at 1 GiB) + the verifier's counters at three sizes + the L2 policy micro-benchmark
mkdir -p gpurun_out
python - > gpurun_out/r02j_verify.txt 2>&1 <<'PY'
import torch, coffeedb_b200 as cdb
from tests import corpora
for tb in (20_000_000, 200_000_000, 1 << 30):
    t, o, i, nd, n = corpora.utf8_corpus_on_device(tb, seed=55)
    ix = cdb.StringIndex(device=0)
    ix.build_device(t.data_ptr(), o.data_ptr(), i.data_ptr(), nd, torch.cuda.current_stream().cuda_stream, keep=(t, o, i))
    print(tb, nd, n, ix.info(), ix.build_stats(), ix.verify_sa(), flush=True)
    ix.close()
PY
timeout 1500 python -m pytest tests/test_gpu_cfg5.py -q -s 2>&1 | tail -25 > gpurun_out/r02j_cfg5.txt
tools/_build/l2_microbench > gpurun_out/r02j_l2.txt 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02j_l2_ncu.csv tools/_build/l2_microbench > /dev/null 2>&1
cat gpurun_out/r02j_verify.txt gpurun_out/r02j_cfg5.txt gpurun_out/r02j_l2.txt
