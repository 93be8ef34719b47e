"""One exact brute-force call per case inside the profiler window (run under
`ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`): the DRAM
traffic of the fused scan (no [nq, n] score matrix) against its algorithmic bytes.  1M x 768 f32, Q = 1024 and Q = 1."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import CONFIGS, build_snapshot, make_queries
from velesdb_b200 import _native as nv

nv.init(0)
dev = torch.device("cuda", 0)
n, k = 1_000_000, 10
cfg = dict(CONFIGS["c1"], n=n, name="probe")
snap, _, _, _ = build_snapshot(torch, cfg, dev)
stream = torch.cuda.current_stream().cuda_stream
for nq in (1024, 1):
    q_d = make_queries(torch, cfg, nq, 99, dev)
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    for _ in range(2):
        snap.bruteforce_batch_device(q_d, k, ids, sc, stream)
    torch.cuda.synchronize()
    l0 = nv.lib().veles_launch_count()
    torch.cuda.profiler.start()
    snap.bruteforce_batch_device(q_d, k, ids, sc, stream)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps({"n": n, "nq": nq, "launches": int(nv.lib().veles_launch_count() - l0),
                      "algorithmic_bytes": n * 768 * 4 + nq * 768 * 4 + nq * k * 8}), flush=True)
