"""FASTA ingest timing on the GPU box (pinned output): native scanner vs the Python line loop (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idelucs_b200.seqset import read_fasta_native, read_fasta_raw
rng = np.random.default_rng(0)
ni, L = 20000, 10000
rows = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(ni, L // 100, 100))]
nl = np.full((L // 100, 1), 10, np.uint8)
with open("/tmp/t.fa", "wb") as fh:
    for i in range(ni):
        fh.write(b">seq_%d synthetic\n" % i)
        fh.write(np.concatenate([rows[i], nl], axis=1).tobytes())
size = os.path.getsize("/tmp/t.fa")
read_fasta_native("/tmp/t.fa")
import torch
t = time.perf_counter(); x = torch.empty(200_000_000, dtype=torch.uint8, pin_memory=torch.cuda.is_available()); print("pinned alloc of 200 MB: %.3f s" % (time.perf_counter() - t)); del x
for env in ({"IDL_FASTA_THREADS": "1"}, {"IDL_FASTA_THREADS": "2"}, {"IDL_FASTA_THREADS": "4"}, {"IDL_FASTA_THREADS": "8"}, {}):
    os.environ.update(env)
    ts = []
    for _ in range(3):
        t = time.perf_counter(); a = read_fasta_native("/tmp/t.fa"); ts.append(time.perf_counter() - t)
    dt = min(ts)
    t = time.perf_counter(); a2 = read_fasta_native("/tmp/t.fa", pinned=False); dtu = time.perf_counter() - t
    print("   unpinned output: %.3f s" % dtu)
    print(env, "native %.3f s = %.2f GB/s (%d cpus)" % (dt, size / dt / 1e9, os.cpu_count()), flush=True)
    for k in env: os.environ.pop(k)
t = time.perf_counter(); b = read_fasta_raw("/tmp/t.fa"); dt = time.perf_counter() - t
print("python line loop %.3f s = %.2f GB/s" % (dt, size / dt / 1e9))
assert a[0] == b[0] and a[1].numpy().tobytes() == b"".join(b[1])
