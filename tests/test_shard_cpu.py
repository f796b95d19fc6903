"""CPU tests of the N > 1 host logic: shard planning, margins, and the gather of
raw records over torch.distributed (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np

from yara_b200 import shard
from yara_b200.api import RAW_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_covers_block_exactly():
    for size in (0, 1, 4095, 4096, 1 << 20, (1 << 30) + 12345):
        for n in (1, 2, 3, 8):
            p = shard.plan(size, n, 8192)
            assert len(p) == n
            assert p[0][0] == 0 and p[-1][1] == size
            for a, b in zip(p, p[1:]):
                assert a[1] == b[0]                      # owned ranges tile the block
            for lo, hi, g0, g1 in p:
                assert g0 <= lo <= hi <= g1 <= size
                assert lo - g0 == min(lo, 8192) and g1 - hi == min(size - hi, 8192)


def test_required_margin_from_golden_image():
    img = open(os.path.join(ROOT, "tests", "golden", "synth_cfg3.img"), "rb").read()
    m = shard.required_margin(img)
    assert m >= 4096 + 64 and m % 4096 == 0


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from yara_b200 import shard
from yara_b200.api import RAW_DTYPE
dist.init_process_group("gloo")
r = dist.get_rank()
raw = np.zeros(3 + r, dtype=RAW_DTYPE)
raw["key"] = np.arange(len(raw)) + 100 * r
raw["string_idx"] = r
out = shard.gather_raw(raw, dist, dst=0)
if r == 0:
    assert len(out) == 7, len(out)
    assert out["string_idx"].tolist() == [0, 0, 0, 1, 1, 1, 1]
    assert out["key"].tolist() == [0, 1, 2, 100, 101, 102, 103]
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
"""


def test_gather_raw_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    p = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "GATHER_OK" in p.stdout


def test_plan_files_deals_every_file_once_and_balances():
    import random
    rng = random.Random(1)
    for n in (1, 2, 3, 8):
        for nfiles in (0, 1, 5, 1000):
            sizes = [rng.randrange(1, 8 << 20) for _ in range(nfiles)]
            p = shard.plan_files(sizes, n)
            assert len(p) == n
            assert sorted(k for ix in p for k in ix) == list(range(nfiles))
            assert all(ix == sorted(ix) for ix in p)
            if nfiles >= 1000:
                loads = [sum(sizes[k] for k in ix) for ix in p]
                assert max(loads) - min(loads) <= max(sizes)
    assert shard.plan_files([5, 5, 5], 2) == [[0, 2], [1]]   # deterministic tie-breaking


FILES_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from yara_b200 import shard
from yara_b200.api import YRB_MATCH_DTYPE
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
sizes = [100, 7000, 30, 7000, 1, 0, 512]
mine = {}
for k in shard.plan_files(sizes, w)[r]:
    m = np.zeros(k %% 3, dtype=YRB_MATCH_DTYPE)
    m["offset"] = np.arange(len(m)) + 10 * k
    m["string_idx"] = k
    mine[k] = (0, m)
out = shard.gather_files(mine, len(sizes), dist, dst=0)
if r == 0:
    assert len(out) == len(sizes) and all(o is not None for o in out)
    for k, (e, m) in enumerate(out):
        assert e == 0 and len(m) == k %% 3 and (m["string_idx"] == k).all()
        assert m["offset"].tolist() == [10 * k + j for j in range(k %% 3)]
    print("FILES_OK")
else:
    assert out is None
dist.destroy_process_group()
"""


def test_gather_files_gloo_world2(tmp_path):
    script = tmp_path / "worker_files.py"
    script.write_text(FILES_WORKER % ROOT)
    p = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29519", str(script)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "FILES_OK" in p.stdout
