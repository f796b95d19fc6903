#!/usr/bin/env python3
"""BASELINE configs[3]: a file batch (log-uniform 64 KiB - 8 MiB blobs) sharded per file
over the GPUs of one node, every rank scanning its share with yrb_scan_batch from pinned
host memory (end to end: H2D inside the timed region, records back on the host).

  python scripts/file_batch_multi.py [--strings 100000] [--files-per-gpu 600] [--steps 5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P scripts/file_batch_multi.py ...

No collective on the data path; the per-file records are gathered with a host gather.
Prints one JSON line on rank 0 (weak scaling: files per GPU fixed)."""
import argparse, json, math, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
import yrh
from yara_b200 import Engine, PinnedBuffer, shard, synth

ap = argparse.ArgumentParser()
ap.add_argument("--strings", type=int, default=100000)
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--files-per-gpu", type=int, default=600)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--check", type=int, default=6, help="files per rank re-scanned singly and compared")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)

rs = synth.ruleset(args.config, args.strings)
t0 = time.time()
image = yrh.b200().compile(rs.source).flatten()
t_compile = time.time() - t0
eng = Engine(image, local)

nfiles = args.files_per_gpu * world
rng = random.Random(3)
sizes = [int(math.exp(rng.uniform(math.log(65536), math.log(8 << 20)))) for _ in range(nfiles)]
mine = shard.plan_files(sizes, world)[rank]
pin = PinnedBuffer(sum((sizes[k] + 15) & ~15 for k in mine) + 16)
files, o = {}, 0
for k in mine:
    b = synth.corpus(rs, sizes[k], 65536, seed=1000 + k, plant_seed=k, tail_plant=False)
    pin.array[o : o + len(b)] = b
    files[k] = pin.array[o : o + len(b)]
    o += (len(b) + 15) & ~15
my_bytes = sum(sizes[k] for k in mine)

def step():
    return shard.scan_files(eng, files, mine, one_allocation=True)

for _ in range(args.warmup):
    out = step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t = time.time()
for _ in range(args.steps):
    out = step()
torch.cuda.synchronize()
dt = torch.tensor([time.time() - t], device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
dt = float(dt.item())

# parity spot check: batch records == single-buffer scan records
for k in mine[: args.check]:
    r = eng.scan_host(files[k])
    a = sorted((int(x["string_idx"]), int(x["offset"]), int(x["match_length"])) for x in out[k][1])
    b = sorted((int(x["string_idx"]), int(x["offset"]), int(x["match_length"])) for x in r.matches)
    assert r.rc == 0 and a == b, "file %d differs" % k

if world > 1:
    gloo = dist.new_group(backend="gloo")
    class _G:   # gather_files only needs these three
        get_world_size = staticmethod(lambda: world)
        get_rank = staticmethod(lambda: rank)
        gather_object = staticmethod(lambda obj, lst, dst=0: dist.gather_object(obj, lst, dst=dst, group=gloo))
    allres = shard.gather_files(out, nfiles, _G, dst=0)
else:
    allres = [out[k] for k in range(nfiles)]

if rank == 0:
    total = sum(sizes)
    nm = sum(len(m) for _, m in allres)
    print(json.dumps({
        "metric": "scanned GB/s, file batch end to end (pinned host -> records on host)", "value": round(total * args.steps / dt / 1e9, 2),
        "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "files_per_s": round(nfiles * args.steps / dt), "scaling": "weak",
        "config": {"workload": "configs[3] shape: %d-string mixed ruleset, %d blobs of 64 KiB-8 MiB (log-uniform), file-sharded" % (args.strings, nfiles),
                   "bytes": total, "tables_size_u32": int(np.frombuffer(image[16:20], dtype=np.uint32)[0]), "matches": nm,
                   "compile_s": round(t_compile, 1)},
    }))
eng.close()
if world > 1:
    dist.destroy_process_group()
