"""Byte-range sharding of ONE block across GPUs / processes.

Rank g owns the Aho-Corasick events of bytes [g*S/G, (g+1)*S/G) and additionally
reads a margin on both sides: 3 bytes of look-back for the automaton state and up
to max(YR_RE_SCAN_LIMIT, 2 x longest literal) bytes for verification, which may
start before / run past the owned range.  Every candidate is verified by exactly
one rank (the owner of its atom event).  What stays global -- dedup of equal
offsets reached from different atoms, the 1M cap, chained strings -- is done once
over the concatenated raw records of all ranks (Engine.resolve_raw) on one GPU.

No collective is on the data path: raw records (32 bytes per match callback) are
gathered to rank 0, by plain D2H + host gather.

File batches (BASELINE configs[3]) shard per file instead: plan_files() deals the
files out by size, every rank scans its share with Engine.scan_batch, and the
per-file records are gathered on the host the same way.
"""
import numpy as np

from .api import RAW_DTYPE

RE_SCAN_LIMIT = 4096


def required_margin(image: bytes) -> int:
    """Smallest read margin that makes any shard boundary invisible for this ruleset."""
    hdr = np.frombuffer(image[:48], dtype=np.uint32)
    num_strings = int(hdr[6])
    off_strings = int(np.frombuffer(image[48 + 24 : 48 + 32], dtype=np.uint64)[0])
    longest = 0
    if num_strings:
        s = np.frombuffer(image, dtype=np.int32, count=num_strings * 10, offset=off_strings).reshape(num_strings, 10)
        longest = int(s[:, 1].max())
    m = max(RE_SCAN_LIMIT, 2 * longest) + 64
    return (m + 4095) & ~4095


def plan(block_size: int, n: int, margin: int):
    """[(own_lo, own_hi, g0, g1)] for n shards; boundaries are multiples of 4096."""
    out = []
    per = ((block_size + n - 1) // n + 4095) & ~4095
    for g in range(n):
        lo, hi = min(block_size, g * per), min(block_size, (g + 1) * per)
        out.append((lo, hi, max(0, lo - margin), min(block_size, hi + margin)))
    return out


def scan_sharded(engine, data: np.ndarray, n: int, margin=None, flags=0):
    """Reference implementation of the sharded flow on ONE engine: every shard is
    scanned on its own (as another GPU would), then all raw records are resolved
    together.  Returns the resolved ScanResult."""
    margin = required_margin(engine.image) if margin is None else margin
    raws = []
    for lo, hi, g0, g1 in plan(len(data), n, margin):
        r = engine.scan_host(data[g0:g1], flags=flags, data_offset=g0, block_size=len(data), own=(lo, hi), defer=True)
        if r.rc != 0 or r.error != 0:
            return r
        raws.append(engine.shard_raw())
    raw = np.concatenate(raws) if raws else np.zeros(0, dtype=RAW_DTYPE)
    return engine.resolve_raw(raw, flags=flags)


def gather_raw(raw: np.ndarray, dist, dst=0):
    """Gathers every rank's raw records on `dst` (torch.distributed, any backend
    that supports gather_object: gloo on CPU, or nccl's object path)."""
    world = dist.get_world_size()
    parts = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(raw.tobytes(), parts, dst=dst)
    if dist.get_rank() != dst:
        return None
    return np.concatenate([np.frombuffer(p, dtype=RAW_DTYPE) for p in parts])


# ------------------------------------------------------------------ file batches
def plan_files(sizes, n: int):
    """Deals files out to n ranks, largest first onto the least loaded rank (ties: lowest rank).
    Returns n ascending index lists; every file appears exactly once."""
    load = [0] * n
    out = [[] for _ in range(n)]
    for k in sorted(range(len(sizes)), key=lambda k: (-int(sizes[k]), k)):
        g = min(range(n), key=lambda g: (load[g], g))
        out[g].append(k)
        load[g] += int(sizes[k])
    return [sorted(ix) for ix in out]


def scan_files(engine, files, indices, flags=0, one_allocation=False):
    """One rank's share: scans files[k] for k in indices as one batch; whatever the batch hands
    back as ITEM_RESCAN goes through the general single-buffer path.  Returns {k: yrb_match records}."""
    from .api import ITEM_RESCAN
    items = [files[k] for k in indices]
    res = engine.scan_batch(items, flags=flags, one_allocation=one_allocation)
    if res.rc != 0:
        raise RuntimeError("yrb_scan_batch failed: %d" % res.rc)
    out = {}
    for j, k in enumerate(indices):
        if res.state[j] == ITEM_RESCAN:
            r = engine.scan_host(items[j], flags=flags)
            if r.rc != 0:
                raise RuntimeError("yrb_scan_host failed: %d" % r.rc)
            out[k] = (r.error, r.matches)
        else:
            out[k] = (0, res.matches(j))
    return out


def gather_files(mine: dict, n_files: int, dist, dst=0):
    """Per-file records of every rank -> list indexed by file on `dst` (plain host gather)."""
    world = dist.get_world_size()
    parts = [None] * world if dist.get_rank() == dst else None
    dist.gather_object({k: (e, m.tobytes()) for k, (e, m) in mine.items()}, parts, dst=dst)
    if dist.get_rank() != dst:
        return None
    from .api import YRB_MATCH_DTYPE
    out = [None] * n_files
    for p in parts:
        for k, (e, b) in p.items():
            assert out[k] is None, "file %d scanned twice" % k
            out[k] = (e, np.frombuffer(b, dtype=YRB_MATCH_DTYPE))
    return out
