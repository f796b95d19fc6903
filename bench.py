#!/usr/bin/env python3
"""bench.py -- scanned GB/s of the YARA byte-scanning hot path on B200.

Workload (BASELINE.json configs[2]): 20,000 declared hex strings with
wildcards/jumps (+ nocase/wide text), ONE 16 GiB block of seeded random bytes with
one planted instance per 64 KiB, instances straddling every shard boundary and
chained strings whose parts lie in different shards.  A "step" = one pass of the
hot path (AC scan -> verify -> resolve -> match records back on the host) over
the whole block.

  value   HBM-resident: the block (N = 1) or every rank's byte-range shard with its
          read margins (N > 1) is already in device memory.
  e2e     the same block in PINNED HOST memory scanned through the unchanged
          libyara API of the drop-in (yr_rules_scan_mem -> H2D -> kernels -> D2H
          -> YR_MATCH materialisation), copies inside the timed region.  For N > 1
          it is ONE yr_rules_scan_mem call on rank 0 that the drop-in fans out over
          the N GPUs ($YR_B200_DEVICES -> yrb_multi_scan_host).

N > 1 (torchrun, one process per GPU), default = STRONG scaling of the 16 GiB block
(BASELINE configs[2] as written): rank r owns bytes [r*S/N, (r+1)*S/N) and reads
a margin either side; no data-path collective.  Every step each rank scans and
verifies its shard (yrb_scan_device, defer_resolve) and leaves its raw records in
an export buffer on its GPU; rank 0 pulls them with peer copies (CUDA IPC; only the
record COUNT crosses the host, through a 24-byte shared-memory mailbox: no NCCL, no
pickle) and runs the ONE global resolve (yrb_resolve_raw_device) -- all inside the
timed region.  The weak-scaling figure (16 GiB per GPU, same flow) is measured in
the same run and reported under "weak".  torch.distributed (gloo) is used only for
start/stop barriers and the max-over-ranks of the timings.

What is timed is also checked (config.parity): windows of the block -- its start,
its end and every shard boundary -- are scanned by the reference libyara on the
CPU (oracle/_ref; the oracle restatement if that is absent) and compared bit for
bit with the GPU scan of the same bytes, and the records the timed whole-block /
sharded scan returned inside those windows must be exactly those.

`--impl reference` times the reference libyara (oracle/_ref, built from
/root/reference) on the host cores, all threads, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scanned GB/s (HBM-resident & end-to-end), 20k-string ruleset"
SHORT_ATOMS_RULE = 'rule short_atoms { strings: $mz = "MZ" $pk = { 50 4B 03 } $w = "Zz" wide condition: any of them }\n'
INTERIOR = 16384   # a window's records are compared this far away from its edges (longest chained instance < 4 KiB)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(device),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, names = [], ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for k, nm in enumerate(names):
                    if r[3 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def workload_name(args, S):
    if getattr(args, "short_atoms", False):
        return "synthetic ruleset shape %d, %d declared strings + a rule with 2- and 3-byte atoms, one %.3g GiB block" % (args.config, args.strings, S / 2**30)
    if args.config == 3 and args.strings == 20000:
        return ("configs[2]: 20,000 hex strings with wildcards/jumps + nocase/wide text over one %.3g GiB block, "
                "byte-range sharded with read margins" % (S / 2**30))
    return "synthetic ruleset shape %d, %d declared strings, one %.3g GiB block" % (args.config, args.strings, S / 2**30)


def reference_arm(args):
    """The reference's own CPU implementation of the path, all host threads."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import yrh
    from yara_b200 import synth

    ncores = os.cpu_count() or 1
    rs = synth.ruleset(args.config, args.strings)
    rules = yrh.ref().compile(rs.source)
    sample = int(args.ref_sample_mib) << 20
    data = synth.corpus(rs, sample, 65536, tail_plant=False)
    times = []
    for it in range(args.warmup + args.steps):
        secs, rc = rules.scan_mt(data, ncores)
        assert rc == 0, rc
        if it >= args.warmup:
            times.append(secs)
    t = float(np.mean(times))
    gbs = sample / t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, int(args.size_gib * 2**30)), "strings_declared": args.strings,
                   "strings_compiled": rules.num_strings, "sample_bytes": sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": ncores, "kind": "reference",
                         "sample": "%d MiB of the same corpus, %d threads x yr_rules_scan_mem over contiguous slices" % (args.ref_sample_mib, ncores)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ corpus on the device
def make_shard(api, synth, rs, device, block, g0, g1, every, parts, with_tail):
    """Block bytes [g0, g1) on `device`: counter-based random bytes, one planted instance per cell, the
    boundary plants of SURVEY.md 8(d) row 3, and the edge plants (offset 0, last byte)."""
    def clip(pl):   # the part of every plant that lies inside [g0, g1): shards see identical bytes in their overlaps
        out = []
        for o, b in pl:
            a, e = max(o, g0), min(o + len(b), g1)
            if e > a:
                out.append((a, b[a - o : e - o]))
        return out

    dbuf = api.DeviceBuffer(device, g1 - g0)
    dbuf.fill_random(7, g0)
    n = dbuf.scatter(clip(synth.plants(rs, g1 - g0, every, 11, start=g0, edge=(g0 == 0))), local_start=g0)
    n += dbuf.scatter(clip(synth.boundary_plants(rs, block, parts)), local_start=g0)
    if with_tail and g1 == block and rs.specs:
        import random
        rng = random.Random(0x7A11)
        inst = rs.specs[-1 if rs.specs[-1].kind != "regex" else 0].instance(rng)
        n += dbuf.scatter([(block - len(inst), inst)], local_start=g0)   # ends at the last byte of the block
    return dbuf, n


# ------------------------------------------------------------------ shared-memory mailbox (N > 1)
class Mailbox:
    """What crosses the host between the ranks of a sharded scan: 24 bytes per rank and step.  One small
    file in /dev/shm per rank: u64 [seq, n_records, ack] + the 64-byte CUDA IPC handle of the rank's export
    buffer (written once).  The records themselves go device to device (yrb_shard_export_* /
    yrb_resolve_raw_device).  No collective, no pickle."""

    SIZE = 192

    def __init__(self, tag, rank, world):
        self.rank, self.world = rank, world
        self.paths = ["/dev/shm/yrb_bench_%s_%d" % (tag, r) for r in range(world)]
        with open(self.paths[rank], "wb") as f:   # always fresh: a stale file would carry old sequence numbers
            f.truncate(self.SIZE)
        self.maps = [None] * world

    def open_all(self):
        for r in range(self.world):
            self.maps[r] = np.memmap(self.paths[r], dtype=np.uint8, mode="r+", shape=(self.SIZE,))
        self.hdr = [m[:64].view(np.uint64) for m in self.maps]

    def put_handle(self, handle: bytes):
        self.maps[self.rank][64:128] = np.frombuffer(handle, dtype=np.uint8)

    def get_handle(self, r) -> bytes:
        return bytes(self.maps[r][64:128])

    def publish(self, step, n):
        """rank r > 0: wait until rank 0 has consumed the previous step, then announce this one"""
        h, h0 = self.hdr[self.rank], self.hdr[0]
        while int(h0[2]) < step:     # acks count completed steps
            pass
        h[1] = n
        h[0] = step + 1

    def collect(self, step):
        """rank 0: every other rank's record count of this step"""
        out = []
        for r in range(1, self.world):
            h = self.hdr[r]
            while int(h[0]) < step + 1:
                pass
            out.append(int(h[1]))
        return out

    def ack(self, step):
        self.hdr[0][2] = step + 1

    def close(self):
        self.maps = None
        try:
            os.unlink(self.paths[self.rank])
        except Exception:
            pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size-gib", type=float, default=16.0, help="the block (strong scaling) / bytes owned per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--strings", type=int, default=20000)
    ap.add_argument("--config", type=int, default=3, help="synthetic ruleset shape (BASELINE configs): 2, 3 (headline), 4, 5")
    ap.add_argument("--every", type=int, default=65536, help="one planted instance per this many bytes")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 walk, 2 filter")
    ap.add_argument("--ref-sample-mib", type=int, default=2048)
    ap.add_argument("--cpu-sample-mib", type=int, default=2048)
    ap.add_argument("--window-mib", type=int, default=64, help="size of each parity window")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--short-atoms", action="store_true",
                    help="append a rule with 2- and 3-byte atoms (\"MZ\", { 50 4B 03 }, \"Zz\" wide): the filter kernel's short-atom variant")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-full", action="store_true",
                    help="parity windows tile the WHOLE block (reference scans on all host cores; minutes for 16 GiB)")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--corpus", default="random", choices=["random", "zero", "text"], help="background bytes (plants always)")
    ap.add_argument("--flow", default="multi", choices=["multi", "ranks"],
                    help="N > 1: 'multi' = rank 0 drives all N GPUs through yrb_multi_scan_device (one process, the drop-in's own "
                         "path; the other ranks only wait); 'ranks' = every rank scans its shard, rank 0 pulls the records (CUDA IPC)")
    args = ap.parse_args()

    if args.impl == "reference":
        return reference_arm(args)

    rank, local_rank, world = dist_env()
    dist = None
    if world > 1:
        # read once per process by the drop-in's glue: large one-block scans fan out over these GPUs (the e2e leg)
        os.environ["YR_B200_DEVICES"] = ",".join(str(d) for d in range(world))
        import torch.distributed as dist
        dist.init_process_group("gloo")   # harness only: barriers and the max over ranks; nothing on the data path
        if args.flow == "multi" and rank != 0:
            # the sharded scan is ONE call of one process (yr_rules_scan_mem of the drop-in -> yrb_multi_*): rank 0
            # drives all N GPUs, every other rank waits for it to finish
            dist.barrier()
            dist.destroy_process_group()
            return

    from yara_b200 import Engine, LibYara, DeviceBuffer, PinnedBuffer, synth
    from yara_b200 import api

    device = local_rank
    S = int(args.size_gib * (1 << 30)) & ~4095

    multi_flow = world > 1 and args.flow == "multi"

    def barrier():
        if dist is not None and not multi_flow:
            dist.barrier()

    # ---- rules: compiled on the host by the (unchanged) libyara compiler once, then loaded from the
    # compiled-rules file like any yara deployment does (yarac -> yr_rules_load); the prefilter tables come
    # from the engine's own cache keyed by the rules image
    setup = {}
    t0 = time.time()
    ly = LibYara()
    ly.set_device(device)
    rs = synth.ruleset(args.config, args.strings)
    if args.short_atoms:
        rs.extra = SHORT_ATOMS_RULE
    cache_dir = os.environ.get("YR_B200_CACHE_DIR") or os.path.join(ROOT, "build", "cache")
    os.makedirs(cache_dir, exist_ok=True)
    os.environ.setdefault("YR_B200_CACHE_DIR", cache_dir)
    yarc = os.path.join(cache_dir, "bench_cfg%d_n%d%s.yarc" % (args.config, args.strings, "_short" if args.short_atoms else ""))
    if os.path.exists(yarc):
        rules = ly.load(yarc)
        setup["rules"] = "yr_rules_load(%s)" % os.path.basename(yarc)
    else:
        rules = ly.compile(rs.source)
        setup["rules"] = "compiled"
        if rank == 0:
            try:
                ly.save(rules, yarc + ".tmp%d" % os.getpid())
                os.replace(yarc + ".tmp%d" % os.getpid(), yarc)
            except Exception:
                pass
    setup["rules_s"] = round(time.time() - t0, 3)
    t0 = time.time()
    image = ly.flatten(rules)
    hdr = np.frombuffer(image[:64], dtype=np.uint32)
    eng = Engine(image, device)
    if args.kernel:
        eng.set_option(api.OPT_SCAN_KERNEL, args.kernel)
    multi = None
    if multi_flow:
        multi = api.MultiEngine(image, list(range(world)))
    setup["flatten_upload_s"] = round(time.time() - t0, 3)
    setup["filter_tables_from"] = api.filter_cache_stats()
    setup["prefilter"] = eng.filter_info()

    from yara_b200 import shard as shard_mod
    MARGIN = shard_mod.required_margin(image)
    PARTS = 8   # boundary plants sit at the eighths of the block whatever N is

    def geometry(scaling):
        """(block, own_lo, own_hi, g0, g1) of this rank"""
        if scaling == "weak":
            block = S * world
            lo, hi = rank * S, (rank + 1) * S
        else:
            block = S
            per = ((block + world - 1) // world + 4095) & ~4095
            lo, hi = min(block, rank * per), min(block, (rank + 1) * per)
        return block, lo, hi, max(0, lo - MARGIN), min(block, hi + MARGIN)

    def background(dbuf, g0):
        if args.corpus == "random":
            return
        n = dbuf.size
        if args.corpus == "zero":
            chunk = np.zeros(min(n, 1 << 26), dtype=np.uint8)
        else:
            words = (b"the quick brown fox jumps over the lazy dog while kernel32.dll loads CreateFileW and "
                     b"http://example.com/index.php?id=1234 returns <html><body>Hello, World!</body></html>\r\n")
            chunk = np.frombuffer((words * ((1 << 26) // len(words) + 1))[: min(n, 1 << 26)], dtype=np.uint8)
        for o in range(0, n, len(chunk)):
            dbuf.write(o, chunk[: min(len(chunk), n - o)])

    def build_corpus(scaling):
        block, lo, hi, g0, g1 = geometry(scaling)
        t0 = time.time()
        if multi_flow:
            plan = multi.plan(block)
            bufs, n_pl = [], 0
            for g, (pg0, pg1, po0, po1) in enumerate(plan):
                b, n = make_shard(api, synth, rs, g, block, pg0, pg1, args.every, PARTS, True)
                bufs.append(b)
                n_pl += n
            return (bufs, plan), n_pl, time.time() - t0
        if args.corpus == "random":
            dbuf, n_pl = make_shard(api, synth, rs, device, block, g0, g1, args.every, PARTS, True)
        else:
            dbuf = DeviceBuffer(device, g1 - g0)
            background(dbuf, g0)
            pl = synth.plants(rs, g1 - g0, args.every, 11, start=g0, edge=(g0 == 0))
            n_pl = dbuf.scatter(pl, local_start=g0)
        return dbuf, n_pl, time.time() - t0

    # ---- one sharded step (N > 1): local scan + verify, records to rank 0, ONE global resolve there
    mbox = None
    if world > 1 and not multi_flow:
        tag = os.environ.get("MASTER_PORT", "0")
        mbox = Mailbox(tag, rank, world)
        barrier()
        mbox.open_all()
        mbox.put_handle(eng.shard_export_create(max(1 << 16, int(S // 1024))))
        barrier()
        peer_ptrs = [eng.shard_export_ptr()] + ([eng.shard_import_open(mbox.get_handle(r)) for r in range(1, world)] if rank == 0 else [])

    step_no = [0]
    ev_ms = [0.0]      # rank 0: CUDA-event time of the gather + global resolve
    wait_ms = [0.0]    # rank 0: time spent waiting for the other ranks' record counts (all steps, warm-up included)

    def step_sharded(dbuf, geo, want_records=False):
        block, lo, hi, g0, g1 = geo
        st = step_no[0]
        step_no[0] += 1
        r = eng.scan_device(dbuf, copy=False, data_offset=g0, block_size=block, own=(lo, hi), defer=True)
        assert r.rc == 0 and r.error == 0, (r.rc, r.error)
        n, err = eng.shard_export_publish()      # records -> this rank's export buffer (device to device)
        assert err == 0
        if rank != 0:
            mbox.publish(st, n)
            return r, None, 0.0
        t0 = time.perf_counter()
        counts = [n] + mbox.collect(st)
        wait_ms[0] += (time.perf_counter() - t0) * 1e3
        res = eng.resolve_raw_device(peer_ptrs, counts, copy=want_records)   # peer loads + the ONE global resolve
        mbox.ack(st)
        assert res.rc == 0 and res.error == 0, (res.rc, res.error)
        ev_ms[0] += res.stats["ms_resolve"]
        return r, res, (time.perf_counter() - t0) * 1e3

    def timed_run(dbuf, geo):
        """W warm-up steps, K timed steps; returns dict of timings (this rank) and the last result"""
        last = None
        if multi_flow:
            bufs, plan = dbuf
            block = geo[0]
            for _ in range(args.warmup):
                last = multi.scan_device(bufs, plan, block, copy=False)
                assert last.rc == 0 and last.error == 0, (last.rc, last.error)
            ms_tot, ms_scan, ms_verify, ms_gr, ms_kernel, launches = [], [], [], [], [], 0
            w0 = time.perf_counter()
            for it in range(args.steps):
                last = multi.scan_device(bufs, plan, block, copy=(it == args.steps - 1))
                assert last.rc == 0 and last.error == 0, (last.rc, last.error)
                st = last.stats
                ms_tot.append(st["ms_total"]); ms_scan.append(st["ms_scan"]); ms_verify.append(st["ms_verify"])
                ms_gr.append(st["ms_resolve"]); ms_kernel.append(st["ms_kernel"]); launches += st["kernel_launches"]
            wall = time.perf_counter() - w0
            # per-phase figures: CUDA events on each device's stream, max over devices; the step itself: host clock
            # around the whole call (workers started -> records of the global resolve on the host)
            return dict(seconds=float(np.sum(ms_tot)) / 1e3, wall=wall, scan=ms_scan, verify=ms_verify, resolve=[], kernel=ms_kernel,
                        launches=launches, local=last, final=last, gather_resolve=ms_gr)
        if world == 1:
            for _ in range(args.warmup):
                last = eng.scan_device(dbuf, copy=False)
                assert last.rc == 0 and last.error == 0, (last.rc, last.error)
            barrier()
            ev_ms, ms_scan, ms_verify, ms_resolve, ms_kernel, launches = [], [], [], [], [], 0
            w0 = time.perf_counter()
            for it in range(args.steps):
                last = eng.scan_device(dbuf, copy=(it == args.steps - 1))
                assert last.rc == 0 and last.error == 0, (last.rc, last.error)
                st = last.stats
                ev_ms.append(st["ms_total"]); ms_scan.append(st["ms_scan"]); ms_verify.append(st["ms_verify"])
                ms_resolve.append(st["ms_resolve"]); ms_kernel.append(st["ms_kernel"]); launches += st["kernel_launches"]
            wall = time.perf_counter() - w0
            barrier()
            return dict(seconds=float(np.sum(ev_ms)) / 1e3, wall=wall, scan=ms_scan, verify=ms_verify, resolve=ms_resolve,
                        kernel=ms_kernel, launches=launches, local=last, final=last, gather_resolve=[])
        for _ in range(args.warmup):
            step_sharded(dbuf, geo)
        barrier()
        ms_scan, ms_verify, ms_kernel, ms_gr, launches = [], [], [], [], 0
        final = None
        w0 = time.perf_counter()
        for it in range(args.steps):
            r, res, gr = step_sharded(dbuf, geo, want_records=(it == args.steps - 1))
            st = r.stats
            ms_scan.append(st["ms_scan"]); ms_verify.append(st["ms_verify"]); ms_kernel.append(st["ms_kernel"])
            launches += st["kernel_launches"] + (res.stats["kernel_launches"] if res is not None else 0)
            ms_gr.append(gr)
            last, final = r, (res if res is not None else final)
        wall = time.perf_counter() - w0   # rank 0: ends when the last global resolve has returned its records
        barrier()
        return dict(seconds=wall, wall=wall, scan=ms_scan, verify=ms_verify, resolve=[], kernel=ms_kernel, launches=launches,
                    local=last, final=final, gather_resolve=ms_gr)

    clocks = ClockSampler(device)  # nvidia-smi needs ~0.1 s to start: it runs from the warm-up to the end of the e2e loop

    # ---- the headline run
    geo = geometry(args.scaling)
    dbuf, n_plants, t_corpus = build_corpus(args.scaling)
    setup["corpus_s"] = round(t_corpus, 2)
    T = timed_run(dbuf, geo)

    # ---- parity of what was just timed
    parity = None
    if not args.no_parity and args.corpus == "random":
        parity = check_parity(args, api, synth, rs, ly, rules, image, eng, device, geo, world, rank, T["final"], MARGIN, PARTS, barrier)

    # ---- the other scaling mode, same flow (N > 1 only)
    other = None
    if world > 1 and not args.no_weak:
        other_mode = "weak" if args.scaling == "strong" else "strong"

        def free(d):
            for b in (d[0] if isinstance(d, tuple) else [d]):
                b.free()

        free(dbuf)
        geo2 = geometry(other_mode)
        dbuf2, _, _ = build_corpus(other_mode)
        T2 = timed_run(dbuf2, geo2)
        other = (other_mode, geo2, T2)
        free(dbuf2)
        dbuf = None

    # ---- end to end through the drop-in libyara API, from pinned host memory
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, api, synth, rs, ly, rules, device, geo, world, rank, dbuf, barrier, PinnedBuffer, DeviceBuffer, PARTS)

    clk = clocks.stop()

    # ---- max over ranks
    def maxr(x):
        if dist is None or multi_flow:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    secs = maxr(T["seconds"])
    wall = maxr(T["wall"])
    scan_ms = maxr(float(np.mean(T["scan"])))
    verify_ms = maxr(float(np.mean(T["verify"])))
    kernel_ms = maxr(float(np.mean(T["kernel"])))
    other_secs = maxr(other[2]["seconds"]) if other else None
    parity_ok = maxr(0.0 if (parity is None or parity.get("ok")) else 1.0) == 0.0

    if mbox is not None:
        barrier()
        mbox.close()
    if multi_flow:
        dist.barrier()   # releases the waiting ranks
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    block, lo, hi, g0, g1 = geo
    total_bytes = float(block) * args.steps
    value = total_bytes / secs / 1e9
    scan_kernel = T["local"].stats["scan_kernel"]
    own_bytes = hi - lo
    kernel_gbs = (float(own_bytes) / (kernel_ms / 1e3) / 1e9) if kernel_ms > 0 else float(own_bytes) / (scan_ms / 1e3) / 1e9

    traffic, traffic_src = None, None
    try:  # dram__bytes_read+write of the dominant kernel per algorithmic byte, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if scan_kernel == 2:
            traffic = float(tj["dram_bytes_per_algorithmic_byte"]) * own_bytes
            traffic_src = "profile-derived, not measured in this run: %s" % tj.get("source", "profiles/r02_traffic.json")
    except Exception:
        pass

    final = T["final"]
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": workload_name(args, block),
            "strings_declared": args.strings, "strings_compiled": int(hdr[6]), "ac_slots": int(hdr[4]),
            "block_bytes": block, "bytes_owned_rank0": own_bytes, "margin_bytes": MARGIN, "plants_rank0": n_plants,
            "matches": int(final.n_matches) if final is not None else None,
            "candidates_rank0": int(T["local"].stats["n_candidates"]),
            "scan_kernel": {1: "walk", 2: "filter"}.get(scan_kernel, "?"), "corpus": args.corpus,
            "l2_policy": "each rank's input (%.3g GiB) is larger than L2 (126 MB); no flush needed" % ((g1 - g0) / 2**30),
            "timing": ("CUDA events on the scanner stream per step (sum over steps); wall-clock of the same loop %.4f s" % wall) if world == 1 else
                      (("host clock around every yrb_multi_scan_device call (worker threads started -> records of the global resolve on "
                        "the host), sum over steps; wall-clock of the loop %.4f s; per-phase figures are CUDA events on each device's "
                        "stream, max over devices" % wall) if multi_flow else
                       ("wall-clock of the K-step loop between barriers, max over ranks (rank 0's ends when the last global "
                        "resolve has returned its records); per-phase figures are CUDA events, max over ranks")),
            "flow": ("one process, N devices: yrb_multi_scan_device (rank 0); the other ranks wait" if multi_flow else
                     ("one process per GPU, records pulled by rank 0 over CUDA IPC" if world > 1 else "one device")),
            "setup": setup,
            "ms_breakdown": {"scan": round(scan_ms, 3), "verify": round(verify_ms, 3),
                             "resolve_and_d2h": round(float(np.mean(T["resolve"])), 3) if T["resolve"] else None,
                             "filter_kernel": round(kernel_ms, 3),
                             "gather_and_global_resolve": round(float(np.mean(T["gather_resolve"])), 3) if T["gather_resolve"] else None,
                             "of_which_waiting_for_the_slowest_rank_mean_all_steps": round(wait_ms[0] / max(1, step_no[0]), 3) if (world > 1 and not multi_flow) else None,
                             "gather_and_global_resolve_cuda_events_mean_all_steps": round(ev_ms[0] / max(1, step_no[0]), 3) if (world > 1 and not multi_flow) else None},
            "parity_checked": bool(parity is not None and parity_ok), "parity": parity,
        },
        "clocks": clk,
        "gpu_launches": int(T["launches"]),
        "roofline": {
            "bound": "hbm", "achieved": kernel_gbs, "peak": peak, "unit": "GB/s", "frac": kernel_gbs / peak, "traffic": traffic,
            "traffic_source": traffic_src,
            "kernel": "AC scan kernel (%s), one launch over rank 0's %.3g GiB" % (
                {1: "ac_walk_kernel", 2: "ac_filter_tma_kernel"}.get(scan_kernel, "?"), own_bytes / 2**30),
            "algorithmic_bytes_per_launch": own_bytes, "peak_source": peak_src,
            "timed_with": "CUDA events around the kernel launch on the scanner's stream (yrb_scan_stats.ms_kernel), mean over the timed steps",
            "whole_step_frac": value / world / peak,
        },
    }
    if other:
        mode, geo2, T2 = other
        line[mode] = {"value": float(geo2[0]) * args.steps / other_secs / 1e9, "unit": "GB/s", "block_bytes": geo2[0],
                      "ms_per_step": other_secs / args.steps * 1e3, "scaling": mode,
                      "note": "same sharded flow (local scan + verify, records to rank 0, one global resolve) inside the timed region"}
    if e2e:
        if final is not None:
            d2h = int(final.n_matches) * 24 + 256        # yrb_match records + status
            if world == 1 and int(final.n_matches) >= 8192 and os.environ.get("YR_B200_GATHER_MATCH_DATA", "1") != "0":
                # large single-device results: the match data comes packed from the device as well (yrb_last_match_data)
                ml = np.clip(final.matches["match_length"].astype(np.int64), 0, 512)
                d2h += int(ml.sum()) + (int(final.n_matches) + 1) * 8
            e2e["d2h_bytes_per_step"] = d2h
        else:
            e2e["d2h_bytes_per_step"] = None
        line["e2e"] = e2e

    if not args.no_cpu and world == 1:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import yrh

            ncores = os.cpu_count() or 1
            sample = min(block, int(args.cpu_sample_mib) << 20)
            host = dbuf.read(0, sample)
            rr = yrh.ref().compile(rs.source)
            secs_c, rc = rr.scan_mt(host, ncores)
            secs1, rc1 = rr.scan_mt(host[: sample // 8], 1)
            line["cpu_baseline"] = {
                "value": sample / secs_c / 1e9, "unit": "GB/s", "cores": ncores, "kind": "reference",
                "sample": "first %d MiB of the same corpus; %d threads x yr_rules_scan_mem (reference libyara 4.2.1, gcc -O2) "
                          "over contiguous slices; 1 thread: %.4f GB/s" % (sample >> 20, ncores, sample / 8 / secs1 / 1e9),
            }
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (e,)}

    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------- parity of the timed scan
def check_parity(args, api, synth, rs, ly, rules, image, eng, device, geo, world, rank, final, MARGIN, PARTS, barrier):
    """Rank 0: windows at the start of the block, at its end and across every shard boundary.  Each is
    regenerated on this rank's GPU, scanned there as a block of its own and compared with the reference
    libyara's scan of the same bytes (all fields); then the records of the TIMED scan (whole block at N = 1,
    sharded flow at N > 1) that lie inside the window's interior must be exactly the window scan's."""
    if rank != 0:
        return {"ok": True}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import yrh

    block = geo[0]
    W = min(int(args.window_mib) << 20, block)
    per = ((block + PARTS - 1) // PARTS + 4095) & ~4095
    starts = [0, max(0, block - W)]
    if world > 1:
        wper = ((block + world - 1) // world + 4095) & ~4095
        starts += [max(0, min(block - W, k * wper - W // 2)) for k in range(1, world)]
    else:
        starts += [max(0, min(block - W, (PARTS // 2) * per - W // 2))]
    if getattr(args, "parity_full", False):   # every byte of the block lies in some window's interior
        starts += list(range(0, max(1, block - W + 1), W - 2 * INTERIOR))
    starts = sorted(set(s & ~4095 for s in starts))

    if os.path.exists(yrh.ref_path()):
        checker, ref_rules = "reference libyara 4.2.1 (oracle/_ref)", yrh.ref().compile(rs.source)
    else:
        checker, ref_rules = "oracle restatement (oracle/yr_oracle.c)", None
    drop = yrh.b200().compile(rs.source) if ref_rules is not None else None

    fm = final.matches
    out = {"checker": checker, "windows": [], "ok": True, "whole_block_matches": int(final.n_matches)}
    t0 = time.time()
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) - 1)))
    ref_extra = [ref_rules] + ([yrh.ref().compile(rs.source) for _ in range(7)] if (ref_rules is not None and len(starts) > 8) else [])

    def prepare(k):
        """device work of window k (serial) + its reference scan handed to the pool (one compiled ruleset per slot:
        a YR_RULES default scanner is not shared between threads)"""
        w0 = starts[k]
        w1 = min(block, w0 + W)
        wbuf, _ = make_shard(api, synth, rs, device, block, w0, w1, args.every, PARTS, True)
        host = wbuf.read(0, w1 - w0)
        g = eng.scan_device(wbuf, copy=True)
        assert g.rc == 0 and g.error == 0
        wbuf.free()
        b = drop.scan(host, cap=1 << 20) if ref_rules is not None else None   # through the drop-in: every YR_MATCH field incl. the data bytes
        rr = ref_extra[k % len(ref_extra)] if ref_rules is not None else None
        fut = pool.submit(rr.scan, host, 0, 0, 0, 1 << 20) if rr is not None else None
        return w0, w1, host, g.matches, b, fut

    DEPTH = len(ref_extra)
    pending = []
    nxt = 0
    total_ref = 0
    while nxt < len(starts) or pending:
        while nxt < len(starts) and len(pending) < DEPTH:
            pending.append(prepare(nxt))
            nxt += 1
        w0, w1, host, gm, b, fut = pending.pop(0)
        if ref_rules is not None:
            a = fut.result()
            same = (a.rc == b.rc and a.matches.tobytes() == b.matches.tobytes() and a.rule_bits.tobytes() == b.rule_bits.tobytes())
            n_ref = len(a.matches)
            # and the engine's own records for the window are the reference's (string, offset, length, xor key)
            ka = sorted(zip(a.matches["string_idx"].tolist(), a.matches["offset"].tolist(), a.matches["match_length"].tolist(), a.matches["xor_key"].tolist()))
            kg = sorted(zip(gm["string_idx"].tolist(), gm["offset"].tolist(), gm["match_length"].tolist(), gm["xor_key"].tolist()))
            same = same and ka == kg
        else:
            o = yrh.oracle_scan(image, host)
            om = yrh.matches_from_flat(o.matches, host)
            n_ref = len(om)
            same = len(om) == len(gm) and all((om[f] == gm[f]).all() for f in ("offset", "string_idx", "match_length", "xor_key"))
        # the timed result inside the window's interior
        lo_i = w0 + (INTERIOR if w0 > 0 else 0)
        hi_i = w1 - (INTERIOR if w1 < block else 0)
        sel = (fm["offset"] >= lo_i) & (fm["offset"] + fm["match_length"] <= hi_i)
        t = fm[sel]
        selw = (gm["offset"] + w0 >= lo_i) & (gm["offset"] + w0 + gm["match_length"] <= hi_i)
        wloc = gm[selw]
        key_t = sorted(zip(t["string_idx"].tolist(), t["offset"].tolist(), t["match_length"].tolist(), t["xor_key"].tolist()))
        key_w = sorted(zip(wloc["string_idx"].tolist(), (wloc["offset"] + w0).tolist(), wloc["match_length"].tolist(), wloc["xor_key"].tolist()))
        inside = key_t == key_w
        total_ref += int(n_ref)
        if len(starts) <= 16 or not (same and inside):
            out["windows"].append({"bytes": [w0, w1], "reference_matches": int(n_ref), "window_equals_reference": bool(same),
                                   "timed_scan_records_in_interior": len(key_t), "timed_scan_equals_window": bool(inside)})
        out["ok"] = out["ok"] and bool(same) and bool(inside)
    pool.shutdown()
    out["n_windows"] = len(starts)
    out["window_bytes"] = W
    out["covers_whole_block"] = bool(getattr(args, "parity_full", False))
    out["reference_matches_all_windows"] = total_ref
    out["seconds"] = round(time.time() - t0, 1)
    return out


# ------------------------------------------------------------------------- end to end
def run_e2e(args, api, synth, rs, ly, rules, device, geo, world, rank, dbuf, barrier, PinnedBuffer, DeviceBuffer, PARTS):
    """yr_rules_scan_mem of the drop-in on the whole block in pinned host memory.  N > 1: ONE call on rank 0
    that the drop-in fans out over the N GPUs; the other ranks wait."""
    block = geo[0] if args.scaling == "strong" else int(args.size_gib * (1 << 30)) & ~4095
    out = None
    if rank == 0:
        hbuf = PinnedBuffer(block)
        # the block's bytes, produced on this rank's GPU piece by piece (identical to what the shards hold)
        piece = 1 << 30
        if world == 1 and dbuf is not None and dbuf.size == block:
            dbuf.read_into(0, block, hbuf.ptr)
        else:
            for o in range(0, block, piece):
                e = min(block, o + piece)
                pb, _ = make_shard(api, synth, rs, device, block, o, e, args.every, PARTS, True)
                pb.read_into(0, e - o, hbuf.ptr + o)
                pb.free()
        rc, nrules = ly.scan_mem(rules, hbuf.ptr, block)  # warm-up (also sizes the scanners' device buffers)
        assert rc == 0, rc
        rc, nrules = ly.scan_mem(rules, hbuf.ptr, block)
        assert rc == 0, rc
        w0 = time.perf_counter()
        for _ in range(args.steps):
            rc, nrules = ly.scan_mem(rules, hbuf.ptr, block)
            assert rc == 0, rc
        secs = time.perf_counter() - w0
        out = {"value": float(block) * args.steps / secs / 1e9, "unit": "GB/s", "h2d_bytes_per_step": block,
               "d2h_bytes_per_step": None, "rules_matching": nrules, "seconds": secs,
               "api": "yr_rules_scan_mem (drop-in libyara_b200.so), pinned host block" +
                      ("" if world == 1 else "; one call on rank 0 fanned out over %d GPUs (YR_B200_DEVICES -> yrb_multi_scan_host)" % world)}
        if world == 1:
            # what PCIe gives a bare copy of the same pinned block
            try:
                tmp = dbuf if (dbuf is not None and dbuf.size == block) else DeviceBuffer(device, block)
                tmp.write(0, hbuf.array[: min(block, 1 << 30)])
                w0 = time.perf_counter()
                tmp.write(0, hbuf.array)
                out["h2d_copy_gbs"] = float(block) / (time.perf_counter() - w0) / 1e9
                out["frac_of_h2d_copy"] = out["value"] / out["h2d_copy_gbs"]
            except Exception:
                pass
        hbuf.free()
    barrier()
    return out


if __name__ == "__main__":
    main()
