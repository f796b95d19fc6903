#!/usr/bin/env python3
"""bench.py -- scanned GB/s of the YARA byte-scanning hot path on B200.

Workload (BASELINE.json configs[2]): 20,000 declared hex strings with
wildcards/jumps (+ nocase/wide text), 16 GiB of seeded random bytes with one
planted instance per 64 KiB.  A "step" = one pass of the hot path (AC scan ->
verify -> resolve -> match records back on the host) over the whole block.

  value   HBM-resident: the block is already in device memory (yrb_scan_device)
  e2e     the same block in PINNED HOST memory scanned through the unchanged
          libyara API of the drop-in (yr_rules_scan_mem -> H2D -> kernels -> D2H
          -> YR_MATCH materialisation), copies inside the timed region

N > 1 (torchrun, one process per GPU): byte-range sharding with read margins,
no data-path collective; rank r owns bytes [r*S, (r+1)*S) of one N*S-byte block
(weak scaling: S = --size-gib per GPU).  torch.distributed is used only for the
barrier and the max-over-ranks of the timings.

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
MARGIN = 8192  # >= YR_RE_SCAN_LIMIT (4096) and 2x the longest literal, plus the 3-byte AC look-back


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
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[2]: 20k hex/wild/jump + nocase/wide strings", "strings_declared": args.strings,
                   "strings_compiled": rules.num_strings, "sample_bytes": sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": ncores, "kind": "reference",
                         "sample": "%d MiB of the same corpus, %d threads x yr_rules_scan_mem over contiguous slices" % (args.ref_sample_mib, ncores)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size-gib", type=float, default=16.0, help="bytes owned per GPU")
    ap.add_argument("--strings", type=int, default=20000)
    ap.add_argument("--config", type=int, default=3, help="synthetic ruleset shape (BASELINE configs): 2, 3 (headline), 4, 5")
    ap.add_argument("--every", type=int, default=65536, help="one planted instance per this many bytes")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 walk, 2 filter")
    ap.add_argument("--ref-sample-mib", type=int, default=2048)
    ap.add_argument("--cpu-sample-mib", type=int, default=2048)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        return reference_arm(args)

    rank, local_rank, world = dist_env()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from yara_b200 import Engine, LibYara, DeviceBuffer, PinnedBuffer, synth
    from yara_b200 import api

    device = local_rank
    S = int(args.size_gib * (1 << 30)) & ~4095
    block = S * world
    own_lo, own_hi = rank * S, (rank + 1) * S
    g0, g1 = max(0, own_lo - MARGIN), min(block, own_hi + MARGIN)

    # ---- rules: compiled on the host by the (unchanged) libyara compiler, uploaded once
    t0 = time.time()
    ly = LibYara()
    ly.set_device(device)
    rs = synth.ruleset(args.config, args.strings)
    rules = ly.compile(rs.source)
    image = ly.flatten(rules)
    hdr = np.frombuffer(image[:64], dtype=np.uint32)
    eng = Engine(image, device)
    if args.kernel:
        eng.set_option(api.OPT_SCAN_KERNEL, args.kernel)
    t_compile = time.time() - t0

    # ---- corpus, generated on the device: counter-based random bytes + planted instances
    t0 = time.time()
    dbuf = DeviceBuffer(device, g1 - g0)
    dbuf.fill_random(7, g0)
    pl = synth.plants(rs, g1 - g0, args.every, 11, start=g0, edge=(rank == 0))
    n_plants = dbuf.scatter(pl, local_start=g0)
    t_corpus = time.time() - t0

    kw = dict(data_offset=g0, block_size=block, own=(own_lo, own_hi)) if world > 1 else {}

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- HBM-resident
    clocks = ClockSampler(device)  # nvidia-smi needs ~0.1 s to start: it runs from the warm-up to the end of the e2e loop
    for _ in range(args.warmup):
        r = eng.scan_device(dbuf, copy=False, **kw)
        assert r.rc == 0 and r.error == 0, (r.rc, r.error)
    barrier()
    ms_total, ms_scan, ms_verify, ms_resolve, launches, wall0 = [], [], [], [], 0, time.time()
    for _ in range(args.steps):
        r = eng.scan_device(dbuf, copy=False, **kw)
        assert r.rc == 0 and r.error == 0, (r.rc, r.error)
        ms_total.append(r.stats["ms_total"])
        ms_scan.append(r.stats["ms_scan"])
        ms_verify.append(r.stats["ms_verify"])
        ms_resolve.append(r.stats["ms_resolve"])
        launches += r.stats["kernel_launches"]
    wall = time.time() - wall0
    barrier()
    n_matches, n_cand, scan_kernel = r.n_matches, r.stats["n_candidates"], r.stats["scan_kernel"]
    dev_s = float(np.sum(ms_total)) / 1e3

    # ---- end to end through the drop-in libyara API, from pinned host memory
    e2e = None
    if not args.no_e2e:
        hbuf = PinnedBuffer(S)
        dbuf.read_into(own_lo - g0, S, hbuf.ptr)
        rc, nrules = ly.scan_mem(rules, hbuf.ptr, S)  # warm-up (also sizes the scanner's device buffers)
        assert rc == 0, rc
        barrier()
        w0 = time.time()
        for _ in range(args.steps):
            rc, nrules = ly.scan_mem(rules, hbuf.ptr, S)
            assert rc == 0, rc
        e2e_s = time.time() - w0
        barrier()
        e2e = {"seconds": e2e_s, "rules_matching": nrules}
        # what PCIe gives a bare copy of the same pinned block (the bytes are already what the device holds)
        try:
            dbuf.write(own_lo - g0, hbuf.array[: min(S, 1 << 30)])
            w0 = time.time()
            dbuf.write(own_lo - g0, hbuf.array)
            e2e["h2d_gbs"] = float(S) / (time.time() - w0) / 1e9
        except Exception:
            pass

    clk = clocks.stop()

    # ---- max over ranks
    if dist is not None:
        import torch

        t = torch.tensor([dev_s, wall, e2e["seconds"] if e2e else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, wall, e2e_max = [float(x) for x in t.tolist()]
        if e2e:
            e2e["seconds"] = e2e_max

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    traffic = None
    try:  # dram__bytes_read+write of the dominant kernel, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if scan_kernel == 2:
            traffic = float(tj["dram_bytes_per_algorithmic_byte"]) * S
    except Exception:
        pass
    total_bytes = float(S) * world * args.steps
    value = total_bytes / dev_s / 1e9
    scan_gbs = float(S) / (float(np.mean(ms_scan)) / 1e3) / 1e9

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": ("configs[2]: 20,000 hex strings with wildcards/jumps + nocase/wide text over %.3g GiB per GPU, "
                         "byte-range sharded with %d-byte read margins" % (S / 2**30, MARGIN)) if args.config == 3 and args.strings == 20000
                        else "synthetic ruleset shape %d, %d declared strings, %.3g GiB per GPU" % (args.config, args.strings, S / 2**30),
            "strings_declared": args.strings, "strings_compiled": int(hdr[6]), "ac_slots": int(hdr[4]),
            "bytes_per_gpu": S, "plants_per_gpu": n_plants, "matches_rank0": int(n_matches), "candidates_rank0": int(n_cand),
            "scan_kernel": {1: "walk", 2: "filter"}.get(scan_kernel, "?"),
            "l2_policy": "input (%.3g GiB) is larger than L2 (126 MB); no flush needed" % (S / 2**30),
            "timing": "CUDA events on the scanner stream per step (sum over steps), max over ranks; wall-clock %.4f s" % wall,
            "setup_s": {"compile_upload": round(t_compile, 2), "corpus": round(t_corpus, 2)},
            "ms_breakdown": {"scan": round(float(np.mean(ms_scan)), 3), "verify": round(float(np.mean(ms_verify)), 3),
                             "resolve_and_d2h": round(float(np.mean(ms_resolve)), 3)},
        },
        "clocks": clk,
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "achieved": scan_gbs, "peak": peak, "unit": "GB/s", "frac": scan_gbs / peak, "traffic": traffic,
            "kernel": "AC scan kernel (%s)" % {1: "ac_walk_kernel", 2: "ac_filter_kernel"}.get(scan_kernel, "?"),
            "algorithmic_bytes_per_launch": S, "peak_source": peak_src,
        },
    }
    if e2e:
        line["e2e"] = {
            "value": float(S) * world * args.steps / e2e["seconds"] / 1e9, "unit": "GB/s",
            "h2d_bytes_per_step": S, "d2h_bytes_per_step": int(n_matches) * 24 + 256,
            "api": "yr_rules_scan_mem (drop-in libyara_b200.so), pinned host block", "rules_matching": e2e["rules_matching"],
        }
        if "h2d_gbs" in e2e and world == 1:
            line["e2e"]["h2d_copy_gbs"] = e2e["h2d_gbs"]   # bare cudaMemcpy of the same pinned block, same run
            line["e2e"]["frac_of_h2d_copy"] = line["e2e"]["value"] / e2e["h2d_gbs"]

    if not args.no_cpu and world == 1:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import yrh

            ncores = os.cpu_count() or 1
            sample = min(S, int(args.cpu_sample_mib) << 20)
            host = hbuf.array[:sample] if e2e else dbuf.read(own_lo - g0, sample)
            rr = yrh.ref().compile(rs.source)
            secs, rc = rr.scan_mt(host, ncores)
            secs1, rc1 = rr.scan_mt(host[: sample // 8], 1)
            line["cpu_baseline"] = {
                "value": sample / secs / 1e9, "unit": "GB/s", "cores": ncores, "kind": "reference",
                "sample": "first %d MiB of the same corpus; %d threads x yr_rules_scan_mem (reference libyara 4.2.1, gcc -O2) "
                          "over contiguous slices; 1 thread: %.4f GB/s" % (sample >> 20, ncores, sample / 8 / secs1 / 1e9),
            }
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (e,)}

    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
