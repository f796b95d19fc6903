"""BASELINE configs[0] (plumbing): the reference `yara` CLI over sample.rules /
sample.file, and the same unmodified CLI linked against the GPU drop-in printing
the same matches (rule names, offsets, string ids, matched bytes) for a synthetic
ruleset as the reference CLI."""
import os
import subprocess

import numpy as np
import pytest

from yara_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def cli(exe, *args):
    path = os.path.join(REFDIR, exe)
    if not os.path.exists(path):
        pytest.skip(exe + " not built")
    return subprocess.run([path] + list(args), capture_output=True, text=True, timeout=600)


def test_config0_reference_cli_plumbing():
    rules, target = os.path.join(REFDIR, "sample", "sample.rules"), os.path.join(REFDIR, "sample", "sample.file")
    if not os.path.exists(rules):
        pytest.skip("sample files not snapshotted")
    p = cli("yara_ref", "-s", "-p", "4", rules, target)
    assert p.returncode == 0, p.stderr


@pytest.mark.gpu
def test_cli_through_gpu_matches_reference_cli(tmp_path):
    rs = synth.ruleset(4, 1500)
    rules = tmp_path / "r.yar"
    rules.write_text(rs.source)
    files = []
    for k in range(4):
        f = tmp_path / ("blob%d.bin" % k)
        synth.corpus(rs, (1 << 20) + 4096 * k, 4096, seed=50 + k, plant_seed=k).tofile(f)
        files.append(str(f))
    for f in files:
        a = cli("yara_ref", "-s", str(rules), f)
        b = cli("yara_b200cli", "-s", str(rules), f)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
        assert len(a.stdout) > 1000
        assert a.stdout == b.stdout
    # directory scan with 4 scanner threads (file queue + N scanners, cli/yara.c:1246-1293)
    a = cli("yara_ref", "-p", "4", str(rules), str(tmp_path))
    b = cli("yara_b200cli", "-p", "4", str(rules), str(tmp_path))
    assert sorted(a.stdout.splitlines()) == sorted(b.stdout.splitlines()) and a.returncode == b.returncode == 0
