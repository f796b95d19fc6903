"""The reference's OWN known-answer test programs (tests/test-rules.c -- ~1,500
assertions over strings, hex strings, regexps, modifiers, counts/offsets/lengths,
run in 3 passes incl. 1024-byte blocks with 256-byte overlap --, test-api.c,
test-async.c), compiled unmodified by oracle/Makefile and linked against the
drop-in libyara_b200.so, so every scan they make goes through the GPU path.
A test program exits non-zero on its first failed assertion."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def run(name, timeout=1500):
    exe = os.path.join(REFDIR, name)
    if not os.path.exists(exe):
        pytest.skip("%s not built (needs /root/reference at build time)" % name)
    env = dict(os.environ, TOP_SRCDIR=os.path.join(REFDIR, "top"))
    p = subprocess.run([exe], cwd=REFDIR, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, "%s failed (rc=%d):\n%s" % (name, p.returncode, tail)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test-rules_b200", "test-api_b200", "test-async_b200"])
def test_reference_program_on_gpu(name):
    run(name)


@pytest.mark.parametrize("name", ["test-api_ref", "test-async_ref"])
def test_reference_program_on_reference(name):
    """sanity of the build recipe itself (CPU, reference library)"""
    run(name, timeout=600)
