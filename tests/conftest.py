"""Shared pytest fixtures.

`-m "not gpu"`: oracle vs the reference's known answers, host logic, C-ABI symbol checks (no GPU needed).
`-m gpu`: parity of the CUDA path (through the C-ABI / host CLI) against the oracle.
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
FIX = ROOT / "tests" / "golden" / "fixtures"
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device")


@pytest.fixture(scope="session")
def oracle_bin():
    """Build (if needed) and return the path of the CPU oracle executable (test infrastructure only)."""
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)
    p = ROOT / "oracle" / "mosdepth_oracle"
    assert p.exists()
    return p


@pytest.fixture(scope="session")
def fixtures():
    return FIX


def run_tool(exe, args, cwd, env=None, check=True):
    e = dict(os.environ)
    for k in list(e):
        if k.startswith("MOSDEPTH_"):
            del e[k]
    if env:
        e.update(env)
    r = subprocess.run([str(exe)] + [str(a) for a in args], cwd=str(cwd), env=e, capture_output=True, text=True)
    if check:
        assert r.returncode == 0, f"{exe} {args} -> rc={r.returncode}\n{r.stderr}"
    return r
