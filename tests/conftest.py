import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _oracle_lib():
    """The C oracle is test infrastructure; build it if the prebuilt library did not travel."""
    so = os.path.join(ROOT, "oracle", "_build", "libgnss_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return so


def pytest_terminal_summary(terminalreporter):
    """Observed parity of the closed-loop tracking comparisons (tests/helpers.windowed_iq_compare): how many epochs the 1e-6
    window covered and the worst error inside / after it; also written to gpurun_out/ for profiles/."""
    try:
        from helpers import PARITY_REPORT
    except Exception:
        return
    if not PARITY_REPORT:
        return
    lines = ["%-72s window %6d / %6d   inside %.2e   after %.2e" % r for r in PARITY_REPORT]
    terminalreporter.write_sep("-", "closed-loop parity windows (1e-6 of |P| inside)")
    for ln in lines:
        terminalreporter.write_line(ln)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_windows.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
    except OSError:
        pass
