import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_terminal_summary(terminalreporter):
    """Print what the parity assertions measured (tests/parity.py) so that the GPU test record shows the deviations
    and the reference's own fp32 noise floor, not just PASSED; also written to gpurun_out/parity_report.txt."""
    import parity
    if not parity.LOG:
        return
    terminalreporter.section("parity measurements (policy: tests/parity.py)")
    for line in parity.LOG:
        terminalreporter.write_line(line)
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.txt"), "w") as f:
            f.write("\n".join(parity.LOG) + "\n")
    except OSError:
        pass


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


@pytest.fixture
def golden():
    return load_golden


def coo_to_csr(indices, values, S):
    """Coalesced COO (rows sorted) -> rowptr, col, W."""
    rows = np.asarray(indices[0])
    rowptr = np.zeros(S + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr), np.asarray(indices[1]), np.asarray(values)
