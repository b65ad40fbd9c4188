import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers_modest import GOLDEN_CASES, GOLDEN_DIR, build_case, load_golden  # noqa: E402,F401


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the shared library is a build artefact (git-ignored): build it when a fresh tree has none
    lib = os.path.join(ROOT, "modest_b200", "libmodest_b200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["bash", os.path.join(ROOT, "modest_b200", "csrc", "build.sh")], check=True)


_case_cache = {}


@pytest.fixture(scope="session")
def golden_case():
    """callable name -> (case, shape, golden npz); inputs are regenerated from the seed and
    checked against the digests stored with the golden outputs."""
    import hashlib

    def sha(a):
        a = np.ascontiguousarray(a)
        return hashlib.sha256(a.tobytes() + str(a.dtype).encode() + str(a.shape).encode()).hexdigest()

    def get(name):
        if name not in _case_cache:
            case, shape = build_case(name)
            g = load_golden(name)
            assert sha(case.query) == str(g["query_sha"]), "synthetic generator drifted from the golden inputs"
            assert sha(np.concatenate(case.history)) == str(g["history_sha"])
            _case_cache[name] = (case, shape, g)
        return _case_cache[name]
    return get
