import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def pytest_sessionstart(session):
    """Make sure the native pieces exist (the driver normally runs __graft_entry__.build() first): the CUDA
    library + CLI (nvcc cross-compiles without a GPU) and the oracle's C library. Never rebuilds when the
    artefacts are up to date, and never falls back to anything if the build fails - the tests then fail."""
    try:
        from phylocsf_b200 import build as b

        if b.stale():
            b.build()
    except Exception as e:  # noqa: BLE001
        print("WARNING: could not build libphylocsf_b200.so: %s" % e)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import subprocess

        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


@pytest.fixture(scope="session")
def params_base(tmp_path_factory):
    """A $PHYLOCSF_BASE-like directory re-emitted from tests/golden (reference file formats)."""
    from tools import golden_params

    d = str(tmp_path_factory.mktemp("phylocsf_base"))
    golden_params.materialize(d)
    golden_params.write_examples(d)
    return d
