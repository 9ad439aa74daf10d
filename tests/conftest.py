import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import sdt_pkg
    return sdt_pkg.load()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def tiny_transcriptome(pkg):
    return pkg.synth.make_transcriptome(20, 7)


def make_dataset(pkg, tr, n_pairs, read_len, seed, ragged=0, n_rate=0):
    reads, lens = pkg.synth.make_reads(tr, n_pairs, read_len, seed, ragged=ragged)
    if n_rate:
        rng = np.random.default_rng(seed)
        reads = reads.copy()
        reads[rng.random(reads.shape) < n_rate] = 4
    return reads, lens
