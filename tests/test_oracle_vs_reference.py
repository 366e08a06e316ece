"""CPU, build container only: pins the oracle restatements (oracle/cl_losses.py) to the REFERENCE'S OWN unmodified
files, imported from /root/reference with the nnunet shim on sys.path, and re-checks the committed golden fixtures.
Skipped where the reference is not mounted (the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

import util

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")


@pytest.fixture(scope="module")
def refvals():
    for p in (os.path.join(util.ROOT, "oracle", "shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import gen_golden
    case = gen_golden.seeded_case()
    return case, gen_golden.reference_values(case)


def test_golden_fixture_is_current(refvals):
    """tests/golden/cl_losses.npz == what the reference produces today"""
    _, vals = refvals
    gold = np.load(os.path.join(util.ROOT, "tests", "golden", "cl_losses.npz"))
    assert set(gold.files) == set(vals.keys())
    for k in gold.files:
        np.testing.assert_allclose(gold[k], vals[k], rtol=1e-6, atol=1e-7, equal_nan=True, err_msg=k)


def test_reference_own_structural_test_passes_on_the_shim():
    """the reference's test/network_architecture/test_MultiHead_Module.py runs unmodified against the shim"""
    import subprocess
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(util.ROOT, "oracle", "shim"), REF]))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        os.path.join(REF, "test", "network_architecture", "test_MultiHead_Module.py")],
                       env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
