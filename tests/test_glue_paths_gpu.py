"""The shape-specialised actuation kernel (`actuate_conv_kernel`, csrc/glue.cuh) must reproduce the generic
`actuate_kernel` BIT FOR BIT: same MLP accumulation order, same `(power * a) * w` gather in ascending actuator
order, padding taps contribute exact zeros.  The library picks the kernel once per process
(PDEB200_ACTUATE_GENERIC), so each variant runs in its own interpreter."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent


def _run(tmp_path, generic):
    out = tmp_path / ("generic.npz" if generic else "conv.npz")
    env = dict(os.environ, PDEB200_ACTUATE_GENERIC="1" if generic else "0")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "helpers" / "glue_paths_dump.py"), str(out)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


def test_specialised_actuation_equals_generic_bitwise(tmp_path):
    a, b = _run(tmp_path, generic=False), _run(tmp_path, generic=True)
    assert sorted(a.files) == sorted(b.files) and len(a.files) == 18
    for k in a.files:
        assert a[k].shape == b[k].shape and np.isfinite(a[k]).all(), k
        assert np.array_equal(a[k], b[k]), (k, float(np.abs(a[k] - b[k]).max()))
