"""Randomised parity sweep (tools/fuzz_probe.py) as part of the GPU suite: ten seconds of random
small texts (random, periodic, long runs, repeated segments; alphabets of 1..255 letters) through
every entry point, compared bit for bit with the oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_random_sweep(engine, seed):
    env = dict(os.environ)
    if seed % 2 == 0:  # chain offsets in every doubling round, bucket overflow always handled as shallow groups
        env.update({"B200SA_CHAIN_MIN_FRAC": "100000000", "B200SA_CHAIN_USE_FRAC": "100000000", "B200SA_MSD_MORE_FRAC": "1", "B200SA_MSD_OVER_FRAC": "1"})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_probe.py"), "5", str(seed)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fuzz ok" in r.stdout
