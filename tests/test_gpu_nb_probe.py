"""GPU: first hardware run of the opt-in negative-binomial path (tests/test_gpu_nb_experimental.py), in a CHILD process.

The kernel instantiation hfg_estep_kernel<THREADS, true> and its host fold were written when no GPU time was left, so the
path is switched off by default (hfg_create wants HFG_EXPERIMENTAL_NB=1) and its tests skip.  This probe runs them in a
child pytest with the variable set: a CUDA error there cannot touch this process's context, a hang is cut by the timeout.
It is marked xfail(strict=False): a failure is the documented state ("not validated yet"), a pass shows up as XPASS --
the signal that the switch can go."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="negative-binomial device path: not validated on hardware yet (DESIGN.md section 7)")
def test_negative_binomial_path_in_a_child_process():
    env = dict(os.environ, HFG_EXPERIMENTAL_NB="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_nb_experimental.py", "-m", "gpu", "-q", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-4000:]  # (all skipped == nothing ran == not a pass)
