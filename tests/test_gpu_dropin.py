"""The drop-in layer (include/libecc_b200_dropin.h) exercised with REAL reference structs by the C harness
tests/dropin/dropin_harness.c (built against the reference's headers in the build container, shipped prebuilt):
  direct   drop-in prj_pt_mul / prj_pt_mul_blind / batch / ECDSA verify_batch vs the reference's own functions;
  preload  the unmodified reference's ec_sign / ec_verify / ECC-CDH running with prj_pt_mul interposed by the GPU."""
import os
import subprocess

import pytest

from common import ROOT

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "oracle", "_ref", "dropin_harness")
DROPIN = os.path.join(ROOT, "libecc_b200", "libecc_b200_dropin.so")


def _need():
    if not os.path.exists(HARNESS):
        pytest.fail("oracle/_ref/dropin_harness is missing: run `make -C oracle all` where /root/reference exists")
    assert os.path.exists(DROPIN), "libecc_b200_dropin.so not built"


def test_dropin_direct_against_reference_structs():
    _need()
    r = subprocess.run([HARNESS, "direct", DROPIN], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "HARNESS OK" in r.stdout


def test_reference_code_runs_on_gpu_through_interposed_prj_pt_mul():
    _need()
    env = dict(os.environ, LD_PRELOAD=DROPIN)
    r = subprocess.run([HARNESS, "preload"], capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "HARNESS OK" in r.stdout
