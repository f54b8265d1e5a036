"""The drop-in layer (include/libecc_b200_dropin.h: prj_pt_mul, ec_verify, ec_verify_batch and the verify_batch slots of all
twelve short-Weierstrass schemes) exercised with REAL reference structs by the C harness
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


@pytest.mark.parametrize("scheme,curve", [("ECFSDSA", "FRP256V1"), ("BIP0340", "SECP256K1"), ("BIP0340", "SECP224R1"),
                                          ("ECFSDSA", "SECP384R1")])
def test_verify_batch_adapters_use_the_multi_scalar_fast_path(scheme, curve):
    """ECFSDSA / BIP0340 verify_batch adapters on real structs: a valid batch is settled by ONE multi-scalar
    multiplication (K6), a batch with invalid signatures falls through to the per-item kernel, which names them; the
    harness checks every verdict.  BIP0340 on SECP224R1 (p = 1 mod 4) is not served by K6 and stays per item."""
    _need()
    import json
    env = dict(os.environ, ECCB200_DROPIN_MSM_MIN="256")
    for invalid_every, settled in ((0, 0 if curve == "SECP224R1" else 3), (64, 0)):
        r = subprocess.run([HARNESS, "bench", DROPIN, curve, "4096", scheme, str(invalid_every)], capture_output=True,
                           text=True, timeout=900, env=env)
        print(r.stdout[-2000:], r.stderr[-1000:])
        assert r.returncode == 0 and "HARNESS OK" in r.stdout
        line = json.loads(r.stdout.split("DROPIN_BENCH ", 1)[1].splitlines()[0])
        assert line["wrong_verdicts"] == 0
        # all four timed calls of a valid batch are settled by the fast path, none of a batch with invalid signatures
        assert line["batches_settled_by_multi_scalar_multiplication"] == (4 if settled else 0)


@pytest.mark.parametrize("scheme", ["ECGDSA", "SM2", "BIGN"])
def test_double_scalar_scheme_adapters_on_a_large_batch(scheme):
    """The verify_batch adapters of the double-scalar schemes on 40 000 real ec_pub_key structs (the big engine slot,
    several marshalling threads, projective keys normalised on the device, 1 signature in 64 corrupted): every verdict
    equals the expected one.  The eight-curve, item-by-item comparison with the reference's ec_verify is the direct mode
    above; tests/test_dropin_host.py replays it on the host build."""
    _need()
    import json
    env = dict(os.environ, HARNESS_POOL="256", HARNESS_REPS="2")
    r = subprocess.run([HARNESS, "bench", DROPIN, "FRP256V1", "40000", scheme, "64"], capture_output=True, text=True,
                       timeout=600, env=env)
    print(r.stdout[-2000:], r.stderr[-1000:])
    assert r.returncode == 0 and "HARNESS OK" in r.stdout
    line = json.loads(r.stdout.split("DROPIN_BENCH ", 1)[1].splitlines()[0])
    assert line["wrong_verdicts"] == 0 and line["items"] == 40000
