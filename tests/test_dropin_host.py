"""The drop-in layer's HOST logic (libecc_b200/csrc/dropin.cpp: struct marshalling, scheme checks, mod-q scalar
preparation, hashing through the reference's src/hash, verdict mapping, forwarding) on the CPU: the C harness
tests/dropin/dropin_harness.c runs against the unmodified reference with the engine entry points served by the host
build of the device algorithms (tests/hostsim/engine_stub.cpp, preloaded in front of the product library).  The same
harness runs on the real engine in tests/test_gpu_dropin.py; this file keeps host-side mistakes from costing GPU minutes.
A subset of the curves keeps the CPU suite short (the GPU run covers all eight)."""
import json
import os
import re
import subprocess

import pytest

from common import ROOT, engine_stub_so

HARNESS = os.path.join(ROOT, "oracle", "_ref", "dropin_harness")
DROPIN = os.path.join(ROOT, "libecc_b200", "libecc_b200_dropin.so")


def _run(args, preload, curves=None, timeout=900, extra_env=None):
    if not os.path.exists(HARNESS) or not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/dropin_harness or libecc_b200_dropin.so not built (python __graft_entry__.py build)")
    env = dict(os.environ, LD_PRELOAD=" ".join(preload))
    if curves:
        env["HARNESS_CURVES"] = curves
    env.update(extra_env or {})
    r = subprocess.run([HARNESS] + args, capture_output=True, text=True, timeout=timeout, env=env)
    print(r.stdout[-3000:], r.stderr[-2000:])
    return r


def test_direct_mode_against_reference_structs_on_the_host_build():
    """prj_pt_mul / batch / every verify_batch adapter / ec_verify shim / generic ec_verify_batch on real reference
    structs, verdicts judged by the reference's own ec_verify: ECDSA, ECFSDSA, ECSDSA, ECOSDSA, ECKCDSA, ECGDSA, ECRDSA,
    SM2, BIGN, DBIGN, BIP0340 (keys at infinity and off the curve, out-of-range and zero signatures, wrong / missing
    ancillary data).  A 256- and a 224-bit curve here (p = 1 mod 4 for BIP0340); the known-answer run below adds the 384-bit
    and the other curves."""
    r = _run(["direct", DROPIN], [engine_stub_so()], curves="FRP256V1,SECP224R1")
    assert r.returncode == 0 and "HARNESS OK" in r.stdout
    assert r.stdout.count("done, failures so far 0") == 2


def test_reference_known_answer_vectors_through_the_drop_in():
    """Every fixed-vector case of the reference's own self tests (src/tests/ec_self_tests_core.h: the table
    `ec_self_tests vectors` walks — ECDSA, DECDSA, ECKCDSA, ECSDSA, ECOSDSA, ECFSDSA, ECGDSA, ECRDSA, SM2, BIGN, DBIGN,
    BIP0340, EdDSA) through the drop-in's ec_verify and a one-item ec_verify_batch: the expected signature verifies, an
    altered message does not, as the reference's own ec_verify says; cases on curves / schemes the layer does not serve
    are forwarded and agree too."""
    r = _run(["kats", DROPIN], [engine_stub_so()], extra_env={"HARNESS_KATS_THIN": "1"})
    assert r.returncode == 0 and "HARNESS OK" in r.stdout
    line = [l for l in r.stdout.splitlines() if l.startswith("kats:")][0]
    served = dict(tok.split(":") for tok in line.split("ec_alg_type:")[1].split())
    # every ECDSA / DECDSA / ECKCDSA / ECSDSA / ECOSDSA / ECFSDSA / BIP0340 vector sits on a served curve
    for alg in ("1", "2", "3", "4", "5", "14", "20"):   # (ECDSA / DECDSA thinned to a third in this CPU run)
        got, total = served[alg].split("/")
        assert got == total and int(total) > 0, (alg, served[alg])
    # ECGDSA (brainpool) and SM2 (sm2p256v1) vectors on served curves ran on the engine as well
    assert int(served["6"].split("/")[0]) >= 2 and int(served["8"].split("/")[0]) >= 1


def test_521_bit_curve_host_logic():
    """nine 64-bit limbs mod q, byte-granular wire fields, BIGN's l = 33 > the BELT digest: the sections of the eight
    double-scalar schemes (the GPU run does every section on every curve)."""
    r = _run(["direct", DROPIN], [engine_stub_so()], curves="SECP521R1", extra_env={"HARNESS_SECTIONS": "sd"})
    assert r.returncode == 0 and "HARNESS OK" in r.stdout


def test_reference_code_runs_through_interposed_symbols_on_the_host_build():
    """LD_PRELOAD interposition of prj_pt_mul / ec_verify under the unmodified reference's sign / verify / ECC-CDH."""
    r = _run(["preload"], [engine_stub_so(), DROPIN], curves="SECP256R1,BRAINPOOLP256R1")
    assert r.returncode == 0 and "HARNESS OK" in r.stdout


def test_concurrent_callers_share_the_engine_slots():
    r = _run(["threads", DROPIN], [engine_stub_so()])
    assert r.returncode == 0 and "HARNESS OK" in r.stdout


@pytest.mark.parametrize("scheme", ["ECGDSA", "SM2", "BIGN", "ECKCDSA"])
def test_batches_larger_than_one_marshalling_chunk(scheme):
    """more than 2048 items: the marshalling loops and the shared mod-q inversion run on several host threads."""
    r = _run(["bench", DROPIN, "FRP256V1", "2600", scheme, "64"], [engine_stub_so()],
             extra_env={"HARNESS_POOL": "96", "HARNESS_REPS": "2"})
    assert r.returncode == 0 and "HARNESS OK" in r.stdout
    line = json.loads(r.stdout.split("DROPIN_BENCH ", 1)[1].splitlines()[0])
    assert line["wrong_verdicts"] == 0 and line["items"] == 2600


def test_ecrdsa_iso14888_3_switch_changes_the_digest_byte_order():
    """ECCB200_ECRDSA_ISO14888_3=1 serves a reference built with USE_ISO14888_3_ECRDSA (sig/ecrdsa.c:545-547); against
    the default build used here it must therefore reject what the reference accepts."""
    r = _run(["bench", DROPIN, "FRP256V1", "64", "ECRDSA", "0"], [engine_stub_so()],
             extra_env={"ECCB200_ECRDSA_ISO14888_3": "1", "HARNESS_POOL": "64", "HARNESS_REPS": "2"})
    assert r.returncode != 0 and "HARNESS FAILED" in r.stdout and "batch verdict -1" in r.stdout
    r = _run(["bench", DROPIN, "FRP256V1", "64", "ECRDSA", "0"], [engine_stub_so()],
             extra_env={"HARNESS_POOL": "64", "HARNESS_REPS": "2"})
    assert r.returncode == 0 and "HARNESS OK" in r.stdout


@pytest.mark.parametrize("curve,mutants,hash_name", [("FRP256V1", "150", "SHA256"), ("SECP224R1", "60", "SHA384"),
                                                    ("SECP384R1", "40", "SHA224")])
def test_mutated_signatures_keys_and_ancillary_data_get_the_reference_verdict(curve, mutants, hash_name):
    """Differential fuzzing of the host logic: reference-made signatures of the twelve served schemes with single bit
    flips, fields forced to 0 / q - 1 / q / all-ones, lengths off by one, altered / shortened / missing ancillary data,
    keys at infinity / off the curve / of another scheme / uninitialised, empty messages - the drop-in's ec_verify must
    return what the reference's ec_verify returns for every mutant (digests longer and shorter than the order: the
    truncation rules of ECDSA / ECGDSA / ECKCDSA); the same mutants in batches of sixteen through ec_verify_batch."""
    r = _run(["fuzz", DROPIN, curve, mutants, hash_name], [engine_stub_so()])
    assert r.returncode == 0 and "HARNESS OK" in r.stdout
    line = [l for l in r.stdout.splitlines() if l.startswith("fuzz ")][0]
    total, accepted, engine = (int(x) for x in re.findall(r"(\d+) (?:mutants|accepted|judged)", line))
    assert total >= 11 * int(mutants) and 0 < accepted < total // 4 and engine > total // 2


@pytest.mark.parametrize("curve", ["SECP256R1", "BRAINPOOLP384R1", "SECP192R1"])
def test_prj_pt_mul_on_random_scalar_widths_and_point_forms(curve):
    """prj_pt_mul of the drop-in on nn scalars of every width (0 to 27 words) and on projective (blinded by the reference),
    infinite, off-curve and aliased points, walking from result to result: return code, infinity flag and coordinates as
    the reference's prj_pt_mul gives them.  (Eleven curves x 1500 iterations were run once: profiles/r02_s3_host_sanitizers.md.)"""
    r = _run(["fuzzmul", DROPIN, curve, "250"], [engine_stub_so()])
    assert r.returncode == 0 and "HARNESS OK" in r.stdout and "250 multiplications agree" in r.stdout
