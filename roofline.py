"""Roofline accounting for bench.py: algorithmic integer-multiply-add work per item and the
measured IMAD peak of the device (imad_peak micro-benchmark inside libecc_b200.so)."""
from __future__ import annotations

import ctypes
import math

# one Montgomery CIOS product of n 64-bit limbs = 2n^2+n wide multiply-accumulates, each = 4 32x32->64 IMADs
IMAD32_PER_MUL = {"SECP256R1": (2 * 4 * 4 + 4) * 4, "FRP256V1": (2 * 4 * 4 + 4) * 4, "SECP384R1": (2 * 6 * 6 + 6) * 4,
                  "BRAINPOOLP256R1": 144, "SECP256K1": 144, "BRAINPOOLP384R1": 312, "SECP521R1": (2 * 9 * 9 + 9) * 4}
# IMAD.WIDE instructions the generated code really issues per product (tools/gen_fp_ptx.py: P-256 mul 96 / sqr 68,
# generic 256-bit 136 / 108, P-384 276 / 210 incl. the m_i products)
IMAD_MUL_SQR = {"SECP256R1": (96, 68), "FRP256V1": (136, 108), "SECP384R1": (276, 210), "BRAINPOOLP256R1": (136, 108),
                "SECP256K1": (136, 108), "BRAINPOOLP384R1": (300, 234), "SECP521R1": (630, 477)}
# averaged over a Jacobian mixed addition (8 mul + 3 sqr): the variable-base / verification kernels
IMAD_EXECUTED_PER_MUL = {c: (8 * m + 3 * q) / 11 for c, (m, q) in IMAD_MUL_SQR.items()}
# averaged over the comb's extended-Jacobian mixed addition (8 mul + 2 sqr): the fixed-base kernel
IMAD_EXECUTED_PER_MUL_FIXED = {c: (8 * m + 2 * q) / 10 for c, (m, q) in IMAD_MUL_SQR.items()}
M_REF = {"SECP256R1": 8724, "FRP256V1": 8724, "SECP384R1": 13076,   # reference ladder (probe)
         "BRAINPOOLP256R1": 8724, "SECP256K1": 8724, "BRAINPOOLP384R1": 13076,
         "SECP521R1": (2 * 521 + 1) * 17 + 3}
QBITS = {"SECP256R1": 256, "FRP256V1": 256, "SECP384R1": 384, "BRAINPOOLP256R1": 256, "SECP256K1": 256,
         "BRAINPOOLP384R1": 384, "SECP521R1": 521}


def work_per_item(workload: str, comb_window: int):
    from bench import WORKLOADS
    curve, kind, _, _ = WORKLOADS[workload]
    nwin = math.ceil(QBITS[curve] / comb_window)
    nib = (QBITS[curve] + 3) // 4
    m_fixed = 10 * (nwin - 1) + 2                               # extended-Jacobian mixed add 8M+2S per window (the
    #                                                             first window is a copy) + 2M back to Jacobian
    cta_inv = 16 + 2 / 4.0                                      # CTA-wide inversion per thread: 2 scans (14) + 2; the
    #                                                             inversion itself (safegcd division steps, run by one
    #                                                             of the CTA's four warps) costs two field products and
    #                                                             ~750 division steps of 32-bit ALU work, not charged
    # variable base: 8-entry table (6 mixed adds + 1 doubling + the abandoned add that detects it), made affine with
    # one shared inversion (6 prefix + 12 back-substitution + 7 x 4 conversion products), 4 doublings per digit
    # (8 products each) and a mixed addition for 15 of 16 digits
    m_var = (6 * 11 + 8 + 5) + (6 + 12 + 28) + cta_inv + nib * 4 * 8 + nib * (15.0 / 16) * 11
    if kind == "fixed":
        m, kernel = m_fixed, "k_smul_fixed"
        m_ref = M_REF[curve]
    elif kind == "var":
        m, kernel = m_var, "k_smul_var"
        m_ref = M_REF[curve]
    else:
        # comb for uG, window for vY with uG folded into the table inversion (+7) and added last (+11), the mod-q
        # inversion, u / v / key validation / final comparison (~12)
        m, kernel = m_fixed + m_var + 7 + 11 + cta_inv + 12, "k_ecdsa_verify"
        m_ref = 2 * M_REF[curve] + 17 + (2 * QBITS[curve] + 1) + 2
    per_mul = (IMAD_EXECUTED_PER_MUL_FIXED if kind == "fixed" else IMAD_EXECUTED_PER_MUL)[curve]
    return {"M_impl": m, "kernel": kernel, "imad32_per_mul": IMAD32_PER_MUL[curve],
            "imad32_per_item": m * IMAD32_PER_MUL[curve], "imad32_ref_per_item": m_ref * IMAD32_PER_MUL[curve],
            "imad_executed_per_item": m * per_mul}


def imad_peak_measured(device: int = 0):
    import libecc_b200
    lib = libecc_b200.load_library()
    lib.eccb200_imad_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    best, clk = ctypes.c_double(), ctypes.c_double()
    if lib.eccb200_imad_peak(device, ctypes.byref(best), ctypes.byref(clk)):
        raise RuntimeError(lib.eccb200_last_error().decode())
    return {"timad32_per_s": best.value / 1e12, "imad_per_clk_per_sm": clk.value,
            "how": "measured: imad_peak micro-benchmark (independent IMAD.WIDE chains, 8 per thread, "
                   "148x8 CTAs x 256 threads, best of 5, CUDA events)"}
