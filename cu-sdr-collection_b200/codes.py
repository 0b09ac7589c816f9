"""PRN replica generation, host side (integer LFSR work).

Product-side counterpart of the reference's ``generateCAcode.m`` /
``makeCaTable.m`` (GPS/GPS_L1CA/include/generateCAcode.m:42-90,
makeCaTable.m:43-67).  Written as a Galois-free bit-LFSR on 0/1 bits (the
reference multiplies ±1 values); the C++ library carries its own copy of the same
generator (csrc/codes.cpp) for the device tables — this module serves the
synthetic-record generator and the host mirror.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

# G2 output delay (chips) per PRN, IS-GPS-200 Table 3-Ia; PRN 33..51 are the
# SBAS entries the reference lists (true_PRN = PRN + 87).
G2_DELAY = (5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258,
            469, 470, 471, 472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862,
            145, 175, 52, 21, 237, 235, 886, 657, 634, 762, 355, 1012, 176, 603, 130, 359,
            595, 68, 386)


def _lfsr(taps) -> np.ndarray:
    """1023-chip maximal sequence of a 10-stage register, all-ones start, output = stage 10."""
    reg = [1] * 10
    out = np.empty(1023, dtype=np.int8)
    for i in range(1023):
        out[i] = reg[9]
        fb = 0
        for t in taps:
            fb ^= reg[t - 1]
        reg = [fb] + reg[:9]
    return out


@lru_cache(maxsize=None)
def ca_code(prn: int) -> np.ndarray:
    """±1 C/A chips (int8, length 1023); bit 1 -> -1... sign convention of the reference:
    chip = +1 where G1 xor G2 == 1 (generateCAcode.m:90, ``-(g1.*g2)`` on ±1 registers
    loaded with -1)."""
    g1 = _lfsr((3, 10))
    g2 = _lfsr((2, 3, 6, 8, 9, 10))
    g2 = np.roll(g2, G2_DELAY[prn - 1])
    x = g1 ^ g2                      # 0/1, IS-GPS-200 logic levels
    return (2 * x.astype(np.int8) - 1)


def ca_first10_octal(prn: int) -> str:
    """First 10 chips as the octal word IS-GPS-200 tabulates (known-answer test)."""
    x = (ca_code(prn)[:10] > 0).astype(int)
    v = 0
    for b in x:
        v = (v << 1) | int(b)
    return format(v, "o")


@lru_cache(maxsize=None)
def glo_code() -> np.ndarray:
    """+-1 GLONASS ST code (511 chips): 9-stage register, feedback from stages 5 and 9, output of
    stage 7, all-ones start (GLO/GLO_GL1/include/generateCAcode.m:95-108; bit 1 <-> -1)."""
    reg = [1] * 9
    out = np.empty(511, dtype=np.int8)
    for i in range(511):
        out[i] = -1 if reg[6] else 1
        fb = reg[4] ^ reg[8]
        reg = [fb] + reg[:8]
    return out


B3I_INIT = (4, 11, 13, 22, 30, 36, 44, 48, 88, 104, 116, 129, 376, 418, 458, 682, 696, 707, 1078, 2069,
            2248, 2574, 2596, 2731, 4294, 4436, 4647, 4978, 4986, 1, 5209, 5539, 6061, 6488, 7130, 7165,
            7403, 5879, 1681, 5080, 5938, 3983, 6208, 7223, 2996, 1814, 6906, 6144, 4713, 7406, 7264, 1766,
            5347, 3515, 7951, 7054, 3884, 6067, 4230, 3803, 869, 3683, 1205)


@lru_cache(maxsize=None)
def b3i_code(prn: int) -> np.ndarray:
    """+-1 BeiDou B3I chips (int8, 10230): truncated 13-stage G1 (taps 1,3,4,13, reloaded at state
    1111111111100) times G2 (taps 1,5,6,7,9,10,12,13) advanced by a per-PRN step count
    (BDS/B3I/include/generateB3Icode.m:33-86; bit 1 <-> -1)."""
    def step(reg, taps):
        fb = 0
        for t in taps:
            fb ^= reg[t - 1]
        return [fb] + reg[:12]
    reset = [1] * 11 + [0, 0]
    a = [1] * 13
    ca = np.empty(10230, dtype=np.int8)
    for i in range(10230):
        ca[i] = a[12]
        a = [1] * 13 if a == reset else step(a, (1, 3, 4, 13))
    b = [1] * 13
    for _ in range(B3I_INIT[prn - 1]):
        b = step(b, (1, 5, 6, 7, 9, 10, 12, 13))
    out = np.empty(10230, dtype=np.int8)
    for i in range(10230):
        out[i] = -1 if (b[12] ^ ca[i]) else 1
        b = step(b, (1, 5, 6, 7, 9, 10, 12, 13))
    return out


# ---- Galileo E1 memory codes -------------------------------------------------------------------------
# The E1-B / E1-C primary codes are ICD tables, not LFSR output; the reference keeps them as data files
# next to its sources and reads them at run time (GAL/GAL_E1C/include/generateE1Bcode.m:44-55,
# generateE1Ccode.m).  The engine therefore takes them from the caller (gc_set_code).
def load_e1_codes(code_dir: str) -> dict:
    """{PRN: (e1b, e1c)} for PRN 1..50 from ``E1b.dat`` / ``E1c.dat`` in ``code_dir`` (the format the reference
    reads with fscanf '%d': 50 x 4092 whitespace-separated 0/1 digits); values are the +-1 primary chips
    ``1 - 2*bit`` (generateE1Bcode.m:55)."""
    import os
    tabs = []
    for name in ("E1b.dat", "E1c.dat"):
        with open(os.path.join(code_dir, name)) as f:
            v = np.array(f.read().split(), dtype=np.int8)
        if v.size < 50 * 4092:
            raise ValueError(f"{name}: expected 50 x 4092 digits, found {v.size}")
        tabs.append((1 - 2 * v[: 50 * 4092]).astype(np.int8).reshape(50, 4092))
    return {prn: (tabs[0][prn - 1], tabs[1][prn - 1]) for prn in range(1, 51)}


def standin_e1_codes(seed: int = 20260101) -> dict:
    """Seeded random +-1 tables with the shape of the E1 memory codes, for synthetic records where the real
    tables are not at hand (tests and benchmarks on the GPU box): {PRN: (e1b, e1c)}."""
    rng = np.random.default_rng(seed)
    t = (1 - 2 * rng.integers(0, 2, size=(2, 50, 4092))).astype(np.int8)
    return {prn: (t[0, prn - 1], t[1, prn - 1]) for prn in range(1, 51)}


def boc11(primary: np.ndarray) -> np.ndarray:
    """BOC(1,1) sub-chips [c -c] of a +-1 primary code (generateE1Bcode.m:58-64)."""
    out = np.empty(2 * primary.size, dtype=np.int8)
    out[0::2] = primary
    out[1::2] = -primary
    return out


E1_SECONDARY = np.array([1, 1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, -1, 1, -1, 1, -1, -1, 1, -1, -1, 1, 1, -1, 1],
                        dtype=np.int8)     # CS25 '380AD90', antipodal (GAL_E1C/include/acquisition.m:135)


def standin_codes(signal: str, seed: int = 20260101) -> dict:
    """Seeded random +-1 stand-ins for the generated 10230-chip codes of GPS L5C (I5/Q5), GAL E5a/E5b (I/Q) and
    BDS B2a (data/pilot): {PRN: (data, pilot, pilot_secondary)} for PRN 1..63.  The engine takes these codes from
    the caller (the MATLAB wrappers pass what generateL5Icode.m etc. return); synthetic records and the parity
    tests only need *some* codes that both sides share.  ``pilot_secondary`` (100 chips) is the per-PRN code
    GAL E5a's fine search uses; GPS L5C pilots carry the fixed NH20 code instead."""
    rng = np.random.default_rng([seed, sum(map(ord, signal))])
    t = (1 - 2 * rng.integers(0, 2, size=(2, 63, 10230))).astype(np.int8)
    sec = (1 - 2 * rng.integers(0, 2, size=(63, 100))).astype(np.int8)
    return {prn: (t[0, prn - 1], t[1, prn - 1], sec[prn - 1]) for prn in range(1, 64)}


NH20 = np.array([1, 1, 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1], dtype=np.int8)


def standin_varb_codes(signal: str, seed: int = 20260101) -> dict:
    """Seeded stand-ins for the codes of the variant-B signals, {PRN: (code,)}: BDS B1I 2046 +-1 chips (what
    generateCAcode53.m returns), GPS L2C the 20460-entry return-to-zero CM sequence generateCMcode.m returns
    (chip, 0, chip, 0, ...)."""
    rng = np.random.default_rng([seed, sum(map(ord, signal))])
    if signal == "BDS_B1I":
        t = (1 - 2 * rng.integers(0, 2, size=(63, 2046))).astype(np.int8)
        return {prn: (t[prn - 1],) for prn in range(1, 59)}
    t = np.zeros((32, 20460), dtype=np.int8)
    t[:, 0::2] = 1 - 2 * rng.integers(0, 2, size=(32, 10230))
    return {prn: (t[prn - 1],) for prn in range(1, 33)}


def standin_l2c_cl_codes(prns, seed: int = 20260101) -> dict:
    """Seeded stand-ins for the GPS L2C CL sequences generateCLcode.m returns: 2*767250 entries, return-to-zero with the
    CL chips in the slots the CM sequence leaves empty (0, chip, 0, chip, ...).  {PRN: (cm, cl)} for the PRNs asked
    for (1.5 MB each), cm as in ``standin_varb_codes``."""
    cm = standin_varb_codes("GPS_L2C", seed)
    out = {}
    for prn in prns:
        rng = np.random.default_rng([seed, 0xC1, int(prn)])
        cl = np.zeros(2 * 767250, dtype=np.int8)
        cl[1::2] = 1 - 2 * rng.integers(0, 2, size=767250)
        out[int(prn)] = (cm[int(prn)][0], cl)
    return out


def boc61_from_boc11(pilot_boc11: np.ndarray) -> np.ndarray:
    """Pilot BOC(6,1) sequence of generatePilotBOC61.m:60-66 ((-1)^ii * Primary(jj), ii = 1..12) from the BOC(1,1)
    sub-chips ([-c +c] per primary chip, so Primary = the odd entries)."""
    prim = np.asarray(pilot_boc11, dtype=np.int8)[1::2]
    sub = np.array([-1, 1] * 6, dtype=np.int8)
    return (prim[:, None] * sub[None, :]).reshape(-1)


def standin_b1c_codes(seed: int = 20260101) -> dict:
    """Seeded stand-ins for the BDS B1C BOC(1,1) sub-chip sequences generateDataBOC11.m / generatePilotBOC11.m return
    (20460 entries, [-c +c] per primary chip): {PRN: (data, pilot)} for PRN 1..63."""
    rng = np.random.default_rng([seed, 0xB1C])
    prim = (1 - 2 * rng.integers(0, 2, size=(2, 63, 10230))).astype(np.int8)
    out = {}
    for prn in range(1, 64):
        comps = []
        for r in range(2):
            c = np.empty(20460, dtype=np.int8)
            c[0::2] = -prim[r, prn - 1]
            c[1::2] = prim[r, prn - 1]
            comps.append(c)
        out[prn] = tuple(comps)
    return out


def icd_codes(signal: str, prns=None, cl: bool = False, boc61: bool = False) -> dict:
    """The signal's REAL primary codes from the library's generators (``gc_generate_code``: the reference's generate*code.m as
    bit-packed registers, ICD tables included) in the {PRN: (component 0, component 1, ...)} layout the scene builders and
    ``Engine(codes=...)`` take.  GAL E1C: (e1b, e1c) primary chips; GPS L5C / GAL E5a / GAL E5b / BDS B2a: (data, pilot, pilot
    secondary code - NH20-free: E5a / E5b their own 100 chips, L5C and B2a a row of ones, used by scene synthesis only);
    BDS B1I: (code,); GPS L2C: (cm,) or (cm, cl) with ``cl``; BDS B1C: (data, pilot) BOC(1,1) or with ``boc61`` also the pilot
    BOC(6,1) sequence.  The engine does not need this dict - it generates whatever it is not given."""
    from .engine import generate_code
    pools = {"GAL_E1C": range(1, 51), "GPS_L5C": range(1, 33), "GAL_E5a": range(1, 51), "GAL_E5b": range(1, 51), "BDS_B2a": range(1, 64),
             "BDS_B1I": range(1, 59), "GPS_L2C": range(1, 33), "BDS_B1C": range(1, 64)}
    out = {}
    for prn in (pools[signal] if prns is None else prns):
        prn = int(prn)
        if signal in ("GPS_L5C", "BDS_B2a"):
            out[prn] = (generate_code(signal, prn, 0), generate_code(signal, prn, 1), np.ones(100, dtype=np.int8))
        elif signal in ("GAL_E5a", "GAL_E5b"):
            out[prn] = (generate_code(signal, prn, 0), generate_code(signal, prn, 1), generate_code(signal, prn, 2))
        elif signal == "BDS_B1I":
            out[prn] = (generate_code(signal, prn, 0),)
        elif signal == "GPS_L2C":
            out[prn] = (generate_code(signal, prn, 0),) + ((generate_code(signal, prn, 1),) if cl else ())
        elif signal == "BDS_B1C":
            out[prn] = (generate_code(signal, prn, 0), generate_code(signal, prn, 1)) + ((generate_code(signal, prn, 2),) if boc61 else ())
        else:
            out[prn] = (generate_code(signal, prn, 0), generate_code(signal, prn, 1))
    return out
