"""Primary-code generators: the library's bit-packed registers (csrc/codegen.h, called through the C ABI without
a GPU) against the oracle's statement-by-statement restatement of the reference's generate*code.m files, for every SV, plus the
known answers of the signal ICDs that survive independent of both (Galileo OS SIS ICD: first 24 chips of E5a-I and E5a-Q code 1,
E1-B code 1) and the structure every spreading code must have."""
import os

import numpy as np
import pytest

import np_oracle as O
from cu_sdr_collection_b200.engine import generate_code

REF = "/root/reference"


def _hex24(c):
    return "%06X" % int("".join("1" if x == -1 else "0" for x in c[:24]), 2)


CASES = [("GPS_L5C", 0, O.generateL5Icode, range(1, 64)), ("GPS_L5C", 1, O.generateL5Qcode, range(1, 64)),
         ("GAL_E5a", 0, O.generateE5aIcode, range(1, 51)), ("GAL_E5a", 1, O.generateE5aQcode, range(1, 51)),
         ("GAL_E5a", 2, O.generateE5aQ_secondary, range(1, 51)),
         ("GAL_E5b", 0, O.generateE5bIcode, range(1, 51)), ("GAL_E5b", 1, O.generateE5bQcode, range(1, 51)),
         ("GAL_E5b", 2, O.generateE5bQ_secondary, range(1, 51)),
         ("BDS_B2a", 0, O.generateB2aDataCode, range(1, 64)), ("BDS_B2a", 1, O.generateB2aPilotCode, range(1, 64)),
         ("BDS_B1I", 0, O.generateCAcode53, range(1, 59)),
         ("GPS_L2C", 0, O.generateCMcode, range(1, 64)),
         ("BDS_B1C", 0, O.generateDataBOC11, range(1, 64)), ("BDS_B1C", 1, O.generatePilotBOC11, range(1, 64)),
         ("BDS_B1C", 2, O.generatePilotBOC61, (1, 7, 63))]


@pytest.mark.parametrize("signal,comp,oracle_fn,svs", CASES, ids=[f"{c[0]}-{c[1]}" for c in CASES])
def test_library_generators_equal_the_oracle_restatement(signal, comp, oracle_fn, svs):
    svs = list(svs)
    if len(svs) > 12:                               # every SV is cheap in the library; the +-1-list oracle takes ~20 ms per code
        svs = svs[:6] + svs[len(svs) // 2: len(svs) // 2 + 3] + svs[-3:]
    for sv in svs:
        got = generate_code(signal, sv, comp)
        want = oracle_fn(sv)
        assert got.shape == want.shape and np.array_equal(got, want), (signal, comp, sv)


def test_l2c_cl_code_and_e1_memory_codes():
    got = generate_code("GPS_L2C", 1, 1)
    want = O.generateCLcode(1)
    assert got.size == 1534500 and np.array_equal(got, want)
    assert not got[0::2].any() and np.all(np.abs(got[1::2]) == 1)                      # [0 c 0 c ...]
    cm = generate_code("GPS_L2C", 1, 0)
    assert not cm[1::2].any() and np.all(np.abs(cm[0::2]) == 1)                        # [c 0 c 0 ...]
    for prn in (1, 2, 50):
        b, c = generate_code("GAL_E1C", prn, 0), generate_code("GAL_E1C", prn, 1)
        assert np.array_equal(b, O.generateE1Bcode(prn)[0::2]) and np.array_equal(c, O.generateE1Ccode(prn)[0::2])


def test_icd_known_answers():
    """Galileo OS SIS ICD: E1-B code 1 starts F5D710..., E5a-I code 1 (start value 30305) starts 3CEA9D, E5a-Q code 1 (25652)
    starts 515537 - first chip = most significant bit, logic 1 = chip -1."""
    assert _hex24(generate_code("GAL_E1C", 1, 0)) == "F5D710"
    assert _hex24(generate_code("GAL_E5a", 1, 0)) == "3CEA9D"
    assert _hex24(generate_code("GAL_E5a", 1, 1)) == "515537"
    # E5a-I codes 2 and 3 (start values 14234, 27213) start 9D8CF1, 45D1C8
    assert _hex24(generate_code("GAL_E5a", 2, 0)) == "9D8CF1" and _hex24(generate_code("GAL_E5a", 3, 0)) == "45D1C8"
    # E5b: base register 1 starts all ones, so the first 14 chips of a code are the complement of the ICD's 14-bit start value of
    # base register 2 - 07220 (octal) for E5b-I code 1, 03331 for E5b-Q code 1
    for comp, start in ((0, 0o07220), (1, 0o03331)):
        first14 = "".join("1" if x == -1 else "0" for x in generate_code("GAL_E5b", 1, comp)[:14])
        assert int(first14, 2) == (~start) & 0x3FFF, (comp, first14)
    # IS-GPS-200 C/A (the generator the library has had since round 1) through the same entry point: PRN 1 starts 1440 octal
    # (generateCAcode.m:90 returns -(g1 .* g2): logic 1 is chip +1 there)
    ca = generate_code("GPS_L1CA", 1, 0)
    assert "".join("1" if x == 1 else "0" for x in ca[:10]) == "1100100000"


@pytest.mark.parametrize("signal,comp,n,svs", [("GPS_L5C", 0, 10230, (1, 2, 3)), ("GAL_E5a", 1, 10230, (1, 2, 3)), ("GAL_E5b", 0, 10230, (1, 2, 3)),
                                               ("BDS_B2a", 0, 10230, (1, 2, 3)), ("BDS_B1I", 0, 2046, (1, 2, 3)), ("GPS_L2C", 0, 20460, (1, 2, 3)),
                                               ("BDS_B1C", 1, 20460, (1, 2, 3))])
def test_code_structure(signal, comp, n, svs):
    """Balance, sharp autocorrelation and low cross-correlation of the generated primary codes."""
    codes = []
    for sv in svs:
        c = generate_code(signal, sv, comp).astype(np.float64)
        assert c.size == n
        if signal == "GPS_L2C":
            c = c[0::2]
        if signal == "BDS_B1C":
            assert np.array_equal(c[0::2], -c[1::2])
            c = c[1::2]
        assert abs(c.sum()) <= 0.02 * c.size + 2, (signal, sv, c.sum())
        codes.append(c)
    L = codes[0].size
    F = [np.fft.fft(c) for c in codes]
    auto = np.fft.ifft(F[0] * np.conj(F[0])).real
    assert abs(auto[0] - L) < 1e-6 and np.max(np.abs(auto[1:])) < 0.08 * L
    cross = np.fft.ifft(F[0] * np.conj(F[1])).real
    assert np.max(np.abs(cross)) < 0.08 * L


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_committed_tables_are_what_the_reference_embeds(tmp_path):
    """tools/extract_icd_tables.py re-run against the mounted reference reproduces the committed tables."""
    import subprocess, sys, shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    keep = {f: open(os.path.join(root, f)).read() for f in ("oracle/icd_tables.py", "cu-sdr-collection_b200/csrc/icd_tables.inc")}
    stamp = {f: os.stat(os.path.join(root, f)) for f in keep}
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "extract_icd_tables.py")], stdout=subprocess.DEVNULL)
    for f, txt in keep.items():
        assert open(os.path.join(root, f)).read() == txt, f + " is stale"
        os.utime(os.path.join(root, f), ns=(stamp[f].st_atime_ns, stamp[f].st_mtime_ns))   # same bytes: no rebuild of the library


# ---- GPS ICD known answers that survive independent of the reference and of both restatements --------------------------------
# IS-GPS-200 Table 3-IIa (L2 CM / L2 CL): initial shift-register state and END state (the state that outputs the last chip, i.e.
# after 10229 / 767249 shifts) in octal, PRN 1..3.  IS-GPS-705 Table 3-Ia/Ib (L5): XB code advance and the XB register state it
# leads to ("initial XB code state", stage 1 leftmost), PRN 1..3.
L2CM_ICD = {1: (0o742417664, 0o552566002), 2: (0o756014035, 0o034445034), 3: (0o002747144, 0o723443711)}
L2CL_ICD = {1: (0o624145772, 0o267724236), 2: (0o506610362, 0o167516066), 3: (0o220360016, 0o771756405)}
L5I_ICD = {1: (266, "0101011100100"), 2: (365, "1100000110101"), 3: (804, "0100000001000")}
L5Q_ICD = {1: (1701, "1001011001100"), 2: (323, "0100011110110"), 3: (5292, "1111000100011")}


def _l2c_register(state, n):
    """The 27-stage modular register of IS-GPS-200 Figure 3-12 as an integer (bit 0 = output stage): the output chip is fed
    back into stages 4, 7, 9, 12, 15, 17, 19, 22, 23, 24, 25 counted from the input end (polynomial 1112225171 octal).  Returns the
    n output bits and the state after n - 1 shifts."""
    mask = sum(1 << (27 - p) for p in (4, 7, 9, 12, 15, 17, 19, 22, 23, 24, 25))
    out = np.empty(n, dtype=np.int8)
    last = state
    for i in range(n):
        last = state
        b = state & 1
        out[i] = b
        state = (state >> 1) | (b << 26)
        if b:
            state ^= mask
    return out, last


def test_l2c_registers_reach_the_icd_end_states_and_match_the_generators():
    for prn, (init, end) in L2CM_ICD.items():
        bits, last = _l2c_register(init, 10230)
        assert last == end, (prn, oct(last))
        chips = 1 - 2 * bits.astype(np.int64)                                        # logic 1 = chip -1
        assert np.array_equal(generate_code("GPS_L2C", prn, 0)[0::2], chips)         # the library (bit-packed register)
        assert np.array_equal(O.generateCMcode(prn)[0::2], chips)                    # the oracle (generateCMcode.m restated)
    for prn, (init, end) in L2CL_ICD.items():
        bits, last = _l2c_register(init, 767250)
        assert last == end, (prn, oct(last))
        if prn == 1:
            assert np.array_equal(generate_code("GPS_L2C", prn, 1)[1::2], 1 - 2 * bits.astype(np.int64))


def _l5_xb_state(advance):
    """XB register of IS-GPS-705 (1 + x + x^3 + x^4 + x^6 + x^7 + x^8 + x^12 + x^13, all ones at the start) after `advance` shifts,
    stage 1 leftmost."""
    reg = [1] * 13
    for _ in range(advance):
        fb = reg[0] ^ reg[2] ^ reg[3] ^ reg[5] ^ reg[6] ^ reg[7] ^ reg[11] ^ reg[12]
        reg = [fb] + reg[:-1]
    return reg


def _l5_code(advance, n=10230):
    """XA (1 + x^9 + x^10 + x^12 + x^13, short-cycled: the state 1111111111101 is followed by all ones) xor XB advanced."""
    xa, xb = [1] * 13, _l5_xb_state(advance)
    out = np.empty(n, dtype=np.int64)
    for i in range(n):
        out[i] = xa[12] ^ xb[12]
        if xa == [1] * 11 + [0, 1]:
            xa = [1] * 13
        else:
            xa = [xa[8] ^ xa[9] ^ xa[11] ^ xa[12]] + xa[:-1]
        xb = [xb[0] ^ xb[2] ^ xb[3] ^ xb[5] ^ xb[6] ^ xb[7] ^ xb[11] ^ xb[12]] + xb[:-1]
    return 1 - 2 * out


def test_l5_xb_advances_reach_the_icd_states_and_match_the_generators():
    for comp, table in ((0, L5I_ICD), (1, L5Q_ICD)):
        for prn, (advance, state) in table.items():
            assert "".join(str(b) for b in _l5_xb_state(advance)) == state, (comp, prn)
            want = _l5_code(advance)
            assert np.array_equal(generate_code("GPS_L5C", prn, comp), want)
            assert np.array_equal((O.generateL5Icode if comp == 0 else O.generateL5Qcode)(prn), want)


def _b1i_icd(tap_a, tap_b, n=2046):
    """BDS-SIS-ICD (B1I): G1 = 1 + X + X^7 + X^8 + X^9 + X^10 + X^11, G2 = 1 + X + X^2 + X^3 + X^4 + X^5 + X^8 + X^9 + X^11, both
    registers start 01010101010, the G2 output is the xor of two phase-selector stages; logic values."""
    g1 = [0, 1] * 5 + [0]
    g2 = list(g1)
    out = np.empty(n, dtype=np.int64)
    for i in range(n):
        out[i] = g1[10] ^ g2[tap_a - 1] ^ g2[tap_b - 1]
        f1 = g1[0] ^ g1[6] ^ g1[7] ^ g1[8] ^ g1[9] ^ g1[10]
        f2 = g2[0] ^ g2[1] ^ g2[2] ^ g2[3] ^ g2[4] ^ g2[7] ^ g2[8] ^ g2[10]
        g1 = [f1] + g1[:-1]
        g2 = [f2] + g2[:-1]
    return out


def test_b1i_codes_follow_the_icd_registers():
    """An integer restatement of the ICD's two 11-stage registers with the ICD's phase assignment of PRN 1 - 10 (1+3, 1+4, 1+5,
    1+6, 1+8, 1+9, 1+10, 1+11, 2+7, 3+4) against the library and the oracle.  generateCAcode53.m:103 ends with
    CAcode = -(g1 .* g2): in the reference (and therefore here) logic 1 is chip +1 for this signal."""
    taps = {1: (1, 3), 2: (1, 4), 3: (1, 5), 4: (1, 6), 5: (1, 8), 6: (1, 9), 7: (1, 10), 8: (1, 11), 9: (2, 7), 10: (3, 4)}
    for prn, (a, b) in taps.items():
        want = 2 * _b1i_icd(a, b) - 1
        assert np.array_equal(generate_code("BDS_B1I", prn, 0), want), prn
        assert np.array_equal(O.generateCAcode53(prn), want), prn


def _b3i_icd(g2_init, n=10230):
    """BDS-SIS-ICD (B3I): G1 = 1 + X + X^3 + X^4 + X^13 (all ones at the start, short-cycled: the state 1111111111100 is followed by
    all ones), G2 = 1 + X + X^5 + X^6 + X^7 + X^9 + X^10 + X^12 + X^13 started from the PRN's 13-bit phase; logic values."""
    g1, g2 = [1] * 13, [int(c) for c in g2_init]
    out = np.empty(n, dtype=np.int64)
    for i in range(n):
        out[i] = g1[12] ^ g2[12]
        g1 = [1] * 13 if g1 == [1] * 11 + [0, 0] else [g1[0] ^ g1[2] ^ g1[3] ^ g1[12]] + g1[:-1]
        g2 = [g2[0] ^ g2[4] ^ g2[5] ^ g2[6] ^ g2[8] ^ g2[9] ^ g2[11] ^ g2[12]] + g2[:-1]
    return out


def test_b3i_codes_follow_the_icd_registers():
    """The ICD's G2 initial phases of PRN 1 - 10 through an integer restatement of its two 13-stage registers, against the
    library's and the oracle's generators (logic 1 = chip -1)."""
    phases = {1: "1010111111111", 2: "1111000101011", 3: "1011110001010", 4: "1111111111011", 5: "1100100011111",
              6: "1001001100100", 7: "1111111010010", 8: "1110111111101", 9: "1010000000010", 10: "0010000011011"}
    for prn, init in phases.items():
        want = 1 - 2 * _b3i_icd(init)
        assert np.array_equal(generate_code("BDS_B3I", prn, 0), want), prn
        assert np.array_equal(O.generateB3Icode(prn), want), prn


def _b2a_icd(g2_init, taps1, taps2, n=10230):
    """BDS-SIS-ICD (B2a): register 1 starts all ones and is reset after 8190 chips, register 2 starts from the PRN's 13-bit
    value; the code is the xor of the two last stages; logic values."""
    g1, g2 = [1] * 13, [int(c) for c in g2_init]
    out = np.empty(n, dtype=np.int64)
    for i in range(n):
        out[i] = g1[12] ^ g2[12]
        if i == 8189:
            g1 = [1] * 13
        else:
            f = 0
            for p in taps1:
                f ^= g1[p - 1]
            g1 = [f] + g1[:-1]
        f = 0
        for p in taps2:
            f ^= g2[p - 1]
        g2 = [f] + g2[:-1]
    return out


def test_b2a_and_glonass_codes_follow_the_icd_registers():
    """B2a data (g1 = 1 + x + x^5 + x^11 + x^13, g2 = 1 + x^3 + x^5 + x^9 + x^11 + x^12 + x^13) and pilot (g1 = 1 + x^3 + x^6 + x^7 +
    x^13, g2 = 1 + x + x^5 + x^7 + x^8 + x^12 + x^13) codes of PRN 1 - 5 from the ICD's register-2 initial values, and the GLONASS
    ST code (1 + x^5 + x^9, nine ones at the start, taken from stage 7) - integer restatements of the ICD figures against the
    library's and the oracle's generators (logic 1 = chip -1)."""
    init = {1: "1000000100101", 2: "1000000110100", 3: "1000010101101", 4: "1000101001111", 5: "1000101010101"}
    for prn, v in init.items():
        d = 1 - 2 * _b2a_icd(v, (1, 5, 11, 13), (3, 5, 9, 11, 12, 13))
        p = 1 - 2 * _b2a_icd(v, (3, 6, 7, 13), (1, 5, 7, 8, 12, 13))
        assert np.array_equal(generate_code("BDS_B2a", prn, 0), d) and np.array_equal(O.generateB2aDataCode(prn), d), prn
        assert np.array_equal(generate_code("BDS_B2a", prn, 1), p) and np.array_equal(O.generateB2aPilotCode(prn), p), prn
    reg, st = [1] * 9, []
    for _ in range(511):
        st.append(reg[6])
        reg = [reg[4] ^ reg[8]] + reg[:-1]
    assert np.array_equal(generate_code("GLO_GL1", 0, 0), 1 - 2 * np.array(st))


def test_b1c_primary_codes_follow_the_icd_weil_construction():
    """BDS-SIS-ICD (B1C): Legendre sequence of length 10243 (L(k) = 1 for the quadratic residues - taken here from the set of squares,
    not from Euler's criterion as the oracle does), Weil code L(k) xor L(k + w), truncated to 10230 chips from position p; the ICD's
    (w, p) of PRN 1 and 2: data (2678, 699), (4802, 694), pilot (796, 7575), (156, 2369).  generateDataBOC11.m:86-89 writes the chip
    1 - 2*W as the sub-chip pair [-c, c]."""
    N = 10243
    L = np.zeros(N, dtype=np.int64)
    L[(np.arange(1, N, dtype=np.int64) ** 2) % N] = 1
    for comp, params in ((0, {1: (2678, 699), 2: (4802, 694)}), (1, {1: (796, 7575), 2: (156, 2369)})):
        for prn, (w, p) in params.items():
            k = (np.arange(10230) + p - 1) % N
            c = 1 - 2 * (L[k] ^ L[(k + w) % N])
            got = generate_code("BDS_B1C", prn, comp)
            assert np.array_equal(got[0::2], -c) and np.array_equal(got[1::2], c), (comp, prn)
