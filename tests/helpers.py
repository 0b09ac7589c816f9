"""Shared test plumbing: the oracle bindings (checker only), seeded scenes, comparisons."""
import ctypes as C
import os

import numpy as np

import np_oracle as O                                   # oracle/ (test infrastructure)
from cu_sdr_collection_b200 import synth
from cu_sdr_collection_b200.settings import Settings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRACK_FIELDS = O.TRACK_FIELDS


class OrcSettings(C.Structure):
    _fields_ = ([(n, C.c_double) for n in ["samplingFreq", "IF", "codeFreqBasis", "codeLength",
                                           "acqSearchBand", "acqSearchStep", "acqThreshold"]] +
                [("acqNonCohTime", C.c_int), ("skipNumberOfBytes", C.c_int)] +
                [(n, C.c_double) for n in ["dllDampingRatio", "dllNoiseBandwidth", "dllCorrelatorSpacing",
                                           "pllDampingRatio", "pllNoiseBandwidth", "intTime", "CNo_accTime"]] +
                [("CNo_VSMinterval", C.c_int), ("freqSpacing", C.c_double), ("glo", C.c_int), ("pilotTRKflag", C.c_int)])


_orc = None


def orc():
    global _orc
    if _orc is None:
        _orc = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libgnss_oracle.so"))
        _orc.orc_CNoVSM.restype = C.c_double
    return _orc


def orc_settings(s) -> OrcSettings:
    return OrcSettings(s.samplingFreq, s.IF, s.codeFreqBasis, s.codeLength, s.acqSearchBand, s.acqSearchStep,
                       s.acqThreshold, s.acqNonCohTime, s.skipNumberOfBytes, s.dllDampingRatio,
                       s.dllNoiseBandwidth, s.dllCorrelatorSpacing, s.pllDampingRatio, s.pllNoiseBandwidth,
                       s.intTime, s.CNo_accTime, s.CNo_VSMinterval, float(getattr(s, "freqSpacing", 0.0)),
                       oracle_mode(s), int(getattr(s, "pilotTRKflag", 0)))


def oracle_mode(s) -> int:
    """0 = GPS L1CA files, 1 = GLO_GL1/GL2, 2 = BDS B3I, 3 = GAL E1C (the `glo` field of the C oracle's settings)."""
    if float(getattr(s, "freqSpacing", 0.0)) != 0.0:
        return 1
    return 2 if int(s.codeLength) == 10230 else 3 if int(s.codeLength) == 4092 else 0


_e1_keep = None


def orc_set_e1_codes(codes: dict):
    """Hand the E1 memory codes ({PRN: (e1b, e1c)} +-1 primary chips) to the C oracle as its 0/1 tables."""
    global _e1_keep
    tabs = [np.zeros((50, 4092), dtype=np.int8), np.zeros((50, 4092), dtype=np.int8)]
    for prn, (b, c) in codes.items():
        tabs[0][prn - 1] = (1 - np.asarray(b, dtype=np.int64)) // 2
        tabs[1][prn - 1] = (1 - np.asarray(c, dtype=np.int64)) // 2
    _e1_keep = tabs
    orc().orc_set_e1_codes(P(tabs[0]), P(tabs[1]))


def oracle_codes(codes: dict) -> dict:
    """{PRN: (e1b_bits, e1c_bits)} 0/1 tables for the NumPy oracle from the +-1 chips."""
    return {prn: ((1 - np.asarray(b, dtype=np.int64)) // 2, (1 - np.asarray(c, dtype=np.int64)) // 2) for prn, (b, c) in codes.items()}


def to_oracle_settings(s: Settings) -> "O.Settings":
    o = O.Settings()
    for k in vars(o):
        if hasattr(s, k):
            setattr(o, k, getattr(s, k))
    return o


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def c_acquisition(raw: np.ndarray, s, prns):
    """C oracle acquisition; raw starts at the skip point."""
    cs = orc_settings(s)
    prn = np.asarray(prns, dtype=np.int32)
    nres = {0: 32, 1: 21, 2: 63, 3: 50}[cs.glo]
    cf, cp, pm = np.zeros(nres), np.zeros(nres), np.zeros(nres)
    cb, ccp = np.zeros(nres, dtype=np.int32), np.zeros(nres, dtype=np.int32)
    sp = C.c_double()
    rc = orc().orc_acquisition(P(raw), C.c_size_t(raw.size // 2), C.byref(cs), P(prn), int(prn.size),
                               P(cf), P(cp), P(pm), P(cb), P(ccp), C.byref(sp))
    assert rc == 0, rc
    return dict(carrFreq=cf, codePhase=cp, peakMetric=pm, coarseBin=cb, coarseCodePhase=ccp, sigPower=sp.value)


def c_tracking(raw: np.ndarray, s, prn, acq_freq, code_phase, n_epochs, parallel=1, code_freq0=None):
    cs = orc_settings(s)
    prn = np.asarray(prn, dtype=np.int32)
    af = np.asarray(acq_freq, dtype=np.float64)
    cp = np.asarray(code_phase, dtype=np.float64)
    nch = prn.size
    nv = n_epochs // s.CNo_VSMinterval
    out = np.zeros((nch, 15, n_epochs))
    vv, vi = np.zeros((nch, nv)), np.zeros((nch, nv))
    done = np.zeros(nch, dtype=np.int32)
    cf0 = None if code_freq0 is None else np.ascontiguousarray(code_freq0, dtype=np.float64)
    rc = orc().orc_tracking(P(raw), C.c_size_t(raw.size), C.byref(cs), nch, P(prn), P(af), P(cp),
                            P(cf0) if cf0 is not None else None, n_epochs,
                            P(out), P(vv), P(vi), P(done), parallel)
    assert rc == 0, rc
    return out, vv, vi, done


def scene(fs, nsat=4, seed=7, cn0=None, IF=20e3):
    sc = synth.default_scene(fs=fs, IF=IF, nsat=nsat, seed=seed)
    if cn0 is not None:
        for s in sc.sats:
            s.cn0 = cn0
    return sc


def iq_scale(out):
    """|I_P + i*Q_P| per epoch: the scale the 1e-6 relative tolerance on I/Q sums refers to."""
    return np.hypot(out[:, 3, :], out[:, 7, :])


def track_rel_err(got, ref):
    """max over epochs of |got-ref| / scale per field; scale = |P| for the six I/Q rows,
    |ref| (floored) for the others."""
    errs = {}
    sc = iq_scale(ref)
    for i, f in enumerate(TRACK_FIELDS):
        d = np.abs(got[:, i, :] - ref[:, i, :])
        if 3 <= i <= 8:
            e = d / np.maximum(sc, 1e-30)
        else:
            e = d / np.maximum(np.abs(ref[:, i, :]), 1e-12)
        e = np.where(d == 0, 0.0, e)
        errs[f] = float(np.nanmax(e))
    return errs


def first_illconditioned_epoch(ref_rem, ref_cf, got_rem, got_cf, abs_sample, fs, spacing, sub=1.0, safety=8.0):
    """First epoch in which a sample's code phase (early, prompt or late) lies closer to a chip edge than the two
    implementations' code phases differ there (times `safety`), or len() if none.  tcode(k) = remCodePhase -/+ spacing +
    k*codeFreq/fs feeds ceil(); the engine's remCodePhase / codeFreq follow the reference's to ~1e-10 chips / ~1e-7 Hz
    (fp32 correlator sums feed the discriminators), so where |tcode(k) - integer| is below that the replica chip of
    one sample - 1e-4..1e-3 of a correlator sum at these SNRs - is not determined by the algorithm but by the 10th
    decimal of the loop state; closed-loop comparisons at 1e-6 are only meaningful up to such an epoch.
    (16.368 Msps / 1.023 Mcps puts 16 samples on a chip and never gets there; 18 Msps / 10.23 Mcps has 18000
    distinct code-phase fractions per epoch and meets such a point every ~1e5 epoch-correlators.)"""
    n = len(ref_rem)
    for e in range(n):
        if not (np.isfinite(ref_rem[e]) and np.isfinite(got_rem[e])):
            return e
        blk = int(abs_sample[e + 1] - abs_sample[e]) if e + 1 < n and abs_sample[e + 1] > 0 else int(np.ceil(fs / ref_cf[e] * 10230)) + 2
        step = ref_cf[e] / fs
        d_rem, d_step = abs(got_rem[e] - ref_rem[e]), abs(got_cf[e] - ref_cf[e]) / fs
        k = np.arange(blk, dtype=np.float64)
        tol = safety * (d_rem + k * d_step) * sub          # zero where the two states are bit-identical: nothing to flag
        for off in (-spacing, 0.0, spacing):
            t = (ref_rem[e] + off + k * step) * sub
            if np.any(np.abs(t - np.rint(t)) < tol):
                return e
    return n


PARITY_REPORT = []      # (label, window epochs, total epochs, worst pre-window error, worst post-window error): printed by conftest


def windowed_iq_compare(label, tr, ref, names, fs, spacing, sub=1.0, rem_scale=1.0, scale_keys=("I_P", "Q_P"), iq_tol=1e-6,
                        post_tol=1e-2, min_frac=0.9, exact=False, abs_sample=None):
    """The closed-loop I/Q comparison of the trackers whose ceil(tcode) can be ill conditioned (18 Msps / 10.23 Mcps, BOC tables):
    every correlator row within `iq_tol` of |P| up to the first ill-conditioned epoch (first_illconditioned_epoch), `post_tol`
    after it; the window must cover at least `min_frac` of the run.  With exact=True (the engine's float64 checking mode) there is
    no window: `iq_tol` holds over the whole run - that is the proof that the window is conditioning and not an error.
    Returns (window, worst error inside, worst error after) and records them for the test report."""
    nE = len(ref["I_P"])
    sc_ = np.hypot(ref[scale_keys[0]], ref[scale_keys[1]])
    err = np.zeros(nE)
    for name in names:
        err = np.maximum(err, np.abs(tr[name] - ref[name]) / sc_)
    if exact:
        ok = nE
    else:
        a = ref["absoluteSample"] if abs_sample is None else abs_sample
        ok = first_illconditioned_epoch(rem_scale * ref["remCodePhase"], rem_scale * ref["codeFreq"], rem_scale * tr["remCodePhase"],
                                        rem_scale * tr["codeFreq"], a, fs, rem_scale * spacing, sub=sub)
    pre = float(err[:ok].max()) if ok else 0.0
    post = float(err[ok:].max()) if ok < nE else 0.0
    PARITY_REPORT.append((label + (" [float64 checking mode]" if exact else ""), ok, nE, pre, post))
    print(f"[parity] {PARITY_REPORT[-1][0]}: 1e-6 window {ok}/{nE} epochs, worst |dIQ|/|P| inside {pre:.2e}, after {post:.2e}")
    assert pre < iq_tol, (label, "inside the window", pre)
    assert post < post_tol, (label, "after the window", post)
    assert ok >= min_frac * nE, (label, "window", ok, nE)
    return ok, pre, post


# IS-GPS-200 table 20-XIV: source data bits d1..d24 entering D25..D30 and which of D29*, D30* joins them
_NAV_SETS = [
    (0, (1, 2, 3, 5, 6, 10, 11, 12, 13, 14, 17, 18, 20, 23)),
    (1, (2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24)),
    (0, (1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22)),
    (1, (2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23)),
    (1, (1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24)),
    (0, (3, 5, 6, 8, 9, 10, 11, 13, 15, 19, 22, 23, 24))]


def nav_message_bits(n_subframes: int, seed: int) -> np.ndarray:
    """A GPS L1 C/A bit stream (0/1) of whole subframes with the TLM preamble 10001011 and valid word parity: random source
    data, transmitted D1..D24 = d xor D30*, D25..D30 from the table above."""
    rng = np.random.default_rng(seed)
    out = []
    star = [0, 0]                                                  # D29*, D30* of the previous word
    for _ in range(n_subframes):
        for w in range(10):
            d = rng.integers(0, 2, size=24)
            if w == 0:
                d[:8] = [1, 0, 0, 0, 1, 0, 1, 1]
            par = [(star[s] + int(np.sum(d[np.array(ix) - 1]))) % 2 for s, ix in _NAV_SETS]
            word = [int(b) ^ star[1] for b in d] + par
            out += word
            star = word[28:30]
    return np.array(out, dtype=np.int64)


def nav_prompt_row(bits: np.ndarray, start: int, n: int, amp: float, sigma: float, seed: int, polarity: int = 1) -> np.ndarray:
    """trackResults.I_P for a channel whose first bit edge of `bits` falls on 1-based index `start`: 20 values per bit,
    binary 1 -> -amp (the mapping navPartyChk.m assumes, D30* set <=> -1), noise sigma, random data before `start`."""
    rng = np.random.default_rng(seed)
    x = np.repeat(1 - 2 * bits, 20).astype(np.float64)
    row = np.empty(n)
    head = np.repeat(1 - 2 * rng.integers(0, 2, size=start // 20 + 1), 20)[-(start - 1):] if start > 1 else np.empty(0)
    row[: start - 1] = head
    row[max(start - 41, 0): start - 1] = 1.0                       # D29*, D30* before the first word: binary 0, as the generator assumed
    m = min(n - (start - 1), x.size)
    row[start - 1: start - 1 + m] = x[:m]
    if start - 1 + m < n:
        row[start - 1 + m:] = np.repeat(1 - 2 * rng.integers(0, 2, size=(n - start - m) // 20 + 2), 20)[: n - (start - 1 + m)]
    return polarity * amp * row + sigma * rng.standard_normal(n)


def oracle_signal_codes(signal: str, prns, cl: bool = False, boc61: bool = False) -> dict:
    """The signal's real primary codes from the ORACLE's restatement of the reference's generators (np_oracle.generate*), in the
    {PRN: (components...)} layout of the scene builders and the oracle's acquisition / tracking functions.  The engine under test
    is given no codes at all: it generates its own on the device (csrc/codegen.cu), so a parity test on these records also proves
    that the two generators agree."""
    gen = {"GPS_L5C": (O.generateL5Icode, O.generateL5Qcode, None), "GAL_E5a": (O.generateE5aIcode, O.generateE5aQcode, O.generateE5aQ_secondary),
           "GAL_E5b": (O.generateE5bIcode, O.generateE5bQcode, O.generateE5bQ_secondary), "BDS_B2a": (O.generateB2aDataCode, O.generateB2aPilotCode, None)}
    out = {}
    for prn in prns:
        prn = int(prn)
        if signal in gen:
            d, p, sec = gen[signal]
            out[prn] = (d(prn).astype(np.int8), p(prn).astype(np.int8), (sec(prn) if sec else np.ones(100)).astype(np.int8))
        elif signal == "GAL_E1C":
            out[prn] = (O.generateE1Bcode(prn)[0::2].astype(np.int8), O.generateE1Ccode(prn)[0::2].astype(np.int8))   # primary chips
        elif signal == "BDS_B1I":
            out[prn] = (O.generateCAcode53(prn).astype(np.int8),)
        elif signal == "GPS_L2C":
            out[prn] = (O.generateCMcode(prn).astype(np.int8),) + ((O.generateCLcode(prn).astype(np.int8),) if cl else ())
        elif signal == "BDS_B1C":
            out[prn] = (O.generateDataBOC11(prn).astype(np.int8), O.generatePilotBOC11(prn).astype(np.int8)) + \
                       ((O.generatePilotBOC61(prn).astype(np.int8),) if boc61 else ())
        else:
            raise ValueError(signal)
    return out


def c_acquisition_variant(case) -> dict:
    """The SECOND witness (oracle/gnss_oracle_ext.c, written from the reference's .m files) of the acquisition variants the first C
    oracle does not cover, run on a tests/golden_cases.py case: GPS L5C / GAL E5a / GAL E5b / BDS B2a (two-replica variant A),
    BDS B1I / GPS L2C (variant B), BDS B1C (variant C).  Returns the acqResults vectors in the NumPy oracle's layout."""
    s, sig, sv, codes = case.so, case.signal, list(case.sv), case.codes
    cs = orc_settings(s)
    prn = np.asarray(sv, dtype=np.int32)
    raw = np.ascontiguousarray(case.raw_acq, dtype=np.int8)
    n_avail = raw.size // 2

    def stack(comp, n):
        return np.ascontiguousarray(np.stack([np.asarray(codes[p][comp], dtype=np.int8)[:n] for p in sv]))
    nres = {"GPS_L5C": 32, "GAL_E5a": 50, "GAL_E5b": 50, "BDS_B2a": max(sv), "BDS_B1I": 58, "GPS_L2C": 32, "BDS_B1C": max(sv)}[sig]
    cf, cp, pm = np.zeros(nres), np.zeros(nres), np.zeros(nres)
    cb, ccp = np.zeros(nres, dtype=np.int32), np.zeros(nres, dtype=np.int32)
    out = (P(cf), P(cp), P(pm), P(cb), P(ccp))
    if sig in ("GPS_L5C", "GAL_E5a", "GAL_E5b", "BDS_B2a"):
        d, p = stack(0, 10230), stack(1, 10230)
        sec = np.ascontiguousarray(np.stack([np.asarray(codes[q][2], dtype=np.int8)[:100] for q in sv]))
        rc = orc().orc_acquisition_fam5(P(raw), C.c_size_t(n_avail), C.byref(cs), {"GPS_L5C": 4, "GAL_E5a": 5, "GAL_E5b": 6, "BDS_B2a": 7}[sig],
                                        P(prn), int(prn.size), P(d), P(p), P(sec), int(nres), *out)
    elif sig in ("BDS_B1I", "GPS_L2C"):
        l2c = sig == "GPS_L2C"
        c0 = stack(0, 20460 if l2c else 2046)
        step = float(s.acqStep) if l2c else float(getattr(s, "stepSize", 0) or 0)
        rc = orc().orc_acquisition_varb(P(raw), C.c_size_t(n_avail), C.byref(cs), int(l2c), C.c_double(step), P(prn), int(prn.size), P(c0), *out)
    else:
        d, p = stack(0, 20460), stack(1, 20460)
        rc = orc().orc_acquisition_b1c(P(raw), C.c_size_t(n_avail), C.byref(cs), C.c_double(float(s.acqStep)), int(s.acqCohT), int(s.pilotACQflag),
                                       P(prn), int(prn.size), P(d), P(p), int(nres), *out)
    assert rc == 0, rc
    return dict(carrFreq=cf, codePhase=cp, peakMetric=pm, coarseBin=cb, coarseCodePhase=ccp)


def c_tracking_variant(case) -> list:
    """The SECOND witness (oracle/gnss_oracle_ext.c) of the trackers the first C oracle does not cover, on a golden case with a
    channel hand-off: GPS L5C / GAL E5a / GAL E5b / BDS B2a (quadrature pilot, carrier-aided code NCO) and BDS B1I.  Returns one
    dict of rows per channel (None for a channel that is off)."""
    s, sig, codes, ch = case.so, case.signal, case.codes, case.ch
    cs = orc_settings(s)
    n_ch = len(ch)
    prn = np.asarray([c["PRN"] for c in ch], dtype=np.int32)
    af = np.asarray([c["acquiredFreq"] for c in ch], dtype=np.float64)
    cp = np.asarray([float(c["codePhase"]) for c in ch], dtype=np.float64)
    raw = np.ascontiguousarray(case.raw_trk, dtype=np.int8)
    n_e = case.nE
    L = int(s.codeLength)
    live = [p for p in prn if p]
    data = np.zeros((n_ch, L), dtype=np.int8)
    pilot = np.zeros((n_ch, L), dtype=np.int8)
    for i, p in enumerate(prn):
        if p:
            data[i] = np.asarray(codes[int(p)][0], dtype=np.int8)[:L]
            if sig != "BDS_B1I":
                pilot[i] = np.asarray(codes[int(p)][1], dtype=np.int8)[:L]
    quad = int(sig != "BDS_B1I" and int(getattr(s, "pilotTRKflag", 0)) == 1)
    cf0 = None if sig == "BDS_B1I" else np.asarray([c.get("codeFreq", s.codeFreqBasis) for c in ch], dtype=np.float64)
    out = np.zeros((n_ch, 17, n_e))
    done = np.zeros(n_ch, dtype=np.int32)
    rc = orc().orc_tracking_codes(P(raw), C.c_size_t(raw.size), C.byref(cs), n_ch, P(prn), P(af), P(cp), P(cf0) if cf0 is not None else None,
                                  P(data), P(pilot) if quad else None, quad, n_e, P(out), P(done))
    assert rc == 0 and live, rc
    names = ["absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L", "dllDiscr", "dllDiscrFilt", "pllDiscr",
             "pllDiscrFilt", "remCodePhase", "remCarrPhase", "Pilot_I_P", "Pilot_Q_P"]
    return [dict({k: out[i, j] for j, k in enumerate(names)}, epochsDone=int(done[i])) if prn[i] else None for i in range(n_ch)]


_TRK_NAMES = ["absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L", "dllDiscr", "dllDiscrFilt", "pllDiscr",
              "pllDiscrFilt", "remCodePhase", "remCarrPhase", "Pilot_I_P", "Pilot_Q_P", "Pilot_I_E", "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_L"]


def c_tracking_l2c(raw, so, ch, codes, n_epochs, cl_phase=None) -> list:
    """Second witness of GPS_L2C/include/tracking.m (oracle/gnss_oracle_ext.c); codes[PRN] = (CM,) or (CM, CL); cl_phase = the
    channels' CLCodePhase when the CL pilot is tracked."""
    cs = orc_settings(so)
    n_ch = len(ch)
    prn = np.asarray([c["PRN"] for c in ch], dtype=np.int32)
    af = np.asarray([c["acquiredFreq"] for c in ch], dtype=np.float64)
    cp = np.asarray([float(c["codePhase"]) for c in ch], dtype=np.float64)
    raw = np.ascontiguousarray(raw, dtype=np.int8)
    L2 = 2 * int(so.codeLength)
    cm = np.zeros((n_ch, L2), dtype=np.int8)
    cl = None
    cl_len = 767250
    if cl_phase is not None:
        cl = np.zeros((n_ch, 2 * cl_len), dtype=np.int8)
    for i, p in enumerate(prn):
        if p:
            cm[i] = np.asarray(codes[int(p)][0], dtype=np.int8)
            if cl is not None:
                cl[i] = np.asarray(codes[int(p)][1], dtype=np.int8)
    clp = np.asarray(cl_phase if cl_phase is not None else [0] * n_ch, dtype=np.int32)
    out = np.zeros((n_ch, 21, n_epochs))
    done = np.zeros(n_ch, dtype=np.int32)
    rc = orc().orc_tracking_l2c(P(raw), C.c_size_t(raw.size), C.byref(cs), n_ch, P(prn), P(af), P(cp), P(cm), P(cl) if cl is not None else None,
                                P(clp), C.c_long(cl_len), n_epochs, P(out), P(done))
    assert rc == 0, rc
    return [dict({k: out[i, j] for j, k in enumerate(_TRK_NAMES)}, epochsDone=int(done[i])) if prn[i] else None for i in range(n_ch)]


def c_tracking_b1c_nb(raw, so, ch, codes, n_epochs, factor=None) -> list:
    """Second witness of BDS/B1C/include/NB_tracking.m (oracle/gnss_oracle_ext.c); codes[PRN] = (data BOC(1,1), pilot BOC(1,1)).
    With ``factor`` (CalcWeighingFactor's result): WB_tracking.m, codes[PRN][2] = the pilot BOC(6,1) sequence, 21 rows."""
    cs = orc_settings(so)
    n_ch = len(ch)
    prn = np.asarray([c["PRN"] for c in ch], dtype=np.int32)
    af = np.asarray([c["acquiredFreq"] for c in ch], dtype=np.float64)
    cp = np.asarray([float(c["codePhase"]) for c in ch], dtype=np.float64)
    cf0 = np.asarray([c.get("codeFreq", so.codeFreqBasis) for c in ch], dtype=np.float64)
    raw = np.ascontiguousarray(raw, dtype=np.int8)
    L2 = 2 * int(so.codeLength)
    d, pl = np.zeros((n_ch, L2), dtype=np.int8), np.zeros((n_ch, L2), dtype=np.int8)
    for i, p in enumerate(prn):
        if p:
            d[i] = np.asarray(codes[int(p)][0], dtype=np.int8)
            pl[i] = np.asarray(codes[int(p)][1], dtype=np.int8)
    done = np.zeros(n_ch, dtype=np.int32)
    if factor is None:
        out = np.zeros((n_ch, 17, n_epochs))
        rc = orc().orc_tracking_b1c_nb(P(raw), C.c_size_t(raw.size), C.byref(cs), n_ch, P(prn), P(af), P(cp), P(cf0), P(d), P(pl), n_epochs, P(out), P(done))
    else:
        p61 = np.zeros((n_ch, 12 * int(so.codeLength)), dtype=np.int8)
        for i, p in enumerate(prn):
            if p:
                p61[i] = np.asarray(codes[int(p)][2], dtype=np.int8)
        out = np.zeros((n_ch, 21, n_epochs))
        rc = orc().orc_tracking_b1c_wb(P(raw), C.c_size_t(raw.size), C.byref(cs), n_ch, P(prn), P(af), P(cp), P(cf0), P(d), P(pl), P(p61),
                                       C.c_double(float(factor)), n_epochs, P(out), P(done))
    assert rc == 0, rc
    return [dict({k: out[i, j] for j, k in enumerate(_TRK_NAMES[:out.shape[1]])}, epochsDone=int(done[i])) if prn[i] else None for i in range(n_ch)]
