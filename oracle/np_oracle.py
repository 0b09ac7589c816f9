"""NumPy float64 restatement of the reference's GPS L1 C/A hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; it is the checker for ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

PARITY UNPINNED: the reference (gnsscusdr/CU-SDR-Collection) is 100 % MATLAB,
ships no tests, golden vectors or sample data, and neither MATLAB nor Octave is
available in this image, so it cannot be executed.  This file restates the
reference loops line by line (file:line cited per function, paths relative to
``/root/reference``) and is pinned only by (1) ICD known answers the code
embeds (IS-GPS-200 first-10-chip octals of the C/A codes), (2) an independent
C restatement (``oracle/gnss_oracle.c``) that must agree with it, and (3)
closed-loop known-answer tests on synthetic IF.

MATLAB semantics reproduced here (each is a parity hazard):
1-based indices in every returned index; ``max`` returns the first maximal
index; ``var`` on complex uses N-1; ``round`` is half-away-from-zero; ``rem``
keeps the sign of the dividend (``fmod``); ``atan`` (two quadrant); ``fft``
unscaled / ``ifft`` scaled by 1/len; the floating-point colon operator builds
its vector symmetrically from both ends (Cleve Moler's ``colonop``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

try:  # scipy's pocketfft handles arbitrary lengths and is multi-threaded
    import scipy.fft as _fft

    def _FFT(x, workers=1):
        return _fft.fft(x, axis=-1, workers=workers)

    def _IFFT(x, workers=1):
        return _fft.ifft(x, axis=-1, workers=workers)
except Exception:  # pragma: no cover
    def _FFT(x, workers=1):
        return np.fft.fft(x, axis=-1)

    def _IFFT(x, workers=1):
        return np.fft.ifft(x, axis=-1)


# ---------------------------------------------------------------------------
# settings (GPS/GPS_L1CA/initSettings.m:44-136, hot-path fields only)
# ---------------------------------------------------------------------------
@dataclass
class Settings:
    msToProcess: int = 60000            # initSettings.m:47
    numberOfChannels: int = 12          # :50
    skipNumberOfBytes: int = 0          # :56
    fileName: str = ""                  # :61
    dataType: str = "schar"             # :63
    fileType: int = 2                   # :68
    IF: float = 20e3                    # :71
    samplingFreq: float = 18e6          # :72
    codeFreqBasis: float = 1.023e6      # :73
    codeLength: float = 1023.0          # :76
    skipAcquisition: int = 0            # :80
    acqSatelliteList: list = field(default_factory=lambda: list(range(1, 33)))  # :83
    acqSearchBand: float = 7000.0       # :86
    acqNonCohTime: int = 20             # :88
    acqThreshold: float = 3.5           # :90
    acqSearchStep: float = 500.0        # :92
    resamplingThreshold: float = 8e6    # :94
    resamplingflag: int = 0             # :96
    dllDampingRatio: float = 0.7        # :100
    dllNoiseBandwidth: float = 1.5      # :101
    dllCorrelatorSpacing: float = 0.5   # :102
    pllDampingRatio: float = 0.7        # :105
    pllNoiseBandwidth: float = 20.0     # :106
    intTime: float = 0.001              # :108
    CNo_accTime: float = 0.001          # :133
    CNo_VSMinterval: int = 40           # :135
    freqSpacing: float = 0.0            # GLO/GLO_GL1/initSettings.m:72 (GLONASS only)
    carrFreqBasis: float = 0.0          # BDS/B3I/initSettings.m:132 (B3I, L5C, E5a, E5b, B2a)
    pilotTRKflag: int = 0               # GAL/GAL_E1C/initSettings.m:113 (E1C, L5C, E5a, E5b, B2a)
    signal: str = "GPS_L1CA"            # which reference folder these settings belong to


def matlab_round(x: float) -> float:
    """MATLAB ``round``: half away from zero."""
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def colonop(a: float, d: float, b: float) -> np.ndarray:
    """MATLAB floating-point ``a:d:b`` (Cleve Moler, 'colonop').

    The vector is built from both ends towards the middle so that the last
    element is ``b`` (when ``a+n*d`` is within tolerance of it).  Used by
    tracking.m:252-268 for the three ``tcode`` vectors; ``tcode(blksize)``
    feeds ``remCodePhase`` (tracking.m:273).
    """
    if d == 0 or (a < b and d < 0) or (b < a and d > 0):
        return np.zeros(0)
    tol = 2.0 * np.finfo(float).eps * max(abs(a), abs(b))
    sig = 1.0 if d > 0 else -1.0
    if a == math.floor(a) and d == 1:
        n = int(math.floor(b) - a)
    elif a == math.floor(a) and d == math.floor(d):
        n = int(math.trunc((b - a) / d))
    else:
        n = int(matlab_round((b - a) / d))
        if sig * (a + n * d - b) > tol:
            n -= 1
    c = a + n * d
    if sig * (c - b) > -tol:
        c = b
    out = np.empty(n + 1)
    k = np.arange(0, n // 2 + 1, dtype=np.float64)
    out[: n // 2 + 1] = a + k * d
    out[n - np.arange(0, n // 2 + 1)] = c - k * d
    if n % 2 == 0:
        out[n // 2] = (a + c) / 2
    return out


# ---------------------------------------------------------------------------
# code generation
# ---------------------------------------------------------------------------
_G2S = [5, 6, 7, 8, 17, 18, 139, 140, 141, 251,
        252, 254, 255, 256, 257, 258, 469, 470, 471, 472,
        473, 474, 509, 512, 513, 514, 515, 516, 859, 860,
        861, 862,
        145, 175, 52, 21, 237, 235, 886, 657,
        634, 762, 355, 1012, 176, 603, 130, 359, 595, 68,
        386]


def generateCAcode(PRN: int) -> np.ndarray:
    """GPS/GPS_L1CA/include/generateCAcode.m:42-90 — ±1 C/A chips (1023)."""
    g2shift = _G2S[PRN - 1]
    g1 = np.zeros(1023)
    reg = -np.ones(10)
    for i in range(1023):
        g1[i] = reg[9]
        saveBit = reg[2] * reg[9]
        reg[1:10] = reg[0:9].copy()
        reg[0] = saveBit
    g2 = np.zeros(1023)
    reg = -np.ones(10)
    for i in range(1023):
        g2[i] = reg[9]
        saveBit = reg[1] * reg[2] * reg[5] * reg[7] * reg[8] * reg[9]
        reg[1:10] = reg[0:9].copy()
        reg[0] = saveBit
    g2 = np.concatenate([g2[1023 - g2shift:], g2[:1023 - g2shift]])
    return -(g1 * g2)


def samples_per_code(s: Settings) -> int:
    """acquisition.m:116-117 / makeCaTable.m:43-44."""
    return int(matlab_round(s.samplingFreq / (s.codeFreqBasis / s.codeLength)))


def makeCaTable(PRN: int, s: Settings) -> np.ndarray:
    """GPS/GPS_L1CA/include/makeCaTable.m:43-67 — C/A code resampled to fs."""
    N = samples_per_code(s)
    ts = 1 / s.samplingFreq
    tc = 1 / s.codeFreqBasis
    caCode = generateCAcode(PRN)
    codeValueIndex = np.ceil((ts * np.arange(1, N + 1, dtype=np.float64)) / tc).astype(np.int64)
    codeValueIndex[-1] = 1023
    return caCode[codeValueIndex - 1]


# ---------------------------------------------------------------------------
# read path (postProcessing.m:59-96)
# ---------------------------------------------------------------------------
def read_acq_signal(raw: np.ndarray, s: Settings) -> np.ndarray:
    """postProcessing.m:83-96 — first max(42, nonCoh+2) code periods as a
    complex-double row vector (``raw`` = int8 file bytes, I,Q interleaved for
    fileType 2)."""
    N = samples_per_code(s)
    coef = 1 if s.fileType == 1 else 2
    codeLen = max(42, s.acqNonCohTime + 2)
    off = coef * s.skipNumberOfBytes                           # fseek in bytes (:74); int16 values are two bytes each
    if getattr(s, "dataType", "schar") == "int16":
        off //= 2
    data = raw[off: off + coef * codeLen * N].astype(np.float64)
    if coef == 2:
        data = data[0::2] + 1j * data[1::2]
    return data


# ---------------------------------------------------------------------------
# acquisition (GPS/GPS_L1CA/include/acquisition.m:113-292)
# ---------------------------------------------------------------------------
def acquisition(longSignal: np.ndarray, s: Settings, workers: int = 1, want_grid: bool = False):
    """Line-by-line restatement of acquisition.m:113-292 (resampling branch
    :50-111 not restated: resamplingflag is 0 in every initSettings.m).

    Returns dict with 1x32 ``carrFreq``, ``codePhase``, ``peakMetric`` plus the
    intermediate ``coarseBin`` (1-based, for every searched PRN) and
    ``coarseCodePhase`` used by the parity tests.
    """
    N = samples_per_code(s)                                        # :116
    ts = 1 / s.samplingFreq                                        # :119
    phasePoints = np.arange(0, 2 * N, dtype=np.float64) * 2 * np.pi * ts   # :122
    nBins = int(matlab_round(s.acqSearchBand * 2 / s.acqSearchStep)) + 1   # :124
    coarseFreqBin = np.zeros(nBins)
    res = dict(carrFreq=np.zeros(32), codePhase=np.zeros(32), peakMetric=np.zeros(32),
               coarseBin=np.zeros(32, dtype=np.int64), coarseCodePhase=np.zeros(32, dtype=np.int64))
    fineSearchStep = 25                                            # :138
    numOfFineBins = int(matlab_round(s.acqSearchStep / fineSearchStep)) + 1  # :140
    finePhasePoints = np.arange(0, 40 * N, dtype=np.float64) * 2 * np.pi * ts  # :148
    x = longSignal[:N]
    sigPower = math.sqrt(np.sum(np.abs(x - np.mean(x)) ** 2) / (N - 1) * N)    # :151
    res["sigPower"] = sigPower
    grids = {}
    for PRN in s.acqSatelliteList:                                 # :155
        caCodesTable = makeCaTable(PRN, s)                         # :158
        caCodes2ms = np.concatenate([caCodesTable, np.zeros(N)])   # :160
        results = np.zeros((nBins, 2 * N))                         # :162
        caCodeFreqDom = np.conj(_FFT(caCodes2ms))                  # :164
        for k in range(1, nBins + 1):                              # :167
            coarseFreqBin[k - 1] = s.IF + s.acqSearchBand - s.acqSearchStep * (k - 1)  # :169
            sigCarr = np.exp(-1j * coarseFreqBin[k - 1] * phasePoints)                  # :172
            # :175-191, batched over the non-coherent blocks
            win = np.stack([longSignal[(m - 1) * N: (m + 1) * N]
                            for m in range(1, s.acqNonCohTime + 1)])
            IQfreqDom = _FFT(sigCarr[None, :] * win, workers)      # :180-183
            coh = np.abs(_IFFT(IQfreqDom * caCodeFreqDom[None, :], workers))  # :186-188
            for m in range(coh.shape[0]):                          # :190 (sequential adds)
                results[k - 1, :] += coh[m]
        rowmax = results.max(axis=1)
        acqCoarseBin = int(np.argmax(rowmax)) + 1                  # :196
        colmax = results.max(axis=0)
        codePhase = int(np.argmax(colmax)) + 1                     # :198
        peakSize = colmax[codePhase - 1]
        res["peakMetric"][PRN - 1] = peakSize / sigPower / s.acqNonCohTime   # :200
        res["coarseBin"][PRN - 1] = acqCoarseBin
        res["coarseCodePhase"][PRN - 1] = codePhase
        if want_grid:
            grids[PRN] = results
        if res["peakMetric"][PRN - 1] > s.acqThreshold:            # :206
            caCode = generateCAcode(PRN)                           # :213
            codeValueIndex = np.floor((ts * np.arange(0, 40 * N, dtype=np.float64))
                                      / (1 / s.codeFreqBasis)).astype(np.int64)   # :215
            caCode40ms = caCode[np.fmod(codeValueIndex, int(s.codeLength))]        # :218
            sig40 = longSignal[codePhase - 1: codePhase - 1 + 40 * N]              # :221
            fineFreqBins = np.zeros(numOfFineBins)
            fineResult = np.zeros(numOfFineBins)
            for j in range(1, numOfFineBins + 1):                  # :224
                fineFreqBins[j - 1] = coarseFreqBin[acqCoarseBin - 1] + \
                    s.acqSearchStep / 2 - fineSearchStep * (j - 1)  # :227
                sigCarr40 = np.exp(-1j * fineFreqBins[j - 1] * finePhasePoints)    # :230
                basebandSig = sig40 * caCode40ms * sigCarr40       # :232
                sumPerCode = basebandSig.reshape(40, N).sum(axis=1)  # :235-238
                maxPower = 0.0
                for c in range(1, 21):                             # :243
                    comPower = abs(np.sum(sumPerCode[c - 1: c + 19]))  # :245
                    maxPower = max(maxPower, comPower)             # :247
                fineResult[j - 1] = maxPower                       # :249
            maxFinBin = int(np.argmax(fineResult)) + 1             # :253
            res["carrFreq"][PRN - 1] = fineFreqBins[maxFinBin - 1]  # :254
            res["codePhase"][PRN - 1] = codePhase                  # :256
            if res["carrFreq"][PRN - 1] == 0:                      # :258
                res["carrFreq"][PRN - 1] = 1
            res.setdefault("fineResult", {})[PRN] = fineResult
    if want_grid:
        res["grids"] = grids
    return res


# ---------------------------------------------------------------------------
# preRun (GPS/GPS_L1CA/include/preRun.m:44-72)
# ---------------------------------------------------------------------------
def preRun(acq: dict, s: Settings):
    """preRun.m:60-72 — strongest peaks first, up to numberOfChannels."""
    chans = [dict(PRN=0, acquiredFreq=0.0, codePhase=0, status="-")
             for _ in range(s.numberOfChannels)]
    # MATLAB sort(...,'descend') is stable: ties keep ascending index order
    order = np.argsort(-acq["peakMetric"], kind="stable")
    n = min(s.numberOfChannels, int(np.sum(acq["carrFreq"] != 0)))
    for ii in range(n):
        p = int(order[ii])
        chans[ii] = dict(PRN=p + 1, acquiredFreq=float(acq["carrFreq"][p]),
                         codePhase=int(acq["codePhase"][p]), status="T")
    return chans


# ---------------------------------------------------------------------------
# loop coefficients / C/N0
# ---------------------------------------------------------------------------
def calcLoopCoef(LBW: float, zeta: float, k: float):
    """GPS/GPS_L1CA/Common/calcLoopCoef.m:41-45."""
    Wn = LBW * 8 * zeta / (4 * zeta ** 2 + 1)
    tau1 = k / (Wn * Wn)
    tau2 = 2.0 * zeta / Wn
    return tau1, tau2


def CNoVSM(I: np.ndarray, Q: np.ndarray, T: float) -> float:
    """GPS/GPS_L1CA/Common/CNoVSM.m:38-47."""
    Z = I ** 2 + Q ** 2
    Zm = np.mean(Z)
    Zv = np.var(Z, ddof=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        Pav = np.sqrt(np.complex128(Zm ** 2 - Zv))   # MATLAB sqrt of negative -> complex
        Nv = 0.5 * (Zm - Pav)
        return float(10 * np.log10(np.abs((1 / T) * Pav / (2 * Nv))))


TRACK_FIELDS = ["absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L",
                "Q_E", "Q_P", "Q_L", "dllDiscr", "dllDiscrFilt", "pllDiscr",
                "pllDiscrFilt", "remCodePhase", "remCarrPhase"]


# ---------------------------------------------------------------------------
# tracking (GPS/GPS_L1CA/include/tracking.m:45-371)
# ---------------------------------------------------------------------------
def tracking(raw: np.ndarray, channel: list, s: Settings):
    """Line-by-line restatement of tracking.m:45-371.  ``raw`` plays the role of the open file: the array of stored
    values (int8 for dataType 'schar', int16 for 'int16'; I,Q interleaved for fileType 2, one value per sample for
    fileType 1); fseek/ftell/fread are restated as a cursor in VALUES - for 'int16' the reference's byte offsets
    dataAdaptCoeff*(skip + (codePhase-1)*2) (:145-148) and ftell/dataAdaptCoeff/2 (:212-213) are the same cursor divided by the
    two bytes of a value.  A short read ends the
    whole call with partially filled results and status left '-'
    (tracking.m:241-245)."""
    nE = s.msToProcess
    out = []
    for _ in range(s.numberOfChannels):                            # :48-86
        tr = dict(status="-", PRN=0)
        tr["absoluteSample"] = np.zeros(nE)
        for f in ("codeFreq", "carrFreq", "dllDiscr", "dllDiscrFilt", "pllDiscr",
                  "pllDiscrFilt", "remCodePhase", "remCarrPhase"):
            tr[f] = np.full(nE, np.inf)
        for f in ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L"):
            tr[f] = np.zeros(nE)
        tr["VSMValue"] = np.zeros(nE // s.CNo_VSMinterval)
        tr["VSMIndex"] = np.zeros(nE // s.CNo_VSMinterval)
        out.append(tr)
    earlyLateSpc = s.dllCorrelatorSpacing                          # :94
    PDIcode = s.intTime                                            # :97
    tau1code, tau2code = calcLoopCoef(s.dllNoiseBandwidth, s.dllDampingRatio, 1.0)   # :100
    PDIcarr = s.intTime                                            # :106
    tau1carr, tau2carr = calcLoopCoef(s.pllNoiseBandwidth, s.pllDampingRatio, 0.25)  # :109
    coef = 1 if s.fileType == 1 else 2                             # :126-130
    is16 = getattr(s, "dataType", "schar") == "int16"
    L = int(s.codeLength)
    for ch in range(s.numberOfChannels):                           # :133
        if channel[ch]["PRN"] == 0:                                # :136
            continue
        tr = out[ch]
        tr["PRN"] = channel[ch]["PRN"]                             # :138
        if is16:                                                   # :145-148, bytes -> int16 values
            pos = (coef * (s.skipNumberOfBytes + (channel[ch]["codePhase"] - 1) * 2)) // 2
        else:
            pos = coef * (s.skipNumberOfBytes + channel[ch]["codePhase"] - 1)   # :150 (byte cursor)
        caCode = generateCAcode(channel[ch]["PRN"])                # :156
        caCode = np.concatenate([[caCode[L - 1]], caCode, [caCode[0]]])     # :158
        codeFreq = s.codeFreqBasis                                 # :163
        remCodePhase = 0.0                                         # :165
        carrFreq = channel[ch]["acquiredFreq"]                     # :167
        carrFreqBasis = channel[ch]["acquiredFreq"]                # :168
        remCarrPhase = 0.0                                         # :170
        oldCodeNco = oldCodeError = 0.0                            # :173-174
        oldCarrNco = oldCarrError = 0.0                            # :177-178
        vsmCnt = 0                                                 # :181
        for loopCnt in range(1, nE + 1):                           # :184
            tr["absoluteSample"][loopCnt - 1] = pos / coef         # :215
            codePhaseStep = codeFreq / s.samplingFreq              # :219
            blksize = int(math.ceil((s.codeLength - remCodePhase) / codePhaseStep))   # :222
            nbytes = coef * blksize
            chunk = raw[pos: pos + nbytes]                         # :226
            samplesRead = chunk.size
            pos += samplesRead
            if samplesRead != nbytes:                              # :241-245
                return out
            if coef == 2:
                rawSignal = chunk[0::2].astype(np.float64) + 1j * chunk[1::2].astype(np.float64)  # :233-235
            else:
                rawSignal = chunk.astype(np.float64)
            tr["remCodePhase"][loopCnt - 1] = remCodePhase         # :249
            tcode = colonop(remCodePhase - earlyLateSpc, codePhaseStep,
                            (blksize - 1) * codePhaseStep + remCodePhase - earlyLateSpc)   # :252
            earlyCode = caCode[np.ceil(tcode).astype(np.int64)]    # :255-256 (+1, 1-based)
            tcode = colonop(remCodePhase + earlyLateSpc, codePhaseStep,
                            (blksize - 1) * codePhaseStep + remCodePhase + earlyLateSpc)   # :259
            lateCode = caCode[np.ceil(tcode).astype(np.int64)]     # :262-263
            tcode = colonop(remCodePhase, codePhaseStep,
                            (blksize - 1) * codePhaseStep + remCodePhase)                  # :266
            promptCode = caCode[np.ceil(tcode).astype(np.int64)]   # :269-270
            remCodePhase = (tcode[blksize - 1] + codePhaseStep) - s.codeLength             # :273
            tr["remCarrPhase"][loopCnt - 1] = remCarrPhase         # :277
            time = np.arange(0, blksize + 1, dtype=np.float64) / s.samplingFreq            # :280
            trigarg = ((carrFreq * 2.0 * np.pi) * time) + remCarrPhase                     # :281
            remCarrPhase = math.fmod(trigarg[blksize], 2 * np.pi)  # :283
            carrsig = np.exp(-1j * trigarg[:blksize])              # :287
            bb = carrsig * rawSignal                               # :291-292
            iBB, qBB = bb.real, bb.imag
            I_E = float(np.sum(earlyCode * iBB)); Q_E = float(np.sum(earlyCode * qBB))     # :295-296
            I_P = float(np.sum(promptCode * iBB)); Q_P = float(np.sum(promptCode * qBB))   # :297-298
            I_L = float(np.sum(lateCode * iBB)); Q_L = float(np.sum(lateCode * qBB))       # :299-300
            with np.errstate(divide="ignore", invalid="ignore"):
                carrError = float(np.arctan(np.float64(Q_P) / np.float64(I_P)) / (2.0 * np.pi))   # :305
            carrNco = oldCarrNco + (tau2carr / tau1carr) * (carrError - oldCarrError) \
                + carrError * (PDIcarr / tau1carr)                 # :308
            oldCarrNco = carrNco; oldCarrError = carrError         # :310-311
            tr["carrFreq"][loopCnt - 1] = carrFreq                 # :314
            carrFreq = carrFreqBasis + carrNco                     # :317
            sE = math.sqrt(I_E * I_E + Q_E * Q_E); sL = math.sqrt(I_L * I_L + Q_L * Q_L)
            with np.errstate(divide="ignore", invalid="ignore"):
                codeError = float((np.float64(sE) - sL) / (np.float64(sE) + sL))           # :322
            codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) \
                + codeError * (PDIcode / tau1code)                 # :326
            oldCodeNco = codeNco; oldCodeError = codeError         # :328-329
            tr["codeFreq"][loopCnt - 1] = codeFreq                 # :332
            codeFreq = s.codeFreqBasis - codeNco                   # :335
            tr["dllDiscr"][loopCnt - 1] = codeError                # :338-341
            tr["dllDiscrFilt"][loopCnt - 1] = codeNco
            tr["pllDiscr"][loopCnt - 1] = carrError
            tr["pllDiscrFilt"][loopCnt - 1] = carrNco
            tr["I_E"][loopCnt - 1] = I_E; tr["I_P"][loopCnt - 1] = I_P; tr["I_L"][loopCnt - 1] = I_L   # :343-348
            tr["Q_E"][loopCnt - 1] = Q_E; tr["Q_P"][loopCnt - 1] = Q_P; tr["Q_L"][loopCnt - 1] = Q_L
            if loopCnt % s.CNo_VSMinterval == 0:                   # :351
                vsmCnt += 1
                lo = loopCnt - s.CNo_VSMinterval
                tr["VSMValue"][vsmCnt - 1] = CNoVSM(tr["I_P"][lo:loopCnt], tr["Q_P"][lo:loopCnt],
                                                    s.CNo_accTime)   # :353
                tr["VSMIndex"][vsmCnt - 1] = loopCnt               # :356
        tr["status"] = channel[ch]["status"]                       # :365
    return out


# ===========================================================================
# GLONASS L1/L2 (GLO/GLO_GL1 and GLO/GLO_GL2: identical code, different freqSpacing)
# paths below relative to /root/reference/GLO/GLO_GL1/
# ===========================================================================
def glo_settings(**kw) -> Settings:
    """GLO/GLO_GL1/initSettings.m:44-146 defaults (hot-path fields)."""
    s = Settings(IF=0.0, samplingFreq=12e6, codeFreqBasis=0.511e6, codeLength=511.0,
                 acqSatelliteList=list(range(-7, 7)), acqSearchBand=5000.0, acqThreshold=2.0,
                 dllNoiseBandwidth=2.0, pllNoiseBandwidth=25.0, freqSpacing=562.5e3)
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def glo_code() -> np.ndarray:
    """include/generateCAcode.m:95-108 — 511-chip ST code, +-1 (PRN == 0 branch)."""
    reg = -np.ones(9)
    code = np.zeros(511)
    for i in range(511):
        code[i] = reg[6]
        save1 = reg[4] * reg[8]
        reg[1:9] = reg[0:8].copy()
        reg[0] = save1
    return code


def generateCAcode_glo(sampFreq: float, numSamples: int) -> np.ndarray:
    """include/generateCAcode.m:110-116 — the code resampled with a floating-point colon vector."""
    code = glo_code()
    stepSize = 511e3 / sampFreq
    samples = colonop(0.0, stepSize, (numSamples * stepSize) - stepSize)
    samples = np.fmod(np.floor(samples), 511).astype(np.int64)
    return code[samples]


def read_acq_signal_glo(raw: np.ndarray, s: Settings) -> np.ndarray:
    """include/postProcessing.m:83-96 — note data = data2 + 1i*data1 (Q + iI)."""
    N = samples_per_code(s)
    codeLen = max(42, s.acqNonCohTime + 2)
    off = 2 * s.skipNumberOfBytes
    data = raw[off: off + 2 * codeLen * N].astype(np.float64)
    return data[1::2] + 1j * data[0::2]


def acquisition_glo(longSignal: np.ndarray, s: Settings, workers: int = 1):
    """include/acquisition.m:121-292.  Result vectors are 1x21, indexed K+8 (0-based K+7)."""
    N = samples_per_code(s)                                        # :121
    ts = 1 / s.samplingFreq
    phasePoints = np.arange(0, 2 * N, dtype=np.float64) * 2 * np.pi * ts       # :127
    nBins = int(matlab_round(s.acqSearchBand * 2 / s.acqSearchStep)) + 1       # :129
    coarseFreqBin = np.zeros(nBins)
    res = dict(carrFreq=np.zeros(21), codePhase=np.zeros(21), peakMetric=np.zeros(21),
               coarseBin=np.zeros(21, dtype=np.int64), coarseCodePhase=np.zeros(21, dtype=np.int64))
    caCode = generateCAcode_glo(s.samplingFreq, N)                 # :145
    caCodeFreqDom = np.conj(_FFT(np.concatenate([caCode, np.zeros(N)])))       # :147-149
    fineSearchStep = 25
    numOfFineBins = int(matlab_round(s.acqSearchStep / fineSearchStep)) + 1
    caCode40ms = generateCAcode_glo(s.samplingFreq, N * 40)        # :164
    finePhasePoints = np.arange(0, 40 * N, dtype=np.float64) * 2 * np.pi * ts  # :166
    x = longSignal[:N]
    sigPower = math.sqrt(np.sum(np.abs(x - np.mean(x)) ** 2) / (N - 1) * N)    # :169
    res["sigPower"] = sigPower
    for K in s.acqSatelliteList:                                   # :173
        results = np.zeros((nBins, 2 * N))
        for k in range(1, nBins + 1):
            coarseFreqBin[k - 1] = s.IF - s.freqSpacing * K + s.acqSearchBand - s.acqSearchStep * (k - 1)   # :181
            sigCarr = np.exp(-1j * coarseFreqBin[k - 1] * phasePoints)
            win = np.stack([longSignal[(m - 1) * N: (m + 1) * N] for m in range(1, s.acqNonCohTime + 1)])
            coh = np.abs(_IFFT(_FFT(sigCarr[None, :] * win, workers) * caCodeFreqDom[None, :], workers))    # :190-202
            for m in range(coh.shape[0]):
                results[k - 1, :] += coh[m]
        acqCoarseBin = int(np.argmax(results.max(axis=1))) + 1     # :208
        colmax = results.max(axis=0)
        codePhase = int(np.argmax(colmax)) + 1                     # :210
        i = K + 7
        res["peakMetric"][i] = colmax[codePhase - 1] / sigPower / s.acqNonCohTime   # :212
        res["coarseBin"][i] = acqCoarseBin
        res["coarseCodePhase"][i] = codePhase
        if res["peakMetric"][i] > s.acqThreshold:                  # :218
            sig40 = longSignal[codePhase - 1: codePhase - 1 + 40 * N]   # :226
            fineFreqBins = np.zeros(numOfFineBins)
            fineResult = np.zeros(numOfFineBins)
            for j in range(1, numOfFineBins + 1):
                fineFreqBins[j - 1] = coarseFreqBin[acqCoarseBin - 1] + s.acqSearchStep / 2 - fineSearchStep * (j - 1)   # :232
                basebandSig = sig40 * caCode40ms * np.exp(-1j * fineFreqBins[j - 1] * finePhasePoints)   # :235-237
                sumPerCode = basebandSig.reshape(40, N).sum(axis=1)   # :240-243
                maxPower = 0.0
                for c in range(1, 21):                             # :248
                    comPower = abs(np.sum(sumPerCode[c - 1: c + 9]) - np.sum(sumPerCode[c + 9: c + 19]))   # :250
                    maxPower = max(maxPower, comPower)
                fineResult[j - 1] = maxPower
            maxFinBin = int(np.argmax(fineResult)) + 1             # :258
            res["carrFreq"][i] = fineFreqBins[maxFinBin - 1]       # :259
            res["codePhase"][i] = codePhase                        # :261
            if res["carrFreq"][i] == 0:
                res["carrFreq"][i] = 1
    return res


def preRun_glo(acq: dict, s: Settings):
    """include/preRun.m:44-72 — channels carry the frequency number K = Kindexes(ii)-8."""
    chans = [dict(K=0, acquiredFreq=0.0, codePhase=0, status="-") for _ in range(s.numberOfChannels)]
    order = np.argsort(-acq["peakMetric"], kind="stable")
    n = min(s.numberOfChannels, int(np.sum(acq["carrFreq"] != 0)))
    for ii in range(n):
        p = int(order[ii])
        chans[ii] = dict(K=p + 1 - 8, acquiredFreq=float(acq["carrFreq"][p]), codePhase=int(acq["codePhase"][p]), status="T")
    return chans


def calcLoopCoefCarr(s: Settings):
    """GLO/GLO_GL1/Common/calcLoopCoefCarr.m:41-56."""
    Wn = 1.2 * s.pllNoiseBandwidth
    return Wn ** 3 * s.intTime ** 2, 2 * Wn ** 2 * s.intTime, 2 * Wn     # pf3, pf2, pf1


def tracking_glo(raw: np.ndarray, channel: list, s: Settings):
    """GLO/GLO_GL1/include/tracking.m:45-358: as GPS L1 C/A except the shared 511-chip code (:88-90), the
    channel test status ~= '-' (:137), rawSignal = Q + 1i*I (:227) and the carrier loop filter (:281-285)."""
    return _tracking_pf(raw, channel, s, "GLO")


def tracking_b3i(raw: np.ndarray, channel: list, s: Settings):
    """BDS/B3I/include/tracking.m:45-352: the GLONASS-style loop with the per-PRN 10230-chip code (:55-56),
    the PRN ~= 0 channel test (:44), no I/Q swap (:96) and the carrier-aided code NCO centre
    channel.codeFreq (:57, :146)."""
    return _tracking_pf(raw, channel, s, "B3I")


def _tracking_pf(raw: np.ndarray, channel: list, s: Settings, mode: str):
    glo = mode == "GLO"
    nE = s.msToProcess
    out = []
    for _ in range(s.numberOfChannels):
        tr = dict(status="-", PRN=None)
        tr["absoluteSample"] = np.zeros(nE)
        for f in ("codeFreq", "carrFreq", "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase"):
            tr[f] = np.full(nE, np.inf)
        for f in ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L"):
            tr[f] = np.zeros(nE)
        tr["VSMValue"] = np.zeros(nE // s.CNo_VSMinterval)
        tr["VSMIndex"] = np.zeros(nE // s.CNo_VSMinterval)
        out.append(tr)
    if glo:
        caCode = glo_code()                                        # GLO :88 (generateCAcode(0, codeFreqBasis, 511))
        caCode = np.concatenate([[caCode[510]], caCode, [caCode[0]]])  # GLO :90
    earlyLateSpc = s.dllCorrelatorSpacing
    PDIcode = s.intTime
    tau1code, tau2code = calcLoopCoef(s.dllNoiseBandwidth, s.dllDampingRatio, 1.0)     # :104
    pf3, pf2, pf1 = calcLoopCoefCarr(s)                            # :110
    for ch in range(s.numberOfChannels):
        if glo:
            if channel[ch]["status"] == "-":                       # GLO :137
                continue
            tr = out[ch]
            tr["PRN"] = channel[ch]["K"]                           # GLO :141
            codeFreqCentre = s.codeFreqBasis
        else:
            if channel[ch]["PRN"] == 0:                            # B3I :44
                continue
            tr = out[ch]
            tr["PRN"] = channel[ch]["PRN"]
            code = generateB3Icode(channel[ch]["PRN"])             # B3I :55
            caCode = np.concatenate([[code[-1]], code, [code[0]]])  # B3I :56
            codeFreqCentre = channel[ch]["codeFreq"]               # B3I :57
        pos = 2 * (s.skipNumberOfBytes + channel[ch]["codePhase"] - 1)     # :148
        codeFreq = codeFreqCentre; remCodePhase = 0.0
        carrFreq = channel[ch]["acquiredFreq"]; carrFreqBasis = channel[ch]["acquiredFreq"]; remCarrPhase = 0.0
        oldCodeNco = oldCodeError = 0.0
        d2CarrError = dCarrError = 0.0                             # :171-172
        vsmCnt = 0
        for loopCnt in range(1, nE + 1):
            tr["absoluteSample"][loopCnt - 1] = pos / 2            # :206
            codePhaseStep = codeFreq / s.samplingFreq
            blksize = int(math.ceil((s.codeLength - remCodePhase) / codePhaseStep))
            chunk = raw[pos: pos + 2 * blksize]
            pos += chunk.size
            if chunk.size != 2 * blksize:                          # :232-236
                return out
            if glo:
                rawSignal = chunk[1::2].astype(np.float64) + 1j * chunk[0::2].astype(np.float64)   # GLO :227 (Q + iI)
            else:
                rawSignal = chunk[0::2].astype(np.float64) + 1j * chunk[1::2].astype(np.float64)   # B3I :96
            tr["remCodePhase"][loopCnt - 1] = remCodePhase
            tE = colonop(remCodePhase - earlyLateSpc, codePhaseStep, (blksize - 1) * codePhaseStep + remCodePhase - earlyLateSpc)
            tL = colonop(remCodePhase + earlyLateSpc, codePhaseStep, (blksize - 1) * codePhaseStep + remCodePhase + earlyLateSpc)
            tP = colonop(remCodePhase, codePhaseStep, (blksize - 1) * codePhaseStep + remCodePhase)
            earlyCode = caCode[np.ceil(tE).astype(np.int64)]
            lateCode = caCode[np.ceil(tL).astype(np.int64)]
            promptCode = caCode[np.ceil(tP).astype(np.int64)]
            remCodePhase = (tP[blksize - 1] + codePhaseStep) - s.codeLength
            tr["remCarrPhase"][loopCnt - 1] = remCarrPhase
            time = np.arange(0, blksize + 1, dtype=np.float64) / s.samplingFreq
            trigarg = ((carrFreq * 2.0 * np.pi) * time) + remCarrPhase
            remCarrPhase = math.fmod(trigarg[blksize], 2 * np.pi)
            bb = np.exp(-1j * trigarg[:blksize]) * rawSignal
            I_E = float(np.sum(earlyCode * bb.real)); Q_E = float(np.sum(earlyCode * bb.imag))
            I_P = float(np.sum(promptCode * bb.real)); Q_P = float(np.sum(promptCode * bb.imag))
            I_L = float(np.sum(lateCode * bb.real)); Q_L = float(np.sum(lateCode * bb.imag))
            with np.errstate(divide="ignore", invalid="ignore"):
                carrError = float(np.arctan(np.float64(Q_P) / np.float64(I_P)) / (2.0 * np.pi))
            d2CarrError = d2CarrError + carrError * pf3            # :282
            dCarrError = d2CarrError + carrError * pf2 + dCarrError   # :283
            carrNco = dCarrError + carrError * pf1                 # :285
            tr["carrFreq"][loopCnt - 1] = carrFreq
            carrFreq = carrFreqBasis + carrNco
            sE = math.sqrt(I_E * I_E + Q_E * Q_E); sL = math.sqrt(I_L * I_L + Q_L * Q_L)
            with np.errstate(divide="ignore", invalid="ignore"):
                codeError = float((np.float64(sE) - sL) / (np.float64(sE) + sL))
            codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code)
            oldCodeNco = codeNco; oldCodeError = codeError
            tr["codeFreq"][loopCnt - 1] = codeFreq
            codeFreq = codeFreqCentre - codeNco                    # GLO :313 codeFreqBasis ; B3I :146 channel.codeFreq
            tr["dllDiscr"][loopCnt - 1] = codeError; tr["dllDiscrFilt"][loopCnt - 1] = codeNco
            tr["pllDiscr"][loopCnt - 1] = carrError; tr["pllDiscrFilt"][loopCnt - 1] = carrNco
            tr["I_E"][loopCnt - 1] = I_E; tr["I_P"][loopCnt - 1] = I_P; tr["I_L"][loopCnt - 1] = I_L
            tr["Q_E"][loopCnt - 1] = Q_E; tr["Q_P"][loopCnt - 1] = Q_P; tr["Q_L"][loopCnt - 1] = Q_L
            if loopCnt % s.CNo_VSMinterval == 0:
                vsmCnt += 1
                lo = loopCnt - s.CNo_VSMinterval
                tr["VSMValue"][vsmCnt - 1] = CNoVSM(tr["I_P"][lo:loopCnt], tr["Q_P"][lo:loopCnt], s.CNo_accTime)
                tr["VSMIndex"][vsmCnt - 1] = loopCnt
        tr["status"] = channel[ch]["status"]
    return out


# ===========================================================================
# BeiDou B3I (BDS/B3I); paths below relative to /root/reference/BDS/B3I/
# ===========================================================================
def b3i_settings(**kw) -> Settings:
    """initSettings.m:44-132 defaults (hot-path fields)."""
    s = Settings(numberOfChannels=15, codeLength=10230.0, codeFreqBasis=10.23e6, acqSatelliteList=list(range(1, 64)),
                 acqSearchBand=5000.0, acqNonCohTime=10, acqThreshold=3.0, dllNoiseBandwidth=2.0, pllNoiseBandwidth=15.0,
                 carrFreqBasis=1268.520e6)
    for k, v in kw.items():
        setattr(s, k, v)
    return s


_B3I_INIT = [4, 11, 13, 22, 30, 36, 44, 48, 88, 104, 116, 129, 376, 418, 458, 682, 696, 707, 1078, 2069,
             2248, 2574, 2596, 2731, 4294, 4436, 4647, 4978, 4986, 1, 5209, 5539, 6061, 6488, 7130, 7165,
             7403, 5879, 1681, 5080, 5938, 3983, 6208, 7223, 2996, 1814, 6906, 6144, 4713, 7406, 7264, 1766,
             5347, 3515, 7951, 7054, 3884, 6067, 4230, 3803, 869, 3683, 1205]
_B3I_CACHE = {}


def generateB3Icode(PRN: int) -> np.ndarray:
    """include/generateB3Icode.m:33-86 — +-1 chips (10230)."""
    if PRN in _B3I_CACHE:
        return _B3I_CACHE[PRN]
    CodeLength = 10230
    ca_reg = -np.ones(13)
    CA = np.zeros(CodeLength)
    reset_state = np.array([-1] * 11 + [1, 1], dtype=np.float64)
    for ind in range(CodeLength):                                 # :40-49
        CA[ind] = ca_reg[-1]
        if np.array_equal(ca_reg, reset_state):
            ca_reg = -np.ones(13)
        else:
            feedback = ca_reg[0] * ca_reg[2] * ca_reg[3] * ca_reg[12]
            ca_reg = np.roll(ca_reg, 1)
            ca_reg[0] = feedback
    cb_reg = -np.ones(13)
    CB = np.zeros(CodeLength)
    fb_pos = [0, 4, 5, 6, 8, 9, 11, 12]
    for _ in range(_B3I_INIT[PRN - 1]):                           # :68-72
        feedback = np.prod(cb_reg[fb_pos])
        cb_reg = np.roll(cb_reg, 1)
        cb_reg[0] = feedback
    for ind in range(CodeLength):                                 # :75-80
        CB[ind] = cb_reg[-1]
        feedback = np.prod(cb_reg[fb_pos])
        cb_reg = np.roll(cb_reg, 1)
        cb_reg[0] = feedback
    _B3I_CACHE[PRN] = CB * CA                                     # :83
    return _B3I_CACHE[PRN]


def makeB3ITable(PRN: int, s: Settings) -> np.ndarray:
    """include/makeB3ITable.m:38-52."""
    N = samples_per_code(s)
    ts = 1 / s.samplingFreq
    tc = 1 / s.codeFreqBasis
    code = generateB3Icode(PRN)
    idx = np.ceil((ts * np.arange(1, N + 1, dtype=np.float64)) / tc).astype(np.int64)
    idx[-1] = 10230
    return code[idx - 1]


def read_acq_signal_b3i(raw: np.ndarray, s: Settings) -> np.ndarray:
    """include/postProcessing.m:80-94 — max(22, acqNonCohTime+1) code periods."""
    N = samples_per_code(s)
    codeLen = max(22, s.acqNonCohTime + 1)
    off = 2 * s.skipNumberOfBytes
    data = raw[off: off + 2 * codeLen * N].astype(np.float64)
    return data[0::2] + 1j * data[1::2]


_NH = np.array([1, 1, 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1], dtype=np.float64)   # acquisition.m:127


def acquisition_b3i(longSignal: np.ndarray, s: Settings, workers: int = 1):
    """include/acquisition.m:107-237 (resampling branch not restated, resamplingFlag == 0)."""
    N = samples_per_code(s)
    ts = 1 / s.samplingFreq
    phasePoints = np.arange(0, 2 * N, dtype=np.float64) * 2 * np.pi * ts
    nBins = int(matlab_round(s.acqSearchBand * 2 / s.acqSearchStep)) + 1
    coarseFreqBin = np.zeros(nBins)
    res = dict(carrFreq=np.zeros(63), codePhase=np.zeros(63), peakMetric=np.zeros(63),
               coarseBin=np.zeros(63, dtype=np.int64), coarseCodePhase=np.zeros(63, dtype=np.int64))
    fineSearchStep = 25
    numOfFineBins = int(matlab_round(s.acqSearchStep / fineSearchStep)) + 1
    finePhasePoints = np.arange(0, 20 * N, dtype=np.float64) * 2 * np.pi * ts       # :133
    x = longSignal[:N]
    sigPower = math.sqrt(np.sum(np.abs(x - np.mean(x)) ** 2) / (N - 1) * N)         # :135
    res["sigPower"] = sigPower
    for PRN in s.acqSatelliteList:                                                    # :139
        table = makeB3ITable(PRN, s)
        codeFreqDom = np.conj(_FFT(np.concatenate([table, np.zeros(N)])))             # :142-146
        results = np.zeros((nBins, 2 * N))
        for k in range(1, nBins + 1):
            coarseFreqBin[k - 1] = s.IF + s.acqSearchBand - s.acqSearchStep * (k - 1)
            sigCarr = np.exp(-1j * coarseFreqBin[k - 1] * phasePoints)
            win = np.stack([longSignal[(m - 1) * N: (m + 1) * N] for m in range(1, s.acqNonCohTime + 1)])
            coh = np.abs(_IFFT(_FFT(sigCarr[None, :] * win, workers) * codeFreqDom[None, :], workers))
            for m in range(coh.shape[0]):
                results[k - 1, :] += coh[m]
        acqCoarseBin = int(np.argmax(results.max(axis=1))) + 1                       # :164
        colmax = results.max(axis=0)
        codePhase = int(np.argmax(colmax)) + 1                                       # :166
        res["peakMetric"][PRN - 1] = colmax[codePhase - 1] / sigPower / s.acqNonCohTime   # :168
        res["coarseBin"][PRN - 1] = acqCoarseBin
        res["coarseCodePhase"][PRN - 1] = codePhase
        if res["peakMetric"][PRN - 1] > s.acqThreshold:                              # :170
            code = generateB3Icode(PRN)
            codeValueIndex = np.floor((ts * np.arange(0, 20 * N, dtype=np.float64)) / (1 / s.codeFreqBasis)).astype(np.int64)
            code20 = code[np.fmod(codeValueIndex, int(s.codeLength))]                # :174-177
            sig20 = longSignal[codePhase - 1: codePhase - 1 + 20 * N]                # :179
            fineFreqBins = np.zeros(numOfFineBins)
            fineResult = np.zeros(numOfFineBins)
            for j in range(1, numOfFineBins + 1):
                fineFreqBins[j - 1] = coarseFreqBin[acqCoarseBin - 1] + s.acqSearchStep / 2 - fineSearchStep * (j - 1)
                basebandSig = sig20 * code20 * np.exp(-1j * fineFreqBins[j - 1] * finePhasePoints)
                sumPerCode = basebandSig.reshape(20, N).sum(axis=1)                  # :188-191
                if (1 <= PRN <= 5) or (59 <= PRN <= 63):                             # :193 GEO: 2 ms bits
                    comPower1 = np.sum(np.abs(sumPerCode.reshape(10, 2).sum(axis=1)))
                    comPower2 = np.sum(np.abs(sumPerCode[[0, 19]])) + np.sum(np.abs(sumPerCode[1:19].reshape(9, 2).sum(axis=1)))
                    maxPower = max(comPower1, comPower2)
                else:                                                                # :199 MEO/IGSO: NH code
                    maxPower = abs(np.sum(sumPerCode * _NH))
                    for comIndex in range(1, 20):
                        NHshift = np.roll(_NH, comIndex)
                        sNH = sumPerCode * NHshift
                        comPower = abs(np.sum(sNH[:comIndex])) + abs(np.sum(sNH[comIndex:]))
                        maxPower = max(maxPower, comPower)
                fineResult[j - 1] = maxPower
            maxFinBin = int(np.argmax(fineResult)) + 1
            res["carrFreq"][PRN - 1] = fineFreqBins[maxFinBin - 1]                   # :216
            res["codePhase"][PRN - 1] = codePhase
            if res["carrFreq"][PRN - 1] == 0:
                res["carrFreq"][PRN - 1] = 1
    return res


def preRun_b3i(acq: dict, s: Settings):
    """include/preRun.m:44-77 — adds the carrier-aided code NCO centre channel.codeFreq (:71-73)."""
    chans = [dict(PRN=0, acquiredFreq=0.0, codePhase=0, codeFreq=0.0, status="-") for _ in range(s.numberOfChannels)]
    order = np.argsort(-acq["peakMetric"], kind="stable")
    n = min(s.numberOfChannels, int(np.sum(acq["carrFreq"] != 0)))
    for ii in range(n):
        p = int(order[ii])
        af = float(acq["carrFreq"][p])
        chans[ii] = dict(PRN=p + 1, acquiredFreq=af, codePhase=int(acq["codePhase"][p]),
                         codeFreq=s.codeFreqBasis + (af - s.IF) / s.carrFreqBasis * s.codeFreqBasis, status="T")
    return chans


# ===========================================================================
# Galileo E1 (GAL/GAL_E1C); paths below relative to /root/reference/GAL/GAL_E1C/
# ===========================================================================
# The primary codes are the ICD memory codes; the reference reads them at run time from
# include/E1b.dat / include/E1c.dat (generateE1Bcode.m:44-55).  They are DATA, not algorithm: the
# restatement takes them as an argument (``codes[PRN] = (e1b_bits, e1c_bits)``, 0/1 arrays of 4092),
# so tests can run on the real tables where the reference tree is mounted and on seeded stand-in
# tables on the GPU box.
def e1c_settings(**kw) -> Settings:
    """initSettings.m:44-140 defaults (hot-path fields)."""
    s = Settings(codeLength=4092.0, acqSatelliteList=list(range(1, 37)), acqSearchBand=7000.0, acqNonCohTime=1,
                 acqSearchStep=150.0, acqThreshold=10.0, resamplingThreshold=50e6, dllCorrelatorSpacing=0.3,
                 pllNoiseBandwidth=15.0, intTime=0.004, CNo_accTime=0.004, CNo_VSMinterval=400)
    s.pilotTRKflag = 1                                             # :113
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def read_e1_dat(path: str) -> np.ndarray:
    """fscanf(fid, '%d', 4092*50) of E1b.dat / E1c.dat (generateE1Bcode.m:47-51) -> [50][4092] 0/1."""
    with open(path) as f:
        v = np.array(f.read().split(), dtype=np.int64)
    return v[: 4092 * 50].reshape(50, 4092)


def generateE1code(bits: np.ndarray) -> np.ndarray:
    """generateE1Bcode.m:55-64 / generateE1Ccode.m: 1-2*bit, then the BOC(1,1) sub-carrier [c -c] -> 8184 values."""
    raw = 1.0 - 2.0 * np.asarray(bits, dtype=np.float64)
    out = np.empty(2 * raw.size)
    out[0::2] = raw
    out[1::2] = -raw
    return out


def makeE1Table(code: np.ndarray, s: Settings) -> np.ndarray:
    """makeE1BTable.m:38-58 / makeE1CTable.m (``code`` = the 8184-value BOC code)."""
    N = samples_per_code(s)
    ts = 1 / s.samplingFreq
    tc = 1 / s.codeFreqBasis / 2                                   # :43
    idx = np.ceil((ts * np.arange(1, N + 1, dtype=np.float64)) / tc).astype(np.int64)   # :51
    idx[-1] = int(s.codeLength) * 2                                # :54
    idx[0] = 1                                                     # :55
    return code[idx - 1]


_E1_SECONDARY = np.array([1, 1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, -1, 1, -1, 1, -1, -1, 1, -1, -1, 1, 1, -1, 1],
                         dtype=np.float64)                         # acquisition.m:135 ('380AD90')


def acquisition_e1c(longSignal: np.ndarray, s: Settings, codes: dict, workers: int = 1):
    """include/acquisition.m:112-292 (resampling branch not restated, resamplingflag == 0)."""
    N = samples_per_code(s)                                        # :114
    ts = 1 / s.samplingFreq
    phasePoints = np.arange(0, 2 * N, dtype=np.float64) * 2 * np.pi * ts   # :120
    nBins = int(matlab_round(s.acqSearchBand * 2 / s.acqSearchStep)) + 1   # :122
    coarseFreqBin = np.zeros(nBins)
    res = dict(carrFreq=np.zeros(50), codePhase=np.zeros(50), peakMetric=np.zeros(50),
               coarseBin=np.zeros(50, dtype=np.int64), coarseCodePhase=np.zeros(50, dtype=np.int64))
    fineSearchStep = 10                                            # :138
    numOfFineBins = int(matlab_round(s.acqSearchStep / fineSearchStep)) + 1   # :140
    finePhasePoints = np.arange(0, 25 * N, dtype=np.float64) * 2 * np.pi * ts   # :148
    x = longSignal[:N]
    sigPower = math.sqrt(np.sum(np.abs(x - np.mean(x)) ** 2) / (N - 1) * N)     # :151
    res["sigPower"] = sigPower
    for PRN in s.acqSatelliteList:                                 # :155
        E1bCode = generateE1code(codes[PRN][0])
        E1cCode = generateE1code(codes[PRN][1])
        E1bFreqDom = np.conj(_FFT(np.concatenate([makeE1Table(E1bCode, s), np.zeros(N)])))   # :158-171
        E1cFreqDom = np.conj(_FFT(np.concatenate([makeE1Table(E1cCode, s), np.zeros(N)])))
        results = np.zeros((nBins, 2 * N))
        for k in range(1, nBins + 1):                              # :174
            coarseFreqBin[k - 1] = s.IF + s.acqSearchBand - s.acqSearchStep * (k - 1)   # :176
            sigCarr = np.exp(-1j * coarseFreqBin[k - 1] * phasePoints)                   # :179
            for m in range(1, s.acqNonCohTime + 1):                # :182
                signal = longSignal[(m - 1) * N: (m + 1) * N]      # :184
                IQfreqDom = _FFT(sigCarr * signal, workers)        # :186-189
                coh = np.abs(_IFFT(IQfreqDom * E1bFreqDom, workers)) + np.abs(_IFFT(IQfreqDom * E1cFreqDom, workers))   # :192-196
                results[k - 1, :] += coh                           # :198
        acqCoarseBin = int(np.argmax(results.max(axis=1))) + 1     # :204
        colmax = results.max(axis=0)
        codePhase = int(np.argmax(colmax)) + 1                     # :206
        res["peakMetric"][PRN - 1] = colmax[codePhase - 1] / sigPower / s.acqNonCohTime   # :208
        res["coarseBin"][PRN - 1] = acqCoarseBin
        res["coarseCodePhase"][PRN - 1] = codePhase
        if res["peakMetric"][PRN - 1] > s.acqThreshold:            # :212
            codeValueIndex = np.floor((ts * np.arange(0, 25 * N, dtype=np.float64)) /
                                      (1 / s.codeFreqBasis / 2)).astype(np.int64)        # :219
            E1cCode25ms = E1cCode[np.fmod(codeValueIndex, int(s.codeLength) * 2)]        # :222
            sig25ms = longSignal[codePhase - 1: codePhase - 1 + 25 * N]                  # :224
            fineFreqBins = np.zeros(numOfFineBins)
            fineResult = np.zeros(numOfFineBins)
            for j in range(1, numOfFineBins + 1):                  # :227
                fineFreqBins[j - 1] = coarseFreqBin[acqCoarseBin - 1] + s.acqSearchStep / 2 - fineSearchStep * (j - 1)   # :230
                basebandSig = sig25ms * E1cCode25ms * np.exp(-1j * fineFreqBins[j - 1] * finePhasePoints)   # :233-235
                sumPerCode = basebandSig.reshape(25, N).sum(axis=1)                      # :237-240
                maxPower = abs(np.sum(sumPerCode * _E1_SECONDARY))                       # :246
                for comIndex in range(1, 25):                                            # :248
                    s2 = sumPerCode * np.roll(_E1_SECONDARY, comIndex)                   # :250-252
                    comPower = abs(np.sum(s2[:comIndex])) + abs(np.sum(s2[comIndex:]))   # :254
                    maxPower = max(maxPower, comPower)
                fineResult[j - 1] = maxPower
            maxFinBin = int(np.argmax(fineResult)) + 1             # :263
            res["carrFreq"][PRN - 1] = fineFreqBins[maxFinBin - 1]
            res["codePhase"][PRN - 1] = codePhase
            if res["carrFreq"][PRN - 1] == 0:
                res["carrFreq"][PRN - 1] = 1
    return res


def tracking_e1c(raw: np.ndarray, channel: list, s: Settings, codes: dict):
    """include/tracking.m:45-383: 4 ms epochs, BOC(1,1) sub-chip tables indexed by ceil(tcode*2)+1 (:236-262),
    pilot component correlated with the same code phase and both discriminators averaged when
    settings.pilotTRKflag == 1 (:297-300, :327-333), three-coefficient carrier filter (:303-305)."""
    nE = int(matlab_round(s.msToProcess / 1000 / s.intTime))       # :48
    nV = int(math.floor(s.msToProcess / 4 / s.CNo_VSMinterval))    # :70-73
    out = []
    for _ in range(s.numberOfChannels):
        tr = dict(status="-", PRN=0)
        tr["absoluteSample"] = np.zeros(nE)
        for f in ("codeFreq", "carrFreq", "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase"):
            tr[f] = np.full(nE, np.inf)
        for f in ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L"):
            tr[f] = np.zeros(nE)
        tr["VSMValue"] = np.zeros(nV)
        tr["VSMIndex"] = np.zeros(nV)
        out.append(tr)
    earlyLateSpc = s.dllCorrelatorSpacing                          # :82
    PDIcode = s.intTime                                            # :85
    tau1code, tau2code = calcLoopCoef(s.dllNoiseBandwidth, s.dllDampingRatio, 1.0)   # :88
    pf3, pf2, pf1 = calcLoopCoefCarr(s)                            # :92
    pilot = int(getattr(s, "pilotTRKflag", 1)) == 1
    L2 = int(s.codeLength) * 2
    for ch in range(s.numberOfChannels):
        if channel[ch]["PRN"] == 0:                                # :116
            continue
        tr = out[ch]
        PRN = channel[ch]["PRN"]
        tr["PRN"] = PRN
        pos = 2 * (s.skipNumberOfBytes + channel[ch]["codePhase"] - 1)    # :120
        c = generateE1code(codes[PRN][0])
        E1bCode = np.concatenate([[c[L2 - 1]], c, [c[0]]])         # :125-126
        if pilot:
            c = generateE1code(codes[PRN][1])
            E1cCode = np.concatenate([[c[L2 - 1]], c, [c[0]]])     # :128-130
        codeFreq = s.codeFreqBasis; remCodePhase = 0.0             # :133-135
        carrFreq = channel[ch]["acquiredFreq"]; carrFreqBasis = channel[ch]["acquiredFreq"]; remCarrPhase = 0.0
        oldCodeNco = oldCodeError = 0.0
        d2CarrError = dCarrError = 0.0
        vsmCnt = 0
        for loopCnt in range(1, nE + 1):                           # :154
            tr["absoluteSample"][loopCnt - 1] = pos / 2            # :207
            codePhaseStep = codeFreq / s.samplingFreq              # :211
            blksize = int(math.ceil((s.codeLength - remCodePhase) / codePhaseStep))   # :214
            chunk = raw[pos: pos + 2 * blksize]
            pos += chunk.size
            if chunk.size != 2 * blksize:                          # :228-232
                return out
            rawSignal = chunk[0::2].astype(np.float64) + 1j * chunk[1::2].astype(np.float64)
            tr["remCodePhase"][loopCnt - 1] = remCodePhase         # :234
            tE = colonop((remCodePhase - earlyLateSpc) * 2, codePhaseStep * 2,
                         ((blksize - 1) * codePhaseStep + remCodePhase - earlyLateSpc) * 2)      # :236-238
            iE = np.ceil(tE).astype(np.int64)                      # tcode2 = ceil(tcode) + 1, 1-based
            tL = colonop((remCodePhase + earlyLateSpc) * 2, codePhaseStep * 2,
                         ((blksize - 1) * codePhaseStep + remCodePhase + earlyLateSpc) * 2)      # :245-247
            iL = np.ceil(tL).astype(np.int64)
            tP = colonop(remCodePhase * 2, codePhaseStep * 2, ((blksize - 1) * codePhaseStep + remCodePhase) * 2)   # :254-256
            iP = np.ceil(tP).astype(np.int64)
            remCodePhase = tP[blksize - 1] / 2 + codePhaseStep - s.codeLength   # :263
            tr["remCarrPhase"][loopCnt - 1] = remCarrPhase
            time = np.arange(0, blksize + 1, dtype=np.float64) / s.samplingFreq
            trigarg = ((carrFreq * 2.0 * np.pi) * time) + remCarrPhase
            remCarrPhase = math.fmod(trigarg[blksize], 2 * np.pi)
            bb = np.exp(-1j * trigarg[:blksize]) * rawSignal
            iB, qB = bb.real, bb.imag
            I_E = float(np.sum(E1bCode[iE] * iB)); Q_E = float(np.sum(E1bCode[iE] * qB))
            I_P = float(np.sum(E1bCode[iP] * iB)); Q_P = float(np.sum(E1bCode[iP] * qB))
            I_L = float(np.sum(E1bCode[iL] * iB)); Q_L = float(np.sum(E1bCode[iL] * qB))
            with np.errstate(divide="ignore", invalid="ignore"):
                carrError = float(np.arctan(np.float64(Q_P) / np.float64(I_P)) / (2.0 * np.pi))   # :296
                sE = math.sqrt(I_E * I_E + Q_E * Q_E); sL = math.sqrt(I_L * I_L + Q_L * Q_L)
                codeError = float((np.float64(sE) - sL) / (np.float64(sE) + sL))                # :322-323
                if pilot:
                    I_Ec = float(np.sum(E1cCode[iE] * iB)); Q_Ec = float(np.sum(E1cCode[iE] * qB))
                    I_Pc = float(np.sum(E1cCode[iP] * iB)); Q_Pc = float(np.sum(E1cCode[iP] * qB))
                    I_Lc = float(np.sum(E1cCode[iL] * iB)); Q_Lc = float(np.sum(E1cCode[iL] * qB))
                    carrErrorE1c = float(np.arctan(np.float64(Q_Pc) / np.float64(I_Pc)) / (2.0 * np.pi))   # :298
                    carrError = (carrError + carrErrorE1c) / 2     # :299
                    sEc = math.sqrt(I_Ec * I_Ec + Q_Ec * Q_Ec); sLc = math.sqrt(I_Lc * I_Lc + Q_Lc * Q_Lc)
                    codeErrorE1c = float((np.float64(sEc) - sLc) / (np.float64(sEc) + sLc))     # :328-329
                    codeError = (codeError + codeErrorE1c) / 2     # :331
            d2CarrError = d2CarrError + carrError * pf3            # :303
            dCarrError = d2CarrError + carrError * pf2 + dCarrError
            carrNco = dCarrError + carrError * pf1
            tr["carrFreq"][loopCnt - 1] = carrFreq
            carrFreq = carrFreqBasis + carrNco                     # :311
            codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code)   # :335-336
            oldCodeNco = codeNco; oldCodeError = codeError
            tr["codeFreq"][loopCnt - 1] = codeFreq
            codeFreq = s.codeFreqBasis - codeNco                   # :343
            tr["dllDiscr"][loopCnt - 1] = codeError; tr["dllDiscrFilt"][loopCnt - 1] = codeNco
            tr["pllDiscr"][loopCnt - 1] = carrError; tr["pllDiscrFilt"][loopCnt - 1] = carrNco
            tr["I_E"][loopCnt - 1] = I_E; tr["I_P"][loopCnt - 1] = I_P; tr["I_L"][loopCnt - 1] = I_L
            tr["Q_E"][loopCnt - 1] = Q_E; tr["Q_P"][loopCnt - 1] = Q_P; tr["Q_L"][loopCnt - 1] = Q_L
            if loopCnt % s.CNo_VSMinterval == 0:                   # :360-368
                vsmCnt += 1
                lo = loopCnt - s.CNo_VSMinterval
                tr["VSMValue"][vsmCnt - 1] = CNoVSM(tr["I_P"][lo:loopCnt], tr["Q_P"][lo:loopCnt], s.CNo_accTime)
                tr["VSMIndex"][vsmCnt - 1] = loopCnt
        tr["status"] = channel[ch]["status"]
    return out


# ===========================================================================
# The four 10230-chip data + pilot signals: GPS L5C, Galileo E5a, Galileo E5b, BeiDou B2a
# (GPS/GPS_L5C, GAL/GAL_E5a, GAL/GAL_E5b, BDS/B2a; "L5C :n" etc. cite <folder>/include/<file>.m)
# ===========================================================================
# Their acquisition.m / tracking.m are one algorithm with per-signal constants; the primary codes come
# from per-signal generators built on ICD tables (generateL5Icode.m, generateE5aIcode.m, ...), which are
# DATA here: ``codes[PRN] = (data_chips, pilot_chips[, pilot_secondary])`` as +-1 arrays.
def fam5_settings(signal: str, **kw) -> Settings:
    """initSettings.m of the signal (hot-path fields)."""
    base = dict(codeLength=10230.0, codeFreqBasis=10.23e6, acqSearchBand=5000.0, acqSearchStep=500.0, acqThreshold=4.5,
                pllNoiseBandwidth=15.0, carrFreqBasis=1176.45e6)
    per = {
        "GPS_L5C": dict(acqSatelliteList=list(range(1, 33)), acqNonCohTime=25, dllNoiseBandwidth=2.0, CNo_VSMinterval=400,
                        resamplingThreshold=50e6),
        "GAL_E5a": dict(acqSatelliteList=list(range(1, 37)), acqNonCohTime=15, dllNoiseBandwidth=1.5, CNo_VSMinterval=100,
                        resamplingThreshold=45e6),
        "GAL_E5b": dict(acqSatelliteList=list(range(1, 37)), acqNonCohTime=15, acqSearchStep=60.0, dllNoiseBandwidth=1.5,
                        pllNoiseBandwidth=25.0, CNo_VSMinterval=100, resamplingThreshold=45e6, carrFreqBasis=1207.14e6),
        "BDS_B2a": dict(acqSatelliteList=list(range(19, 31)) + list(range(32, 47)) + [59, 60], acqNonCohTime=15,
                        acqThreshold=5.0, dllNoiseBandwidth=2.0, CNo_VSMinterval=200, resamplingThreshold=50e6),
    }[signal]
    s = Settings(**{**base, **per})
    s.signal = signal
    s.pilotTRKflag = 1 if signal in ("GAL_E5a", "GAL_E5b") else 0      # L5C initSettings.m:113, B2a :112: 0
    for k, v in kw.items():
        setattr(s, k, v)
    return s


_FAM5_RESLEN = {"GPS_L5C": 32, "GAL_E5a": 50, "GAL_E5b": 50, "BDS_B2a": None}    # B2a: max(acqSatelliteList) (B2a :128-132)
_FAM5_MINPER = {"GPS_L5C": 42, "GAL_E5a": 102, "GAL_E5b": 102, "BDS_B2a": 12}    # postProcessing.m codeLen = max(., nonCoh+2)


def read_acq_signal_fam5(raw: np.ndarray, s: Settings) -> np.ndarray:
    """postProcessing.m: max(42 | 102 | 12, acqNonCohTime+2) code periods (L5C :88, E5a :88, E5b :89, B2a :86)."""
    N = samples_per_code(s)
    codeLen = max(_FAM5_MINPER[s.signal], s.acqNonCohTime + 2)
    off = 2 * s.skipNumberOfBytes
    data = raw[off: off + 2 * codeLen * N].astype(np.float64)
    return data[0::2] + 1j * data[1::2]


def make_code_table(code: np.ndarray, s: Settings) -> np.ndarray:
    """makeL5ITable.m / makeL5QTable.m / makeE5aITable.m ...: index = ceil(ts*(1:N)/tc), last forced to codeLength."""
    N = samples_per_code(s)
    ts = 1 / s.samplingFreq
    tc = 1 / s.codeFreqBasis
    idx = np.ceil((ts * np.arange(1, N + 1, dtype=np.float64)) / tc).astype(np.int64)
    idx[-1] = int(s.codeLength)
    return np.asarray(code, dtype=np.float64)[idx - 1]


def acquisition_fam5(longSignal: np.ndarray, s: Settings, codes: dict, workers: int = 1):
    """L5C acquisition.m:118-300, E5a :118-290, E5b :118-230, B2a :116-300 (resampling branch not restated)."""
    sig = s.signal
    N = samples_per_code(s)
    ts = 1 / s.samplingFreq
    phasePoints = np.arange(0, 2 * N, dtype=np.float64) * 2 * np.pi * ts
    nBins = int(matlab_round(s.acqSearchBand * 2 / s.acqSearchStep)) + 1
    coarseFreqBin = np.zeros(nBins)
    nRes = _FAM5_RESLEN[sig] or max(s.acqSatelliteList)
    res = dict(carrFreq=np.zeros(nRes), codePhase=np.zeros(nRes), peakMetric=np.zeros(nRes),
               coarseBin=np.zeros(nRes, dtype=np.int64), coarseCodePhase=np.zeros(nRes, dtype=np.int64))
    NHcode = np.array([1, 1, 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1], dtype=np.float64)   # L5C :134
    fineSearchStep = 5 if sig in ("GAL_E5a", "GAL_E5b") else 25                 # E5a :136; L5C :136; B2a :135
    numOfFineBins = int(matlab_round(s.acqSearchStep / fineSearchStep)) + 1
    nPer = {"GPS_L5C": 20, "GAL_E5a": 100, "GAL_E5b": 100, "BDS_B2a": max(10, s.acqNonCohTime)}[sig]   # L5C :146; E5a :142; B2a :140
    finePhasePoints = np.arange(0, nPer * N, dtype=np.float64) * 2 * np.pi * ts
    x = longSignal[:N]
    sigPower = math.sqrt(np.sum(np.abs(x - np.mean(x)) ** 2) / (N - 1) * N)
    res["sigPower"] = sigPower
    for PRN in s.acqSatelliteList:
        dcode = np.asarray(codes[PRN][0], dtype=np.float64)
        pcode = np.asarray(codes[PRN][1], dtype=np.float64)
        IFreqDom = np.conj(_FFT(np.concatenate([make_code_table(dcode, s), np.zeros(N)])))     # L5C :158-167
        QFreqDom = np.conj(_FFT(np.concatenate([make_code_table(pcode, s), np.zeros(N)])))
        results = np.zeros((nBins, 2 * N))
        for k in range(1, nBins + 1):
            coarseFreqBin[k - 1] = s.IF + s.acqSearchBand - s.acqSearchStep * (k - 1)
            sigCarr = np.exp(-1j * coarseFreqBin[k - 1] * phasePoints)
            for m in range(1, s.acqNonCohTime + 1):
                IQfreqDom = _FFT(sigCarr * longSignal[(m - 1) * N: (m + 1) * N], workers)
                coh = np.abs(_IFFT(IQfreqDom * IFreqDom, workers)) + np.abs(_IFFT(IQfreqDom * QFreqDom, workers))   # L5C :186-190
                results[k - 1, :] += coh
        acqCoarseBin = int(np.argmax(results.max(axis=1))) + 1
        colmax = results.max(axis=0)
        codePhase = int(np.argmax(colmax)) + 1
        res["peakMetric"][PRN - 1] = colmax[codePhase - 1] / sigPower / s.acqNonCohTime
        res["coarseBin"][PRN - 1] = acqCoarseBin
        res["coarseCodePhase"][PRN - 1] = codePhase
        if res["peakMetric"][PRN - 1] > s.acqThreshold:
            if sig == "GAL_E5b":                                                # E5b :203-205: no fine search
                res["carrFreq"][PRN - 1] = coarseFreqBin[acqCoarseBin - 1]
                res["codePhase"][PRN - 1] = codePhase
                continue
            codeValueIndex = np.floor((ts * np.arange(1, nPer * N + 1, dtype=np.float64)) / (1 / s.codeFreqBasis)).astype(np.int64)   # L5C :196 (1-based sample index)
            longP = pcode[np.fmod(codeValueIndex, int(s.codeLength))]          # L5C :198
            longD = dcode[np.fmod(codeValueIndex, int(s.codeLength))]          # B2a :210
            sigF = longSignal[codePhase - 1: codePhase - 1 + nPer * N]         # L5C :200
            fineFreqBins = np.zeros(numOfFineBins)
            fineResult = np.zeros(numOfFineBins)
            sec = NHcode if sig == "GPS_L5C" else np.asarray(codes[PRN][2], dtype=np.float64) if sig == "GAL_E5a" else None
            for j in range(1, numOfFineBins + 1):
                fineFreqBins[j - 1] = coarseFreqBin[acqCoarseBin - 1] + s.acqSearchStep / 2 - fineSearchStep * (j - 1)
                carr = np.exp(-1j * fineFreqBins[j - 1] * finePhasePoints)
                sumPerCode = (longP * carr * sigF).reshape(nPer, N).sum(axis=1)                 # L5C :208-212
                if sig == "BDS_B2a":                                                            # B2a :216-228
                    sum1 = (longD * carr * sigF).reshape(nPer, N).sum(axis=1)
                    fineResult[j - 1] = np.sum(np.abs(sum1)) + np.sum(np.abs(sumPerCode))
                    continue
                maxPower = 0.0
                for c in range(nPer):                                                           # L5C :214-219
                    maxPower = max(maxPower, abs(np.sum(sumPerCode * np.roll(sec, c))))
                fineResult[j - 1] = maxPower
            maxFinBin = int(np.argmax(fineResult)) + 1
            res["carrFreq"][PRN - 1] = fineFreqBins[maxFinBin - 1]
            res["codePhase"][PRN - 1] = codePhase
            if res["carrFreq"][PRN - 1] == 0:
                res["carrFreq"][PRN - 1] = 1
    return res


def preRun_fam5(acq: dict, s: Settings):
    """preRun.m:44-78 of the four folders: strongest peaks first; channel.codeFreq is the carrier-aided code
    NCO centre codeFreqBasis + (acquiredFreq - IF)/carrFreqBasis*codeFreqBasis (L5C :69-71)."""
    chans = [dict(PRN=0, acquiredFreq=0.0, codePhase=0, codeFreq=0.0, status="-") for _ in range(s.numberOfChannels)]
    order = np.argsort(-acq["peakMetric"], kind="stable")
    n = min(s.numberOfChannels, int(np.sum(acq["carrFreq"] != 0)))
    for ii in range(n):
        p = int(order[ii])
        af = float(acq["carrFreq"][p])
        chans[ii] = dict(PRN=p + 1, acquiredFreq=af, codePhase=int(acq["codePhase"][p]),
                         codeFreq=s.codeFreqBasis + (af - s.IF) / s.carrFreqBasis * s.codeFreqBasis, status="T")
    return chans


def tracking_fam5(raw: np.ndarray, channel: list, s: Settings, codes: dict):
    """L5C tracking.m:45-424 (E5a/E5b/B2a: the same loop): 1 ms epochs, code NCO centred on channel.codeFreq,
    three-coefficient carrier filter, and with pilotTRKflag == 1 the pilot replica on the same code phase: its
    prompt is rotated by -pi/2 before the atan (:277-281), both discriminators are averaged, and Pilot_I_P /
    Pilot_Q_P are recorded (:323-324)."""
    nE = s.msToProcess
    pilot = int(getattr(s, "pilotTRKflag", 0)) == 1
    out = []
    for _ in range(s.numberOfChannels):
        tr = dict(status="-", PRN=0)
        tr["absoluteSample"] = np.zeros(nE)
        for f in ("codeFreq", "carrFreq", "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase"):
            tr[f] = np.full(nE, np.inf)
        for f in ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L", "Pilot_I_P", "Pilot_Q_P"):
            tr[f] = np.zeros(nE)
        tr["VSMValue"] = np.zeros(nE // s.CNo_VSMinterval)
        tr["VSMIndex"] = np.zeros(nE // s.CNo_VSMinterval)
        out.append(tr)
    earlyLateSpc = s.dllCorrelatorSpacing
    PDIcode = s.intTime
    tau1code, tau2code = calcLoopCoef(s.dllNoiseBandwidth, s.dllDampingRatio, 1.0)
    pf3, pf2, pf1 = calcLoopCoefCarr(s)
    Lc = int(s.codeLength)
    rot = np.exp(-1j * np.pi / 2)                                  # :278
    for ch in range(s.numberOfChannels):
        if channel[ch]["PRN"] == 0:
            continue
        tr = out[ch]
        PRN = channel[ch]["PRN"]
        tr["PRN"] = PRN
        pos = 2 * (s.skipNumberOfBytes + channel[ch]["codePhase"] - 1)
        c = np.asarray(codes[PRN][0], dtype=np.float64)
        ICode = np.concatenate([[c[Lc - 1]], c, [c[0]]])           # :164-165
        if pilot:
            c = np.asarray(codes[PRN][1], dtype=np.float64)
            QCode = np.concatenate([[c[Lc - 1]], c, [c[0]]])       # :167-169
        codeFreq = channel[ch]["codeFreq"]; remCodePhase = 0.0     # :173-175
        carrFreq = channel[ch]["acquiredFreq"]; carrFreqBasis = channel[ch]["acquiredFreq"]; remCarrPhase = 0.0
        oldCodeNco = oldCodeError = 0.0
        d2CarrError = dCarrError = 0.0
        vsmCnt = 0
        for loopCnt in range(1, nE + 1):
            tr["absoluteSample"][loopCnt - 1] = pos / 2
            codePhaseStep = codeFreq / s.samplingFreq
            blksize = int(math.ceil((s.codeLength - remCodePhase) / codePhaseStep))
            chunk = raw[pos: pos + 2 * blksize]
            pos += chunk.size
            if chunk.size != 2 * blksize:
                return out
            rawSignal = chunk[0::2].astype(np.float64) + 1j * chunk[1::2].astype(np.float64)
            tr["remCodePhase"][loopCnt - 1] = remCodePhase
            tE = colonop(remCodePhase - earlyLateSpc, codePhaseStep, (blksize - 1) * codePhaseStep + remCodePhase - earlyLateSpc)
            tL = colonop(remCodePhase + earlyLateSpc, codePhaseStep, (blksize - 1) * codePhaseStep + remCodePhase + earlyLateSpc)
            tP = colonop(remCodePhase, codePhaseStep, (blksize - 1) * codePhaseStep + remCodePhase)
            iE = np.ceil(tE).astype(np.int64); iL = np.ceil(tL).astype(np.int64); iP = np.ceil(tP).astype(np.int64)
            remCodePhase = (tP[blksize - 1] + codePhaseStep) - s.codeLength
            tr["remCarrPhase"][loopCnt - 1] = remCarrPhase
            time = np.arange(0, blksize + 1, dtype=np.float64) / s.samplingFreq
            trigarg = ((carrFreq * 2.0 * np.pi) * time) + remCarrPhase
            remCarrPhase = math.fmod(trigarg[blksize], 2 * np.pi)
            bb = np.exp(-1j * trigarg[:blksize]) * rawSignal
            iB, qB = bb.real, bb.imag
            I_E = float(np.sum(ICode[iE] * iB)); Q_E = float(np.sum(ICode[iE] * qB))
            I_P = float(np.sum(ICode[iP] * iB)); Q_P = float(np.sum(ICode[iP] * qB))
            I_L = float(np.sum(ICode[iL] * iB)); Q_L = float(np.sum(ICode[iL] * qB))
            with np.errstate(divide="ignore", invalid="ignore"):
                carrError = float(np.arctan(np.float64(Q_P) / np.float64(I_P)) / (2.0 * np.pi))
                sE = math.sqrt(I_E * I_E + Q_E * Q_E); sL = math.sqrt(I_L * I_L + Q_L * Q_L)
                codeError = float((np.float64(sE) - sL) / (np.float64(sE) + sL))
                if pilot:
                    I_EQ = float(np.sum(QCode[iE] * iB)); Q_EQ = float(np.sum(QCode[iE] * qB))
                    I_PQ = float(np.sum(QCode[iP] * iB)); Q_PQ = float(np.sum(QCode[iP] * qB))
                    I_LQ = float(np.sum(QCode[iL] * iB)); Q_LQ = float(np.sum(QCode[iL] * qB))
                    QI = (I_PQ + 1j * Q_PQ) * rot                                               # :278
                    carrErrorQ = float(np.arctan(np.float64(QI.imag) / np.float64(QI.real)) / (2.0 * np.pi))   # :279
                    carrError = (carrError + carrErrorQ) / 2                                    # :280
                    sEq = math.sqrt(I_EQ ** 2 + Q_EQ ** 2); sLq = math.sqrt(I_LQ ** 2 + Q_LQ ** 2)
                    codeErrorQ = float((np.float64(sEq) - sLq) / (np.float64(sEq) + sLq))     # :298-299
                    codeError = (codeError + codeErrorQ) / 2
            d2CarrError = d2CarrError + carrError * pf3
            dCarrError = d2CarrError + carrError * pf2 + dCarrError
            carrNco = dCarrError + carrError * pf1
            tr["carrFreq"][loopCnt - 1] = carrFreq
            carrFreq = carrFreqBasis + carrNco
            codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code)
            oldCodeNco = codeNco; oldCodeError = codeError
            tr["codeFreq"][loopCnt - 1] = codeFreq
            codeFreq = channel[ch]["codeFreq"] - codeNco           # :309
            tr["dllDiscr"][loopCnt - 1] = codeError; tr["dllDiscrFilt"][loopCnt - 1] = codeNco
            tr["pllDiscr"][loopCnt - 1] = carrError; tr["pllDiscrFilt"][loopCnt - 1] = carrNco
            tr["I_E"][loopCnt - 1] = I_E; tr["I_P"][loopCnt - 1] = I_P; tr["I_L"][loopCnt - 1] = I_L
            tr["Q_E"][loopCnt - 1] = Q_E; tr["Q_P"][loopCnt - 1] = Q_P; tr["Q_L"][loopCnt - 1] = Q_L
            if pilot:
                tr["Pilot_I_P"][loopCnt - 1] = I_PQ; tr["Pilot_Q_P"][loopCnt - 1] = Q_PQ   # :323-324
            if loopCnt % s.CNo_VSMinterval == 0:                   # :328-335 (B2a computes DataCNo/PLD from the same rows instead)
                vsmCnt += 1
                lo = loopCnt - s.CNo_VSMinterval
                tr["VSMValue"][vsmCnt - 1] = CNoVSM(tr["I_P"][lo:loopCnt], tr["Q_P"][lo:loopCnt], s.CNo_accTime)
                tr["VSMIndex"][vsmCnt - 1] = loopCnt
        tr["status"] = channel[ch]["status"]
    return out


# ===========================================================================
# Acquisition variant B: BeiDou B1I (BDS/B1I) and GPS L2C (GPS/GPS_L2C)
# ===========================================================================
def varb_settings(signal: str, **kw) -> Settings:
    """initSettings.m of BDS/B1I and GPS/GPS_L2C (hot-path fields); acqSearchBand is in kHz there."""
    if signal == "BDS_B1I":
        s = Settings(codeFreqBasis=2.046e6, codeLength=2046.0, acqSatelliteList=list(range(6, 59)), acqSearchBand=10.0,
                     acqThreshold=2.0, dllNoiseBandwidth=4.0, pllNoiseBandwidth=35.0, CNo_VSMinterval=400)
        s.stepSize = 125.0
    else:
        s = Settings(samplingFreq=8e6, codeFreqBasis=0.5115e6, codeLength=10230.0, acqSearchBand=10.0, acqThreshold=1.5,
                     dllNoiseBandwidth=4.0, dllCorrelatorSpacing=0.25, pllNoiseBandwidth=10.0, intTime=0.02,
                     CNo_accTime=0.02, CNo_VSMinterval=40)
        s.acqStep = (1000 / 2) / 20 / 2
    s.signal = signal
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def read_acq_signal_varb(raw: np.ndarray, s: Settings) -> np.ndarray:
    """B1I postProcessing.m:98: 11 code periods; L2C postProcessing.m:88-90: max(42, acqCohT+2) code periods."""
    N = samples_per_code(s)
    codeLen = 11 if s.signal == "BDS_B1I" else max(42, int(getattr(s, "acqCohT", 20)) + 2)
    off = 2 * s.skipNumberOfBytes
    data = raw[off: off + 2 * codeLen * N].astype(np.float64)
    return data[0::2] + 1j * data[1::2]


def _varb_second_peak(corrVec: np.ndarray, codePhase: int, chip: int, N1: int) -> float:
    """B1I acquisition.m:127-141 / L2C :89-103 with MATLAB's 1-based colon ranges."""
    e1, e2 = codePhase - chip, codePhase + chip
    if e1 < 2:
        rng = np.arange(e2, N1 + e1 + 1)
    elif e2 >= N1:
        rng = np.arange(e2 - N1 + 1, e1 + 1)
    else:
        rng = np.concatenate([np.arange(1, e1 + 1), np.arange(e2, N1 + 1)])
    return float(np.max(corrVec[rng - 1]))


def acquisition_b1i(longSignal: np.ndarray, s: Settings, codes: dict, workers: int = 1):
    """BDS/B1I/include/acquisition.m:4-176 (resampling branch not restated).  codes[PRN][0] = the 2046 chips
    generateCAcode53.m returns."""
    Ncodes, Nblocks = 2, 4                                         # :6-7
    spb = int(matlab_round(s.samplingFreq / (s.codeFreqBasis / (Nblocks * s.codeLength))))   # :9
    signal1, signal2 = longSignal[:spb], longSignal[spb: 2 * spb]  # :13-14
    ts = 1 / s.samplingFreq
    phasePoints = np.arange(0, spb, dtype=np.float64) * 2 * np.pi * ts
    freqResolution = s.samplingFreq / spb                           # :20
    nBins = int(matlab_round(s.acqSearchBand * 1e3 / freqResolution)) + 1   # :22
    stepSize = getattr(s, "stepSize", None)
    if not stepSize:                                                # :24-39
        stepSize = 0.5 / (Nblocks * s.codeLength / s.codeFreqBasis)
    elif stepSize != freqResolution:
        steps = np.arange(1, freqResolution / 2 + 1e-12, 0.25)
        steps = steps[np.fmod(freqResolution, steps) == 0]
        diff = steps - stepSize
        k = int(np.argmin(np.abs(diff)))
        stepSize = steps[k - 1] if diff[k] > 0 else steps[k]
    Nshifts = int(matlab_round(freqResolution / stepSize))          # :40
    spc2 = int(matlab_round(s.samplingFreq / (s.codeFreqBasis / (Ncodes * s.codeLength))))   # makeCaTableDMA.m:9
    tc = 1 / s.codeFreqBasis
    idx = np.ceil((ts * np.arange(1, spc2 + 1, dtype=np.float64)) / tc).astype(np.int64)
    idx[-1] = Ncodes * 2046
    res = dict(carrFreq=np.zeros(58), codePhase=np.zeros(58), peakMetric=np.zeros(58),
               coarseBin=np.zeros(58, dtype=np.int64), coarseCodePhase=np.zeros(58, dtype=np.int64))
    initFreq = s.IF + (s.acqSearchBand / 2) * 1000                  # :50
    chip = int(matlab_round(s.samplingFreq / s.codeFreqBasis))      # :126
    for PRN in s.acqSatelliteList:
        c = np.asarray(codes[PRN][0], dtype=np.float64)
        table = np.concatenate([c, c])[idx - 1]                     # makeCaTableDMA.m:15-18
        codeFreqDom = np.conj(_FFT(np.concatenate([table, np.zeros(spb // Ncodes)])))   # :58
        prevmax = 0.0
        corrVec = np.zeros(spb)
        frequencyBinIndex = freqShift = 0
        for binIter in range(1, Nshifts + 1):
            f0 = initFreq + (binIter - 1) * (freqResolution / Nshifts)   # :64
            sigCarr = np.exp(-1j * f0 * phasePoints)
            F1 = _FFT(sigCarr * signal1, workers)
            F2 = _FFT(sigCarr * signal2, workers)
            sh = np.arange(nBins if binIter == 1 else nBins - 1)    # :79-81 the last bin is skipped for binIter > 1
            A1 = np.abs(_IFFT(np.stack([np.roll(F1, k) for k in sh]) * codeFreqDom[None, :], workers))   # :83-90
            A2 = np.abs(_IFFT(np.stack([np.roll(F2, k) for k in sh]) * codeFreqDom[None, :], workers))
            p1, p2 = A1.max(axis=1), A2.max(axis=1)
            for k in sh:                                            # :101-116
                if p1[k] > prevmax or p2[k] > prevmax:
                    if p1[k] > p2[k]:
                        prevmax = p1[k]; corrVec = A1[k]
                    else:
                        prevmax = p2[k]; corrVec = A2[k]
                    freqShift = binIter
                    frequencyBinIndex = int(k) + 1
        codePhase = int(np.argmax(corrVec)) + 1                     # :125
        maxPeak = corrVec[codePhase - 1]
        second = _varb_second_peak(corrVec, codePhase, chip, spb // Nblocks)
        res["peakMetric"][PRN - 1] = maxPeak / second               # :142
        res["coarseBin"][PRN - 1] = frequencyBinIndex
        res["coarseCodePhase"][PRN - 1] = codePhase
        if maxPeak / second > s.acqThreshold:                       # :145-150
            res["codePhase"][PRN - 1] = codePhase
            res["carrFreq"][PRN - 1] = initFreq - freqResolution * (frequencyBinIndex - 1) + (freqResolution / Nshifts) * (freqShift - 1)
    return res


def acquisition_l2c(longSignal: np.ndarray, s: Settings, codes: dict, workers: int = 1):
    """GPS/GPS_L2C/include/acquisition.m:4-145.  codes[PRN][0] = the 20460-entry return-to-zero CM sequence
    generateCMcode.m returns; with settings.pilotTRKflag == 1 the CL code phase search of :100-137 runs on the acquired
    PRNs (codes[PRN][1] = the 1534500-entry return-to-zero CL sequence of generateCLcode.m) and CLCodePhase is returned."""
    Nblocks = 2
    N = samples_per_code(s)
    chip = int(matlab_round(s.samplingFreq / s.codeFreqBasis))      # :7
    spb = N * Nblocks
    signal = longSignal[:spb]
    ts = 1 / s.samplingFreq
    phasePoints = np.arange(0, spb, dtype=np.float64) * 2 * np.pi * ts
    freqResolution = s.samplingFreq / spb
    nBins = int(matlab_round(s.acqSearchBand * 1e3 / freqResolution)) + 1
    Nshifts = int(matlab_round(freqResolution / s.acqStep))         # :24
    tc = 1 / (s.codeFreqBasis * 2)                                   # makeCMTable.m:8
    idx = np.ceil((ts * np.arange(0, N, dtype=np.float64)) / tc).astype(np.int64)
    idx[-1] = int(s.codeLength) * 2
    idx[0] = 1
    res = dict(carrFreq=np.zeros(32), codePhase=np.zeros(32), peakMetric=np.zeros(32),
               coarseBin=np.zeros(32, dtype=np.int64), coarseCodePhase=np.zeros(32, dtype=np.int64),
               CLCodePhase=np.zeros(32, dtype=np.int64))
    initFreq = s.IF + (s.acqSearchBand / 2) * 1000
    for PRN in s.acqSatelliteList:
        cm = np.asarray(codes[PRN][0], dtype=np.float64)
        cmFreqDom = np.conj(_FFT(np.concatenate([cm[idx - 1], np.zeros(N)])))   # :44-48
        prevmax = 0.0
        corrVec = np.zeros(spb)
        frequencyBinIndex = freqShift = 0
        for binIter in range(1, Nshifts + 1):
            f0 = initFreq - (binIter - 1) * (freqResolution / Nshifts)   # :62
            F = _FFT(np.exp(-1j * f0 * phasePoints) * signal, workers)
            last = nBins if binIter == 1 else nBins - 1
            for k0 in range(0, last, 64):                            # batches of rows to bound memory
                sh = np.arange(k0, min(last, k0 + 64))
                A = np.abs(_IFFT(np.stack([np.roll(F, k) for k in sh]) * cmFreqDom[None, :], workers))
                pk = A.max(axis=1)
                for i, k in enumerate(sh):                           # :78-83
                    if pk[i] > prevmax:
                        prevmax = pk[i]; corrVec = A[i].copy(); frequencyBinIndex = int(k) + 1; freqShift = binIter
        codePhase = int(np.argmax(corrVec)) + 1
        maxPeak = corrVec[codePhase - 1]
        second = _varb_second_peak(corrVec, codePhase, chip, spb // Nblocks)
        res["peakMetric"][PRN - 1] = maxPeak / second
        res["coarseBin"][PRN - 1] = frequencyBinIndex
        res["coarseCodePhase"][PRN - 1] = codePhase
        if maxPeak / second > s.acqThreshold:
            res["carrFreq"][PRN - 1] = initFreq - freqResolution * (frequencyBinIndex - 1) - (freqResolution / Nshifts) * (freqShift - 1)   # :95
            res["codePhase"][PRN - 1] = codePhase
            if int(getattr(s, "pilotTRKflag", 0)) == 1:              # :100-137
                signal0DC = longSignal[codePhase - 1: codePhase - 1 + N]
                signal0DC = signal0DC - np.mean(signal0DC)
                phasePointsCL = np.arange(0, N, dtype=np.float64) * 2 * np.pi * ts
                sigCarr = np.exp(-1j * res["carrFreq"][PRN - 1] * phasePointsCL)
                CLCode = np.asarray(codes[PRN][1], dtype=np.float64)
                cvi = np.ceil((ts * np.arange(0, N, dtype=np.float64)) / tc).astype(np.int64)   # :124
                cvi[0] = 1
                cvi[-1] = int(s.codeLength) if getattr(s, "acqCohT", 20) <= 10 else int(s.codeLength) * 2   # :128-133
                powerArray = np.zeros(75)
                for ind in range(1, 76):
                    CLCodeSample = CLCode[cvi + int(s.codeLength) * 2 * (ind - 1) - 1]
                    powerArray[ind - 1] = abs(np.sum(signal0DC * CLCodeSample * sigCarr))
                res["CLCodePhase"][PRN - 1] = int(np.argmax(powerArray)) + 1
    return res


def tracking_b1i(raw: np.ndarray, channel: list, s: Settings, codes: dict):
    """BDS/B1I/include/tracking.m: the B3I loop (three-coefficient carrier filter, 1 ms epochs) with the 2046-chip code
    generateCAcode53(PRN) (:50) and the code NCO centred on settings.codeFreqBasis (:52, :139)."""
    ch = [dict(c, codeFreq=s.codeFreqBasis) for c in channel]
    s2 = Settings(**{k: getattr(s, k) for k in Settings.__dataclass_fields__})
    s2.pilotTRKflag = 0
    return tracking_fam5(raw, ch, s2, {p: (c[0], c[0]) for p, c in codes.items()})


# ===========================================================================
# Acquisition variant C: BeiDou B1C (BDS/B1C)
# ===========================================================================
def b1c_settings(**kw) -> Settings:
    """BDS/B1C/initSettings.m (acquisition fields)."""
    s = Settings(numberOfChannels=15, codeLength=10230.0, codeFreqBasis=1.023e6, acqSatelliteList=list(range(1, 63)),
                 acqSearchBand=5000.0, acqThreshold=10.0, intTime=0.01)
    s.acqCohT, s.acqStep, s.pilotACQflag, s.signal = 10, 1000 / 10 / 2, 1, "BDS_B1C"
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def acquisition_b1c(longSignal: np.ndarray, s: Settings, codes: dict, workers: int = 1):
    """BDS/B1C/include/acquisition.m:128-276 (resampling branch :50-126 not restated).  codes[PRN] = (data, pilot) BOC(1,1)
    sub-chip sequences as generateDataBOC11.m / generatePilotBOC11.m return them (20460 entries)."""
    N = samples_per_code(s)
    xLen = int(matlab_round(N / 10 * s.acqCohT))                   # :131
    Lc = int(matlab_round(N / 10 * (10 + s.acqCohT)))              # :133
    sig = longSignal[:Lc]                                          # :136
    ts = 1 / s.samplingFreq
    phasePoints = np.arange(0, Lc, dtype=np.float64) * 2 * np.pi * ts
    nBins = int(matlab_round(s.acqSearchBand * 2 / s.acqStep)) + 1   # :142
    nRes = max(s.acqSatelliteList)
    res = dict(carrFreq=np.zeros(nRes), codePhase=np.zeros(nRes), peakMetric=np.zeros(nRes),
               coarseBin=np.zeros(nRes, dtype=np.int64), coarseCodePhase=np.zeros(nRes, dtype=np.int64))
    fineStep = 25
    nFine = int(matlab_round(s.acqStep / 25)) * 2 + 1              # :155
    finePhasePoints = np.arange(0, N, dtype=np.float64) * 2 * np.pi * ts
    x = sig[:xLen]
    sigPower = math.sqrt(np.sum(np.abs(x - np.mean(x)) ** 2) / (xLen - 1) * xLen)   # :163
    initFreq = s.IF + s.acqSearchBand                              # :166
    IQfreqDom = _FFT(np.exp(-1j * initFreq * phasePoints) * sig, workers)   # :168-172
    tc = 1 / s.codeFreqBasis / 2                                   # makeDataTable.m:9
    idx = np.ceil((ts * np.arange(1, N + 1, dtype=np.float64)) / tc).astype(np.int64)
    idx[-1] = int(s.codeLength) * 2
    idx[0] = 1
    pilot = int(getattr(s, "pilotACQflag", 1)) == 1
    for PRN in s.acqSatelliteList:
        DataTab = np.asarray(codes[PRN][0], dtype=np.float64)[idx - 1]
        DataF = np.conj(_FFT(np.concatenate([DataTab[:xLen], np.zeros(Lc - xLen)])))   # :176-179
        if pilot:
            PilotTab = np.asarray(codes[PRN][1], dtype=np.float64)[idx - 1]
            PilotF = np.conj(_FFT(np.concatenate([PilotTab[:xLen], np.zeros(Lc - xLen)])))
        results = np.zeros((nBins, Lc))
        for k0 in range(0, nBins, 32):
            sh = np.arange(k0, min(nBins, k0 + 32))
            S = np.stack([np.roll(IQfreqDom, k) for k in sh])      # :203
            r = np.abs(_IFFT(S * DataF[None, :], workers))         # :205-207
            if pilot:
                r = (r * math.sqrt(11) + np.abs(_IFFT(S * PilotF[None, :], workers)) * math.sqrt(29)) / math.sqrt(40)   # :211-214
            results[sh] = r
        frequencyBinIndex = int(np.argmax(results.max(axis=1))) + 1   # :221
        selFreq = initFreq - (frequencyBinIndex - 1) * s.acqStep
        colmax = results.max(axis=0)
        codePhase = int(np.argmax(colmax)) + 1                     # :225
        res["peakMetric"][PRN - 1] = colmax[codePhase - 1] / sigPower
        if codePhase + N - 1 > longSignal.size:                    # :229-231
            codePhase -= N
        res["coarseBin"][PRN - 1] = frequencyBinIndex
        res["coarseCodePhase"][PRN - 1] = codePhase
        if res["peakMetric"][PRN - 1] > s.acqThreshold:            # :234
            signal0DC = longSignal[codePhase - 1: codePhase - 1 + N]
            xCarrier = signal0DC * DataTab
            fineFrq = np.zeros(nFine); fineRes = np.zeros(nFine)
            for j in range(1, nFine + 1):
                fineFrq[j - 1] = selFreq + s.acqStep - fineStep * (j - 1)   # :244
                carr = np.exp(-1j * fineFrq[j - 1] * finePhasePoints)
                fineRes[j - 1] = abs(np.sum(xCarrier * carr))
                if pilot:
                    fineRes[j - 1] = (fineRes[j - 1] * 11 + abs(np.sum(signal0DC * PilotTab * carr)) * 29) / 40   # :247-249
            res["carrFreq"][PRN - 1] = fineFrq[int(np.argmax(fineRes))]
            if res["carrFreq"][PRN - 1] == 0:
                res["carrFreq"][PRN - 1] = 1
            res["codePhase"][PRN - 1] = codePhase
    return res


def tracking_l2c(raw: np.ndarray, channel: list, s: Settings, codes: dict):
    """GPS/GPS_L2C/include/tracking.m:45-405: 20 ms epochs in half-chip units - the return-to-zero CM table of 2*codeLength entries, code NCO at
    2*codeFreqBasis, spacing*2 (:93-94, :171) - fseek to codePhase (not codePhase-1, :153), fractional absoluteSample
    (:223), and remCodePhase, codeFreq, dllDiscr, dllDiscrFilt recorded halved (:250, :376, :382-383).  With
    settings.pilotTRKflag == 1 the CL pilot is correlated too: CLCode(tcode2 + codeLength*(CLCodePhase-1)) from the padded
    CL sequence codes[PRN][1], CLCodePhase = channel.CLCodePhase stepping 1..75 every epoch (:259-286, :363-366), both
    discriminator pairs averaged (:335-339, :359-361) and six Pilot_* rows recorded (:396-402)."""
    pilotOn = int(getattr(s, "pilotTRKflag", 0)) == 1
    nE = int(matlab_round(s.msToProcess / 1000 / s.intTime))       # :51
    nV = int(math.floor(s.msToProcess / s.CNo_VSMinterval / 20))   # :80-83
    out = []
    for _ in range(s.numberOfChannels):
        tr = dict(status="-", PRN=0)
        tr["absoluteSample"] = np.zeros(nE)
        for f in ("codeFreq", "carrFreq", "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase"):
            tr[f] = np.full(nE, np.inf)
        for f in ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L"):
            tr[f] = np.zeros(nE)
        if pilotOn:                                                # :72-83
            for f in ("Pilot_I_P", "Pilot_I_E", "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_P", "Pilot_Q_L"):
                tr[f] = np.zeros(nE)
        tr["VSMValue"] = np.zeros(nV); tr["VSMIndex"] = np.zeros(nV)
        out.append(tr)
    earlyLateSpc = s.dllCorrelatorSpacing * 2                      # :93
    codeLength = int(s.codeLength) * 2                             # :94
    PDIcode = s.intTime
    tau1code, tau2code = calcLoopCoef(s.dllNoiseBandwidth, s.dllDampingRatio, 1.0)
    pf3, pf2, pf1 = calcLoopCoefCarr(s)
    for ch in range(s.numberOfChannels):
        if channel[ch]["PRN"] == 0:
            continue
        tr = out[ch]
        tr["PRN"] = channel[ch]["PRN"]
        pos = 2 * (s.skipNumberOfBytes + channel[ch]["codePhase"])   # :153
        c = np.asarray(codes[channel[ch]["PRN"]][0], dtype=np.float64)
        cmCode = np.concatenate([[c[codeLength - 1]], c, [c[0]]])  # :155-156
        if pilotOn:                                                # :160-167
            CLCodePhase = int(channel[ch]["CLCodePhase"])
            c = np.asarray(codes[channel[ch]["PRN"]][1], dtype=np.float64)
            CLCode = np.concatenate([[c[-1]], c, [c[0]]])
        codeFreq = s.codeFreqBasis * 2; remCodePhase = 0.0         # :171-173
        carrFreq = channel[ch]["acquiredFreq"]; carrFreqBasis = channel[ch]["acquiredFreq"]; remCarrPhase = 0.0
        oldCodeNco = oldCodeError = 0.0
        d2CarrError = dCarrError = 0.0
        vsmCnt = 0
        for loopCnt in range(1, nE + 1):
            codePhaseStep = codeFreq / s.samplingFreq              # :220
            tr["absoluteSample"][loopCnt - 1] = (pos / 2) / 1 + 1 - remCodePhase / codePhaseStep   # :223
            blksize = int(math.ceil((codeLength - remCodePhase) / codePhaseStep))   # :226
            chunk = raw[pos: pos + 2 * blksize]
            pos += chunk.size
            if chunk.size != 2 * blksize:
                return out
            rawSignal = chunk[0::2].astype(np.float64) + 1j * chunk[1::2].astype(np.float64)
            tr["remCodePhase"][loopCnt - 1] = remCodePhase / 2     # :250
            tE = colonop(remCodePhase - earlyLateSpc, codePhaseStep, (blksize - 1) * codePhaseStep + remCodePhase - earlyLateSpc)
            tL = colonop(remCodePhase + earlyLateSpc, codePhaseStep, (blksize - 1) * codePhaseStep + remCodePhase + earlyLateSpc)
            tP = colonop(remCodePhase, codePhaseStep, (blksize - 1) * codePhaseStep + remCodePhase)
            iE = np.ceil(tE).astype(np.int64); iL = np.ceil(tL).astype(np.int64); iP = np.ceil(tP).astype(np.int64)
            remCodePhase = (tP[blksize - 1] + codePhaseStep) - codeLength   # :288
            tr["remCarrPhase"][loopCnt - 1] = remCarrPhase
            time = np.arange(0, blksize + 1, dtype=np.float64) / s.samplingFreq
            trigarg = ((carrFreq * 2.0 * np.pi) * time) + remCarrPhase
            remCarrPhase = math.fmod(trigarg[blksize], 2 * np.pi)
            bb = np.exp(-1j * trigarg[:blksize]) * rawSignal
            iB, qB = bb.real, bb.imag
            I_E = float(np.sum(cmCode[iE] * iB)); Q_E = float(np.sum(cmCode[iE] * qB))
            I_P = float(np.sum(cmCode[iP] * iB)); Q_P = float(np.sum(cmCode[iP] * qB))
            I_L = float(np.sum(cmCode[iL] * iB)); Q_L = float(np.sum(cmCode[iL] * qB))
            if pilotOn:                                            # :261-285, :319-324
                o = codeLength * (CLCodePhase - 1)
                I_ECL = float(np.sum(CLCode[iE + o] * iB)); Q_ECL = float(np.sum(CLCode[iE + o] * qB))
                I_PCL = float(np.sum(CLCode[iP + o] * iB)); Q_PCL = float(np.sum(CLCode[iP + o] * qB))
                I_LCL = float(np.sum(CLCode[iL + o] * iB)); Q_LCL = float(np.sum(CLCode[iL + o] * qB))
            with np.errstate(divide="ignore", invalid="ignore"):
                carrError = float(np.arctan(np.float64(Q_P) / np.float64(I_P)) / (2.0 * np.pi))
                sE = math.sqrt(I_E ** 2 + Q_E ** 2); sL = math.sqrt(I_L ** 2 + Q_L ** 2)
                codeError = float((np.float64(sE) - sL) / (np.float64(sE) + sL))
                if pilotOn:
                    carrErrorCL = float(np.arctan(np.float64(Q_PCL) / np.float64(I_PCL)) / (2.0 * np.pi))   # :335
                    carrError = (carrError + carrErrorCL) / 2                                              # :339
                    sEc = math.sqrt(I_ECL ** 2 + Q_ECL ** 2); sLc = math.sqrt(I_LCL ** 2 + Q_LCL ** 2)
                    codeErrorCL = float((np.float64(sEc) - sLc) / (np.float64(sEc) + sLc))                # :359
                    codeError = (codeError + codeErrorCL) / 2                                              # :361
                    CLCodePhase = CLCodePhase + 1                                                          # :363-366
                    if CLCodePhase >= 76:
                        CLCodePhase = 1
            d2CarrError = d2CarrError + carrError * pf3
            dCarrError = d2CarrError + carrError * pf2 + dCarrError
            carrNco = dCarrError + carrError * pf1
            tr["carrFreq"][loopCnt - 1] = carrFreq
            carrFreq = carrFreqBasis + carrNco
            codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code)
            oldCodeNco = codeNco; oldCodeError = codeError
            if pilotOn:                                            # :396-402
                tr["Pilot_I_E"][loopCnt - 1] = I_ECL; tr["Pilot_I_P"][loopCnt - 1] = I_PCL; tr["Pilot_I_L"][loopCnt - 1] = I_LCL
                tr["Pilot_Q_E"][loopCnt - 1] = Q_ECL; tr["Pilot_Q_P"][loopCnt - 1] = Q_PCL; tr["Pilot_Q_L"][loopCnt - 1] = Q_LCL
            tr["codeFreq"][loopCnt - 1] = codeFreq / 2             # :376
            codeFreq = s.codeFreqBasis * 2 - codeNco               # :379
            tr["dllDiscr"][loopCnt - 1] = codeError / 2; tr["dllDiscrFilt"][loopCnt - 1] = codeNco / 2   # :382-383
            tr["pllDiscr"][loopCnt - 1] = carrError; tr["pllDiscrFilt"][loopCnt - 1] = carrNco
            tr["I_E"][loopCnt - 1] = I_E; tr["I_P"][loopCnt - 1] = I_P; tr["I_L"][loopCnt - 1] = I_L
            tr["Q_E"][loopCnt - 1] = Q_E; tr["Q_P"][loopCnt - 1] = Q_P; tr["Q_L"][loopCnt - 1] = Q_L
            if loopCnt % s.CNo_VSMinterval == 0:
                vsmCnt += 1
                lo = loopCnt - s.CNo_VSMinterval
                tr["VSMValue"][vsmCnt - 1] = CNoVSM(tr["I_P"][lo:loopCnt], tr["Q_P"][lo:loopCnt], s.CNo_accTime)
                tr["VSMIndex"][vsmCnt - 1] = loopCnt
        tr["status"] = channel[ch]["status"]
    return out


def tracking_b1c_nb(raw: np.ndarray, channel: list, s: Settings, codes: dict):
    return _tracking_b1c(raw, channel, s, codes, False, 0.0)


def CalcWeighingFactor(s: Settings) -> float:
    """BDS/B1C/include/CalcWeighingFactor.m:45-82: data-channel weight of the full-band code discriminator from the BOC(1,1)
    and QMBOC power spectra integrated over the front-end bandwidth settings.FEBW.  MATLAB's `integral` (adaptive
    Gauss-Kronrod, RelTol 1e-6) and scipy's `quad` agree to ~1e-9 relative here; the factor is a settings-derived scalar
    that the engine takes from its caller."""
    from scipy.integrate import quad
    fc = s.codeFreqBasis; Tc = 1 / fc; Br = s.FEBW
    def boc(f, m):                                                  # BOC(m,1) spectrum, m = 1 -> pi/2, m = 6 -> pi/12
        a = np.pi / (2 * m)
        return Tc * (np.sin(a * f / fc) * np.sin(np.pi * f / fc) / np.cos(a * f / fc) * fc / f / np.pi) ** 2
    G11 = lambda f: boc(f, 1)
    Gp = lambda f: 29 / 33 * boc(f, 1) + 4 / 33 * boc(f, 6)
    def I(g):                                                       # even integrand, removable singularity at 0
        return 2 * quad(g, 0, Br / 2, limit=400, epsabs=0, epsrel=1e-12)[0]
    P11 = I(G11); P11_2 = I(lambda f: G11(f) * f ** 2)
    Pp = I(Gp); Pp_2 = I(lambda f: Gp(f) * f ** 2)
    rms11 = (P11_2 / P11) ** 0.5; rmsp = (Pp_2 / Pp) ** 0.5
    temp1 = 11 * P11 * rms11 ** 2
    temp2 = 33 * Pp * rmsp ** 2
    return float(temp1 / (temp1 + temp2))


def tracking_b1c_wb(raw: np.ndarray, channel: list, s: Settings, codes: dict, factor: float):
    """BDS/B1C/include/WB_tracking.m:47-474 (settings.pilotTRKflag == 2): as NB_tracking plus the pilot BOC(6,1) table
    codes[PRN][2] (122760 entries) indexed by ceil(tcode*6)+1 (:283-305), 18 sums, composite pilot correlations
    -sqrt(4/33)*p61 +- sqrt(29/33)*p11 cross terms (:339-344), carrier error (data + 3*pilot)/4 (:356), code error weighted by
    `factor` = CalcWeighingFactor(settings) (:374), six composite Pilot_* rows recorded (:409-414)."""
    return _tracking_b1c(raw, channel, s, codes, True, factor)


def _tracking_b1c(raw: np.ndarray, channel: list, s: Settings, codes: dict, wb: bool, factor: float):
    """BDS/B1C/include/NB_tracking.m:47-365 (settings.pilotTRKflag == 1): 10 ms epochs, BOC(1,1) sub-chip tables indexed by
    ceil(tcode*2)+1 (:225-246), code NCO centred on channel.codeFreq, the pilot in quadrature with atan(-I/Q) (:301), carrier
    and code discriminators weighted 11/40 : 29/40 (:302, :318), code discriminators scaled by (1 - spacing) (:313-317),
    Pilot_I_P / Pilot_Q_P recorded.  The DataCNo / PLD block (Calc_CNo_PLD.m) is host post-processing of these rows."""
    nE = int(matlab_round(s.msToProcess / 1000 / s.intTime))       # :49
    out = []
    for _ in range(s.numberOfChannels):
        tr = dict(status="-", PRN=0)
        tr["absoluteSample"] = np.zeros(nE)
        for f in ("codeFreq", "carrFreq", "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase"):
            tr[f] = np.full(nE, np.inf)
        for f in ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L", "Pilot_I_P", "Pilot_Q_P"):
            tr[f] = np.zeros(nE)
        if wb:
            for f in ("Pilot_I_E", "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_L"):
                tr[f] = np.zeros(nE)
        out.append(tr)
    spc = s.dllCorrelatorSpacing
    codeLength = int(s.codeLength)
    PDIcode = s.intTime
    tau1code, tau2code = calcLoopCoef(s.dllNoiseBandwidth, s.dllDampingRatio, 1.0)
    pf3, pf2, pf1 = calcLoopCoefCarr(s)
    for ch in range(s.numberOfChannels):
        if channel[ch]["PRN"] == 0:
            continue
        tr = out[ch]
        PRN = channel[ch]["PRN"]
        tr["PRN"] = PRN
        pos = 2 * (s.skipNumberOfBytes + channel[ch]["codePhase"] - 1)
        c = np.asarray(codes[PRN][0], dtype=np.float64)
        D = np.concatenate([[c[2 * codeLength - 1]], c, [c[0]]])   # :155-156
        c = np.asarray(codes[PRN][1], dtype=np.float64)
        P11 = np.concatenate([[c[2 * codeLength - 1]], c, [c[0]]])
        if wb:                                                     # WB_tracking.m:181-183
            c = np.asarray(codes[PRN][2], dtype=np.float64)
            P61 = np.concatenate([[c[12 * codeLength - 1]], c, [c[0]]])
        codeFreq = channel[ch]["codeFreq"]; remCodePhase = 0.0     # :163
        carrFreq = channel[ch]["acquiredFreq"]; carrFreqBasis = channel[ch]["acquiredFreq"]; remCarrPhase = 0.0
        oldCodeNco = oldCodeError = 0.0
        d2CarrError = dCarrError = 0.0
        for loopCnt in range(1, nE + 1):
            tr["absoluteSample"][loopCnt - 1] = pos / 2
            step = codeFreq / s.samplingFreq
            blksize = int(math.ceil((codeLength - remCodePhase) / step))
            chunk = raw[pos: pos + 2 * blksize]
            pos += chunk.size
            if chunk.size != 2 * blksize:
                return out
            rawSignal = chunk[0::2].astype(np.float64) + 1j * chunk[1::2].astype(np.float64)
            tr["remCodePhase"][loopCnt - 1] = remCodePhase
            tE = colonop((remCodePhase - spc) * 2, step * 2, ((blksize - 1) * step + remCodePhase - spc) * 2)
            tL = colonop((remCodePhase + spc) * 2, step * 2, ((blksize - 1) * step + remCodePhase + spc) * 2)
            tP = colonop(remCodePhase * 2, step * 2, ((blksize - 1) * step + remCodePhase) * 2)
            iE = np.ceil(tE).astype(np.int64); iL = np.ceil(tL).astype(np.int64); iP = np.ceil(tP).astype(np.int64)
            remCodePhase = tP[blksize - 1] / 2 + step - codeLength   # :248
            tr["remCarrPhase"][loopCnt - 1] = remCarrPhase
            time = np.arange(0, blksize + 1, dtype=np.float64) / s.samplingFreq
            trigarg = ((carrFreq * 2.0 * np.pi) * time) + remCarrPhase
            remCarrPhase = math.fmod(trigarg[blksize], 2 * np.pi)
            bb = np.exp(-1j * trigarg[:blksize]) * rawSignal
            iB, qB = bb.real, bb.imag
            I_E = float(np.sum(D[iE] * iB)); Q_E = float(np.sum(D[iE] * qB))
            I_P = float(np.sum(D[iP] * iB)); Q_P = float(np.sum(D[iP] * qB))
            I_L = float(np.sum(D[iL] * iB)); Q_L = float(np.sum(D[iL] * qB))
            pI_E = float(np.sum(P11[iE] * iB)); pQ_E = float(np.sum(P11[iE] * qB))
            pI_P = float(np.sum(P11[iP] * iB)); pQ_P = float(np.sum(P11[iP] * qB))
            pI_L = float(np.sum(P11[iL] * iB)); pQ_L = float(np.sum(P11[iL] * qB))
            if wb:
                jE = np.ceil(tE * 6).astype(np.int64); jL = np.ceil(tL * 6).astype(np.int64); jP = np.ceil(tP * 6).astype(np.int64)   # :283,294,305
                sI_E = float(np.sum(P61[jE] * iB)); sQ_E = float(np.sum(P61[jE] * qB))
                sI_P = float(np.sum(P61[jP] * iB)); sQ_P = float(np.sum(P61[jP] * qB))
                sI_L = float(np.sum(P61[jL] * iB)); sQ_L = float(np.sum(P61[jL] * qB))
                a61 = -math.sqrt(4 / 33); b11 = math.sqrt(29 / 33)
                cI_E = a61 * sI_E + b11 * pQ_E; cQ_E = a61 * sQ_E - b11 * pI_E    # :339-344
                cI_P = a61 * sI_P + b11 * pQ_P; cQ_P = a61 * sQ_P - b11 * pI_P
                cI_L = a61 * sI_L + b11 * pQ_L; cQ_L = a61 * sQ_L - b11 * pI_L
                with np.errstate(divide="ignore", invalid="ignore"):
                    carrError = float(np.arctan(np.float64(Q_P) / np.float64(I_P)) / (2.0 * np.pi))
                    p_carrError = float(np.arctan(np.float64(cQ_P) / np.float64(cI_P)) / (2.0 * np.pi))       # :353
                    carrError = (carrError * 1 + p_carrError * 3) / 4                                         # :356
                    sE = math.sqrt(I_E ** 2 + Q_E ** 2); sL = math.sqrt(I_L ** 2 + Q_L ** 2)
                    codeError = float((np.float64(sE) - sL) / (np.float64(sE) + sL)) * (1 - spc)              # :366-367
                    sEp = math.sqrt(cI_E ** 2 + cQ_E ** 2); sLp = math.sqrt(cI_L ** 2 + cQ_L ** 2)
                    p_codeError = float((np.float64(sEp) - sLp) / (np.float64(sEp) + sLp)) * (1 - spc)        # :371-372
                    codeError = codeError * factor + p_codeError * (1 - factor)                               # :374
            if not wb:
                with np.errstate(divide="ignore", invalid="ignore"):
                    carrError = float(np.arctan(np.float64(Q_P) / np.float64(I_P)) / (2.0 * np.pi))
                    p11_carrError = float(np.arctan(-np.float64(pI_P) / np.float64(pQ_P)) / (2.0 * np.pi))   # :301
                    carrError = (carrError * 11 + p11_carrError * 29) / 40                                   # :302
                    sE = math.sqrt(I_E ** 2 + Q_E ** 2); sL = math.sqrt(I_L ** 2 + Q_L ** 2)
                    codeError = float((np.float64(sE) - sL) / (np.float64(sE) + sL)) * (1 - spc)             # :313-314
                    sEp = math.sqrt(pI_E ** 2 + pQ_E ** 2); sLp = math.sqrt(pI_L ** 2 + pQ_L ** 2)
                    p11_codeError = float((np.float64(sEp) - sLp) / (np.float64(sEp) + sLp)) * (1 - spc)     # :315-317
                    codeError = (codeError * 11 + p11_codeError * 29) / 40                                   # :318
            d2CarrError = d2CarrError + carrError * pf3
            dCarrError = d2CarrError + carrError * pf2 + dCarrError
            carrNco = dCarrError + carrError * pf1
            tr["carrFreq"][loopCnt - 1] = carrFreq
            carrFreq = carrFreqBasis + carrNco
            codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code)
            oldCodeNco = codeNco; oldCodeError = codeError
            tr["codeFreq"][loopCnt - 1] = codeFreq
            codeFreq = channel[ch]["codeFreq"] - codeNco
            tr["dllDiscr"][loopCnt - 1] = codeError; tr["dllDiscrFilt"][loopCnt - 1] = codeNco
            tr["pllDiscr"][loopCnt - 1] = carrError; tr["pllDiscrFilt"][loopCnt - 1] = carrNco
            tr["I_E"][loopCnt - 1] = I_E; tr["I_P"][loopCnt - 1] = I_P; tr["I_L"][loopCnt - 1] = I_L
            tr["Q_E"][loopCnt - 1] = Q_E; tr["Q_P"][loopCnt - 1] = Q_P; tr["Q_L"][loopCnt - 1] = Q_L
            if wb:                                                 # WB_tracking.m:409-414: the composite pilot correlations
                tr["Pilot_I_E"][loopCnt - 1] = cI_E; tr["Pilot_Q_E"][loopCnt - 1] = cQ_E
                tr["Pilot_I_P"][loopCnt - 1] = cI_P; tr["Pilot_Q_P"][loopCnt - 1] = cQ_P
                tr["Pilot_I_L"][loopCnt - 1] = cI_L; tr["Pilot_Q_L"][loopCnt - 1] = cQ_L
            else:
                tr["Pilot_I_P"][loopCnt - 1] = pI_P; tr["Pilot_Q_P"][loopCnt - 1] = pQ_P
        tr["status"] = channel[ch]["status"]
    return out


# ---------------------------------------------------------------------------
# navigation-bit front end (GPS/GPS_L1CA/include/NAVdecoding.m:69-170, Common/navPartyChk.m)
# ---------------------------------------------------------------------------
def navPartyChk(ndat) -> int:
    """Common/navPartyChk.m: ndat = [D29* D30* d1..d24 D25..D30] as +-1 (32 values); returns -D30* when the six parity
    equations of IS-GPS-200 table 20-XIV hold, else 0."""
    n = [0] + [int(v) for v in ndat]                               # 1-based like the reference
    if n[2] != 1:
        for i in range(3, 27):
            n[i] = -n[i]
    sets = [
        (1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22, 25),
        (2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23, 26),
        (1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24),
        (2, 4, 6, 7, 8, 10, 11, 15, 16, 17, 18, 19, 22, 23, 25),
        (2, 3, 5, 7, 8, 9, 11, 12, 16, 17, 18, 19, 20, 23, 24, 26),
        (1, 5, 7, 8, 10, 11, 12, 13, 15, 17, 21, 24, 25, 26)]
    parity = [int(np.prod([n[i] for i in st])) for st in sets]
    if sum(int(parity[k] == n[27 + k]) for k in range(6)) == 6:
        return -1 * n[2]
    return 0


def nav_sync(I_P: np.ndarray, msToProcess: int):
    """NAVdecoding.m:69-170: (subFrameStart, navBits).  subFrameStart is the 1-based index of the first preamble-like pattern
    with another one 6000 ms later whose TLM and HOW words pass the parity check (0 when there is none); navBits the 1501
    bits summed over 20 ms from subFrameStart-20 (the last bit of the previous subframe, then five subframes) (None when that range leaves the record)."""
    searchStartOffset = 0                                          # :66
    preamble_bits = np.array([1, -1, -1, -1, 1, -1, 1, 1])
    preamble_ms = np.kron(preamble_bits, np.ones(20))             # :73
    bits = np.asarray(I_P, dtype=np.float64)[searchStartOffset:].copy()
    bits[bits > 0] = 1                                             # :82-83
    bits[bits <= 0] = -1
    n = bits.size
    # xcorr(bits, preamble_ms), non-negative lags (:86, :93-96): sum_j bits(l+j) preamble_ms(j)
    full = np.correlate(np.concatenate([bits, np.zeros(preamble_ms.size)]), preamble_ms, mode="valid")[:n]
    index = np.nonzero(np.abs(full) > 153)[0] + 1 + searchStartOffset
    index = index[(index > 40) & (index < msToProcess - (20 * 60 - 1))]   # :101
    subFrameStart = 0
    iset = set(int(v) for v in index)
    for i in index:                                                # :104-141
        if int(i) + 6000 in iset:
            b = np.asarray(I_P, dtype=np.float64)[int(i) - 40 - 1: int(i) + 20 * 60 - 1]
            b = b.reshape(-1, 20)
            sums = np.array([float(np.sum(row)) for row in b])
            sb = np.where(sums > 0, 1, -1)
            if navPartyChk(sb[0:32]) != 0 and navPartyChk(sb[30:62]) != 0:
                subFrameStart = int(i)
                break
    if subFrameStart == 0:
        return 0, None
    lo, hi = subFrameStart - 20 - 1, subFrameStart + 1500 * 20 - 1
    if lo < 0 or hi > np.asarray(I_P).size:
        return subFrameStart, None
    s = np.asarray(I_P, dtype=np.float64)[lo:hi].reshape(-1, 20)
    navBits = (np.array([float(np.sum(row)) for row in s]) > 0).astype(np.uint8)   # :152-166
    return subFrameStart, navBits


def unpack_cplx(data: np.ndarray) -> np.ndarray:
    """GPS/GPS_L2C/include/unpack_cplx.m:17-31: every byte of the 2-bit packed record becomes four schar values I1 Q1 I2 Q2
    through four 256-entry tables.  The tables are periodic patterns - LUT_I_long1 = [1 -1 1 -1 3 -3 3 -3] repeated,
    LUT_Q_long1 = [1 1 -1 -1 1 1 -1 -1 3 3 -3 -3 3 3 -3 -3] repeated, LUT_I_long2 / LUT_Q_long2 the same patterns stretched by 16 -
    restated here as such; returns the int8 I,Q interleaved record the reference then processes with fileType 2."""
    b = np.asarray(data, dtype=np.uint8).astype(np.int64)
    pat_i = np.array([1, -1, 1, -1, 3, -3, 3, -3])
    pat_q = np.array([1, 1, -1, -1, 1, 1, -1, -1, 3, 3, -3, -3, 3, 3, -3, -3])
    idx = np.arange(256)
    lut_i1, lut_q1 = pat_i[idx % 8], pat_q[idx % 16]
    lut_i2, lut_q2 = pat_i[(idx // 16) % 8], pat_q[(idx // 16) % 16]
    out = np.empty(4 * b.size, dtype=np.int8)
    out[0::4] = lut_i1[b]; out[1::4] = lut_q1[b]; out[2::4] = lut_i2[b]; out[3::4] = lut_q2[b]
    return out


# ---------------------------------------------------------------------------------------------------------------------------------
# BDS/B2a/include/Calc_CNo_PLD.m (B1C's twin differs in the pilot branch only) and the block of tracking.m that calls it
def Calc_CNo_PLD(trackResults: dict, s, loopCnt: int, signal: str = "BDS_B2a"):
    """[CNo, PllDetector] = Calc_CNo_PLD(trackResults, settings, loopCnt) - BDS/B2a/include/Calc_CNo_PLD.m:38-100;
    BDS/B1C/include/Calc_CNo_PLD.m:80-90 picks the pilot rows as recorded for pilotTRKflag == 2 and swapped for 1."""
    CNo = np.zeros(3)
    PllDetector = np.zeros(2)
    T = s.intTime                                                              # :41
    n = int(s.CNo_VSMinterval)                                                 # settings.CNoInterval
    I_P = np.asarray(trackResults["I_P"][loopCnt - n: loopCnt], dtype=np.float64)   # :44-45
    Q_P = np.asarray(trackResults["Q_P"][loopCnt - n: loopCnt], dtype=np.float64)

    def est(I_P, Q_P):
        Z = I_P ** 2 + Q_P ** 2                                                # :49
        Zm = np.mean(Z)                                                        # :51
        Zv = np.var(Z, ddof=1)                                                 # :52
        with np.errstate(invalid="ignore", divide="ignore"):
            Pav = np.sqrt(np.complex128(Zm ** 2 - Zv))                         # :54 (complex for a negative argument, as in MATLAB)
            Nv = 0.5 * (Zm - Pav)                                              # :56
            cno = np.abs((1 / T) * Pav / (2 * Nv))                             # :58
            NBP = (np.sum(I_P[I_P > 0]) - np.sum(I_P[I_P < 0])) ** 2 + np.sum(Q_P) ** 2   # :63
            NBD = (np.sum(I_P[I_P > 0]) - np.sum(I_P[I_P < 0])) ** 2 - np.sum(Q_P) ** 2   # :64
            return float(cno), float(NBD / NBP)                                # :66

    with np.errstate(invalid="ignore", divide="ignore"):
        DataCNo, PllDetector[0] = est(I_P, Q_P)
        CNo[0] = 10 * np.log10(DataCNo)                                        # :59
        PilotCNo = 0.0                                                         # :70
        flag = int(s.pilotTRKflag)
        if flag == 2 and signal == "BDS_B1C":                                  # B1C Calc_CNo_PLD.m:80-83
            PilotCNo, PllDetector[1] = est(np.asarray(trackResults["Pilot_I_P"][loopCnt - n: loopCnt]),
                                           np.asarray(trackResults["Pilot_Q_P"][loopCnt - n: loopCnt]))
            CNo[1] = 10 * np.log10(PilotCNo)
        elif flag == 1:                                                        # :72-75: Q_P = Pilot_I_P, I_P = Pilot_Q_P
            PilotCNo, PllDetector[1] = est(np.asarray(trackResults["Pilot_Q_P"][loopCnt - n: loopCnt]),
                                           np.asarray(trackResults["Pilot_I_P"][loopCnt - n: loopCnt]))
            CNo[1] = 10 * np.log10(PilotCNo)
        CNo[2] = 10 * np.log10(DataCNo + PilotCNo)                             # :100
    return CNo, PllDetector


def cno_pld_rows(trackResults: dict, s, done: int, signal: str = "BDS_B2a") -> dict:
    """DataCNo / DataPLD / PilotCNo / PilotPLD / total C/N0 as BDS/B2a/include/tracking.m:409-432 fills them: every CNoInterval
    epochs, the C/N0 values smoothed 0.5/0.5 with the previous interval's (tempCNoValue starts at zeros(1,3), :192)."""
    n = int(s.CNo_VSMinterval)
    nE = len(trackResults["I_P"])
    nv = nE // n
    pilot = int(s.pilotTRKflag) >= 1
    res = {"DataCNo": np.zeros(nv), "DataPLD": np.zeros(nv), "PilotCNo": np.zeros(nv), "PilotPLD": np.zeros(nv), "TotalCNo": np.zeros(nv)}
    temp = np.zeros(3)                                                         # :192
    for loopCnt in range(1, done + 1):
        if loopCnt % n == 0:                                                   # :409
            CNoValue, PllDetector = Calc_CNo_PLD(trackResults, s, loopCnt, signal)   # :411-412
            c = loopCnt // n                                                   # :414
            res["DataCNo"][c - 1] = CNoValue[0] * 0.5 + temp[0] * 0.5          # :418-419
            res["DataPLD"][c - 1] = PllDetector[0]                             # :421
            if pilot:                                                          # :424-430
                res["PilotCNo"][c - 1] = CNoValue[1] * 0.5 + temp[1] * 0.5
                res["TotalCNo"][c - 1] = CNoValue[2] * 0.5 + temp[2] * 0.5
                res["PilotPLD"][c - 1] = PllDetector[1]
            temp = CNoValue                                                    # :432
    return res


# ---------------------------------------------------------------------------------------------------------------------------------
# Code generators of the nine signals whose primary codes the reference builds at run time.  Each function follows
# its .m file statement by statement (+-1 registers, prod() feedback, circshift) - deliberately NOT the bit-packed form the library
# uses (csrc/codegen.h), so that the two are independent witnesses.  The ICD constant tables come from oracle/icd_tables.py.
import icd_tables as _icd


def _circshift(v, k):
    """circshift(v', k)' for a row vector."""
    k %= len(v)
    return v[-k:] + v[:-k] if k else list(v)


def _prod(v, pos):
    p = 1
    for i in pos:
        p *= v[i - 1]
    return p


def _l5_code(PRN, advance, codeLength=10230):
    """GPS/GPS_L5C/include/generateL5Icode.m:44-133 (generateL5Qcode.m is the same with its own advance table)."""
    xa_FeedbackPos = [9, 10, 12, 13]                                       # :47
    xa_reg = [-1] * 13                                                     # :49
    XA = [0] * codeLength
    reset_state = [-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 1, -1]      # :53
    for ind in range(codeLength):                                          # :55
        XA[ind] = xa_reg[-1]
        if xa_reg == reset_state:
            xa_reg = [-1] * 13
        else:
            feedback = _prod(xa_reg, xa_FeedbackPos)
            xa_reg = _circshift(xa_reg, 1)
            xa_reg[0] = feedback
    xbi_FeedbackPos = [1, 3, 4, 6, 7, 8, 12, 13]                           # :104
    xbi_reg = [-1] * 13
    XBI = [0] * codeLength
    resetPos = advance[PRN - 1]                                            # :110
    for _ in range(resetPos):                                              # :112-118
        feedback = _prod(xbi_reg, xbi_FeedbackPos)
        xbi_reg = _circshift(xbi_reg, 1)
        xbi_reg[0] = feedback
    for ind in range(codeLength):                                          # :121-129
        XBI[ind] = xbi_reg[-1]
        feedback = _prod(xbi_reg, xbi_FeedbackPos)
        xbi_reg = _circshift(xbi_reg, 1)
        xbi_reg[0] = feedback
    return np.array(XBI, dtype=np.int64) * np.array(XA, dtype=np.int64)    # :132


def generateL5Icode(PRN, codeLength=10230):
    return _l5_code(PRN, _icd.L5I_ADVANCE, codeLength)


def generateL5Qcode(PRN, codeLength=10230):
    return _l5_code(PRN, _icd.L5Q_ADVANCE, codeLength)


def _dec2bin(x):
    """dec2bin(x) - 48: the binary digits without leading zeros."""
    return [int(c) for c in bin(int(x))[2:]]


def _gal_e5_primary(start_value, Feedback_Reg1, Feedback_Reg2):
    """GAL/GAL_E5a/include/generateE5aIcode.m:45-97 (the E5a-Q, E5b-I and E5b-Q files differ in the tables and the octal
    feedback strings only)."""
    Register1 = [1] * 14                                                   # :47
    taps1_coef = _dec2bin(int(Feedback_Reg1, 8))[:14]                      # :68-71 dec2bin(base2dec(.,8))-48, then (1:14)
    taps2_coef = _dec2bin(int(Feedback_Reg2, 8))[:14]
    StartValues = _dec2bin(start_value)                                    # :74-75
    Register2 = [0] * 14                                                   # :76
    Register2[14 - len(StartValues):] = StartValues                        # :77 Register2(end-length+1:end)
    Pri = [0] * 10230
    for ind in range(10230):                                               # :80
        RegOut1 = [a * b for a, b in zip(Register1, taps1_coef)]
        RegOut2 = [a * b for a, b in zip(Register2, taps2_coef)]
        Pri[ind] = (1 - 2 * RegOut1[0]) * (1 - 2 * RegOut2[0])             # :84
        feedback1, feedback2 = 1, 1
        for v in RegOut1:
            feedback1 *= (1 - 2 * v)                                       # :87 prod(1 - 2*RegOut1)
        for v in RegOut2:
            feedback2 *= (1 - 2 * v)
        feedback1 = 0 if feedback1 == 1 else 1                             # :90-99
        feedback2 = 0 if feedback2 == 1 else 1
        Register1 = _circshift(Register1, -1)                              # :101 shift left
        Register2 = _circshift(Register2, -1)
        Register1[-1] = feedback1                                          # :104
        Register2[-1] = feedback2
    return np.array(Pri, dtype=np.int64)


def generateE5aIcode(PRN):
    return _gal_e5_primary(_icd.E5AI_START[PRN - 1], "40503", "50661")


def generateE5aQcode(PRN):
    return _gal_e5_primary(_icd.E5AQ_START[PRN - 1], "40503", "50661")


def generateE5bIcode(PRN):
    return _gal_e5_primary(_icd.E5BI_START[PRN - 1], "64021", "51445")


def generateE5bQcode(PRN):
    return _gal_e5_primary(_icd.E5BQ_START[PRN - 1], "64021", "43143")


def _gal_secondary(code):
    """GAL/GAL_E5a/include/generateE5aQ_secondary.m:73-87: 25 hex characters -> 100 chips, 1 - 2*bit."""
    first_segement = [0] * (13 * 4)
    second_segement = [0] * (12 * 4)
    t = _dec2bin(int(code[:13], 16))
    first_segement[len(first_segement) - len(t):] = t
    t = _dec2bin(int(code[13:], 16))
    second_segement[len(second_segement) - len(t):] = t
    return 1 - 2 * np.array(first_segement + second_segement, dtype=np.int64)


def generateE5aQ_secondary(PRN):
    return _gal_secondary(_icd.E5AQ_SECONDARY[PRN - 1])


def generateE5bQ_secondary(PRN):
    return _gal_secondary(_icd.E5BQ_SECONDARY[PRN - 1])


def _b2a_code(reg2_ini, reg1_FeedbackPos, reg2_FeedbackPos, codeLength=10230):
    """BDS/B2a/include/generateB2aDataCode.m:111-138 (generateB2aPilotCode.m: other taps and initial states)."""
    register1 = [-1] * 13                                                  # :116
    bits = [(reg2_ini >> (12 - i)) & 1 for i in range(13)]                 # row PRN of B2aData_reg2_ini
    register2 = [1 - 2 * b for b in bits]                                  # :117
    code = [0] * codeLength
    reset_index = 8190                                                     # :120
    for ind in range(1, codeLength + 1):
        code[ind - 1] = register1[-1] * register2[-1]                      # :123
        feedback1 = _prod(register1, reg1_FeedbackPos)
        register1 = _circshift(register1, 1)
        register1[0] = feedback1
        feedback2 = _prod(register2, reg2_FeedbackPos)
        register2 = _circshift(register2, 1)
        register2[0] = feedback2
        if ind == reset_index:                                             # :135-137
            register1 = [-1] * 13
    return np.array(code, dtype=np.int64)


def generateB2aDataCode(PRN):
    return _b2a_code(_icd.B2AD_REG2[PRN - 1], [1, 5, 11, 13], [3, 5, 9, 11, 12, 13])


def generateB2aPilotCode(PRN):
    return _b2a_code(_icd.B2AP_REG2[PRN - 1], [3, 6, 7, 13], [1, 5, 7, 8, 12, 13])


def generateCAcode53(PRN):
    """BDS/B1I/include/generateCAcode53.m:38-103."""
    g1 = [0] * 2046
    reg = [-x for x in [-1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1]]            # :43
    for i in range(2046):
        g1[i] = reg[10]
        saveBit = reg[0] * reg[6] * reg[7] * reg[8] * reg[9] * reg[10]     # :48
        reg[1:11] = reg[0:10]
        reg[0] = saveBit
    g2 = [0] * 2046
    reg = [-x for x in [-1, 1, -1, 1, -1, 1, -1, 1, -1, 1, -1]]
    g2s1, g2s2, g2s3 = _icd.B1I_G2S1, _icd.B1I_G2S2, _icd.B1I_G2S3
    for i in range(2046):
        if PRN > 37:                                                       # :83-90
            g2[i] = reg[g2s1[PRN - 1] - 1] * reg[g2s2[PRN - 1] - 1] * reg[g2s3[PRN - 37 - 1] - 1]
        else:
            g2[i] = reg[g2s1[PRN - 1] - 1] * reg[g2s2[PRN - 1] - 1]
        saveBit = reg[0] * reg[1] * reg[2] * reg[3] * reg[4] * reg[7] * reg[8] * reg[10]
        reg[1:11] = reg[0:10]
        reg[0] = saveBit
    return -(np.array(g1, dtype=np.int64) * np.array(g2, dtype=np.int64))  # :102


def _l2c_code(code_init, CodeLength):
    """GPS/GPS_L2C/include/generateCMcode.m:88-106: 27-stage register, taps multiplied by the output chip."""
    RegPos = [4, 7, 9, 12, 15, 17, 19, 22, 23, 24, 25]                     # :41
    reg = _dec2bin(code_init)                                              # :92 dec2bin(oct2dec(code_init)) - 48
    reg = [0] * (27 - len(reg)) + reg                                      # :93
    reg = [-1 if b else 1 for b in reg]                                    # :94-95
    code = np.zeros(CodeLength, dtype=np.int64)
    for index in range(CodeLength):
        c = reg[-1]
        code[index] = c                                                    # :99
        reg = _circshift(reg, 1)                                           # :100
        for p in RegPos:
            reg[p - 1] *= c                                                # :101
    return code


def generateCMcode(PRN, codeLength=10230):
    """The return-to-zero CM sequence [c1 0 c2 0 ...] (generateCMcode.m:108-111).  PRN 1..63."""
    CM = _l2c_code(_icd.L2CM_INIT[PRN - 1], codeLength)
    out = np.zeros(2 * codeLength, dtype=np.int64)
    out[0::2] = CM
    return out


def generateCLcode(PRN, CLCodeLength=767250):
    """[0 c1 0 c2 ...] (generateCLcode.m:107-110)."""
    CL = _l2c_code(_icd.L2CL_INIT[PRN - 1], CLCodeLength)
    out = np.zeros(2 * CLCodeLength, dtype=np.int64)
    out[1::2] = CL
    return out


def JacobiSymbol_prime(a, N):
    """JacobiSymbol(a, N) of BDS/B1C/include/JacobiSymbol.m for an odd prime N (all its callers pass 10243 or 3607): the
    Legendre symbol, by Euler's criterion."""
    if a % N == 0:
        return 0
    return 1 if pow(a, (N - 1) // 2, N) == 1 else -1


def _weil_primary(w, p, N=10243, length=10230):
    """BDS/B1C/include/generateDataBOC11.m:66-82."""
    legendre = np.zeros(N, dtype=np.int64)
    for ind in range(1, N):                                                # :68-70
        legendre[ind] = JacobiSymbol_prime(ind, N)
    legendre[legendre == -1] = 0                                           # :71
    Primary = np.zeros(length, dtype=np.int64)
    for ind in range(length):                                              # :77-80
        k = (ind + p - 1) % N
        Primary[ind] = legendre[k] ^ legendre[(k + w) % N]
    return 1 - 2 * Primary                                                 # :82


def generateDataBOC11(PRN):
    P = _weil_primary(_icd.B1CD_W[PRN - 1], _icd.B1CD_P[PRN - 1])
    out = np.zeros(2 * P.size, dtype=np.int64)
    out[0::2] = -P                                                         # :86-89
    out[1::2] = P
    return out


def generatePilotBOC11(PRN):
    P = _weil_primary(_icd.B1CP_W[PRN - 1], _icd.B1CP_P[PRN - 1])
    out = np.zeros(2 * P.size, dtype=np.int64)
    out[0::2] = -P
    out[1::2] = P
    return out


def generatePilotBOC61(PRN):
    """BDS/B1C/include/generatePilotBOC61.m:103-110: twelve entries (-1)^ii * chip per chip."""
    P = _weil_primary(_icd.B1CP_W[PRN - 1], _icd.B1CP_P[PRN - 1])
    sign = np.array([(-1) ** ii for ii in range(1, 13)], dtype=np.int64)
    return (P[:, None] * sign[None, :]).reshape(-1)


def _e1_primary(hexrow):
    bits = bin(int(hexrow, 16))[2:].zfill(4092)
    return 1 - 2 * np.array([int(c) for c in bits], dtype=np.int64)        # generateE1Bcode.m:55


def generateE1Bcode(PRN):
    """GAL/GAL_E1C/include/generateE1Bcode.m:44-64 with the memory code of E1b.dat: BOC(1,1) sub-chips [c -c]."""
    c = _e1_primary(_icd.E1B_HEX[PRN - 1])
    out = np.zeros(2 * c.size, dtype=np.int64)
    out[0::2] = c
    out[1::2] = -c
    return out


def generateE1Ccode(PRN):
    c = _e1_primary(_icd.E1C_HEX[PRN - 1])
    out = np.zeros(2 * c.size, dtype=np.int64)
    out[0::2] = c
    out[1::2] = -c
    return out
