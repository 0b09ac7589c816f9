"""``settings`` struct of the reference (GPS/GPS_L1CA/initSettings.m:44-136), hot-path fields.

Field names are the reference's own so a MATLAB user finds every knob where it was; the nested
``settings.CNo.*`` struct is flattened to ``CNo_accTime`` / ``CNo_VSMinterval``.
"""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class Settings:
    signal: str = "GPS_L1CA"            # which reference folder these settings belong to
    msToProcess: int = 60000            # initSettings.m:47
    numberOfChannels: int = 12          # :50
    skipNumberOfBytes: int = 0          # :56
    fileName: str = "../../../L1_IF20KHz_FS18MHz.bin"   # :61
    dataType: str = "schar"             # :63
    fileType: int = 2                   # :68
    IF: float = 20e3                    # :71
    samplingFreq: float = 18e6          # :72
    codeFreqBasis: float = 1.023e6      # :73
    codeLength: float = 1023.0          # :76
    skipAcquisition: int = 0            # :80
    acqSatelliteList: list = field(default_factory=lambda: list(range(1, 33)))   # :83
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
    freqSpacing: float = 0.0            # GLONASS only: FDMA channel spacing (GLO_GL1/initSettings.m:72)
    carrFreqBasis: float = 0.0          # B3I: RF carrier used to aid the code NCO (BDS/B3I/initSettings.m:132)
    pilotTRKflag: int = 0               # Galileo E1: track the pilot component too (GAL/GAL_E1C/initSettings.m:113)
    codeDir: str = ""                   # Galileo E1: directory holding E1b.dat / E1c.dat (the reference keeps them in include/)
    stepSize: float = 0.0               # BDS B1I: sub-bin step request (BDS/B1I/initSettings.m:96; 0 = [] = derive it)
    acqStep: float = 0.0                # GPS L2C: sub-bin step (GPS/GPS_L2C/initSettings.m:94)
    acqCohT: int = 20                   # GPS L2C: coherent time in ms (:91); BDS B1C: BDS/B1C/initSettings.m:97 (10)
    pilotACQflag: int = 0               # BDS B1C: the pilot replica joins the acquisition (BDS/B1C/initSettings.m:74)
    FEBW: float = 27e6                  # BDS B1C: front-end bandwidth (BDS/B1C/initSettings.m:59), input of CalcWeighingFactor.m
    wbFactor: float | None = None       # BDS B1C full band: CalcWeighingFactor(settings) if the caller already has it

    @property
    def is_glonass(self) -> bool:
        return self.signal in ("GLO_GL1", "GLO_GL2")

    @property
    def is_varb(self) -> bool:
        """Acquisition variant B (circularly shifted spectra, best row kept): BDS B1I, GPS L2C."""
        return self.signal in ("BDS_B1I", "GPS_L2C")

    @property
    def is_fam5(self) -> bool:
        """One of the four 10230-chip data + pilot signals (L5C, E5a, E5b, B2a)."""
        return self.signal in ("GPS_L5C", "GAL_E5a", "GAL_E5b", "BDS_B2a")


# GLO/GLO_GL1/initSettings.m:44-146 (GLO_GL2 differs in freqSpacing and fileName only).  The GLONASS
# folders call the record offset `skipNumberOfSamples`; it is the same quantity as skipNumberOfBytes.
_GLO_DEFAULTS = dict(fileName="../../../GL1_IF0KHz_FS12MHz.bin", IF=0.0, samplingFreq=12e6, codeFreqBasis=0.511e6,
                     codeLength=511.0, acqSatelliteList=list(range(-7, 7)), acqSearchBand=5000.0, acqThreshold=2.0,
                     dllNoiseBandwidth=2.0, pllNoiseBandwidth=25.0, freqSpacing=562.5e3)


# BDS/B3I/initSettings.m:44-132
_B3I_DEFAULTS = dict(numberOfChannels=15, fileName="../../../B3I_IF20KHz_FS18MHz.bin", codeLength=10230.0,
                     codeFreqBasis=10.23e6, acqSatelliteList=list(range(1, 64)), acqSearchBand=5000.0, acqNonCohTime=10,
                     acqThreshold=3.0, resamplingThreshold=45e6, dllNoiseBandwidth=2.0, pllNoiseBandwidth=15.0,
                     carrFreqBasis=1268.520e6)


# GAL/GAL_E1C/initSettings.m:44-140
_E1C_DEFAULTS = dict(codeLength=4092.0, acqSatelliteList=list(range(1, 37)), acqSearchBand=7000.0, acqNonCohTime=1,
                     acqSearchStep=150.0, acqThreshold=10.0, resamplingThreshold=50e6, dllCorrelatorSpacing=0.3,
                     pllNoiseBandwidth=15.0, intTime=0.004, pilotTRKflag=1, CNo_accTime=0.004, CNo_VSMinterval=400)


# GPS/GPS_L5C, GAL/GAL_E5a, GAL/GAL_E5b, BDS/B2a initSettings.m (hot-path fields).  B2a's settings.CNoInterval (:128)
# takes the place of CNo.VSMinterval.
_FAM5_BASE = dict(codeLength=10230.0, codeFreqBasis=10.23e6, acqSearchBand=5000.0, acqSearchStep=500.0, acqThreshold=4.5,
                  pllNoiseBandwidth=15.0, carrFreqBasis=1176.45e6)
_FAM5 = {
    "GPS_L5C": dict(acqSatelliteList=list(range(1, 33)), acqNonCohTime=25, dllNoiseBandwidth=2.0, CNo_VSMinterval=400,
                    resamplingThreshold=50e6, pilotTRKflag=0, fileName="../../../L5_IF20KHz_FS18MHz.bin"),
    "GAL_E5a": dict(acqSatelliteList=list(range(1, 37)), acqNonCohTime=15, dllNoiseBandwidth=1.5, CNo_VSMinterval=100,
                    resamplingThreshold=45e6, pilotTRKflag=1, fileName="../../../L5_IF20KHz_FS18MHz.bin"),
    "GAL_E5b": dict(acqSatelliteList=list(range(1, 37)), acqNonCohTime=15, acqSearchStep=60.0, dllNoiseBandwidth=1.5,
                    pllNoiseBandwidth=25.0, CNo_VSMinterval=100, resamplingThreshold=45e6, carrFreqBasis=1207.14e6,
                    pilotTRKflag=1, fileName="../../../E5b_IF20KHz_FS18MHz.bin"),
    "BDS_B2a": dict(acqSatelliteList=list(range(19, 31)) + list(range(32, 47)) + [59, 60], acqNonCohTime=15,
                    acqThreshold=5.0, dllNoiseBandwidth=2.0, CNo_VSMinterval=200, resamplingThreshold=50e6, pilotTRKflag=0,
                    fileName="../../../L5_IF20KHz_FS18MHz.bin"),
}
FAM5_SIGNALS = tuple(_FAM5)


# BDS/B1I/initSettings.m and GPS/GPS_L2C/initSettings.m (hot-path fields); acqSearchBand is in kHz in these two folders
_B1I_DEFAULTS = dict(codeFreqBasis=2.046e6, codeLength=2046.0, acqSatelliteList=list(range(6, 59)), acqSearchBand=10.0,
                     acqThreshold=2.0, resamplingThreshold=9e6, stepSize=125.0, dllNoiseBandwidth=4.0, pllNoiseBandwidth=35.0,
                     CNo_VSMinterval=400, fileName="../../../B1I_IF20KHz_FS18MHz.bin")
_L2C_DEFAULTS = dict(samplingFreq=8e6, codeFreqBasis=0.5115e6, codeLength=10230.0, acqSearchBand=10.0, acqThreshold=1.5,
                     resamplingThreshold=6e6, acqStep=12.5, acqCohT=20, dllNoiseBandwidth=4.0, dllCorrelatorSpacing=0.25,
                     pllNoiseBandwidth=10.0, intTime=0.02, CNo_accTime=0.02, CNo_VSMinterval=40,
                     fileName="../../../L2_IF20KHz_FS8MHz.bin")


# BDS/B1C/initSettings.m (hot-path fields of the acquisition)
_B1C_DEFAULTS = dict(numberOfChannels=15, codeLength=10230.0, codeFreqBasis=1.023e6, acqSatelliteList=list(range(1, 63)),
                     acqSearchBand=5000.0, acqCohT=10, acqStep=50.0, acqThreshold=10.0, resamplingThreshold=15e6, pilotACQflag=1,
                     pilotTRKflag=1, intTime=0.01, CNo_VSMinterval=50, CNo_accTime=0.01, carrFreqBasis=1575.42e6,
                     dllNoiseBandwidth=1.0, dllCorrelatorSpacing=0.06, pllNoiseBandwidth=18.0,
                     fileName="../../../B1C_IF20KHz_FS18MHz.bin")


def varb_step(s: "Settings") -> float:
    """Sub-bin step of the variant-B acquisitions: settings.acqStep for L2C; for B1I settings.stepSize resolved the way
    BDS/B1I/include/acquisition.m:24-39 does ([] -> 0.5/(4 ms); == freqResolution -> itself; else the nearest divisor of
    freqResolution on a 0.25 Hz raster that does not exceed it)."""
    import math
    import numpy as np
    if s.signal == "GPS_L2C":
        return float(s.acqStep)
    spb = int(math.floor(s.samplingFreq / (s.codeFreqBasis / (4 * s.codeLength)) + 0.5))
    res = s.samplingFreq / spb
    if not s.stepSize:
        return 0.5 / (4 * s.codeLength / s.codeFreqBasis)
    if s.stepSize == res:
        return float(s.stepSize)
    steps = np.arange(1, res / 2 + 1e-12, 0.25)
    steps = steps[np.fmod(res, steps) == 0]
    diff = steps - s.stepSize
    k = int(np.argmin(np.abs(diff)))
    return float(steps[k - 1] if diff[k] > 0 else steps[k])


def init_settings(signal: str = "GPS_L1CA", **overrides) -> Settings:
    """``settings = initSettings()`` of the given signal folder (GPS/GPS_L1CA/init.m:56,
    GLO/GLO_GL1, GLO/GLO_GL2) with optional field overrides."""
    s = Settings(signal=signal)
    if signal in ("GLO_GL1", "GLO_GL2"):
        for k, v in _GLO_DEFAULTS.items():
            setattr(s, k, list(v) if isinstance(v, list) else v)
        if signal == "GLO_GL2":
            s.freqSpacing = 437.5e3
            s.fileName = "../../../GL2_IF0KHz_FS12MHz.bin"
    elif signal == "BDS_B3I":
        for k, v in _B3I_DEFAULTS.items():
            setattr(s, k, list(v) if isinstance(v, list) else v)
    elif signal in _FAM5:
        for k, v in {**_FAM5_BASE, **_FAM5[signal]}.items():
            setattr(s, k, list(v) if isinstance(v, list) else v)
    elif signal == "BDS_B1C":
        for k, v in _B1C_DEFAULTS.items():
            setattr(s, k, list(v) if isinstance(v, list) else v)
    elif signal in ("BDS_B1I", "GPS_L2C"):
        for k, v in (_B1I_DEFAULTS if signal == "BDS_B1I" else _L2C_DEFAULTS).items():
            setattr(s, k, list(v) if isinstance(v, list) else v)
    elif signal == "GAL_E1C":
        for k, v in _E1C_DEFAULTS.items():
            setattr(s, k, list(v) if isinstance(v, list) else v)
    elif signal != "GPS_L1CA":
        raise ValueError(f"signal {signal!r} is not implemented")
    for k, v in overrides.items():
        if k == "skipNumberOfSamples":
            k = "skipNumberOfBytes"
        if k == "resamplingFlag":
            k = "resamplingflag"
        if not hasattr(s, k):
            raise AttributeError(f"settings has no field {k!r}")
        setattr(s, k, v)
    return s


def num_to_process(s: Settings) -> int:
    """Integration periods tracking() runs: msToProcess for the 1 ms signals, round(msToProcess/1000/intTime)
    for Galileo E1 (GAL/GAL_E1C/include/tracking.m:48)."""
    import math
    if s.signal in ("GAL_E1C", "GPS_L2C", "BDS_B1C"):  # GPS_L2C/include/tracking.m:51, BDS/B1C/include/NB_tracking.m:49
        x = s.msToProcess / 1000 / s.intTime
        return int(math.floor(x + 0.5))
    return int(s.msToProcess)


def samples_per_code(s: Settings) -> int:
    """round(samplingFreq / (codeFreqBasis / codeLength)) — acquisition.m:116-117 (MATLAB round)."""
    x = s.samplingFreq / (s.codeFreqBasis / s.codeLength)
    import math
    return int(math.floor(x + 0.5))
