"""``settings`` struct of the reference (GPS/GPS_L1CA/initSettings.m:44-136), hot-path fields.

Field names are the reference's own so a MATLAB user finds every knob where it was; the nested
``settings.CNo.*`` struct is flattened to ``CNo_accTime`` / ``CNo_VSMinterval``.
"""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class Settings:
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


def init_settings(**overrides) -> Settings:
    """``settings = initSettings()`` (GPS/GPS_L1CA/init.m:56) with optional field overrides."""
    s = Settings()
    for k, v in overrides.items():
        if not hasattr(s, k):
            raise AttributeError(f"settings has no field {k!r}")
        setattr(s, k, v)
    return s


def samples_per_code(s: Settings) -> int:
    """round(samplingFreq / (codeFreqBasis / codeLength)) — acquisition.m:116-117 (MATLAB round)."""
    x = s.samplingFreq / (s.codeFreqBasis / s.codeLength)
    import math
    return int(math.floor(x + 0.5))
