"""Seeded synthetic GPS L1 C/A IF records (8-bit complex, I,Q interleaved ``schar``).

The reference ships no sample data (README.md:10-11 points at an unreachable drive), so
every test and the benchmark run on records made here.  Signal model per satellite:
``A * D(t) * c((f_code*t + tau0) mod 1023) * exp(i*(2*pi*(IF+fd)*t + phi0))`` with
carrier-coherent code Doppler, 50 bps data bits aligned to code periods, plus complex
white Gaussian noise of ``sigma`` LSB per component, rounded and clipped to int8.

Two generators with the same model: :func:`make_record` (NumPy, CPU, for tests) and
:func:`make_record_torch` (any torch device, chunked, for the 60 s benchmark record).
The code generator used here is the package's own (:mod:`codes`), not the oracle's.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .codes import E1_SECONDARY, NH20, b3i_code, boc11, ca_code, glo_code

L1 = 1575.42e6


@dataclass
class Sat:
    prn: int
    doppler: float          # Hz
    code_phase: float       # chips at t=0 (signal code phase; receiver sees code start at (1023-cp) chips)
    cn0: float              # dB-Hz
    phi0: float = 0.0       # rad
    bit_seed: int = 0
    bit_offset: int = 0     # code periods (0..19) to the first bit edge


@dataclass
class Scene:
    fs: float = 16.368e6
    IF: float = 20e3
    sigma: float = 20.0
    seed: int = 20260101
    sats: list = field(default_factory=list)
    # GLONASS scenes: Sat.prn holds the frequency number K; every satellite uses the same 511-chip code
    # on carrier IF - freqSpacing*K (+ Doppler) as seen AFTER the reference's I/Q swap, data symbols are
    # 10 ms meander halves (+b, -b) of 20 ms bits.
    glonass: bool = False
    freqSpacing: float = 562.5e3
    # BeiDou B3I scenes: 10230-chip codes at 10.23 Mcps; PRN 6-58 carry the 20-bit Neumann-Hoffman
    # secondary code on 20 ms bits, GEO PRNs (1-5, 59-63) carry 2 ms bits.
    b3i: bool = False
    # Galileo E1 scenes: E1-B (250 sps data symbols, one per 4 ms code period) minus E1-C (25-chip secondary
    # code) on the same carrier, each a 4092-chip memory code with the BOC(1,1) sub-carrier; ``codes`` =
    # {PRN: (e1b, e1c)} +-1 primary chips (codes.load_e1_codes or codes.standin_e1_codes).
    e1c: bool = False
    codes: dict = None
    # GPS L5C / GAL E5a / GAL E5b / BDS B2a scenes (``fam5`` = the signal name): 10230-chip codes at 10.23 Mcps, the
    # data component in phase (one random symbol per 1 ms code period) and the pilot in quadrature carrying its
    # secondary code (NH20 for L5C, the PRN's 100-chip code otherwise); ``codes`` = {PRN: (data, pilot, secondary)}.
    fam5: str = ""
    # BDS B1I / GPS L2C scenes (``varb`` = the signal name): one caller-supplied code per SV, ``codes`` = {PRN: (code,)};
    # B1I: 2046 chips at 2.046 Mcps, 20 ms bits with the NH20 secondary code; L2C: the 20460-entry return-to-zero CM
    # sequence at 1.023 M entries/s (the CL slots stay empty), one data symbol per 20 ms code period.
    varb: str = ""


def default_scene(fs: float = 16.368e6, IF: float = 20e3, nsat: int = 8, seed: int = 20260101) -> Scene:
    rng = np.random.default_rng(seed)
    prns = rng.choice(np.arange(1, 33), size=nsat, replace=False)
    sats = []
    for i, p in enumerate(prns):
        sats.append(Sat(prn=int(p), doppler=float(rng.uniform(-5000, 5000)),
                        code_phase=float(rng.uniform(0, 1023)), cn0=float(rng.uniform(38, 50)),
                        phi0=float(rng.uniform(0, 2 * np.pi)), bit_seed=int(rng.integers(1 << 30)),
                        bit_offset=int(rng.integers(0, 20))))
    return Scene(fs=fs, IF=IF, seed=seed, sats=sats)


def _amp(cn0_db: float, sigma: float, fs: float) -> float:
    # complex noise power 2*sigma^2 over bandwidth fs  ->  N0 = 2*sigma^2/fs ; C = A^2
    return float(np.sqrt(10 ** (cn0_db / 10) * 2 * sigma * sigma / fs))


def nav_bits(sat: Sat, nbits: int) -> np.ndarray:
    return np.random.default_rng(sat.bit_seed).integers(0, 2, size=nbits).astype(np.float64) * 2 - 1


def default_scene_glo(fs: float = 12e6, IF: float = 0.0, nsat: int = 5, seed: int = 20260101, freqSpacing: float = 562.5e3) -> Scene:
    rng = np.random.default_rng(seed)
    ks = rng.choice(np.arange(-7, 7), size=nsat, replace=False)
    sats = [Sat(prn=int(k), doppler=float(rng.uniform(-4000, 4000)), code_phase=float(rng.uniform(0, 511)),
                cn0=float(rng.uniform(40, 50)), phi0=float(rng.uniform(0, 2 * np.pi)), bit_seed=int(rng.integers(1 << 30)),
                bit_offset=int(rng.integers(0, 20))) for k in ks]
    return Scene(fs=fs, IF=IF, seed=seed, sats=sats, glonass=True, freqSpacing=freqSpacing)


_NH20 = np.array([1, 1, 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1], dtype=np.float64)


def default_scene_b3i(fs: float = 18e6, IF: float = 20e3, nsat: int = 5, seed: int = 20260101) -> Scene:
    rng = np.random.default_rng(seed)
    prns = rng.choice(np.arange(1, 64), size=nsat, replace=False)
    sats = [Sat(prn=int(p), doppler=float(rng.uniform(-4000, 4000)), code_phase=float(rng.uniform(0, 10230)),
                cn0=float(rng.uniform(40, 50)), phi0=float(rng.uniform(0, 2 * np.pi)), bit_seed=int(rng.integers(1 << 30)),
                bit_offset=int(rng.integers(0, 20))) for p in prns]
    return Scene(fs=fs, IF=IF, seed=seed, sats=sats, b3i=True)


def default_scene_e1c(codes: dict, fs: float = 18e6, IF: float = 20e3, nsat: int = 4, seed: int = 20260101) -> Scene:
    rng = np.random.default_rng(seed)
    prns = rng.choice(np.arange(1, 37), size=nsat, replace=False)
    sats = [Sat(prn=int(p), doppler=float(rng.uniform(-4000, 4000)), code_phase=float(rng.uniform(0, 4092)),
                cn0=float(rng.uniform(42, 50)), phi0=float(rng.uniform(0, 2 * np.pi)), bit_seed=int(rng.integers(1 << 30)),
                bit_offset=int(rng.integers(0, 25))) for p in prns]
    return Scene(fs=fs, IF=IF, seed=seed, sats=sats, e1c=True, codes=codes)


def default_scene_fam5(signal: str, codes: dict, fs: float = 18e6, IF: float = 20e3, nsat: int = 4, seed: int = 20260101) -> Scene:
    rng = np.random.default_rng(seed)
    pool = np.arange(19, 31) if signal == "BDS_B2a" else np.arange(1, 33)
    prns = rng.choice(pool, size=nsat, replace=False)
    sats = [Sat(prn=int(p), doppler=float(rng.uniform(-4000, 4000)), code_phase=float(rng.uniform(0, 10230)),
                cn0=float(rng.uniform(42, 50)), phi0=float(rng.uniform(0, 2 * np.pi)), bit_seed=int(rng.integers(1 << 30)),
                bit_offset=int(rng.integers(0, 100))) for p in prns]
    return Scene(fs=fs, IF=IF, seed=seed, sats=sats, fam5=signal, codes=codes)


def default_scene_varb(signal: str, codes: dict, fs: float, IF: float = 20e3, nsat: int = 3, seed: int = 20260101) -> Scene:
    rng = np.random.default_rng(seed)
    pool = np.arange(6, 59) if signal == "BDS_B1I" else np.arange(1, 33)
    clen = 2046 if signal == "BDS_B1I" else 20460       # (B1C: 20460 BOC sub-chips per 10 ms)
    prns = rng.choice(pool, size=nsat, replace=False)
    sats = [Sat(prn=int(p), doppler=float(rng.uniform(-4000, 4000)), code_phase=float(rng.uniform(0, clen)),
                cn0=float(rng.uniform(42, 50)), phi0=float(rng.uniform(0, 2 * np.pi)), bit_seed=int(rng.integers(1 << 30)),
                bit_offset=int(rng.integers(0, 20))) for p in prns]
    return Scene(fs=fs, IF=IF, seed=seed, sats=sats, varb=signal, codes=codes)


def _fam5_secondary(scene: Scene, prn: int) -> np.ndarray:
    return NH20.astype(np.float64) if scene.fam5 == "GPS_L5C" else np.asarray(scene.codes[prn][2], dtype=np.float64)


def make_record(scene: Scene, nsamples: int, start: int = 0) -> np.ndarray:
    """int8 array of length 2*nsamples (I0,Q0,I1,Q1,...), samples start..start+nsamples-1."""
    n = np.arange(start, start + nsamples, dtype=np.float64)
    t = n / scene.fs
    sig = np.zeros(nsamples, dtype=np.complex128)
    for s in scene.sats:
        if scene.glonass:
            clen, crate, carrier = 511, 511e3, 1602e6
            fc = scene.IF - scene.freqSpacing * s.prn + s.doppler
            chipseq = glo_code().astype(np.float64)
        elif scene.b3i:
            clen, crate, carrier = 10230, 10.23e6, 1268.52e6
            fc = scene.IF + s.doppler
            chipseq = b3i_code(s.prn).astype(np.float64)
        elif scene.varb == "BDS_B1C":
            # data (amplitude sqrt(11/40)) and pilot (sqrt(29/40), in quadrature) BOC(1,1) sub-chip codes, 10 ms periods
            clen, crate, carrier = 20460, 2.046e6, 1575.42e6
            fc = scene.IF + s.doppler
            fcode = crate * (1 + s.doppler / carrier)
            chips = fcode * t + s.code_phase
            period = np.floor(chips / clen).astype(np.int64)
            idx = np.floor(chips - period * float(clen)).astype(np.int64) % clen
            cD = np.asarray(scene.codes[s.prn][0], dtype=np.float64)[idx]
            cP = np.asarray(scene.codes[s.prn][1], dtype=np.float64)[idx]
            bits = nav_bits(s, int(period.max()) + 60)
            dD = bits[period + s.bit_offset]
            dP = bits[::-1][period + s.bit_offset]
            ph = 2 * np.pi * (fc * t % 1.0) + s.phi0
            if len(scene.codes[s.prn]) > 2:
                # full-band signal: the pilot is QMBOC - BOC(1,1) in quadrature at sqrt(29/44), BOC(6,1) in phase at -sqrt(1/11) -
                # next to the data component at 1/2; the combination B1C WB_tracking.m:339-344 is built for
                idx6 = np.floor(6.0 * (chips - period * float(clen))).astype(np.int64) % (6 * clen)
                c61 = np.asarray(scene.codes[s.prn][2], dtype=np.float64)[idx6]
                sig += _amp(s.cn0, scene.sigma, scene.fs) * (0.5 * dD * cD + 1j * np.sqrt(29 / 44) * dP * cP - np.sqrt(1 / 11) * dP * c61) * np.exp(1j * ph)
                continue
            sig += _amp(s.cn0, scene.sigma, scene.fs) * (np.sqrt(11 / 40) * dD * cD + 1j * np.sqrt(29 / 40) * dP * cP) * np.exp(1j * ph)
            continue
        elif scene.varb:
            b1i = scene.varb == "BDS_B1I"
            clen, crate, carrier = (2046, 2.046e6, 1561.098e6) if b1i else (20460, 1.023e6, 1227.6e6)
            fc = scene.IF + s.doppler
            fcode = crate * (1 + s.doppler / carrier)
            chips = fcode * t + s.code_phase
            period = np.floor(chips / clen).astype(np.int64)
            idx = np.floor(chips - period * float(clen)).astype(np.int64) % clen
            code = np.asarray(scene.codes[s.prn][0], dtype=np.float64)[idx]
            if b1i:
                d = nav_bits(s, int(period.max() // 20) + 3)[(period + s.bit_offset) // 20] * _NH20[(period + s.bit_offset) % 20]
            else:
                d = nav_bits(s, int(period.max()) + 30)[period + s.bit_offset]
            ph = 2 * np.pi * (fc * t % 1.0) + s.phi0
            if not b1i and len(scene.codes[s.prn]) > 1:
                # L2C with its CL pilot: CM (data) and CL chips time-multiplexed - the CL sequence fills the half chips the
                # return-to-zero CM sequence leaves empty and runs over 75 CM periods; segment (bit_offset % 75) at period 0
                cl = np.asarray(scene.codes[s.prn][1], dtype=np.float64)
                idx_cl = (idx + ((period + s.bit_offset) % 75) * clen)
                code = d * code + cl[idx_cl]
                d = 1.0
            sig += _amp(s.cn0, scene.sigma, scene.fs) * (np.sqrt(2.0) if not b1i else 1.0) * d * code * np.exp(1j * ph)
            continue
        elif scene.fam5:
            clen, crate = 10230, 10.23e6
            carrier = 1207.14e6 if scene.fam5 == "GAL_E5b" else 1176.45e6
            fc = scene.IF + s.doppler
            fcode = crate * (1 + s.doppler / carrier)
            chips = fcode * t + s.code_phase
            period = np.floor(chips / clen).astype(np.int64)
            idx = np.floor(chips - period * float(clen)).astype(np.int64) % clen
            cD = np.asarray(scene.codes[s.prn][0], dtype=np.float64)[idx]
            cP = np.asarray(scene.codes[s.prn][1], dtype=np.float64)[idx]
            sec = _fam5_secondary(scene, s.prn)
            dD = nav_bits(s, int(period.max()) + 130)[period + s.bit_offset]
            dP = sec[(period + s.bit_offset) % sec.size]
            ph = 2 * np.pi * (fc * t % 1.0) + s.phi0
            sig += _amp(s.cn0, scene.sigma, scene.fs) * (dD * cD + 1j * dP * cP) / np.sqrt(2.0) * np.exp(1j * ph)
            continue
        elif scene.e1c:
            clen, crate, carrier = 4092, 1.023e6, L1
            fc = scene.IF + s.doppler
            fcode = crate * (1 + s.doppler / carrier)
            chips = fcode * t + s.code_phase
            period = np.floor(chips / clen).astype(np.int64)
            sub = np.floor(2.0 * (chips - period * float(clen))).astype(np.int64) % (2 * clen)
            cB = boc11(scene.codes[s.prn][0]).astype(np.float64)[sub]
            cC = boc11(scene.codes[s.prn][1]).astype(np.float64)[sub]
            dB = nav_bits(s, int(period.max()) + 30)[period + s.bit_offset]
            dC = E1_SECONDARY.astype(np.float64)[(period + s.bit_offset) % 25]
            ph = 2 * np.pi * (fc * t % 1.0) + s.phi0
            sig += _amp(s.cn0, scene.sigma, scene.fs) * (dB * cB - dC * cC) / np.sqrt(2.0) * np.exp(1j * ph)
            continue
        else:
            clen, crate, carrier = 1023, 1.023e6, L1
            fc = scene.IF + s.doppler
            chipseq = ca_code(s.prn).astype(np.float64)
        fcode = crate * (1 + s.doppler / carrier)
        chips = fcode * t + s.code_phase
        period = np.floor(chips / clen).astype(np.int64)
        idx = np.floor(chips - period * float(clen)).astype(np.int64) % clen
        code = chipseq[idx]
        bits = nav_bits(s, int(period.max() // 20) + 3)
        d = bits[(period + s.bit_offset) // 20]
        if scene.glonass:                       # meander: second 10 ms of every bit is inverted
            d = d * np.where(((period + s.bit_offset) % 20) < 10, 1.0, -1.0)
        if scene.b3i:
            if (1 <= s.prn <= 5) or (59 <= s.prn <= 63):     # GEO: 2 ms bits
                gbits = nav_bits(s, int(period.max() // 2) + 12)
                d = gbits[(period + s.bit_offset) // 2]
            else:                                             # NH secondary code on every 20 ms bit
                d = d * _NH20[(period + s.bit_offset) % 20]
        ph = 2 * np.pi * (fc * t % 1.0) + s.phi0
        sig += _amp(s.cn0, scene.sigma, scene.fs) * d * code * np.exp(1j * ph)
    # noise is a function of (seed, absolute chunk) so records can be made piecewise
    rng = np.random.default_rng([scene.seed, start])
    sig += scene.sigma * (rng.standard_normal(nsamples) + 1j * rng.standard_normal(nsamples))
    out = np.empty(2 * nsamples, dtype=np.int8)
    re, im = (sig.imag, sig.real) if scene.glonass else (sig.real, sig.imag)   # GLONASS files are read as Q + 1i*I
    out[0::2] = np.clip(np.rint(re), -127, 127).astype(np.int8)
    out[1::2] = np.clip(np.rint(im), -127, 127).astype(np.int8)
    return out


def make_record_torch(scene: Scene, nsamples: int, device="cuda", chunk: int = 1 << 24):
    """Same model on a torch device, generated in chunks; returns an int8 tensor (2*nsamples).

    Noise comes from torch's generator (seeded), so the bytes differ from
    :func:`make_record`; parity tests always feed the *same bytes* to both sides."""
    import torch

    out = torch.empty(2 * nsamples, dtype=torch.int8, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(scene.seed)
    if scene.glonass:
        clen, crate, carrier = 511, 511e3, 1602e6
        codes = {s.prn: torch.tensor(glo_code().astype(np.float32), device=device) for s in scene.sats}
    elif scene.e1c:
        clen, crate, carrier = 4092, 1.023e6, L1
        codes = {s.prn: torch.tensor(boc11(scene.codes[s.prn][0]).astype(np.float32), device=device) for s in scene.sats}
        pcodes = {s.prn: torch.tensor(boc11(scene.codes[s.prn][1]).astype(np.float32), device=device) for s in scene.sats}
        sec = torch.tensor(E1_SECONDARY.astype(np.float32), device=device)
    else:
        clen, crate, carrier = 1023, 1.023e6, L1
        codes = {s.prn: torch.tensor(ca_code(s.prn).astype(np.float32), device=device) for s in scene.sats}
    bits = {s.prn: torch.tensor(nav_bits(s, int(nsamples / scene.fs * (250 if scene.e1c else 50)) + 30).astype(np.float32),
                                device=device) for s in scene.sats}
    for c0 in range(0, nsamples, chunk):
        m = min(chunk, nsamples - c0)
        n = torch.arange(c0, c0 + m, dtype=torch.float64, device=device)
        t = n / scene.fs
        re = torch.zeros(m, dtype=torch.float32, device=device)
        im = torch.zeros(m, dtype=torch.float32, device=device)
        for s in scene.sats:
            fcode = crate * (1 + s.doppler / carrier)
            chips = fcode * t + s.code_phase
            period = torch.floor(chips / clen)
            idx = torch.floor(chips - period * float(clen)).to(torch.int64) % clen
            pint = period.to(torch.int64) + s.bit_offset
            if scene.e1c:
                sub = torch.floor(2.0 * (chips - period * float(clen))).to(torch.int64) % (2 * clen)
                a = _amp(s.cn0, scene.sigma, scene.fs) * 0.7071067811865476 * \
                    (bits[s.prn][pint] * codes[s.prn][sub] - sec[pint % 25] * pcodes[s.prn][sub])
                ph = (2 * np.pi) * torch.frac((scene.IF + s.doppler) * t) + s.phi0
                re += a * torch.cos(ph).to(torch.float32)
                im += a * torch.sin(ph).to(torch.float32)
                continue
            d = bits[s.prn][pint // 20]
            if scene.glonass:
                d = d * torch.where((pint % 20) < 10, 1.0, -1.0).to(torch.float32)
            a = _amp(s.cn0, scene.sigma, scene.fs) * d * codes[s.prn][idx]
            fc = (scene.IF - scene.freqSpacing * s.prn + s.doppler) if scene.glonass else (scene.IF + s.doppler)
            ph = (2 * np.pi) * torch.frac(fc * t) + s.phi0
            re += a * torch.cos(ph).to(torch.float32)
            im += a * torch.sin(ph).to(torch.float32)
        re += scene.sigma * torch.randn(m, device=device, generator=g)
        im += scene.sigma * torch.randn(m, device=device, generator=g)
        if scene.glonass:
            re, im = im, re                     # GLONASS files are read as Q + 1i*I
        out[2 * c0: 2 * (c0 + m): 2] = torch.clamp(torch.round(re), -127, 127).to(torch.int8)
        out[2 * c0 + 1: 2 * (c0 + m): 2] = torch.clamp(torch.round(im), -127, 127).to(torch.int8)
    return out


def pack_cplx2(x: np.ndarray) -> np.ndarray:
    """Complex samples with I, Q in {+-1, +-3} -> the 2-bit packed bytes that include/unpack_cplx.m reads (two samples per
    byte; per component a sign bit and a magnitude bit: first sample I = bits (0, 2), Q = bits (1, 3), second sample I = bits
    (4, 6), Q = bits (5, 7); its four 256-entry tables, unpack_cplx.m:17-20, are exactly that)."""
    x = np.asarray(x)
    re, im = np.rint(x.real).astype(np.int64), np.rint(x.imag).astype(np.int64)
    if not (np.all(np.isin(np.abs(re), (1, 3))) and np.all(np.isin(np.abs(im), (1, 3))) and np.all(re == x.real) and np.all(im == x.imag)):
        from .engine import GnssCorrError
        raise GnssCorrError("2-bit packed records hold I, Q in {+-1, +-3}")
    if x.size % 2:
        re, im = np.append(re, 1), np.append(im, 1)
    nib = (re < 0).astype(np.uint8) | ((im < 0).astype(np.uint8) << 1) | ((np.abs(re) == 3).astype(np.uint8) << 2) | ((np.abs(im) == 3).astype(np.uint8) << 3)
    return (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)


def quantize2(iq8: np.ndarray, step: float = 20.0) -> np.ndarray:
    """An int8 I,Q record quantised to the four levels of a 2-bit front end (+-1 within +-step, +-3 beyond), as complex."""
    v = iq8.astype(np.float64)
    q = np.where(np.abs(v) > step, 3.0, 1.0) * np.where(v < 0, -1.0, 1.0)
    return q[0::2] + 1j * q[1::2]
