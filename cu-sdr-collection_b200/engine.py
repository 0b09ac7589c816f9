"""ctypes binding of ``libgnsscorr.so`` (the C ABI in ``include/gnsscorr.h``) and the ``Engine``
object the MATLAB-mirroring functions sit on."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .settings import Settings, varb_step

_HERE = os.path.dirname(os.path.abspath(__file__))
GC_TRACK_NFIELDS = 15


class GnssCorrError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(_HERE, "libgnsscorr.so")


class gc_config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("signal", C.c_int32),
                ("file_type", C.c_int32), ("sample_bytes", C.c_int32), ("code_length", C.c_int32),
                ("acq_noncoh_time", C.c_int32), ("cno_vsm_interval", C.c_int32),
                ("skip_number_of_bytes", C.c_int64),
                ("sampling_freq", C.c_double), ("IF", C.c_double), ("code_freq_basis", C.c_double),
                ("acq_search_band", C.c_double), ("acq_search_step", C.c_double), ("acq_threshold", C.c_double),
                ("dll_damping_ratio", C.c_double), ("dll_noise_bandwidth", C.c_double),
                ("dll_correlator_spacing", C.c_double), ("pll_damping_ratio", C.c_double),
                ("pll_noise_bandwidth", C.c_double), ("int_time", C.c_double), ("cno_acc_time", C.c_double),
                ("freq_spacing", C.c_double), ("pilot_trk_flag", C.c_int32), ("acq_coh_t", C.c_int32),
                ("pilot_acq_flag", C.c_int32), ("reserved1", C.c_int32), ("carr_freq_basis", C.c_double)]


GC_SIG_GPS_L1CA, GC_SIG_GLO_G1G2, GC_SIG_BDS_B3I, GC_SIG_GAL_E1C = 0, 1, 2, 3
GC_SIG_GPS_L5C, GC_SIG_GAL_E5A, GC_SIG_GAL_E5B, GC_SIG_BDS_B2A = 4, 5, 6, 7
GC_SIG_BDS_B1I, GC_SIG_GPS_L2C, GC_SIG_BDS_B1C = 8, 9, 10
_FAM5_IDS = {"GPS_L5C": GC_SIG_GPS_L5C, "GAL_E5a": GC_SIG_GAL_E5A, "GAL_E5b": GC_SIG_GAL_E5B, "BDS_B2a": GC_SIG_BDS_B2A,
             "BDS_B1I": GC_SIG_BDS_B1I, "GPS_L2C": GC_SIG_GPS_L2C, "BDS_B1C": GC_SIG_BDS_B1C}
GC_SV_NONE = -2147483648
GC_PARAM_B1C_WB_FACTOR = 1
GC_PARAM_TRACK_EXACT_SUMS = 2     # float64 checking mode of the tracking kernel (include/gnsscorr.h)
GC_PARAM_TRACK_FAST_DISC = 3      # fp32 discriminators (default: float64)


class gc_stats(C.Structure):
    _fields_ = [("acq_total_ms", C.c_float), ("acq_fwd_ms", C.c_float), ("acq_corr_ms", C.c_float),
                ("acq_fine_ms", C.c_float), ("track_kernel_ms", C.c_float),
                ("acq_launches", C.c_int32), ("track_launches", C.c_int32), ("fft_len", C.c_int32),
                ("acq_path", C.c_int32), ("n_acquired", C.c_int32),
                ("corr_rows_ms", C.c_float), ("corr_cols_ms", C.c_float), ("corr_row_launches", C.c_int32)]


EXPORTS = ["gc_abi_version", "gc_build_arch", "gc_acq_result_len", "gc_create", "gc_destroy",
           "gc_last_error", "gc_set_code", "gc_set_record_host", "gc_set_record_device", "gc_acquire",
           "gc_acquire_host", "gc_track_nfields", "gc_track", "gc_track_file", "gc_get_stats", "gc_get_stream",
           "gc_set_param", "gc_get_cl_code_phase", "gc_set_cl_code_phase", "gc_nav_sync", "gc_acquire_track",
           "gc_acquire_device", "gc_acquire_device_async", "gc_get_cno_pld", "gc_code_entries", "gc_generate_code", "gc_generate_code_device", "gc_multi_create", "gc_multi_destroy", "gc_multi_last_error", "gc_multi_n_gpus",
           "gc_multi_handle", "gc_multi_set_code", "gc_multi_set_param", "gc_multi_set_cl_code_phase",
           "gc_multi_get_cl_code_phase", "gc_multi_set_record_host", "gc_multi_acquire", "gc_multi_acquire_host",
           "gc_multi_track", "gc_multi_track_file", "gc_multi_get_times"]

_lib = None


def load_lib():
    """Load the CUDA library; there is no fallback, so a missing build is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise GnssCorrError(f"{p} is missing - build it with __graft_entry__.build() "
                            "(make -C cu-sdr-collection_b200/csrc); there is no CPU fallback")
    lib = C.CDLL(p)
    vp, i32p, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.gc_abi_version.restype = C.c_int
    lib.gc_build_arch.restype = C.c_char_p
    lib.gc_acq_result_len.argtypes = [C.c_int32]
    lib.gc_create.argtypes = [C.POINTER(vp), C.POINTER(gc_config)]
    lib.gc_destroy.argtypes = [vp]
    lib.gc_destroy.restype = None
    lib.gc_last_error.argtypes = [vp]
    lib.gc_last_error.restype = C.c_char_p
    lib.gc_set_code.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int32]
    lib.gc_set_record_host.argtypes = [vp, vp, C.c_size_t]
    lib.gc_set_record_device.argtypes = [vp, vp, C.c_size_t]
    lib.gc_acquire.argtypes = [vp, C.c_int32, i32p, dp, dp, dp, i32p, i32p]
    lib.gc_acquire_host.argtypes = [vp, vp, C.c_size_t, C.c_int32, i32p, dp, dp, dp, i32p, i32p]
    lib.gc_track_nfields.argtypes = [vp]
    lib.gc_set_param.argtypes = [vp, C.c_int32, C.c_double]
    lib.gc_get_cl_code_phase.argtypes = [vp, i32p]
    lib.gc_set_cl_code_phase.argtypes = [vp, C.c_int32, i32p]
    lib.gc_track.argtypes = [vp, C.c_int32, i32p, dp, dp, dp, C.c_int32, dp, dp, dp, i32p]
    lib.gc_track_file.argtypes = [vp, C.c_char_p, C.c_int32, i32p, dp, dp, dp, C.c_int32, dp, dp, dp, i32p]
    lib.gc_get_stats.argtypes = [vp, C.POINTER(gc_stats)]
    lib.gc_acquire_track.argtypes = [vp, C.c_int32, i32p, C.c_int32, C.c_int32, dp, dp, dp, i32p, dp, dp, dp, dp, dp, i32p]
    lib.gc_nav_sync.argtypes = [vp, C.c_int32, C.c_int32, dp, i32p, C.POINTER(C.c_uint8), i32p]
    lib.gc_get_stream.argtypes = [vp]
    lib.gc_get_stream.restype = C.c_void_p
    lib.gc_acquire_device.argtypes = [vp, C.c_int32, i32p, vp]
    lib.gc_acquire_device_async.argtypes = [vp, C.c_int32, i32p, vp]
    lib.gc_get_cno_pld.argtypes = [vp, C.c_int32, C.c_int32, dp]
    lib.gc_code_entries.argtypes = [C.c_int32, C.c_int32]
    lib.gc_generate_code.argtypes = [C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32]
    lib.gc_generate_code_device.argtypes = [C.c_int32, C.c_int32, C.c_int32, i32p, C.c_int32, vp]
    lib.gc_multi_create.argtypes = [C.POINTER(vp), C.POINTER(gc_config), C.c_int32]
    lib.gc_multi_destroy.argtypes = [vp]
    lib.gc_multi_destroy.restype = None
    lib.gc_multi_last_error.argtypes = [vp]
    lib.gc_multi_last_error.restype = C.c_char_p
    lib.gc_multi_n_gpus.argtypes = [vp]
    lib.gc_multi_handle.argtypes = [vp, C.c_int32]
    lib.gc_multi_handle.restype = vp
    lib.gc_multi_set_code.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int32]
    lib.gc_multi_set_param.argtypes = [vp, C.c_int32, C.c_double]
    lib.gc_multi_set_cl_code_phase.argtypes = [vp, C.c_int32, i32p]
    lib.gc_multi_get_cl_code_phase.argtypes = [vp, i32p]
    lib.gc_multi_set_record_host.argtypes = [vp, vp, C.c_size_t]
    lib.gc_multi_acquire.argtypes = [vp, C.c_int32, i32p, dp, dp, dp, i32p, i32p]
    lib.gc_multi_acquire_host.argtypes = [vp, vp, C.c_size_t, C.c_int32, i32p, dp, dp, dp, i32p, i32p]
    lib.gc_multi_track.argtypes = [vp, C.c_int32, i32p, dp, dp, dp, C.c_int32, dp, dp, dp, i32p]
    lib.gc_multi_track_file.argtypes = [vp, C.c_char_p, C.c_int32, i32p, dp, dp, dp, C.c_int32, dp, dp, dp, i32p]
    lib.gc_multi_get_times.argtypes = [vp, dp, dp]
    _lib = lib
    return lib


def signal_id(s: Settings) -> int:
    if s.signal in _FAM5_IDS:
        return _FAM5_IDS[s.signal]
    return (GC_SIG_GLO_G1G2 if s.is_glonass else GC_SIG_BDS_B3I if s.signal == "BDS_B3I" else
            GC_SIG_GAL_E1C if s.signal == "GAL_E1C" else GC_SIG_GPS_L1CA)


def config_from_settings(s: Settings, device: int = 0) -> gc_config:
    if s.resamplingflag != 0:   # (B3I spells it resamplingFlag, BDS/B3I/initSettings.m:88)
        raise GnssCorrError("resamplingflag == 1 is outside the accelerated path "
                            "(acquisition.m:50-111); run the reference for that case")
    if s.fileType not in (1, 2, 3) or s.dataType not in ("schar", "int16") or (s.fileType == 3 and s.dataType != "schar"):
        raise GnssCorrError("fileType must be 1 (real), 2 (I/Q) or 3 (2-bit packed I/Q, unpack_cplx.m) and dataType 'schar' or 'int16' (initSettings.m:63-68)")
    sig = signal_id(s)
    return gc_config(abi_version=5, device=device, carr_freq_basis=float(getattr(s, 'carrFreqBasis', 0.0) or 0.0), pilot_trk_flag=int(s.pilotTRKflag), acq_coh_t=int(s.acqCohT),
                     pilot_acq_flag=int(s.pilotACQflag), signal=sig, freq_spacing=float(s.freqSpacing),
                     file_type=s.fileType, sample_bytes=2 if s.dataType == "int16" else 1,
                     code_length=int(s.codeLength), acq_noncoh_time=int(s.acqNonCohTime),
                     cno_vsm_interval=int(s.CNo_VSMinterval), skip_number_of_bytes=int(s.skipNumberOfBytes),
                     sampling_freq=s.samplingFreq, IF=s.IF, code_freq_basis=s.codeFreqBasis,
                     acq_search_band=s.acqSearchBand, acq_search_step=varb_step(s) if s.is_varb else s.acqStep if s.signal == "BDS_B1C" else s.acqSearchStep,
                     acq_threshold=s.acqThreshold, dll_damping_ratio=s.dllDampingRatio,
                     dll_noise_bandwidth=s.dllNoiseBandwidth, dll_correlator_spacing=s.dllCorrelatorSpacing,
                     pll_damping_ratio=s.pllDampingRatio, pll_noise_bandwidth=s.pllNoiseBandwidth,
                     int_time=s.intTime, cno_acc_time=s.CNo_accTime)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Engine:
    """One engine = one GPU: owns the FFT plan, replica spectra and the resident IF record."""

    def __init__(self, settings: Settings, device: int = 0, codes: dict | None = None):
        """``codes``: Galileo E1 only - {PRN: (e1b, e1c)} +-1 primary chips (codes.load_e1_codes /
        codes.standin_e1_codes); default: read E1b.dat / E1c.dat from ``settings.codeDir`` as the reference
        does from its include/ folder (generateE1Bcode.m:44)."""
        self.lib = load_lib()
        self.settings = settings
        self._h = C.c_void_p()
        cfg = config_from_settings(settings, device)
        rc = self.lib.gc_create(C.byref(self._h), C.byref(cfg))
        if rc != 0:
            raise GnssCorrError(f"gc_create failed ({rc}): {self.lib.gc_last_error(None).decode()}")
        self._keep = None
        # codes= overrides the library's own generators (gc_generate_code: the reference's generate*code.m as device kernels);
        # without it every code is generated inside the library the first time an SV is used
        if settings.signal == "GAL_E1C" and codes is None and settings.codeDir:
            from .codes import load_e1_codes
            codes = load_e1_codes(settings.codeDir)
        if codes is not None and (settings.is_fam5 or settings.is_varb or settings.signal in ("BDS_B1C", "GAL_E1C")):
            self.set_codes(codes)

    def set_codes(self, codes: dict):
        nmax = self.lib.gc_acq_result_len(signal_id(self.settings))
        for prn, comps in codes.items():
            if prn > nmax:
                continue
            for comp, chips in enumerate(comps):
                sig = self.settings.signal
                if comp == 2 and not (sig == "GAL_E5a" or (sig == "BDS_B1C" and int(self.settings.pilotTRKflag) == 2)):
                    continue                                     # E5a: per-PRN secondary code; B1C full band: pilot BOC(6,1)
                if comp >= 1 and self.settings.is_varb and not (sig == "GPS_L2C" and comp == 1 and int(self.settings.pilotTRKflag) == 1):
                    continue                                     # B1I / L2C: one code per SV (+ the CL sequence with the L2C pilot)
                a = np.ascontiguousarray(chips, dtype=np.int8)
                self._check(self.lib.gc_set_code(self._h, int(prn), comp, a.ctypes.data, a.size), "gc_set_code")

    @property
    def sample_dtype(self):
        """numpy dtype of one stored value of the record: settings.dataType 'schar' / 'int16' (initSettings.m:63)."""
        if self.settings.fileType == 3:
            return np.uint8                                          # 2-bit packed: two complex samples per byte
        return np.int16 if self.settings.dataType == "int16" else np.int8

    def close(self):
        if self._h:
            self.lib.gc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise GnssCorrError(f"{what} failed ({rc}): {self.lib.gc_last_error(self._h).decode()}")

    # ---- record -------------------------------------------------------------------------
    def set_record(self, data):
        """Make an IF record resident.  ``data``: int8 numpy array (copied host->device) or an
        int8 CUDA torch tensor (adopted without a copy; kept alive by the engine)."""
        if isinstance(data, np.ndarray):
            a = np.ascontiguousarray(data, dtype=self.sample_dtype)
            self._check(self.lib.gc_set_record_host(self._h, a.ctypes.data, a.nbytes), "gc_set_record_host")
            self._keep = None
        else:  # torch tensor on the GPU
            assert data.is_cuda and data.element_size() == np.dtype(self.sample_dtype).itemsize and data.is_contiguous()
            self._check(self.lib.gc_set_record_device(self._h, data.data_ptr(), data.numel() * data.element_size()), "gc_set_record_device")
            self._keep = data

    # ---- acquisition --------------------------------------------------------------------
    def acquire(self, sv_list=None, host_iq=None):
        s = self.settings
        sv = np.asarray(list(sv_list if sv_list is not None else s.acqSatelliteList), dtype=np.int32)
        n = self.lib.gc_acq_result_len(signal_id(s))
        carr, cph, pm = np.zeros(n), np.zeros(n), np.zeros(n)
        cbin, ccp = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        if host_iq is not None:
            a = np.ascontiguousarray(host_iq, dtype=self.sample_dtype)       # the file's own sample format
            rc = self.lib.gc_acquire_host(self._h, a.ctypes.data, {1: a.size, 2: a.size // 2, 3: a.size * 2}[s.fileType], sv.size, _ip(sv),
                                          _dp(carr), _dp(cph), _dp(pm), _ip(cbin), _ip(ccp))
            self._check(rc, "gc_acquire_host")
        else:
            rc = self.lib.gc_acquire(self._h, sv.size, _ip(sv), _dp(carr), _dp(cph), _dp(pm), _ip(cbin), _ip(ccp))
            self._check(rc, "gc_acquire")
        r = dict(carrFreq=carr, codePhase=cph, peakMetric=pm, coarseBin=cbin, coarseCodePhase=ccp)
        if s.signal == "GPS_L2C" and int(s.pilotTRKflag) == 1:       # acqResults.CLCodePhase (GPS_L2C acquisition.m:136)
            cl = np.zeros(32, dtype=np.int32)
            self._check(self.lib.gc_get_cl_code_phase(self._h, _ip(cl)), "gc_get_cl_code_phase")
            r["CLCodePhase"] = cl
        return r

    def acquire_device(self, sv_list, d_results):
        """gc_acquire_device: the search on the resident record with acqResults left on the GPU in ``d_results`` - a CUDA
        float64 tensor of 4 * gc_acq_result_len elements, [peakMetric | codePhase | carrFreq | coarseBin] - ready for the
        NCCL all-gather of a one-process-per-GPU run (entries of SVs not in ``sv_list`` are zero)."""
        sv = np.asarray(list(sv_list), dtype=np.int32)
        n = self.lib.gc_acq_result_len(signal_id(self.settings))
        assert d_results.is_cuda and d_results.is_contiguous() and d_results.numel() == 4 * n and d_results.element_size() == 8
        self._check(self.lib.gc_acquire_device(self._h, sv.size, _ip(sv), d_results.data_ptr()), "gc_acquire_device")
        return d_results

    def acquire_device_async(self, sv_list, d_results):
        """gc_acquire_device_async: as ``acquire_device`` but the call returns once the search is enqueued; ``d_results`` is
        complete in stream order on the engine's stream (``stream_ptr``), so a collective enqueued behind that stream follows
        without a host round trip."""
        sv = np.asarray(list(sv_list), dtype=np.int32)
        n = self.lib.gc_acq_result_len(signal_id(self.settings))
        assert d_results.is_cuda and d_results.is_contiguous() and d_results.numel() == 4 * n and d_results.element_size() == 8
        self._check(self.lib.gc_acquire_device_async(self._h, sv.size, _ip(sv), d_results.data_ptr()), "gc_acquire_device_async")
        return d_results

    # ---- tracking -----------------------------------------------------------------------
    def set_param(self, key: int, value: float):
        self._check(self.lib.gc_set_param(self._h, int(key), float(value)), "gc_set_param")

    def track(self, prn, acq_freq, code_phase, n_epochs, path=None, code_freq0=None, cl_code_phase=None):
        prn = np.asarray(prn, dtype=np.int32)
        if cl_code_phase is not None:                                # channel.CLCodePhase (GPS_L2C tracking.m:162)
            cl = np.ascontiguousarray(cl_code_phase, dtype=np.int32)
            self._check(self.lib.gc_set_cl_code_phase(self._h, cl.size, _ip(cl)), "gc_set_cl_code_phase")
        af = np.asarray(acq_freq, dtype=np.float64)
        cp = np.asarray(code_phase, dtype=np.float64)
        cf0 = None if code_freq0 is None else np.ascontiguousarray(code_freq0, dtype=np.float64)
        cf0p = _dp(cf0) if cf0 is not None else None
        nch = prn.size
        nv = n_epochs // int(self.settings.CNo_VSMinterval)
        out = np.empty((nch, int(self.lib.gc_track_nfields(self._h)), n_epochs))
        vv, vi = np.zeros((nch, nv)), np.zeros((nch, nv))
        done = np.zeros(nch, dtype=np.int32)
        if path is not None:
            rc = self.lib.gc_track_file(self._h, os.fsencode(path), nch, _ip(prn), _dp(af), _dp(cp), cf0p, n_epochs,
                                        _dp(out), _dp(vv), _dp(vi), _ip(done))
            self._check(rc, "gc_track_file")
        else:
            rc = self.lib.gc_track(self._h, nch, _ip(prn), _dp(af), _dp(cp), cf0p, n_epochs,
                                   _dp(out), _dp(vv), _dp(vi), _ip(done))
            self._check(rc, "gc_track")
        return out, vv, vi, done

    def acquire_track(self, n_channels: int, n_epochs: int, sv_list=None):
        """acquisition -> preRun -> tracking on the resident record in one library call (GPS L1 C/A).  Returns
        ``(acqResults, channel, out, vsmValue, vsmIndex, epochsDone)``."""
        s = self.settings
        sv = np.asarray(list(sv_list if sv_list is not None else s.acqSatelliteList), dtype=np.int32)
        n = self.lib.gc_acq_result_len(signal_id(s))
        carr, cph, pm = np.zeros(n), np.zeros(n), np.zeros(n)
        csv = np.zeros(n_channels, dtype=np.int32)
        caf, ccp = np.zeros(n_channels), np.zeros(n_channels)
        nv = n_epochs // int(s.CNo_VSMinterval)
        out = np.empty((n_channels, int(self.lib.gc_track_nfields(self._h)), n_epochs))
        vv, vi = np.zeros((n_channels, nv)), np.zeros((n_channels, nv))
        done = np.zeros(n_channels, dtype=np.int32)
        rc = self.lib.gc_acquire_track(self._h, sv.size, _ip(sv), n_channels, n_epochs, _dp(carr), _dp(cph), _dp(pm),
                                       _ip(csv), _dp(caf), _dp(ccp), _dp(out), _dp(vv), _dp(vi), _ip(done))
        self._check(rc, "gc_acquire_track")
        if s.is_glonass:
            channel = [dict(K=int(csv[i]) if csv[i] != GC_SV_NONE else 0, acquiredFreq=float(caf[i]), codePhase=int(ccp[i]),
                            status="T" if csv[i] != GC_SV_NONE else "-") for i in range(n_channels)]
        else:
            channel = [dict(PRN=int(csv[i]), acquiredFreq=float(caf[i]), codePhase=int(ccp[i]), status="T" if csv[i] else "-") for i in range(n_channels)]
        return dict(carrFreq=carr, codePhase=cph, peakMetric=pm), channel, out, vv, vi, done

    def cno_pld(self, n_channels: int, n_epochs: int):
        """DataCNo, DataPLD, PilotCNo, PilotPLD, total C/N0 of the last track() (BDS B2a / B1C; Calc_CNo_PLD.m on the device):
        array [n_channels][5][n_epochs // CNoInterval]."""
        nv = n_epochs // int(self.settings.CNo_VSMinterval)
        out = np.zeros((n_channels, 5, nv))
        self._check(self.lib.gc_get_cno_pld(self._h, n_channels, nv, _dp(out)), "gc_get_cno_pld")
        return out

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t the engine launches on (wrap with torch.cuda.ExternalStream to time it)."""
        return int(self.lib.gc_get_stream(self._h) or 0)

    def set_record_host_ptr(self, ptr: int, nbytes: int):
        """Host pointer variant (e.g. a pinned torch tensor's data_ptr): copies host->device."""
        self._check(self.lib.gc_set_record_host(self._h, ptr, nbytes), "gc_set_record_host")
        self._keep = None

    def stats(self) -> dict:
        st = gc_stats()
        self._check(self.lib.gc_get_stats(self._h, C.byref(st)), "gc_get_stats")
        return {k: getattr(st, k) for k, _ in gc_stats._fields_}


def codes_for_settings(settings: Settings, codes: dict | None):
    """The (PRN, component, chips) triples gc_set_code takes for this signal (Engine.set_codes / MultiEngine.set_codes)."""
    sig = settings.signal
    for prn, comps in (codes or {}).items():
        for comp, chips in enumerate(comps):
            if comp == 2 and not (sig == "GAL_E5a" or (sig == "BDS_B1C" and int(settings.pilotTRKflag) == 2)):
                continue
            if comp >= 1 and settings.is_varb and not (sig == "GPS_L2C" and comp == 1 and int(settings.pilotTRKflag) == 1):
                continue
            yield int(prn), comp, np.ascontiguousarray(chips, dtype=np.int8)


class MultiEngine:
    """Several GPUs behind one handle (gc_multi_*): the SV list is dealt round-robin and the channels in contiguous blocks
    over ``n_gpus`` B200s inside the library; same calls and results as ``Engine``."""

    def __init__(self, settings: Settings, n_gpus: int = 0, device: int = 0, codes: dict | None = None):
        self.lib = load_lib()
        self.settings = settings
        self._m = C.c_void_p()
        cfg = config_from_settings(settings, device)
        rc = self.lib.gc_multi_create(C.byref(self._m), C.byref(cfg), int(n_gpus))
        if rc != 0:
            raise GnssCorrError(f"gc_multi_create failed ({rc}): {self.lib.gc_multi_last_error(None).decode()}")
        if settings.signal == "GAL_E1C" and codes is None and settings.codeDir:
            from .codes import load_e1_codes
            codes = load_e1_codes(settings.codeDir)
        nmax = self.lib.gc_acq_result_len(signal_id(settings))
        for prn, comp, a in codes_for_settings(settings, codes):
            if prn <= nmax:
                self._check(self.lib.gc_multi_set_code(self._m, prn, comp, a.ctypes.data, a.size), "gc_multi_set_code")

    @property
    def n_gpus(self) -> int:
        return int(self.lib.gc_multi_n_gpus(self._m))

    def _check(self, rc, what):
        if rc != 0:
            raise GnssCorrError(f"{what} failed ({rc}): {self.lib.gc_multi_last_error(self._m).decode()}")

    def close(self):
        if self._m:
            self.lib.gc_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_param(self, key: int, value: float):
        self._check(self.lib.gc_multi_set_param(self._m, int(key), float(value)), "gc_multi_set_param")

    def set_record(self, data: np.ndarray):
        a = np.ascontiguousarray(data)
        self._check(self.lib.gc_multi_set_record_host(self._m, a.ctypes.data, a.nbytes), "gc_multi_set_record_host")

    def set_record_host_ptr(self, ptr: int, nbytes: int):
        self._check(self.lib.gc_multi_set_record_host(self._m, ptr, nbytes), "gc_multi_set_record_host")

    def acquire(self, sv_list=None, host_iq=None):
        s = self.settings
        sv = np.asarray(list(sv_list if sv_list is not None else s.acqSatelliteList), dtype=np.int32)
        n = self.lib.gc_acq_result_len(signal_id(s))
        carr, cph, pm = np.zeros(n), np.zeros(n), np.zeros(n)
        cbin, ccp = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        if host_iq is not None:
            a = np.ascontiguousarray(host_iq)
            ns = {1: a.size, 2: a.size // 2, 3: a.size * 2}[s.fileType]
            self._check(self.lib.gc_multi_acquire_host(self._m, a.ctypes.data, ns, sv.size, _ip(sv), _dp(carr), _dp(cph), _dp(pm),
                                                       _ip(cbin), _ip(ccp)), "gc_multi_acquire_host")
        else:
            self._check(self.lib.gc_multi_acquire(self._m, sv.size, _ip(sv), _dp(carr), _dp(cph), _dp(pm), _ip(cbin), _ip(ccp)),
                        "gc_multi_acquire")
        r = dict(carrFreq=carr, codePhase=cph, peakMetric=pm, coarseBin=cbin, coarseCodePhase=ccp)
        if s.signal == "GPS_L2C" and int(s.pilotTRKflag) == 1:
            cl = np.zeros(32, dtype=np.int32)
            self._check(self.lib.gc_multi_get_cl_code_phase(self._m, _ip(cl)), "gc_multi_get_cl_code_phase")
            r["CLCodePhase"] = cl
        return r

    def track(self, prn, acq_freq, code_phase, n_epochs, path=None, code_freq0=None, cl_code_phase=None):
        prn = np.asarray(prn, dtype=np.int32)
        if cl_code_phase is not None:
            cl = np.ascontiguousarray(cl_code_phase, dtype=np.int32)
            self._check(self.lib.gc_multi_set_cl_code_phase(self._m, cl.size, _ip(cl)), "gc_multi_set_cl_code_phase")
        af = np.asarray(acq_freq, dtype=np.float64)
        cp = np.asarray(code_phase, dtype=np.float64)
        cf0 = None if code_freq0 is None else np.ascontiguousarray(code_freq0, dtype=np.float64)
        nch = prn.size
        nv = n_epochs // int(self.settings.CNo_VSMinterval)
        h0 = self.lib.gc_multi_handle(self._m, 0)
        out = np.empty((nch, int(self.lib.gc_track_nfields(h0)), n_epochs))
        vv, vi = np.zeros((nch, nv)), np.zeros((nch, nv))
        done = np.zeros(nch, dtype=np.int32)
        args = (nch, _ip(prn), _dp(af), _dp(cp), _dp(cf0) if cf0 is not None else None, n_epochs, _dp(out), _dp(vv), _dp(vi), _ip(done))
        if path is not None:
            self._check(self.lib.gc_multi_track_file(self._m, os.fsencode(path), *args), "gc_multi_track_file")
        else:
            self._check(self.lib.gc_multi_track(self._m, *args), "gc_multi_track")
        return out, vv, vi, done

    def cno_pld(self, n_channels: int, n_epochs: int):
        """As Engine.cno_pld: every GPU holds the rows of its own block of channels, in channel order."""
        nv = n_epochs // int(self.settings.CNo_VSMinterval)
        out = np.zeros((n_channels, 5, nv))
        per = (n_channels + self.n_gpus - 1) // self.n_gpus
        for g in range(self.n_gpus):
            c0 = g * per
            nc = min(per, n_channels - c0)
            if nc <= 0:
                break
            part = np.zeros((nc, 5, nv))
            rc = self.lib.gc_get_cno_pld(self.lib.gc_multi_handle(self._m, g), nc, nv, _dp(part))
            if rc != 0:
                raise GnssCorrError(f"gc_get_cno_pld failed ({rc}) on GPU {g}")
            out[c0:c0 + nc] = part
        return out

    def times(self):
        a, t = C.c_double(), C.c_double()
        self._check(self.lib.gc_multi_get_times(self._m, C.byref(a), C.byref(t)), "gc_multi_get_times")
        return dict(acq_ms=a.value, track_ms=t.value)


def generate_code(signal: str, sv: int, component: int = 0, device: int | None = None) -> np.ndarray:
    """The library's own primary-code generators (``gc_generate_code``; ``device`` given: the same generators as a CUDA kernel,
    ``gc_generate_code_device``).  ``signal``: a settings.signal name; returns the int8 entries in gc_set_code's layout."""
    lib = load_lib()
    sid = signal_id(Settings(signal=signal)) if not isinstance(signal, int) else int(signal)
    n = lib.gc_code_entries(sid, component)
    if n <= 0:
        raise GnssCorrError(f"{signal} has no code component {component}")
    out = np.zeros(n, dtype=np.int8)
    if device is None:
        rc = lib.gc_generate_code(sid, int(sv), component, out.ctypes.data, n)
    else:
        svl = np.asarray([sv], dtype=np.int32)
        rc = lib.gc_generate_code_device(int(device), sid, 1, _ip(svl), component, out.ctypes.data)
    if rc != n:
        raise GnssCorrError(f"gc_generate_code({signal}, {sv}, {component}) failed ({rc})")
    return out
