function cfg = gnsscorr_config(settings)
%GNSSCORR_CONFIG  settings struct (initSettings.m) -> the field names of gc_config (gnsscorr.h).
cfg.device = 0;
if isfield(settings, 'freqSpacing')      % GLO/GLO_GL1, GLO/GLO_GL2
    cfg.signal = 1;  cfg.freq_spacing = settings.freqSpacing;
else                                    % GPS/GPS_L1CA
    cfg.signal = 0;  cfg.freq_spacing = 0;
end
cfg.file_type = settings.fileType;
cfg.sample_bytes = 1;
cfg.code_length = settings.codeLength;
cfg.acq_noncoh_time = settings.acqNonCohTime;
cfg.cno_vsm_interval = settings.CNo.VSMinterval;
if isfield(settings, 'skipNumberOfSamples')   % the GLONASS folders' name for the same offset
    cfg.skip_number_of_bytes = settings.skipNumberOfSamples;
else
    cfg.skip_number_of_bytes = settings.skipNumberOfBytes;
end
cfg.sampling_freq = settings.samplingFreq;
cfg.IF = settings.IF;
cfg.code_freq_basis = settings.codeFreqBasis;
cfg.acq_search_band = settings.acqSearchBand;
cfg.acq_search_step = settings.acqSearchStep;
cfg.acq_threshold = settings.acqThreshold;
cfg.dll_damping_ratio = settings.dllDampingRatio;
cfg.dll_noise_bandwidth = settings.dllNoiseBandwidth;
cfg.dll_correlator_spacing = settings.dllCorrelatorSpacing;
cfg.pll_damping_ratio = settings.pllDampingRatio;
cfg.pll_noise_bandwidth = settings.pllNoiseBandwidth;
cfg.int_time = settings.intTime;
cfg.cno_acc_time = settings.CNo.accTime;
end
