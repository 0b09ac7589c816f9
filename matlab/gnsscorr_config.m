function cfg = gnsscorr_config(settings)
%GNSSCORR_CONFIG  settings struct (initSettings.m) -> the field names of gc_config (gnsscorr.h).
cfg.device = 0;
cfg.file_type = settings.fileType;
cfg.sample_bytes = 1;
cfg.code_length = settings.codeLength;
cfg.acq_noncoh_time = settings.acqNonCohTime;
cfg.cno_vsm_interval = settings.CNo.VSMinterval;
cfg.skip_number_of_bytes = settings.skipNumberOfBytes;
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
