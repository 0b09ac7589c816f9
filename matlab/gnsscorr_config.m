function cfg = gnsscorr_config(settings, signal)
%GNSSCORR_CONFIG  settings struct (initSettings.m) -> the field names of gc_config (gnsscorr.h).
%   signal: 'GPS_L1CA' (default), 'GLO' (GLO_GL1 / GLO_GL2), 'BDS_B3I', 'GAL_E1C', 'GPS_L5C', 'GAL_E5a', 'GAL_E5b'
%   or 'BDS_B2a' - the wrapper of each signal
%   folder passes its own.
if nargin < 2, signal = 'GPS_L1CA'; end
cfg.device = 0;
switch signal
    case 'GLO',     cfg.signal = 1;  cfg.freq_spacing = settings.freqSpacing;
    case 'BDS_B3I', cfg.signal = 2;  cfg.freq_spacing = 0;
    case 'GAL_E1C', cfg.signal = 3;  cfg.freq_spacing = 0;
    case 'GPS_L5C', cfg.signal = 4;  cfg.freq_spacing = 0;
    case 'GAL_E5a', cfg.signal = 5;  cfg.freq_spacing = 0;
    case 'GAL_E5b', cfg.signal = 6;  cfg.freq_spacing = 0;
    case 'BDS_B2a', cfg.signal = 7;  cfg.freq_spacing = 0;
    case 'BDS_B1I', cfg.signal = 8;  cfg.freq_spacing = 0;
    case 'GPS_L2C', cfg.signal = 9;  cfg.freq_spacing = 0;
    case 'BDS_B1C', cfg.signal = 10; cfg.freq_spacing = 0;
    otherwise,      cfg.signal = 0;  cfg.freq_spacing = 0;
end
cfg.file_type = settings.fileType;
if strcmp(settings.dataType, 'int16'), cfg.sample_bytes = 2; else, cfg.sample_bytes = 1; end   % initSettings.m:63
cfg.code_length = settings.codeLength;
if isfield(settings, 'acqNonCohTime')
    cfg.acq_noncoh_time = settings.acqNonCohTime;
else
    cfg.acq_noncoh_time = 1;           % variant-B folders (B1I, L2C) have no non-coherent sum
end
cfg.acq_coh_t = 0;  cfg.pilot_acq_flag = 0;   % BDS B1C only (set by its wrapper)
if isfield(settings, 'CNo')
    cfg.cno_vsm_interval = settings.CNo.VSMinterval;
    cfg.cno_acc_time = settings.CNo.accTime;
else                                  % BDS/B2a: settings.CNoInterval (initSettings.m:128), Calc_CNo_PLD on the host
    cfg.cno_vsm_interval = settings.CNoInterval;
    cfg.cno_acc_time = settings.intTime;
end
if isfield(settings, 'skipNumberOfSamples')   % the GLONASS folders' name for the same offset
    cfg.skip_number_of_bytes = settings.skipNumberOfSamples;
else
    cfg.skip_number_of_bytes = settings.skipNumberOfBytes;
end
cfg.sampling_freq = settings.samplingFreq;
cfg.IF = settings.IF;
cfg.code_freq_basis = settings.codeFreqBasis;
cfg.acq_search_band = settings.acqSearchBand;
if isfield(settings, 'acqSearchStep')
    cfg.acq_search_step = settings.acqSearchStep;
else
    cfg.acq_search_step = 0;           % set by the variant-B wrappers (resolved sub-bin step)
end
cfg.acq_threshold = settings.acqThreshold;
cfg.dll_damping_ratio = settings.dllDampingRatio;
cfg.dll_noise_bandwidth = settings.dllNoiseBandwidth;
cfg.dll_correlator_spacing = settings.dllCorrelatorSpacing;
cfg.pll_damping_ratio = settings.pllDampingRatio;
cfg.pll_noise_bandwidth = settings.pllNoiseBandwidth;
cfg.int_time = settings.intTime;
if isfield(settings, 'pilotTRKflag')
    cfg.pilot_trk_flag = settings.pilotTRKflag;
else
    cfg.pilot_trk_flag = 0;
end
% GPUs the library deals the PRNs / channels over (gc_multi_create): settings.gnsscorrGpus if the user set it, else the
% environment variable GNSSCORR_NGPUS, else 1; 0 = every visible GPU.  The reference's settings struct is otherwise unchanged.
if isfield(settings, 'gnsscorrGpus')
    cfg.n_gpus = settings.gnsscorrGpus;
elseif ~isempty(getenv('GNSSCORR_NGPUS'))
    cfg.n_gpus = str2double(getenv('GNSSCORR_NGPUS'));
else
    cfg.n_gpus = 1;
end
end
