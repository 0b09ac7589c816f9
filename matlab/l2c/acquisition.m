function acqResults = acquisition(longSignal, settings)
%ACQUISITION  Drop-in for GPS/GPS_L2C/include/acquisition.m:4-145 (same signature, 1x32 result vectors): the CM-code
%variant-B search on a B200 and, with settings.pilotTRKflag == 1, the 75-way CL code phase search of :100-137
%(acqResults.CLCodePhase).
fastPath = settings.resamplingflag == 0 && settings.fileType == 2 && ...
           strcmp(settings.dataType, 'schar') && ~isreal(longSignal) && ...
           all(real(longSignal) == round(real(longSignal))) && ...
           all(imag(longSignal) == round(imag(longSignal))) && ...
           max(abs(real(longSignal))) <= 128 && max(abs(imag(longSignal))) <= 128;
if ~fastPath
    acqResults = acquisition_reference(longSignal, settings);
    return
end
cfg = gnsscorr_config(settings, 'GPS_L2C');
cfg.acq_search_step = settings.acqStep;
cfg.acq_noncoh_time = 1;
sv = settings.acqSatelliteList;
codes.sv = double(sv(:).');
codes.data = zeros(2 * settings.codeLength, numel(sv), 'int8');      % generateCMcode returns the return-to-zero sequence
for k = 1:numel(sv), codes.data(:, k) = int8(generateCMcode(sv(k), settings)); end
codes.pilot = codes.data;
if settings.pilotTRKflag == 1                                          % the return-to-zero CL sequences (generateCLcode.m)
    codes.cl = zeros(2 * settings.CLCodeLength, numel(sv), 'int8');
    for k = 1:numel(sv), codes.cl(:, k) = int8(generateCLcode(sv(k), settings)); end
    cfg.acq_coh_t = settings.acqCohT;
end
iq = zeros(1, 2 * numel(longSignal), 'int8');
iq(1:2:end) = int8(real(longSignal));
iq(2:2:end) = int8(imag(longSignal));
r = gnsscorr_mex('acquire', cfg, iq, double(sv), codes);
acqResults.carrFreq   = r.carrFreq;
acqResults.codePhase  = r.codePhase;
acqResults.peakMetric = r.peakMetric;
if settings.pilotTRKflag == 1, acqResults.CLCodePhase = r.CLCodePhase; end   % acquisition.m:136
fprintf('(');
for PRN = sv
    if acqResults.carrFreq(PRN) ~= 0, fprintf('%02d ', PRN); else, fprintf('. '); end
end
fprintf(')\n');
end
