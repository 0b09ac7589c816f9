function acqResults = acquisition(longSignal, settings)
%ACQUISITION  Drop-in for BDS/B1C/include/acquisition.m:128-276 (same signature, 1 x max(acqSatelliteList) result
%vectors): variant C - one carrier wipe-off and FFT of (10 + acqCohT) ms, Doppler bins by circular shift, data and
%pilot BOC(1,1) replicas combined with the 11/40 and 29/40 power split, 25 Hz fine search - on a B200.  The codes
%come from the reference's own generateDataBOC11 / generatePilotBOC11.  NB_tracking / WB_tracking are not accelerated yet.
fastPath = settings.resamplingflag == 0 && settings.fileType == 2 && ...
           strcmp(settings.dataType, 'schar') && ~isreal(longSignal) && ...
           all(real(longSignal) == round(real(longSignal))) && ...
           all(imag(longSignal) == round(imag(longSignal))) && ...
           max(abs(real(longSignal))) <= 128 && max(abs(imag(longSignal))) <= 128;
if ~fastPath
    acqResults = acquisition_reference(longSignal, settings);
    return
end
cfg = gnsscorr_config(settings, 'BDS_B1C');
cfg.acq_search_step = settings.acqStep;
cfg.acq_coh_t = settings.acqCohT;
cfg.pilot_acq_flag = settings.pilotACQflag;
sv = settings.acqSatelliteList;
codes.sv = double(sv(:).');
codes.data  = zeros(2 * settings.codeLength, numel(sv), 'int8');
codes.pilot = zeros(2 * settings.codeLength, numel(sv), 'int8');
for k = 1:numel(sv)
    codes.data(:, k)  = int8(generateDataBOC11(settings, sv(k)));
    codes.pilot(:, k) = int8(generatePilotBOC11(settings, sv(k)));
end
iq = zeros(1, 2 * numel(longSignal), 'int8');
iq(1:2:end) = int8(real(longSignal));
iq(2:2:end) = int8(imag(longSignal));
r = gnsscorr_mex('acquire', cfg, iq, double(sv), codes);
n = max(sv);
acqResults.carrFreq   = r.carrFreq(1:n);
acqResults.codePhase  = r.codePhase(1:n);
acqResults.peakMetric = r.peakMetric(1:n);
fprintf('(');
for PRN = sv
    if acqResults.carrFreq(PRN) ~= 0, fprintf('%02d ', PRN); else, fprintf('. '); end
end
fprintf(')\n');
end
