function acqResults = acquisition(longSignal, settings)
%ACQUISITION  Drop-in for BDS/B1I/include/acquisition.m (same signature, 1x58 result vectors) that runs the
%variant-B search (one spectrum per sub-bin shift, Doppler bins by circular shift, two 4 ms blocks, peak / second-peak
%metric) on a B200.  The 2046-chip codes come from the reference's own generateCAcode53.
%
%   acqResults = acquisition(longSignal, settings)
fastPath = settings.resamplingflag == 0 && settings.fileType == 2 && ...
           strcmp(settings.dataType, 'schar') && ~isreal(longSignal) && ...
           all(real(longSignal) == round(real(longSignal))) && ...
           all(imag(longSignal) == round(imag(longSignal))) && ...
           max(abs(real(longSignal))) <= 128 && max(abs(imag(longSignal))) <= 128;
if ~fastPath
    acqResults = acquisition_reference(longSignal, settings);
    return
end
% sub-bin step exactly as acquisition.m:24-39 resolves settings.stepSize
Nblocks = 4;
samplesPerBlock = round(settings.samplingFreq / (settings.codeFreqBasis / (Nblocks * settings.codeLength)));
freqResolution = settings.samplingFreq / samplesPerBlock;
if isempty(settings.stepSize)
    stepSize = 0.5 / (Nblocks * settings.codeLength / settings.codeFreqBasis);
elseif settings.stepSize == freqResolution
    stepSize = settings.stepSize;
else
    steps = 1:0.25:freqResolution/2;
    steps = steps(rem(freqResolution, steps) == 0);
    stepDiff = steps - settings.stepSize;
    [~, minDiv] = min(abs(stepDiff));
    if stepDiff(minDiv) > 0, stepSize = steps(minDiv - 1); else, stepSize = steps(minDiv); end
end
cfg = gnsscorr_config(settings, 'BDS_B1I');
cfg.acq_search_step = stepSize;
sv = settings.acqSatelliteList;
codes.sv = double(sv(:).');
codes.data = zeros(2046, numel(sv), 'int8');
for k = 1:numel(sv), codes.data(:, k) = int8(generateCAcode53(sv(k))); end
codes.pilot = codes.data;
iq = zeros(1, 2 * numel(longSignal), 'int8');
iq(1:2:end) = int8(real(longSignal));
iq(2:2:end) = int8(imag(longSignal));
tstart = tic;
r = gnsscorr_mex('acquire', cfg, iq, double(sv), codes);
acqResults.carrFreq   = r.carrFreq;
acqResults.codePhase  = r.codePhase;
acqResults.peakMetric = r.peakMetric;
acqResults.timeVec    = repmat(toc(tstart) / numel(sv), 1, numel(sv));   % the reference records a per-PRN wall time (:172)
fprintf('(');
for PRN = sv
    if acqResults.carrFreq(PRN) ~= 0, fprintf('%02d ', PRN); else, fprintf('. '); end
end
fprintf(')\n');
end
