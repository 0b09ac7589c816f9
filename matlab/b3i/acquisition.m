function acqResults = acquisition(longSignal, settings)
%ACQUISITION  Drop-in for BDS/B3I/include/acquisition.m (same signature; 1x63 result vectors) that
%runs the parallel code-phase search on a B200 through gnsscorr_mex -> libgnsscorr.so.
%
%   acqResults = acquisition(longSignal, settings)
%
% Put this folder ahead of the signal's include/ on the MATLAB path (init.m:39-40 adds include
% then Common).  Cases outside the accelerated path are handed to the original function, which
% must then be reachable as acquisition_reference (a renamed copy of the reference file).
fastPath = settings.resamplingFlag == 0 && settings.fileType == 2 && ...
           strcmp(settings.dataType, 'schar') && ~isreal(longSignal) && ...
           all(real(longSignal) == round(real(longSignal))) && ...
           all(imag(longSignal) == round(imag(longSignal))) && ...
           max(abs(real(longSignal))) <= 128 && max(abs(imag(longSignal))) <= 128;
if ~fastPath
    acqResults = acquisition_reference(longSignal, settings);
    return
end
iq = zeros(1, 2 * numel(longSignal), 'int8');
iq(1:2:end) = int8(real(longSignal));
iq(2:2:end) = int8(imag(longSignal));
r = gnsscorr_mex('acquire', gnsscorr_config(settings, 'BDS_B3I'), iq, double(settings.acqSatelliteList));
acqResults.carrFreq   = r.carrFreq;
acqResults.codePhase  = r.codePhase;
acqResults.peakMetric = r.peakMetric;
% the reference's console line, printed once the search has returned
fprintf('(');
for PRN = settings.acqSatelliteList
    if acqResults.carrFreq(PRN) ~= 0
        fprintf('%02d ', PRN);
    else
        fprintf('. ');
    end
end
fprintf(')\n');
end
