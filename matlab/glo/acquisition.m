function acqResults = acquisition(longSignal, settings)
%ACQUISITION  Drop-in for GLO/GLO_GL1/include/acquisition.m and GLO/GLO_GL2/include/acquisition.m
%(same signature; results are 1x21 vectors indexed K+8) running on a B200 through gnsscorr_mex.
%
% The GLONASS postProcessing.m builds longSignal = Q + 1i*I (postProcessing.m:94); the engine wants
% the file's byte order back and applies the swap itself.
fastPath = settings.resamplingflag == 0 && settings.fileType == 2 && ...
           strcmp(settings.dataType, 'schar') && ~isreal(longSignal) && ...
           all(real(longSignal) == round(real(longSignal))) && ...
           all(imag(longSignal) == round(imag(longSignal)));
if ~fastPath
    acqResults = acquisition_reference(longSignal, settings);
    return
end
iq = zeros(1, 2 * numel(longSignal), 'int8');
iq(1:2:end) = int8(imag(longSignal));     % I
iq(2:2:end) = int8(real(longSignal));     % Q
r = gnsscorr_mex('acquire', gnsscorr_config(settings, 'GLO'), iq, double(settings.acqSatelliteList));
acqResults.carrFreq   = r.carrFreq;
acqResults.codePhase  = r.codePhase;
acqResults.peakMetric = r.peakMetric;
fprintf('(');
for K = settings.acqSatelliteList
    if acqResults.carrFreq(K + 8) ~= 0
        fprintf('%02d ', K);
    else
        fprintf('. ');
    end
end
fprintf(')\n');
end
