function acqResults = acquisition(longSignal, settings)
%ACQUISITION  Drop-in for GAL/GAL_E1C/include/acquisition.m (same signature; 1x50 result vectors) that
%runs the parallel code-phase search (E1B + E1C BOC(1,1) replicas, 25-period fine search against the
%pilot secondary code) on a B200 through gnsscorr_mex -> libgnsscorr.so.
%
%   acqResults = acquisition(longSignal, settings)
%
% Put this folder ahead of the signal's include/ on the MATLAB path.  The E1 memory codes stay where the
% reference keeps them (include/E1b.dat, E1c.dat): they are read with the reference's own
% generateE1Bcode / generateE1Ccode and handed to the library.  Cases outside the accelerated path go
% to the original function, which must then be reachable as acquisition_reference.
fastPath = settings.resamplingflag == 0 && settings.fileType == 2 && ...
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
r = gnsscorr_mex('acquire', gnsscorr_config(settings, 'GAL_E1C'), iq, double(settings.acqSatelliteList), ...
                 e1codes(settings.acqSatelliteList));
acqResults.carrFreq   = r.carrFreq;
acqResults.codePhase  = r.codePhase;
acqResults.peakMetric = r.peakMetric;
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
