function acqResults = acquisition(longSignal, settings)
%ACQUISITION  Drop-in for GPS/GPS_L1CA/include/acquisition.m (same signature and results) that
%runs the parallel code-phase search on a B200 through gnsscorr_mex -> libgnsscorr.so.
%
%   acqResults = acquisition(longSignal, settings)
%
% Put this folder ahead of the signal's include/ on the MATLAB path (init.m:39-40 adds include
% then Common).  Cases outside the accelerated path are handed to the original function, which
% must then be reachable as acquisition_reference (a renamed copy of the reference file).
% The engine takes the file's own samples: int8 ('schar') or int16, I,Q pairs (fileType 2) or real values (fileType 1).
is16 = strcmp(settings.dataType, 'int16');
if is16, cls = 'int16'; lim = 32768; else, cls = 'int8'; lim = 128; end
fastPath = settings.resamplingflag == 0 && (is16 || strcmp(settings.dataType, 'schar')) && ...
           isreal(longSignal) == (settings.fileType == 1) && ...
           all(real(longSignal) == round(real(longSignal))) && ...
           all(imag(longSignal) == round(imag(longSignal))) && ...
           max(abs(real(longSignal))) <= lim && max(abs(imag(longSignal))) <= lim;
if ~fastPath
    acqResults = acquisition_reference(longSignal, settings);
    return
end
if settings.fileType == 1
    iq = cast(longSignal, cls);
else
    iq = zeros(1, 2 * numel(longSignal), cls);
    iq(1:2:end) = cast(real(longSignal), cls);
    iq(2:2:end) = cast(imag(longSignal), cls);
end
r = gnsscorr_mex('acquire', gnsscorr_config(settings), iq, double(settings.acqSatelliteList));
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
