function acqResults = acquisition(longSignal, settings)
%ACQUISITION  Drop-in for the acquisition.m of GPS/GPS_L5C, GAL/GAL_E5a, GAL/GAL_E5b and BDS/B2a (same
%signature and result vectors: 1x32 L5C, 1x50 E5a/E5b, 1 x max(acqSatelliteList) B2a) that runs the
%two-replica parallel code-phase search and the signal's fine search on a B200.
%
%   acqResults = acquisition(longSignal, settings)
%
% Copy this folder next to the signal folder and set SIGNAL below (or keep one copy per signal); put it ahead
% of the signal's include/ on the MATLAB path.  The primary codes come from the reference's own generators
% (generateL5Icode.m, generateE5aIcode.m, generateB2aDataCode.m, ...), which stay untouched.
SIGNAL = fam5_signal(settings);
if isfield(settings, 'resamplingflag'), rs = settings.resamplingflag; else, rs = settings.resamplingFlag; end
fastPath = rs == 0 && settings.fileType == 2 && ...
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
r = gnsscorr_mex('acquire', gnsscorr_config(settings, SIGNAL), iq, double(settings.acqSatelliteList), ...
                 fam5_codes(SIGNAL, settings.acqSatelliteList, settings));
if strcmp(SIGNAL, 'BDS_B2a')
    n = max(settings.acqSatelliteList);             % BDS/B2a/include/acquisition.m:128-132
else
    n = numel(r.carrFreq);
end
acqResults.carrFreq   = r.carrFreq(1:n);
acqResults.codePhase  = r.codePhase(1:n);
acqResults.peakMetric = r.peakMetric(1:n);
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
