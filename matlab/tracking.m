function [trackResults, channel] = tracking(fid, channel, settings)
%TRACKING  Drop-in for GPS/GPS_L1CA/include/tracking.m (same signature and trackResults struct)
%that runs the correlate-and-dump loops of all channels on a B200.
%
%   [trackResults, channel] = tracking(fid, channel, settings)
%
% A MEX file cannot use MATLAB's fid, so the file name is recovered with fopen(fid) and the
% library reads the record itself.
fastPath = (settings.fileType == 1 || settings.fileType == 2) && ...
           (strcmp(settings.dataType, 'schar') || strcmp(settings.dataType, 'int16'));
if ~fastPath
    [trackResults, channel] = tracking_reference(fid, channel, settings);
    return
end
fname = fopen(fid);
nCh = settings.numberOfChannels;
n   = settings.msToProcess;
r = gnsscorr_mex('track', gnsscorr_config(settings), fname, double([channel(1:nCh).PRN]), ...
                 double([channel(1:nCh).acquiredFreq]), double([channel(1:nCh).codePhase]), n);
names = {'absoluteSample', 'codeFreq', 'carrFreq', 'I_P', 'I_E', 'I_L', 'Q_E', 'Q_P', 'Q_L', ...
         'dllDiscr', 'dllDiscrFilt', 'pllDiscr', 'pllDiscrFilt', 'remCodePhase', 'remCarrPhase'};
shortRead = false;
for ch = nCh:-1:1
    t = struct('status', '-');
    t.absoluteSample = r.out(:, 1, ch).';
    t.codeFreq = r.out(:, 2, ch).';
    t.carrFreq = r.out(:, 3, ch).';
    t.I_P = r.out(:, 4, ch).';  t.I_E = r.out(:, 5, ch).';  t.I_L = r.out(:, 6, ch).';
    t.Q_E = r.out(:, 7, ch).';  t.Q_P = r.out(:, 8, ch).';  t.Q_L = r.out(:, 9, ch).';
    for k = 10:15
        t.(names{k}) = r.out(:, k, ch).';
    end
    t.CNo.VSMValue = r.vsmValue(:, ch).';
    t.CNo.VSMIndex = r.vsmIndex(:, ch).';
    if channel(ch).PRN ~= 0
        t.PRN = channel(ch).PRN;
        if r.epochsDone(ch) == n
            t.status = channel(ch).status;       % only after a complete run, as in the reference
        else
            shortRead = true;
        end
    else
        t.PRN = [];
    end
    trackResults(ch) = t; %#ok<AGROW>
end
if shortRead
    disp('Not able to read the specified number of samples  for tracking, exiting!')
end
end
