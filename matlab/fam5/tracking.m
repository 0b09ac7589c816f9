function [trackResults, channel] = tracking(fid, channel, settings)
%TRACKING  Drop-in for the tracking.m of GPS/GPS_L5C, GAL/GAL_E5a, GAL/GAL_E5b and BDS/B2a (same signature and
%trackResults struct: code NCO centred on channel.codeFreq, Pilot_I_P / Pilot_Q_P when the pilot is tracked,
%B2a's DataCNo / DataPLD block from the device-side Calc_CNo_PLD, r.cnoPld) that runs the correlate-and-dump loops
%of all channels on a B200.
%
%   [trackResults, channel] = tracking(fid, channel, settings)
SIGNAL = fam5_signal(settings);
fastPath = settings.fileType == 2 && strcmp(settings.dataType, 'schar');
if ~fastPath
    [trackResults, channel] = tracking_reference(fid, channel, settings);
    return
end
fname = fopen(fid);
nCh = settings.numberOfChannels;
n   = settings.msToProcess;
prn = double([channel(1:nCh).PRN]);
r = gnsscorr_mex('track', gnsscorr_config(settings, SIGNAL), fname, prn, double([channel(1:nCh).acquiredFreq]), ...
                 double([channel(1:nCh).codePhase]), n, double([channel(1:nCh).codeFreq]), fam5_codes(SIGNAL, prn, settings));
names = {'absoluteSample', 'codeFreq', 'carrFreq', 'I_P', 'I_E', 'I_L', 'Q_E', 'Q_P', 'Q_L', ...
         'dllDiscr', 'dllDiscrFilt', 'pllDiscr', 'pllDiscrFilt', 'remCodePhase', 'remCarrPhase'};
pilot = size(r.out, 2) == 17;
shortRead = false;
for ch = nCh:-1:1
    t = struct('status', '-');
    for k = 1:15
        t.(names{k}) = r.out(:, k, ch).';
    end
    if pilot
        t.Pilot_I_P = r.out(:, 16, ch).';
        t.Pilot_Q_P = r.out(:, 17, ch).';
    elseif ~strcmp(SIGNAL, 'GPS_L5C') && ~strcmp(SIGNAL, 'BDS_B2a')
        t.Pilot_I_P = zeros(1, n);  t.Pilot_Q_P = zeros(1, n);     % E5a/E5b create the fields unconditionally (tracking.m:57-58)
    end
    if strcmp(SIGNAL, 'BDS_B2a')                                    % BDS/B2a/include/tracking.m:66-72, 336-352
        nv = floor(n / settings.CNoInterval);
        t.DataCNo = zeros(1, nv);  t.DataPLD = zeros(1, nv);
        if pilot, t.PilotCNo = zeros(1, nv);  t.PilotPLD = zeros(1, nv);  t.B2a_CNo = zeros(1, nv); end
        % Calc_CNo_PLD.m:38-100 and the 0.5/0.5 smoothing of tracking.m:409-431 were evaluated on the GPU from the same rows
        t.DataCNo = r.cnoPld(:, 1, ch).';  t.DataPLD = r.cnoPld(:, 2, ch).';
        if pilot
            t.PilotCNo = r.cnoPld(:, 3, ch).';  t.PilotPLD = r.cnoPld(:, 4, ch).';  t.B2a_CNo = r.cnoPld(:, 5, ch).';
        end
    else
        t.CNo.VSMValue = r.vsmValue(:, ch).';
        t.CNo.VSMIndex = r.vsmIndex(:, ch).';
    end
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
