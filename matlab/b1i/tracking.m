function [trackResults, channel] = tracking(fid, channel, settings)
%TRACKING  Drop-in for BDS/B1I/include/tracking.m (same signature and trackResults struct) that runs the
%correlate-and-dump loops of all channels on a B200; codes from the reference's generateCAcode53.
fastPath = settings.fileType == 2 && strcmp(settings.dataType, 'schar');
if ~fastPath
    [trackResults, channel] = tracking_reference(fid, channel, settings);
    return
end
fname = fopen(fid);
nCh = settings.numberOfChannels;
n   = settings.msToProcess;
prn = double([channel(1:nCh).PRN]);
sv = unique(prn(prn > 0));
codes.sv = sv;
codes.data = zeros(2046, numel(sv), 'int8');
for k = 1:numel(sv), codes.data(:, k) = int8(generateCAcode53(sv(k))); end
codes.pilot = codes.data;
cfg = gnsscorr_config(settings, 'BDS_B1I');
cfg.acq_search_step = 125;                       % unused by tracking, must divide freqResolution
r = gnsscorr_mex('track', cfg, fname, prn, double([channel(1:nCh).acquiredFreq]), double([channel(1:nCh).codePhase]), n, [], codes);
names = {'absoluteSample', 'codeFreq', 'carrFreq', 'I_P', 'I_E', 'I_L', 'Q_E', 'Q_P', 'Q_L', ...
         'dllDiscr', 'dllDiscrFilt', 'pllDiscr', 'pllDiscrFilt', 'remCodePhase', 'remCarrPhase'};
shortRead = false;
for ch = nCh:-1:1
    t = struct('status', '-');
    for k = 1:15, t.(names{k}) = r.out(:, k, ch).'; end
    t.CNo.VSMValue = r.vsmValue(:, ch).';
    t.CNo.VSMIndex = r.vsmIndex(:, ch).';
    if channel(ch).PRN ~= 0
        t.PRN = channel(ch).PRN;
        if r.epochsDone(ch) == n, t.status = channel(ch).status; else, shortRead = true; end
    else
        t.PRN = [];
    end
    trackResults(ch) = t; %#ok<AGROW>
end
if shortRead
    disp('Not able to read the specified number of samples  for tracking, exiting!')
end
end
