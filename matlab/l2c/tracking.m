function [trackResults, channel] = tracking(fid, channel, settings)
%TRACKING  Drop-in for GPS/GPS_L2C/include/tracking.m (same signature and trackResults struct: NumToProcess =
%round(msToProcess/1000/intTime) 20 ms epochs, fractional absoluteSample, code quantities recorded in chips) that runs
%the correlate-and-dump loops of all channels on a B200; with settings.pilotTRKflag == 1 the CL pilot is correlated too
%(channel.CLCodePhase, six Pilot_* rows, tracking.m:72-83, 259-366, 396-402).
fastPath = settings.fileType == 2 && strcmp(settings.dataType, 'schar');
if ~fastPath
    [trackResults, channel] = tracking_reference(fid, channel, settings);
    return
end
fname = fopen(fid);
nCh = settings.numberOfChannels;
n   = round(settings.msToProcess / 1000 / settings.intTime);          % tracking.m:51
prn = double([channel(1:nCh).PRN]);
sv = unique(prn(prn > 0));
codes.sv = sv;
codes.data = zeros(2 * settings.codeLength, numel(sv), 'int8');
for k = 1:numel(sv), codes.data(:, k) = int8(generateCMcode(sv(k), settings)); end
codes.pilot = codes.data;
cfg = gnsscorr_config(settings, 'GPS_L2C');
cfg.acq_search_step = settings.acqStep;
pilotOn = settings.pilotTRKflag == 1;
if pilotOn
    codes.cl = zeros(2 * settings.CLCodeLength, numel(sv), 'int8');
    for k = 1:numel(sv), codes.cl(:, k) = int8(generateCLcode(sv(k), settings)); end
    clp = ones(1, nCh);
    for ch = 1:nCh, if channel(ch).PRN ~= 0, clp(ch) = channel(ch).CLCodePhase; end, end
    codes.clCodePhase = double(clp);
end
r = gnsscorr_mex('track', cfg, fname, prn, double([channel(1:nCh).acquiredFreq]), double([channel(1:nCh).codePhase]), n, [], codes);
names = {'absoluteSample', 'codeFreq', 'carrFreq', 'I_P', 'I_E', 'I_L', 'Q_E', 'Q_P', 'Q_L', ...
         'dllDiscr', 'dllDiscrFilt', 'pllDiscr', 'pllDiscrFilt', 'remCodePhase', 'remCarrPhase'};
shortRead = false;
for ch = nCh:-1:1
    t = struct('status', '-');
    for k = 1:15, t.(names{k}) = r.out(:, k, ch).'; end
    % the engine runs the loop in half chips like the reference; recorded values are chips (tracking.m:223, 250, 376, 382-383)
    e = double(r.epochsDone(ch));
    step = t.codeFreq(1:e) / settings.samplingFreq;
    t.absoluteSample(1:e) = t.absoluteSample(1:e) + 1 - t.remCodePhase(1:e) ./ step;
    t.remCodePhase = t.remCodePhase / 2;  t.codeFreq = t.codeFreq / 2;
    t.dllDiscr = t.dllDiscr / 2;          t.dllDiscrFilt = t.dllDiscrFilt / 2;
    if pilotOn                                                         % tracking.m:396-402
        t.Pilot_I_P = r.out(:, 16, ch).';  t.Pilot_Q_P = r.out(:, 17, ch).';
        t.Pilot_I_E = r.out(:, 18, ch).';  t.Pilot_I_L = r.out(:, 19, ch).';
        t.Pilot_Q_E = r.out(:, 20, ch).';  t.Pilot_Q_L = r.out(:, 21, ch).';
    end
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
