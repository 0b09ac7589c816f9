function [trackResults, channel] = WB_tracking(fid, channel, settings)
%WB_TRACKING  Drop-in for BDS/B1C/include/WB_tracking.m (same signature and trackResults struct, settings.pilotTRKflag == 2;
%postProcessing.m:34-38 dispatches here) that runs the 10 ms full-band loops of all channels on a B200: data BOC(1,1),
%pilot BOC(1,1) and pilot BOC(6,1) replicas (18 sums), composite pilot correlations, code error weighted by
%CalcWeighingFactor(settings).
fastPath = settings.fileType == 2 && strcmp(settings.dataType, 'schar') && settings.pilotTRKflag == 2;
if ~fastPath
    [trackResults, channel] = WB_tracking_reference(fid, channel, settings);
    return
end
fname = fopen(fid);
nCh = settings.numberOfChannels;
n   = round(settings.msToProcess / 1000 / settings.intTime);          % WB_tracking.m:49
prn = double([channel(1:nCh).PRN]);
sv = unique(prn(prn > 0));
codes.sv = sv;
codes.data  = zeros(2 * settings.codeLength, numel(sv), 'int8');
codes.pilot = zeros(2 * settings.codeLength, numel(sv), 'int8');
for k = 1:numel(sv)
    codes.data(:, k)  = int8(generateDataBOC11(settings, sv(k)));
    codes.pilot(:, k) = int8(generatePilotBOC11(settings, sv(k)));
end
codes.boc61 = zeros(12 * settings.codeLength, numel(sv), 'int8');
for k = 1:numel(sv), codes.boc61(:, k) = int8(generatePilotBOC61(settings, sv(k))); end
cfg = gnsscorr_config(settings, 'BDS_B1C');
cfg.acq_search_step = settings.acqStep;  cfg.acq_coh_t = settings.acqCohT;  cfg.pilot_acq_flag = settings.pilotACQflag;
cfg.wb_factor = CalcWeighingFactor(settings);                      % WB_tracking.m:124 (adaptive quadrature, stays in MATLAB)
r = gnsscorr_mex('track', cfg, fname, prn, double([channel(1:nCh).acquiredFreq]), double([channel(1:nCh).codePhase]), n, ...
                 double([channel(1:nCh).codeFreq]), codes);
names = {'absoluteSample', 'codeFreq', 'carrFreq', 'I_P', 'I_E', 'I_L', 'Q_E', 'Q_P', 'Q_L', ...
         'dllDiscr', 'dllDiscrFilt', 'pllDiscr', 'pllDiscrFilt', 'remCodePhase', 'remCarrPhase'};
nv = floor(n / settings.CNoInterval);
shortRead = false;
for ch = nCh:-1:1
    t = struct('status', '-');
    for k = 1:15, t.(names{k}) = r.out(:, k, ch).'; end
    t.Pilot_I_P = r.out(:, 16, ch).';  t.Pilot_Q_P = r.out(:, 17, ch).';     % composite pilot correlations (WB_tracking.m:409-414)
    t.Pilot_I_E = r.out(:, 18, ch).';  t.Pilot_I_L = r.out(:, 19, ch).';
    t.Pilot_Q_E = r.out(:, 20, ch).';  t.Pilot_Q_L = r.out(:, 21, ch).';
    t.DataCNo = zeros(1, nv);  t.DataPLD = zeros(1, nv);  t.PilotCNo = zeros(1, nv);  t.PilotPLD = zeros(1, nv);  t.B1C_CNo = zeros(1, nv);
    % Calc_CNo_PLD.m and the 0.5/0.5 smoothing were evaluated on the GPU from the same rows (r.cnoPld: nIntervals x 5 x nCh)
    t.DataCNo = r.cnoPld(:, 1, ch).';   t.DataPLD = r.cnoPld(:, 2, ch).';
    t.PilotCNo = r.cnoPld(:, 3, ch).';  t.PilotPLD = r.cnoPld(:, 4, ch).';  t.B1C_CNo = r.cnoPld(:, 5, ch).';
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
