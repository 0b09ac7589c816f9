function [trackResults, channel] = tracking(fid, channel, settings)
%TRACKING  Drop-in for GLO/GLO_GL1/include/tracking.m (and GLO_GL2): same signature and struct; a
%channel is live when channel.status ~= '-' and is identified by its frequency number channel.K,
%which is what ends up in trackResults.PRN (tracking.m:137-141).
fname = fopen(fid);
nCh = settings.numberOfChannels;
n   = settings.msToProcess;
K = nan(1, nCh);                              % NaN marks a channel that is off
for ch = 1:nCh
    if channel(ch).status ~= '-'
        K(ch) = channel(ch).K;
    end
end
r = gnsscorr_mex('track', gnsscorr_config(settings, 'GLO'), fname, K, ...
                 double([channel(1:nCh).acquiredFreq]), double([channel(1:nCh).codePhase]), n);
names = {'absoluteSample', 'codeFreq', 'carrFreq', 'I_P', 'I_E', 'I_L', 'Q_E', 'Q_P', 'Q_L', ...
         'dllDiscr', 'dllDiscrFilt', 'pllDiscr', 'pllDiscrFilt', 'remCodePhase', 'remCarrPhase'};
shortRead = false;
for ch = nCh:-1:1
    t = struct('status', '-');
    for k = 1:15
        t.(names{k}) = r.out(:, k, ch).';
    end
    t.CNo.VSMValue = r.vsmValue(:, ch).';
    t.CNo.VSMIndex = r.vsmIndex(:, ch).';
    t.PRN = [];
    if ~isnan(K(ch))
        t.PRN = K(ch);
        if r.epochsDone(ch) == n
            t.status = channel(ch).status;
        else
            shortRead = true;
        end
    end
    trackResults(ch) = t; %#ok<AGROW>
end
if shortRead
    disp('Not able to read the specified number of samples  for tracking, exiting!')
end
end
