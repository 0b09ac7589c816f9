function signal = fam5_signal(settings)
%FAM5_SIGNAL  Which of the four 10230-chip data + pilot folders these settings belong to, from fields that
%differ between their initSettings.m: B2a has CNoInterval, E5b the 1207.14 MHz carrier, E5a 50-entry results
%with CNo.VSMinterval 100, L5C the rest.
if isfield(settings, 'CNoInterval')
    signal = 'BDS_B2a';
elseif settings.carrFreqBasis == 1207.14e6
    signal = 'GAL_E5b';
elseif max(settings.acqSatelliteList) > 32 || settings.CNo.VSMinterval == 100
    signal = 'GAL_E5a';
else
    signal = 'GPS_L5C';
end
end
