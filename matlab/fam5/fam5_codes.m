function codes = fam5_codes(signal, svList, settings)
%FAM5_CODES  +-1 primary chips of the data and pilot codes for the listed PRNs from the reference's own
%generators, as gnsscorr_mex takes them (int8, 10230 x numel(sv)); GAL E5a adds the PRN's 100-chip pilot
%secondary code for the fine search.
svList = unique(svList(svList > 0));
codes.sv = double(svList(:).');
codes.data  = zeros(10230, numel(svList), 'int8');
codes.pilot = zeros(10230, numel(svList), 'int8');
if strcmp(signal, 'GAL_E5a'), codes.secondary = zeros(100, numel(svList), 'int8'); end
for k = 1:numel(svList)
    p = svList(k);
    switch signal
        case 'GPS_L5C', d = generateL5Icode(p, settings);      q = generateL5Qcode(p, settings);       % GPS_L5C tracking.m:164-168
        case 'GAL_E5a', d = generateE5aIcode(p, 2);            q = generateE5aQcode(p, 1);             % GAL_E5a tracking.m:150-154
                        codes.secondary(:, k) = int8(generateE5aQ_secondary(p));                       % GAL_E5a acquisition.m:190
        case 'GAL_E5b', d = generateE5bIcode(p, 2);            q = generateE5bQcode(p, 1);
        otherwise,      d = generateB2aDataCode(p, settings);  q = generateB2aPilotCode(p, settings);
    end
    codes.data(:, k)  = int8(d(:));
    codes.pilot(:, k) = int8(q(:));
end
end
