function codes = e1codes(svList)
%E1CODES  +-1 primary chips of the E1-B / E1-C memory codes for the listed PRNs, as gnsscorr_mex takes them.
% generateE1Bcode / generateE1Ccode (reference, unchanged) return the BOC(1,1) sub-chip sequence
% [c -c c -c ...]; the primary chips are its odd elements.
svList = unique(svList(svList > 0));
codes.sv = double(svList(:).');
codes.data  = zeros(4092, numel(svList), 'int8');
codes.pilot = zeros(4092, numel(svList), 'int8');
for k = 1:numel(svList)
    b = generateE1Bcode(svList(k));
    c = generateE1Ccode(svList(k));
    codes.data(:, k)  = int8(b(1:2:end));
    codes.pilot(:, k) = int8(c(1:2:end));
end
end
