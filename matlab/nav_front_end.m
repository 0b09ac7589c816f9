function [subFrameStart, navBitsBin] = nav_front_end(I_P_InputBits, settings)
%NAV_FRONT_END  Bit and frame synchronisation of GPS/GPS_L1CA/include/NAVdecoding.m:69-170 on a B200 (preamble
%cross-correlation, 6000 ms spacing test, TLM / HOW parity on 20 ms bit sums, the 1501 navigation bits) for ONE channel's
%trackResults.I_P.  Returns what NAVdecoding.m holds at line 170: subFrameStart (inf when no valid preamble was found) and
%navBitsBin = dec2bin(navBits), so the function body of NAVdecoding.m:60-170 can be replaced by
%
%    [subFrameStart, navBitsBin] = nav_front_end(I_P_InputBits, settings);
%    if subFrameStart == inf, disp('Could not find valid preambles in channel! '); return, end
%
%and its ephemeris decoding (:172-185) continues unchanged.
r = gnsscorr_mex('navsync', gnsscorr_config(settings), double(I_P_InputBits(:)));
if r.subFrameStart(1) == 0
    subFrameStart = inf;  navBitsBin = '';
    return
end
subFrameStart = double(r.subFrameStart(1));
if ~r.bitsValid(1)
    error('Index exceeds the number of array elements.');   % what NAVdecoding.m:152 does when the five subframes do not fit
end
navBitsBin = dec2bin(r.navBits(:, 1));
end
