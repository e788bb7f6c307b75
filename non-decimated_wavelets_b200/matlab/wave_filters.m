function [low_d, hi_d] = wave_filters(wname)
%WAVE_FILTERS  Daubechies analysis taps, 'db1' ... 'db10'.
%   Same call as the reference's Functions/wave_filters.m; the table itself lives in the
%   B200 library (csrc/nddwt_taps.h, re-derived by spectral factorisation) and is fetched through
%   the MEX gateway, so MATLAB and the CUDA kernels can never disagree on a tap.
if ~ischar(wname)
    error('Unknown Wavelet Name');
end
[low_d, hi_d] = nd_dwt_mex('taps', lower(wname));
end
