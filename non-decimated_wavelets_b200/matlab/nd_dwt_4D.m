classdef nd_dwt_4D
%ND_DWT_4D  4-D multi-level non-decimated (periodic) Daubechies wavelet transform on B200.
%   obj = nd_dwt_4D(wname, sizes, 'pres_l2_norm',0|1, 'compute','mat'|'mex'|'gpu'|'gpu_off', ...
%                      'precision','double'|'single')
%   y = obj.dec(x, level);   x = obj.rec(y);
%   Interface of the reference's Functions/nd_dwt_4D.m; the work is done by libnddwt_b200
%   (direct separable circular filtering, no FFT, no stored Fourier-domain filters).
    properties
        f_dec;          % filter descriptor (wavelet names + sizes)
        plan_h;         % device plans built from it in the constructor: [real, complex] handles for nd_dwt_mex
        ngpus = 1;      % extension: 'ngpus', n  splits the last dimension over n GPUs (host arrays)
        sizes;
        f_size;
        wname;
        pres_l2_norm;
        compute;
        precision;
        method = 'fft';  % accepted for compatibility ('fft' | 'conv'); both run the spatial kernels
    end
    methods
        function obj = nd_dwt_4D(wname, sizes, varargin)
            obj = nddwt_b200_setup(obj, 4, wname, sizes, varargin, {'method'});
        end
        function y = dec(obj, x, level)
            y = nddwt_b200_apply(obj, x, 0, level);
        end
        function y = rec(obj, x)
            y = nddwt_b200_apply(obj, x, 1, 0);
        end
    end
end
