classdef harr_nddwt_4D
%HARR_NDDWT_4D  4-D non-decimated Haar transform (== nd_dwt_4D with 'db1') on B200.
%   obj = harr_nddwt_4D(sizes, 'pres_l2_norm',0|1, 'compute',..., 'precision',...)
%   Interface of the reference's Functions/harr_nddwt_4D.m.
    properties
        f_dec;          % filter descriptor (wavelet names + sizes)
        plan_h;         % device plans built from it in the constructor: [real, complex] handles for nd_dwt_mex
        ngpus = 1;      % extension: 'ngpus', n  splits the last dimension over n GPUs (host arrays)
        sizes;
        f_size;
        wname;
        scale;
        pres_l2_norm;
        compute;
        precision;
    end
    methods
        function obj = harr_nddwt_4D(sizes, varargin)
            obj = nddwt_b200_setup(obj, 4, 'db1', sizes, varargin, {});
            if obj.pres_l2_norm
                obj.scale = 1 / 2;
            else
                obj.scale = 1 / sqrt(2);
            end
        end
        function y = dec(obj, x, level)
            if nargin < 3, level = 1; end

            y = nddwt_b200_apply(obj, x, 0, level);
        end
        function y = rec(obj, x)
            y = nddwt_b200_apply(obj, x, 1, 0);
        end
    end
end
