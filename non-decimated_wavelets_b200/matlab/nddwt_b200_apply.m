function y = nddwt_b200_apply(obj, x, direction, level)
%NDDWT_B200_APPLY  dec (direction 0) / rec (direction 1) through the CUDA MEX gateway.
%   Every `compute` value of the reference ('mat','mex','gpu','gpu_off') runs the same sm_100a
%   kernels; gpuArray inputs are gathered for the gateway and the result is returned as a gpuArray
%   again so scripts written for 'gpu' keep working.
was_gpu = isa(x, 'gpuArray');
if was_gpu
    x = gather(x);
end
if strcmpi(obj.precision, 'single')
    x = single(x);
else
    x = double(x);
end
nd = 2 ^ numel(obj.sizes);
if direction ~= 0
    nb = size(x, numel(obj.sizes) + 1);
    if numel(obj.sizes) == 1, nb = size(x, 2); end
    level = 1 + (nb - nd) / (nd - 1);
    if level < 1 || level ~= floor(level)
        error('FIlter size and image size not consistant');
    end
end
y = nd_dwt_mex(x, obj.f_dec, direction, level, obj.pres_l2_norm);
if was_gpu || strcmpi(obj.compute, 'gpu')
    y = gpuArray(y);
end
end
