function y = nddwt_b200_apply(obj, x, direction, level)
%NDDWT_B200_APPLY  dec (direction 0) / rec (direction 1) through the CUDA MEX gateway.
%   Every `compute` value of the reference ('mat','mex','gpu','gpu_off') runs the same sm_100a kernels; they
%   only say where the arrays live, as in the reference:
%     'gpu'      x is (or becomes) a gpuArray and STAYS on the device: the gateway hands its device pointer
%                to the kernels and returns a gpuArray (no gather / upload per call);
%     'gpu_off'  host in, host out (the gateway copies in and out, like nd_dwt_1D.m:139-141,192-194);
%     'mat','mex' host arrays; a gpuArray passed anyway is processed on the device and returned as gpuArray.
%   Plans with 'ngpus' > 1 take host arrays (every GPU loads its own slab of the last dimension).
on_gpu = isa(x, 'gpuArray');
if strcmpi(obj.compute, 'gpu') && ~on_gpu && obj.ngpus == 1
    x = gpuArray(x);
    on_gpu = true;
end
if on_gpu && obj.ngpus > 1
    x = gather(x);
    on_gpu = false;
end
if strcmpi(obj.precision, 'single')
    if ~isa(x, 'single') && ~(on_gpu && strcmp(classUnderlying(x), 'single')), x = single(x); end
else
    if ~isa(x, 'double') && ~(on_gpu && strcmp(classUnderlying(x), 'double')), x = double(x); end
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
h = obj.plan_h(1 + ~isreal(x));          % real / complex plan, built in the constructor
y = nd_dwt_mex(x, h, direction, level, obj.pres_l2_norm);
if strcmpi(obj.compute, 'gpu_off') && isa(y, 'gpuArray')
    y = gather(y);
end
end
