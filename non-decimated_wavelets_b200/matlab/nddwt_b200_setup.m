function obj = nddwt_b200_setup(obj, ndim, wname, sizes, opts, extra_keys)
%NDDWT_B200_SETUP  Shared constructor logic of the nd_dwt_*D / harr_nddwt_*D objects.
%   Validates sizes / wavelet names / name-value options exactly like the reference classes
%   (same keys, defaults, error and warning texts) and builds the small descriptor `f_dec`
%   that replaces the reference's stored Fourier-domain filter array.
if numel(sizes) ~= ndim
    if ndim == 1
        error('1D array length must be a scalar');
    end
    error('The sizes vector must be length %d', ndim);
end
obj.sizes = reshape(double(sizes), 1, []);
if mod(numel(opts), 2)
    error('Optional inputs must come in pairs');
end
if ischar(wname)
    names = repmat({wname}, 1, ndim);
elseif iscell(wname) && ndim == 1
    error('Wavelet Name Must be a string');
elseif iscell(wname) && numel(wname) == ndim
    names = reshape(wname, 1, []);
else
    error(['You must specify one filter name per dimension in a cell array, or a single ' ...
           'string for the same filter to be used in all dimensions']);
end
obj.wname = names;
obj.pres_l2_norm = 0;
obj.precision = 'double';
obj.compute = 'mat';
ngpus = 1;
atrous = false;
for k = 1:2:numel(opts)
    key = lower(opts{k});
    val = opts{k + 1};
    if strcmp(key, 'pres_l2_norm')
        obj.pres_l2_norm = double(logical(val));
    elseif strcmp(key, 'compute')
        obj.compute = val;
    elseif strcmp(key, 'precision')
        obj.precision = val;
    elseif strcmp(key, 'atrous')         % extension: true a-trous transform, dilation 2^(j-1) at level j
        atrous = logical(val);
    elseif strcmp(key, 'ngpus')          % extension: slabs along the last dimension over several GPUs (host arrays)
        ngpus = double(val);
    elseif any(strcmp(key, extra_keys))
        obj.(key) = val;
    else
        warning(sprintf('Unknown optional input #%d ingoring!', k));
    end
end
% taps + "filter longer than the data" check (errors surface in the constructor, as before)
obj.f_size = struct();
for i = 1:ndim
    lo = wave_filters(names{i});
    obj.f_size.(sprintf('s%d', i)) = numel(lo);
    if numel(lo) > obj.sizes(i)
        error('Dimension %d of Data is shorter than the wavelet filter being used', i);
    end
end
% The "stored filters": a descriptor (names + sizes, a few bytes instead of the reference's 2^d*numel complex
% f_dec, nd_dwt_2D.m:305-308) and the device-side plans built from it ONCE, here, like the reference builds
% f_dec in its constructor.  plan_h = [handle for real data, handle for complex data] (uint64), passed in
% nd_dwt_mex's f position on every dec / rec; the plans own taps, scratch and -- ngpus > 1 -- the peer
% mappings and halo inboxes of all GPUs.  They live until nd_dwt_mex('release', h) / clear mex.
obj.f_dec = struct('wname', {names}, 'sizes', obj.sizes);
is_single = strcmpi(obj.precision, 'single');
obj.plan_h = [nd_dwt_mex('plan', obj.f_dec, is_single, false, obj.pres_l2_norm, ngpus), ...
              nd_dwt_mex('plan', obj.f_dec, is_single, true, obj.pres_l2_norm, ngpus)];
obj.ngpus = ngpus;
if atrous
    dil = 2 .^ (0:15);
    dil = dil(dil <= obj.sizes(end));
    for h = obj.plan_h
        nd_dwt_mex('dilations', h, dil);
    end
end
end
