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
for k = 1:2:numel(opts)
    key = lower(opts{k});
    val = opts{k + 1};
    if strcmp(key, 'pres_l2_norm')
        obj.pres_l2_norm = double(logical(val));
    elseif strcmp(key, 'compute')
        obj.compute = val;
    elseif strcmp(key, 'precision')
        obj.precision = val;
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
% what travels in nd_dwt_mex's second argument: names + sizes (a few bytes, not 2^d*numel complex)
obj.f_dec = struct('wname', {names}, 'sizes', obj.sizes);
end
