function set_shrink(obj, table)
%SET_SHRINK  Fused coefficient-domain soft threshold for an nd_dwt_*D / harr_nddwt_*D object (extension).
%   set_shrink(obj, table)   table: [levels x 2^d] thresholds, row j = level j (1 = finest), column 1 ignored
%                            (the approximation band is never thresholded); [] switches the shrink off.
%   Every later obj.dec(x, level) returns the thresholded coefficients: c .* max(0, 1 - t ./ abs(c)), written
%   once by the analysis kernels -- the step between dec and rec of an iterative loop x = rec(shrink(dec(x))).
for h = obj.plan_h
    nd_dwt_mex('shrink', h, double(table));
end
end
