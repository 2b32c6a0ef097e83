function [T, rmse, n_used, status, best, idx] = pcreg_icp(model, src, T0, opts, varargin)
%% pcreg_icp -- batched multi-start ICP against a GPU-resident model handle.
%   model  = pcreg_mex('model_create', single(pcModel.Location));   % once per model cloud
%   T0     : 4x4xH initial poses (row-vector convention, quickTF.m)
%   opts   : struct with fields mode (0 plain | 1 KNN-trim | 2 weighted), iters, k_frac, R_w, thDist2, nn (0 brute | 1 grid)
%   w_src  : optional per-point weights (varargin{1})
    H = size(T0, 3);
    [T, rmse, n_used, status, best, idx] = pcreg_mex('icp', model, src, reshape(double(T0), 16, H), opts, varargin{:});
    T = reshape(T, 4, 4, H);
end
