function [pts_aligned, coeff_unambig, c] = AlignPoints_KNN(pts, varargin)
%% AlignPoints_KNN -- drop-in for the reference's AlignPoints_KNN.m (GPU, libpcreg_b200 via pcreg_mex).
% Optional varargin{1:2} = C1, C2 exactly as in the reference (AlignPoints_KNN.m:8-14).
    if length(varargin) == 2
        [pts_aligned, coeff_unambig, c] = pcreg_mex('align', 1, pts, varargin{1}, varargin{2});
    else
        [pts_aligned, coeff_unambig, c] = pcreg_mex('align', 1, pts);
    end
end
