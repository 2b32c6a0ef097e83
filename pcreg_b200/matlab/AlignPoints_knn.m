function [pts_aligned, coeff_unambig, c] = AlignPoints_knn(pts, K)
%% AlignPoints_knn -- drop-in for the reference's AlignPoints_knn.m (absolute K; GPU via pcreg_mex).
% NOTE: differs from AlignPoints_KNN.m only by letter case (visualizeGTMatches.m:108); keep both shims on
% a case-sensitive file system.
    [pts_aligned, coeff_unambig, c] = pcreg_mex('align', 2, pts, K);
end
