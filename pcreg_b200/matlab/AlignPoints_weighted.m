function [pts_aligned, coeff_unambig, c] = AlignPoints_weighted(pts)
%% AlignPoints_weighted -- drop-in for the reference's AlignPoints_weighted.m, computed on the GPU (libpcreg_b200 via pcreg_mex).
% Put this directory ahead of the reference on the MATLAB path.  Same outputs as the reference,
% including [] where the reference returns [].
    [pts_aligned, coeff_unambig, c] = pcreg_mex('align', 3, pts);
end
