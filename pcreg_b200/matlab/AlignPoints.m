function [pts_aligned, coeff_unambig] = AlignPoints(pts)
%% AlignPoints -- drop-in for the reference's AlignPoints.m, computed on the GPU (libpcreg_b200 via pcreg_mex).
% Put this directory ahead of the reference on the MATLAB path.  Same outputs as the reference,
% including [] where the reference returns [].
    [pts_aligned, coeff_unambig] = pcreg_mex('align', 0, pts);
end
