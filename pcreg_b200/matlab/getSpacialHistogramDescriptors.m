function [feat, desc] = getSpacialHistogramDescriptors(pts, sample_pts, options)
%% getSpacialHistogramDescriptors -- drop-in for the reference's getSpacialHistogramDescriptors.m on the GPU.
% Put this directory ahead of the reference on the MATLAB path.  Same inputs (options.min_pts, max_pts, R, thVar, k,
% ALIGN_POINTS) and outputs (feat: surviving keypoints, desc: one 980-bin spherical histogram per row).  Both parfor
% loops of the reference (one full scan of pts per keypoint, twice) become one batched call; the neighbourhoods never
% leave the GPU.  pts may also be a model handle from pcreg_mex('model_create', ...).
    NUM_R = 10; NUM_THETA = 7; NUM_PHI = 14;            % getSpacialHistogramDescriptors.m:38-40
    R = options.R;
    r_bins = nthroot(0:R^3/NUM_R:R^3, 3);               % :155-158, evaluated by MATLAB itself so the edges are the reference's
    phi_bins = -pi:2*pi/NUM_PHI:pi;
    theta_bins = 0:pi/NUM_THETA:pi;
    own = ~isa(pts, 'uint64');
    if own
        h = pcreg_mex('model_create', pts, 0);
    else
        h = pts;
    end
    [feat, desc] = pcreg_mex('spatial_histogram', h, double(sample_pts), options, r_bins, theta_bins, phi_bins);
    if own
        pcreg_mex('model_destroy', h);
    end
end
