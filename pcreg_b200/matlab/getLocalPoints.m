function [pts_sphere, dists] = getLocalPoints(pts, R, c, min_points, max_points)
%% getLocalPoints -- drop-in for the reference's getLocalPoints.m, computed on the GPU (libpcreg_b200 via pcreg_mex).
% Put this directory ahead of the reference on the MATLAB path.  Same outputs as the reference: the points within
% radius R of c, RELATIVE to c, in the order of pts, and their distances; [] , [] when the count is outside
% [min_points, max_points].  pts may also be a model handle from pcreg_mex('model_create', ...) -- the reference
% scans the same cloud once per keypoint (getSpacialHistogramDescriptors.m:50,68), a resident handle uploads it once,
% and c may then hold many centres (Kx3): use the third output of pcreg_mex('local_points', ...) to split the rows.
    own = ~isa(pts, 'uint64');
    if own
        h = pcreg_mex('model_create', pts, 0);
    else
        h = pts;
    end
    [pts_sphere, dists] = pcreg_mex('local_points', h, double(c), R, min_points, max_points);
    if own
        pcreg_mex('model_destroy', h);
        pts_sphere = cast(pts_sphere, class(pts));      % the reference propagates the class of pts
        dists = cast(dists, class(pts));
    end
end
