function T = estimateTransform(pts1, pts2)
%% estimateTransform -- drop-in for the reference's estimateTransform.m: [pts2, 1] * T = [pts1, 1];
% [] where the reference's rank guard fires (estimateTransform.m:11-14).
    T = pcreg_mex('estimate_transform', double(pts1), double(pts2));
end
