function [T, inlierIdx, numSuccess, maxInliers, pct] = ransac(pts1, pts2, ransacCoef, funcFindTransf, funcDist)
%% ransac -- drop-in for the reference's ransac.m (same five outputs).  The function handles are accepted
% for signature compatibility; the fit is estimateTransform and the distance is calcDists (the only pair the
% reference ever passes: getInliersRANSAC.m:34 and the 8 other call sites).  Sample triplets are drawn HERE
% with randperm exactly as ransac.m:42-43 does, then scored on the GPU in one batch.
    ptNum = size(pts1, 1);
    iterNum = ransacCoef.iterNum;
    triplets = zeros(iterNum, 3);
    for p = 1:iterNum
        sampleIdx = randperm(ptNum);
        triplets(p, :) = sampleIdx(1:3);
    end
    [T, inlierIdx, numSuccess, maxInliers, pct] = pcreg_mex('ransac', double(pts1), double(pts2), ransacCoef, triplets);
end
