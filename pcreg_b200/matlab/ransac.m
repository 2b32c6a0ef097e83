function [T, inlierIdx, numSuccess, maxInliers, pct] = ransac(pts1, pts2, ransacCoef, funcFindTransf, funcDist)
%% ransac -- drop-in for the reference's ransac.m (same five outputs).  The function handles are accepted
% for signature compatibility; the fit is estimateTransform and the distance is calcDists (the only pair the
% reference ever passes: getInliersRANSAC.m:34 and the 8 other call sites).  Sample triplets are drawn HERE
% with randperm exactly as ransac.m:42-43 does, then scored on the GPU in one batch.
    % the GPU path is the reference's only configuration: 3-point samples, estimateTransform, calcDists -- anything else is an
    % error here rather than a silently different result (ransac.m:23,43-45)
    if isfield(ransacCoef, 'minPtNum') && ransacCoef.minPtNum ~= 3
        error('pcreg:ransac', 'pcreg ransac drop-in supports minPtNum = 3 only (got %d)', ransacCoef.minPtNum);
    end
    if nargin >= 4 && isa(funcFindTransf, 'function_handle') && ~strcmp(func2str(funcFindTransf), 'estimateTransform')
        error('pcreg:ransac', 'pcreg ransac drop-in fits with estimateTransform only (got @%s)', func2str(funcFindTransf));
    end
    if nargin >= 5 && isa(funcDist, 'function_handle') && ~strcmp(func2str(funcDist), 'calcDists')
        error('pcreg:ransac', 'pcreg ransac drop-in scores with calcDists only (got @%s)', func2str(funcDist));
    end
    if ~isfield(ransacCoef, 'REFINE')
        error('pcreg:ransac', 'ransacCoef.REFINE is required (ransac.m:53 reads it)');
    end
    ptNum = size(pts1, 1);
    iterNum = ransacCoef.iterNum;
    triplets = zeros(iterNum, 3);
    for p = 1:iterNum
        sampleIdx = randperm(ptNum);
        triplets(p, :) = sampleIdx(1:3);
    end
    [T, inlierIdx, numSuccess, maxInliers, pct] = pcreg_mex('ransac', double(pts1), double(pts2), ransacCoef, triplets);
end
