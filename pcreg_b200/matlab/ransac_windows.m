function [T, inlierIdx, numSuccess, maxInliers, pct] = ransac_windows(pts1, pts2, ransacCoef, seeds)
%% ransac_windows -- ransac.m for a whole batch of matching windows in ONE GPU call.
% Replaces the parfor over windows of slideMatchingWindow_v2.m:178-198 / completeExperiment.m:265-278:
%     parfor w = 1:W, [T{w}, inl{w}, ...] = ransac(pts1{w}, pts2{w}, ransacCoef, @estimateTransform, @calcDists); end
% pts1, pts2: 1xW cell arrays of P_w x 3 matched locations (model / surface); ransacCoef as for ransac.m
% (iterNum, thDist, thInlrRatio, REFINE); seeds: W x 1 (default randi(2^31, W, 1): the samples are drawn on the
% device by the documented counter-based sampler of include/pcreg.h, one stream per window).
% Outputs: T 1xW cell (4x4 or [] exactly where ransac.m returns []), inlierIdx 1xW cell (column vectors, 1-based,
% relative to the window), numSuccess / maxInliers / pct W x 1.
    W = numel(pts1);
    if nargin < 4, seeds = randi(2^31, W, 1); end
    counts = cellfun(@(p) size(p, 1), pts1(:));
    offsets = [0; cumsum(counts)];
    [T16, mask, numSuccess, maxInliers, pct] = pcreg_mex('ransac_batch', double(vertcat(pts1{:})), double(vertcat(pts2{:})), ...
                                                         offsets, ransacCoef, double(seeds(:)));
    T = cell(1, W); inlierIdx = cell(1, W);
    for w = 1:W
        if any(isnan(T16(:, w)))
            T{w} = []; inlierIdx{w} = [];
        else
            T{w} = reshape(T16(:, w), 4, 4);
            inlierIdx{w} = find(mask(offsets(w) + 1 : offsets(w + 1)));
        end
    end
end
