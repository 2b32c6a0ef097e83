function matches = getMatches(descSurface, descModel, par)
%% getMatches -- drop-in for the reference's getMatches.m on the GPU.
% Put this directory ahead of the reference on the MATLAB path.  Same inputs (par.UNNORMALIZE, norm_factor,
% CHANGE_METRIC, metric_factor, Method, MatchThreshold, MaxRatio, Metric, Unique, VERBOSE) and output (matches:
% P x 2 uint32 index pairs into descSurface / descModel, as matchFeatures returns them).  The descriptor weighting
% (getMatches.m:22-41) and the whole matchFeatures call run in one GPU call.  par.Method is accepted and ignored: the
% search is exhaustive -- the exact result that matchFeatures' 'Approximate' kd-forest (the reference's setting)
% approximates, and the same as its 'Exhaustive' method up to the summation order of the scores.
    if isfield(par, 'VERBOSE')
        VERBOSE = par.VERBOSE;
    else
        VERBOSE = 1;                                    % getMatches.m:5-9
    end
    t0 = tic;
    matches = pcreg_mex('get_matches', double(descSurface), double(descModel), par);
    if VERBOSE
        fprintf('Calculated matches in %0.1f seconds...\n', toc(t0));      % getMatches.m:53-55
    end
end
