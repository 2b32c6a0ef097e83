function pts_tf = quickTF(pts, TF)
%% quickTF -- drop-in for the reference's quickTF.m (quickTF.m:1-8): [pts 1] * TF on the GPU, class of pts kept.
% Worth it for whole clouds (AutoAlignPointclouds2.m:25 applies it to the 16 M-point model); small sets are as fast on the host.
    pts_tf = pcreg_mex('quick_tf', pts, double(TF), 0);
end
