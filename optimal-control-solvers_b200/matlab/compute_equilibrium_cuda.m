function [xStar, lamStar, uStar, resnorm, residual, exitflag] = ...
   compute_equilibrium_cuda(prob, xGuess, lamGuess, uGuess, lb, ub, r, params)
% COMPUTE_EQUILIBRIUM_CUDA  functions/compute_equilibrium.m on the GPU, batched over parameter sets.
%
% Same arguments and outputs as the reference's compute_equilibrium(prob, xGuess, lamGuess, uGuess, lb, ub, r);
% the optional params (np x nsets) solves one steady state per column (r may then be 1 x nsets, the guesses
% (. x nsets)).  lsqnonlin is replaced by one Levenberg-Marquardt solve per GPU thread
% (ocs_b200_compute_equilibrium); prob is a TestOCProblem or a CudaOCProblem.
   if nargin < 8, params = []; end
   integrator = CudaRK4Integrator([0 1]);     % only the bound problem is used, not the grid
   [xStar, lamStar, uStar, resnorm, residual, exitflag] = ...
      integrator.compute_equilibrium(prob, xGuess, lamGuess, uGuess, lb, ub, r, params);
   delete(integrator);
end
