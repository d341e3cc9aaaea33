classdef CudaRK4Integrator < Integrator
   % CUDARK4INTEGRATOR  Drop-in replacement for RK4Integrator (Integrator/RK4Integrator.m) that
   % runs the forward state sweep and the backward discrete-adjoint sweep on an NVIDIA B200
   % through libocs_b200.so (C ABI: include/ocs_b200.h, MEX gateway: ocs_b200_mex.c).
   %
   %   integrator = CudaRK4Integrator(tspan);                       % as RK4Integrator(tspan)
   %   integrator = CudaRK4Integrator(tspan, 'Device', 0:7);        % the batch sharded over eight GPUs, one object
   %   soln = single_shooting(prob, x0, tspan, nCONTROL_PTS, 'Integrator', integrator);
   %
   % Same property (t) and methods as the reference class; every array may carry a trailing
   % batch dimension (one page/column per trajectory):
   %   [x, J]      = compute_states(obj, prob, x0, u)      x0: n x Bt,  u: nc x G [x Bt]
   %   [lam, dJdu] = compute_adjoints(obj, prob, u [, lamT])
   % Batched fused path (compute_u and compute_dJdv folded into the sweeps):
   %   obj.attach_control(control);  out = obj.evaluate(prob, x0, v)   % J, dJdv, lam0, xT, status
   %
   % prob is either the reference's TestOCProblem (recognised and mapped to the built-in device
   % code) or a CudaOCProblem produced by make_from_symbolic_cuda.  There is no CPU fallback:
   % an OCProblem whose F / dFdx_times_vec / dFdu_times_vec only exist as MATLAB methods is
   % rejected with an error.

   properties
      t        % tspan augmented with the interval midpoints (RK4Integrator.m:4)
      h        % diff(tspan)
      nSTEPS
   end

   properties (SetAccess = protected)
      handle = uint64(0)   % ocs_b200_handle*, owned by this object
      control = []         % attached Control object (its B matrix lives on the device)
      boundProb = []       % problem whose device code is currently bound
      boundParams = []
      nStates = 0
      nControls = 0
      Bt = 0               % batch size of the last compute_states
      N2 = 0               % tail steps (CudaRK4InfiniteIntegrator)
   end

   methods
      function obj = CudaRK4Integrator(tspan, varargin)
         p = inputParser;
         p.addParameter('Device', 0);          % a CUDA ordinal, or a vector of them: trajectories are then sharded
                                               % over those GPUs inside the one handle (ocs_b200_create_multi)
         p.addParameter('Math', 'fast');       % 'fast' (FMA contraction) or 'strict' (MATLAB-like rounding)
         p.addParameter('tspanExtra', []);
         parse(p, varargin{:});
         mathMode = double(strcmp(p.Results.Math, 'strict'));
         obj.handle = ocs_b200_mex('create', tspan(:)', p.Results.tspanExtra(:)', p.Results.Device, mathMode);
         obj.nSTEPS = length(tspan) - 1;
         obj.N2 = max(length(p.Results.tspanExtra) - 1, 0);
         [obj.t, obj.h] = ocs_b200_mex('get_grid', obj.handle, 0, obj.nSTEPS);
      end

      function delete(obj)
         if obj.handle ~= 0
            ocs_b200_mex('destroy', obj.handle);
            obj.handle = uint64(0);
         end
      end

      function attach_control(obj, control)
         % Fuse u = reshape(v,nControls,[])*B and dJdv = dJdu*B' into the sweeps.  B is built by the
         % (unchanged) reference Control class on obj.t, so control-grid indices are MATLAB's own.
         obj.control = control;
         if isempty(control)
            ocs_b200_mex('set_control_basis', obj.handle, []);
         else
            ocs_b200_mex('set_control_basis', obj.handle, control.B);
            % compute_uFunc on the device (eval_solution): the interpolant each reference class builds
            if isa(control, 'PWConstantControl')          % PWConstantControl.m:58-61 'previous'
               ocs_b200_mex('set_control_ufunc', obj.handle, 1, control.intervalStarts);
            elseif isa(control, 'PWLinearControl')        % PWLinearControl.m:74-77 'linear'
               ocs_b200_mex('set_control_ufunc', obj.handle, 2, control.controlPts);
            elseif isa(control, 'ChebyshevControl')       % basis recurrence ChebyshevControl.m:23-30
               ocs_b200_mex('set_control_ufunc', obj.handle, 3, [obj.t(1), obj.t(end)]);
            end
         end
      end

      function soln = eval_solution(obj, tq, first, count, segment)
         % soln.x(tq), soln.lam(tq), soln.u(tq) of functions/single_shooting.m:128-130 (pchip interpolants of the
         % node states / costates, the control's uFunc) for trajectories first .. first+count-1 (1-based),
         % evaluated on the device.  Arrays are rows x numel(tq) x count.
         if nargin < 3 || isempty(first), first = 1; end
         if nargin < 4 || isempty(count), count = obj.Bt - first + 1; end
         if nargin < 5, segment = 0; end
         [soln.x, soln.lam, soln.u] = ocs_b200_mex('eval_solution', obj.handle, segment, first - 1, count, tq(:)', ...
                                                   obj.nStates, obj.nControls);
      end

      function [x, J] = compute_states(obj, prob, x0, u)
         obj.bind(prob);
         ocs_b200_mex('set_control_u', obj.handle, u);
         [J, ~, ~] = ocs_b200_mex('compute_states', obj.handle, x0);
         obj.Bt = size(x0, 2);
         if nargout > 0 && (obj.Bt == 1 || nargout > 1)
            x = ocs_b200_mex('fetch_states', obj.handle, 0, 0, obj.Bt, obj.nStates + 1, obj.nSTEPS);
         else
            x = [];
         end
      end

      function [lam, dJdu] = compute_adjoints(obj, prob, u, lamT)
         obj.bind(prob);
         if nargin < 4
            lamT = [];
         end
         m = obj.nStates + 1;
         ocs_b200_mex('compute_adjoints', obj.handle, lamT, m, obj.Bt, 0);
         if nargout > 1
            [lam, dJdu] = ocs_b200_mex('fetch_adjoints', obj.handle, 0, 0, obj.Bt, m, obj.nSTEPS, obj.nControls);
         else
            lam = ocs_b200_mex('fetch_adjoints', obj.handle, 0, 0, obj.Bt, m, obj.nSTEPS, obj.nControls);
         end
      end

      function out = evaluate(obj, prob, x0, v)
         % One single_shooting objective/gradient evaluation (single_shooting.m:137-150) for every
         % column of x0: v is (nControls*K) x 1 (shared) or x Bt.
         obj.bind(prob);
         ocs_b200_mex('set_control_v', obj.handle, v);
         [out.J, out.xT, out.status, out.lam0, out.dJdv] = ocs_b200_mex('evaluate', obj.handle, x0, [], ...
                                                                        size(obj.control.B, 1)*obj.nControls);
         obj.Bt = size(x0, 2);
      end

      function out = single_shooting(obj, prob, x0, v0, Lb, Ub, varargin)
         % The fmincon solve of functions/single_shooting.m:114 for every column of x0, on the device
         % (spectral projected gradient with the box [Lb, Ub] of control.compute_nlp_bounds).
         p = inputParser;
         p.addParameter('FreeInitStates', []);    % 1-based, as in single_shooting.m:28
         p.addParameter('FreeStateBounds', []);
         p.addParameter('MaxIter', 300);
         p.addParameter('TolX', 1e-6);
         % [A, b] = control.compute_lincon(prob.ControlBounds) of single_shooting.m:106-111 for a control whose
         % coefficient box does not bound u(t) (ChebyshevControl): ControlBounds (nc x 2) at every grid point
         p.addParameter('GridControlBounds', []);
         parse(p, varargin{:});
         obj.bind(prob);
         ocs_b200_mex('set_solve_control_bounds', obj.handle, p.Results.GridControlBounds);
         free = int32(p.Results.FreeInitStates(:)') - 1;
         fb = p.Results.FreeStateBounds;
         if isempty(free), flb = []; fub = []; else, flb = fb(:,1); fub = fb(:,2); end
         [out.v, out.J, out.iters, out.converged] = ocs_b200_mex('single_shooting', obj.handle, x0, v0, Lb, Ub, ...
            free, flb, fub, p.Results.MaxIter, p.Results.TolX);
         obj.Bt = size(x0, 2);
      end

      function [value, index] = reduce_J(obj, op)
         % Reduction of the last sweep's J over the batch on the device(s): op 'sum', 'min' or 'max';
         % index (1-based) is the arg min / arg max -- the best of a sweep over control candidates or parameter sets.
         codes = struct('sum', 0, 'min', 1, 'max', 2);
         [value, index] = ocs_b200_mex('reduce_J', obj.handle, codes.(op));
      end

      function viol = solve_violation(obj)
         % Largest violation of 'GridControlBounds' per trajectory after the last constrained single_shooting.
         viol = ocs_b200_mex('get_solve_violation', obj.handle);
      end

      function [xStar, lamStar, uStar, resnorm, residual, exitflag] = compute_equilibrium(obj, prob, xGuess, lamGuess, uGuess, lb, ub, r, params)
         % functions/compute_equilibrium.m for every parameter set at once (one GPU thread per set).
         % params (np x nsets) overrides prob's parameters for a sweep; guesses may be columns or (. x nsets).
         obj.bind(prob);
         if nargin < 9 || isempty(params), params = obj.boundParams; end
         [y, resnorm, residual, exitflag] = ocs_b200_mex('compute_equilibrium', obj.handle, params, ...
            [xGuess; lamGuess; uGuess], lb, ub, r, 4000, 1e-13);     % MaxIter 4000: compute_equilibrium.m:24
         n = obj.nStates;
         xStar = y(1:n, :);  lamStar = y(n+1:2*n, :);  uStar = y(2*n+1:end, :);
      end

      function out = fb_sweep(obj, prob, x0, u0, varargin)
         % The sweep loop of functions/fb_sweep.m:79-115 on the RK4 grid, for every column of x0.
         p = inputParser;
         p.addParameter('uRelTol', 1e-7);   % fb_sweep.m:16-17
         p.addParameter('uAbsTol', 1e-7);
         p.addParameter('nSWEEPS', 50);     % fb_sweep.m:20
         p.addParameter('relax', 1);        % 1 = the reference's pure replacement (fb_sweep.m:96)
         p.addParameter('costate', 'discrete');   % 'continuous' = v1 adjointRHS by RK4 (compute_x_lam.m:11-12)
         parse(p, varargin{:});
         obj.bind(prob);
         [out.J, out.sweeps, out.converged, out.maxChange] = ocs_b200_mex('fb_sweep', obj.handle, x0, u0, ...
            prob.ControlBounds, p.Results.uRelTol, p.Results.uAbsTol, p.Results.nSWEEPS, p.Results.relax, ...
            double(strcmp(p.Results.costate, 'continuous')));
         obj.Bt = size(x0, 2);
         out.u = ocs_b200_mex('get_control_u', obj.handle, 0, obj.Bt, obj.nControls, length(obj.t));
      end
   end

   methods (Access = protected)
      function bind(obj, prob)
         if isa(prob, 'CudaOCProblem')
            params = prob.param_array();
            if ~isequal(obj.boundProb, prob)
               ocs_b200_mex('set_problem_source', obj.handle, prob.Source, prob.nStates, prob.nControls, ...
                            size(params, 1), prob.IncludeDir);
               obj.boundProb = prob;  obj.boundParams = [];
               obj.nStates = prob.nStates;  obj.nControls = prob.nControls;
               obj.after_bind();
               if ~isempty(obj.control), obj.attach_control(obj.control); end
            end
         elseif isa(prob, 'TestOCProblem')      % tests/TestOCProblem.m -> built-in device code, params {c, m, r}
            params = [prob.c(:)'; prob.m(:)'; prob.r(:)'];
            if ~isequal(obj.boundProb, prob)
               ocs_b200_mex('set_problem_builtin', obj.handle, 0);
               obj.boundProb = prob;  obj.boundParams = [];
               obj.nStates = 1;  obj.nControls = 1;
               obj.after_bind();
               if ~isempty(obj.control), obj.attach_control(obj.control); end
            end
         else
            error('ocs_b200:problem', ['CudaRK4Integrator needs device code for the problem: pass a ' ...
                  'CudaOCProblem (make_from_symbolic_cuda) or a TestOCProblem. There is no CPU fallback.']);
         end
         if ~isequal(params, obj.boundParams)
            ocs_b200_mex('set_params', obj.handle, params);
            obj.boundParams = params;
         end
      end

      function after_bind(~)
      end
   end
end
