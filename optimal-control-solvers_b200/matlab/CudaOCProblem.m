classdef CudaOCProblem < OCProblem
   % CUDAOCPROBLEM  An OCProblem (OCProblem/OCProblem.m) whose right-hand sides exist as generated
   % CUDA device functions (the `struct Problem` printed by make_from_symbolic_cuda), compiled with
   % NVRTC for sm_100a when a CudaRK4Integrator first sees the object.

   properties
      ControlBounds
      MinMax = 'Min'      % single_shooting.m:11-15
      Source              % CUDA source text of `namespace ocs_gen { struct Problem ... }`
      IncludeDir = ''     % directory of rk4_kernels.cuh ('' = next to libocs_b200.so)
      nStates
      nControls
      ParamNames          % cellstr, the order of the generated parameter vector p
      Params              % struct; a field may hold one value or one value per trajectory
   end

   methods
      function obj = CudaOCProblem(source, nStates, nControls, params, ControlBounds)
         obj.Source = source;
         obj.nStates = nStates;
         obj.nControls = nControls;
         obj.ParamNames = fieldnames(params);
         obj.Params = params;
         obj.ControlBounds = ControlBounds;
      end

      function P = param_array(obj)
         % np x 1 (shared) or np x Bt (a parameter sweep)
         np = numel(obj.ParamNames);
         nb = 1;
         for k = 1:np
            nb = max(nb, numel(obj.Params.(obj.ParamNames{k})));
         end
         P = zeros(np, nb);
         for k = 1:np
            P(k,:) = obj.Params.(obj.ParamNames{k})(:)';   % scalar expansion
         end
      end

      % The plugin methods only exist on the device: the reference's CPU integrators cannot use this class.
      function dxdt = F(~, ~, ~, ~) %#ok<STOUT>
         error('ocs_b200:deviceOnly', 'CudaOCProblem.F exists as CUDA device code only; use CudaRK4Integrator.');
      end
      function value = dFdx_times_vec(~, ~, ~, ~, ~) %#ok<STOUT>
         error('ocs_b200:deviceOnly', 'CudaOCProblem.dFdx_times_vec exists as CUDA device code only.');
      end
      function value = dFdu_times_vec(~, ~, ~, ~, ~) %#ok<STOUT>
         error('ocs_b200:deviceOnly', 'CudaOCProblem.dFdu_times_vec exists as CUDA device code only.');
      end
   end
end
