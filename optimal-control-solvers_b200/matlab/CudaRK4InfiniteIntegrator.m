classdef CudaRK4InfiniteIntegrator < CudaRK4Integrator
   % CUDARK4INFINITEINTEGRATOR  Drop-in replacement for RK4InfiniteIntegrator
   % (Integrator/RK4InfiniteIntegrator.m:12-30): the main horizon tspan chained with a tail
   % tspanExtra under the constant control uStar; J = J1 + J2 and the tail's lam(:,1) is the
   % terminal costate of the main horizon.  Both sweeps run on the device.
   %
   %   integrator = CudaRK4InfiniteIntegrator(tspan, tspanExtra, uStar);

   properties
      uStar
   end

   methods
      function obj = CudaRK4InfiniteIntegrator(tspan, tspanExtra, uStar, varargin)
         obj@CudaRK4Integrator(tspan, 'tspanExtra', tspanExtra, varargin{:});
         obj.uStar = uStar(:);
      end

      function [lam, dJdu] = compute_adjoints(obj, prob, u)
         % RK4InfiniteIntegrator.m:27-30 takes no lamT: the tail sweep supplies it.
         if nargout > 1
            [lam, dJdu] = compute_adjoints@CudaRK4Integrator(obj, prob, u);
         else
            lam = compute_adjoints@CudaRK4Integrator(obj, prob, u);
         end
      end
   end

   methods (Access = protected)
      function after_bind(obj)
         ocs_b200_mex('set_tail_control', obj.handle, obj.uStar);
      end
   end
end
