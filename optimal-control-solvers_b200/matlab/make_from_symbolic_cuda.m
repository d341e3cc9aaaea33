function prob = make_from_symbolic_cuda(symObjective, symStateRHS, nStates, nControls, params, ControlBounds)
% MAKE_FROM_SYMBOLIC_CUDA  CUDA backend of make_from_symbolic (functions/make_from_symbolic.m).
%
% Same inputs as the reference front end (:1-8): a symbolic objective integrand and state
% right-hand side written in t, x1..xn, lam1..lamn, u1..um and the fields of params.  Instead of
% matlabFunction handles (:27-31) it prints the v2 plugin triple
%     F = [g; f],   dFdx_times_vec = (dg/dx)'v_x + v_J df/dx,   dFdu_times_vec = (dg/du)'v_x + v_J df/du
% and the clamped control characterisation (:20-23,111) as CUDA __device__ functions (ccode) and
% returns a CudaOCProblem that CudaRK4Integrator compiles with NVRTC for sm_100a.
%
% This MATLAB twin of optimal-control-solvers_b200/codegen/make_from_symbolic.py emits the un-hoisted form
% (ntt = 0 for every hoisting mode); the Python backend additionally tabulates batch-uniform
% sub-expressions.  Untested in this repository's image (no MATLAB); the Python backend is the
% one covered by tests.

param_names = fieldnames(params);
t = sym('t');
x = sym('x', [nStates, 1]);
lam = sym('lam', [nStates, 1]);
u = sym('u', [nControls, 1]);
vx = sym('v', [nStates, 1]);
vJ = sym('vJ');
f = symObjective;
g = symStateRHS(:);

F = [g; f];
X = jacobian(g, x).'*vx + vJ*jacobian(f, x).';
U = jacobian(g, u).'*vx + vJ*jacobian(f, u).';
H = f + lam.'*g;                                   % make_from_symbolic.m:11
dHdu = jacobian(H, u).';                           % :17
try
   sol = solve(dHdu == 0, u);                      % :20
   if nControls == 1
      CC = sol;
   else
      CC = sym(zeros(nControls, 1));
      for c = 1:nControls
         CC(c) = sol.(char(u(c)));
      end
   end
   haveCC = ~isempty(CC);
catch
   haveCC = false;
end

   function s = body(exprs, out)
      s = '';
      for k = 1:numel(exprs)
         c = ccode(exprs(k));
         c = regexprep(c, '^\s*\w+\s*=\s*', '');
         c = regexprep(c, ';\s*$', '');
         s = sprintf('%s        %s[%d] = %s;\n', s, out, k-1, rename(c));
      end
   end

   function c = rename(c)
      for k = nStates:-1:1
         c = regexprep(c, sprintf('\\<lam%d\\>', k), sprintf('lam[%d]', k-1));
         c = regexprep(c, sprintf('\\<x%d\\>', k), sprintf('y[%d]', k-1));
         c = regexprep(c, sprintf('\\<v%d\\>', k), sprintf('v[%d]', k-1));
      end
      c = regexprep(c, '\<vJ\>', sprintf('v[%d]', nStates));
      for k = nControls:-1:1
         c = regexprep(c, sprintf('\\<u%d\\>', k), sprintf('u[%d]', k-1));
      end
      for k = 1:numel(param_names)
         c = regexprep(c, sprintf('\\<%s\\>', param_names{k}), sprintf('p[%d]', k-1));
      end
   end

sig = 'const double *tt, double t, const double *y, const double *u, const double *p';
src = sprintf('namespace ocs_gen {\nstruct Problem {\n');
src = [src sprintf('    static constexpr int n = %d, nc = %d, np = %d;\n', nStates, nControls, numel(param_names))];
src = [src sprintf('    __host__ __device__ static constexpr int ntt(int) { return 0; }\n')];
src = [src sprintf('    template <int HM> __device__ __forceinline__ static void table(double, const double *, const double *, double *) {}\n')];
src = [src sprintf('    template <int HM> __device__ __forceinline__ static void F(%s, double *o)\n    {\n        (void)tt; (void)t; (void)y; (void)u; (void)p;\n%s    }\n', sig, body(F, 'o'))];
src = [src sprintf('    template <int HM> __device__ __forceinline__ static void X(%s, const double *v, double *o)\n    {\n        (void)tt; (void)t; (void)y; (void)u; (void)p; (void)v;\n%s    }\n', sig, body(X, 'o'))];
src = [src sprintf('    template <int HM> __device__ __forceinline__ static void U(%s, const double *v, double *o)\n    {\n        (void)tt; (void)t; (void)y; (void)u; (void)p; (void)v;\n%s    }\n', sig, body(U, 'o'))];
src = [src sprintf('    template <int HM> __device__ __forceinline__ static void ControlChar(const double *tt, double t, const double *y, const double *lam, const double *p, const double *bounds, double *o)\n    {\n        (void)tt; (void)t; (void)y; (void)lam; (void)p;\n')];
if haveCC
   src = [src body(CC, 'o')];
   for c = 1:nControls                                  % value = min(umax, max(umin, value))   :111
      src = [src sprintf('        o[%d] = fmin(bounds[%d], fmax(bounds[%d], o[%d]));\n', c-1, 2*c-1, 2*c-2, c-1)]; %#ok<AGROW>
   end
else
   src = [src sprintf('        for (int c = 0; c < nc; ++c) o[c] = nan("");\n')];
end
src = [src sprintf('    }\n};\n}  // namespace ocs_gen\n')];

prob = CudaOCProblem(src, nStates, nControls, params, ControlBounds);
end
