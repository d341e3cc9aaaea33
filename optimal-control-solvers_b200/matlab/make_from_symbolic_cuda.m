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
% sub-expressions.  Two halves: the derivations below need the Symbolic Math Toolbox and cannot run in this
% repository's image; the string emission lives in emit_cuda_problem.m, which the tests execute under the
% MATLAB-subset interpreter on the strings `ccode` would return and compile with the sweep kernels.

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

% hand the expressions over as C strings, exactly as ccode prints them; emit_cuda_problem.m (no toolbox needed,
% executed by this repository's tests on such strings) renames the variables and prints the device functions
cF = cell(numel(F), 1);
for k = 1:numel(F)
   cF{k} = ccode(F(k));
end
cX = cell(numel(X), 1);
for k = 1:numel(X)
   cX{k} = ccode(X(k));
end
cU = cell(numel(U), 1);
for k = 1:numel(U)
   cU{k} = ccode(U(k));
end
cCC = {};
if haveCC
   cCC = cell(nControls, 1);
   for k = 1:nControls
      cCC{k} = ccode(CC(k));
   end
end
src = emit_cuda_problem(cF, cX, cU, cCC, nStates, nControls, param_names);

prob = CudaOCProblem(src, nStates, nControls, params, ControlBounds);
end
