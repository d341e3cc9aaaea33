% tests/backprop_test.m of the reference with ONE line changed: the integrator is the drop-in
% CudaRK4Integrator instead of RK4Integrator.  Everything else -- TestOCProblem, PWLinearControl, the
% finite-difference loop -- is the reference's own code, executed by oracle/mlab.
T = 10;
nSteps = 500;
tspan = linspace(0, T, nSteps+1);

x0 = 1;
p.c = 1.5;
p.m = 3;
p.r = .05;
ControlBounds = [0, 1];
prob = TestOCProblem(p, ControlBounds);

integrator = CudaRK4Integrator(tspan);

nControlPts = 21;
v = .5 + .5*rand(nControlPts, 1);
control = PWLinearControl(integrator.t, nControlPts, 1);

epsilon = 1e-4; % finite difference step-size

u = control.compute_u(v);
[x, J] = integrator.compute_states(prob, x0, u);
[lam, dJdu] = integrator.compute_adjoints(prob, u);
dJdv = control.compute_dJdv(dJdu);

dJdv_FD = zeros(nControlPts, 1);
for i=1:nControlPts
   v(i) = v(i) + epsilon;
   u = control.compute_u(v);
   [~, J_FD] = integrator.compute_states(prob, x0, u);
   dJdv_FD(i) = (J_FD - J)/epsilon;
   v(i) = v(i) - epsilon;
end

maxDev = max(dJdv_FD - dJdv);
meanDev = sum(abs(dJdv_FD - dJdv))/nControlPts;

% batched fused path
integrator.attach_control(control);
X0 = [1, 1.5, 2];
out = integrator.evaluate(prob, X0, v);

% infinite horizon (solve_test_problem.m:33)
tspanExtra = linspace(T, 2*T, nSteps+1);
uStar = 0.723368780097983;
integratorInf = CudaRK4InfiniteIntegrator(tspan, tspanExtra, uStar);
uInf = control.compute_u(v);
[xInf, JInf] = integratorInf.compute_states(prob, x0, uInf);
[lamInf, dJduInf] = integratorInf.compute_adjoints(prob, uInf);

% device-side single shooting with the reference's bounds helper
[Lb, Ub] = control.compute_nlp_bounds(prob.ControlBounds);
v0 = control.compute_initial_v(.5);
ss = integrator.single_shooting(prob, X0, v0, Lb, Ub, 'MaxIter', 300, 'TolX', 1e-6);

% soln.x / soln.lam / soln.u (single_shooting.m:128-130) of the three minimisers, interpolated on the device
tQuery = linspace(0, T, 41);
solnSS = integrator.eval_solution(tQuery);

% equilibrium as tests/solve_test_problem.m:21-29 computes it, on the device (one set, then a sweep over r)
lb = [0; -Inf; prob.ControlBounds(1)];
ub = [Inf; Inf; prob.ControlBounds(2)];
[xEq, lamEq, uEq, resnormEq, residualEq, flagEq] = compute_equilibrium_cuda(prob, 2.7, 2.2, .7, lb, ub, p.r);
rSweep = [.04, .05, .06];
[xEqS, lamEqS, uEqS] = integrator.compute_equilibrium(prob, 2.7, 2.2, .7, lb, ub, rSweep, ...
                                                       [p.c p.c p.c; p.m p.m p.m; rSweep]);

% fb_sweep on the grid, discrete adjoint and the v1 continuous costate
fbD = integrator.fb_sweep(prob, X0, .5*ones(1, length(integrator.t)), 'nSWEEPS', 60);
fbC = integrator.fb_sweep(prob, X0, .5*ones(1, length(integrator.t)), 'nSWEEPS', 60, 'costate', 'continuous');

delete(integrator);
delete(integratorInf);
