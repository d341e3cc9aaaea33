% PWLinearControl K=101 (the default control of single_shooting as run by solve_test_problem.m)
% with an explicit terminal costate (4th argument of compute_adjoints).
T = 10; nSteps = 500;
tspan = linspace(0, T, nSteps+1);
x0 = 1.7;
p.c = 1.5; p.m = 3; p.r = .05;
prob = TestOCProblem(p, [0, 1]);
integrator = RK4Integrator(tspan);
K = 101;
control = PWLinearControl(integrator.t, K, 1);
v = rand(K, 1);
u = control.compute_u(v);
[x, J] = integrator.compute_states(prob, x0, u);
[lam, dJdu] = integrator.compute_adjoints(prob, u);
dJdv = control.compute_dJdv(dJdu);
lamT = [0.3; 1];
[lam_T, dJdu_T] = integrator.compute_adjoints(prob, u, lamT);
B = control.B;
controlPts = control.controlPts;
v0 = control.compute_initial_v(.25);
