% Driver (MATLAB syntax, run by oracle/mlab against the UNMODIFIED reference classes):
% PWConstantControl K=100 on the backprop_test grid -- control-grid indices and sweeps.
T = 10; nSteps = 500;
tspan = linspace(0, T, nSteps+1);
x0 = 1;
p.c = 1.5; p.m = 3; p.r = .05;
prob = TestOCProblem(p, [0, 1]);
integrator = RK4Integrator(tspan);
K = 100;
control = PWConstantControl(integrator.t, K, 1);
v = .5 + .5*rand(K, 1);
u = control.compute_u(v);
[x, J] = integrator.compute_states(prob, x0, u);
[lam, dJdu] = integrator.compute_adjoints(prob, u);
dJdv = control.compute_dJdv(dJdu);
B = control.B;
t = integrator.t;
h = integrator.h;
intervalStarts = control.intervalStarts;
[Lb, Ub] = control.compute_nlp_bounds([0, 1]);
v0 = control.compute_initial_v(.25);
