% RK4InfiniteIntegrator as solve_test_problem.m builds it (tspan, tspanExtra, uStar), driven
% explicitly (the reference script never passes it on).
T = 10; nSteps = 500;
tspan = linspace(0, T, nSteps+1);
tspanExtra = linspace(T, 2*T, nSteps+1);
x0 = 1;
p.c = 1.5; p.m = 3; p.r = .05;
prob = TestOCProblem(p, [0, 1]);
uStar = 0.723368780097983;
integrator = RK4InfiniteIntegrator(tspan, tspanExtra, uStar);
K = 101;
control = PWLinearControl(integrator.t, K, 1);
v = uStar + .1*(rand(K, 1) - .5);
u = control.compute_u(v);
[x, J] = integrator.compute_states(prob, x0, u);
[lam, dJdu] = integrator.compute_adjoints(prob, u);
dJdv = control.compute_dJdv(dJdu);
lam2 = integrator.integrator2.compute_adjoints(prob, integrator.uStar);
x2 = integrator.integrator2.xK(:,:,1);
