% ChebyshevControl degree 16 (K = 17).  MATLAB itself refuses to instantiate the class as shipped
% (it lacks compute_uFunc); the interpreter does not enforce abstractness, the numerics are the file's.
T = 10; nSteps = 400;
tspan = linspace(0, T, nSteps+1);
x0 = 0.8;
p.c = 1.5; p.m = 3; p.r = .05;
prob = TestOCProblem(p, [0, 1]);
integrator = RK4Integrator(tspan);
K = 17;
control = ChebyshevControl(integrator.t, K, 1);
v = [.5; .05*(rand(K-1, 1) - .5)];
u = control.compute_u(v);
[x, J] = integrator.compute_states(prob, x0, u);
[lam, dJdu] = integrator.compute_adjoints(prob, u);
dJdv = control.compute_dJdv(dJdu);
B = control.B;
v0 = control.compute_initial_v(.25);
