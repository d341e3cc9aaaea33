% BUILD_MEX  Compiles the MEX gateway against libocs_b200.so (built by optimal-control-solvers_b200/build.py).
here = fileparts(mfilename('fullpath'));
pkg = fileparts(here);
mex('-R2017b', ['-I' fullfile(pkg, '..', 'include')], fullfile(here, 'ocs_b200_mex.c'), ...
    ['-L' pkg], '-locs_b200', ['LDFLAGS=$LDFLAGS -Wl,-rpath,' pkg], '-outdir', here);
