#!/bin/sh
# Regenerates the committed golden fixtures from the UNMODIFIED reference (strict-IEEE build, oracle/Makefile).
# Run in the build container (needs /root/reference): sh tests/golden/make_golden.sh
set -e
cd "$(dirname "$0")"
make -C ../../oracle ref >/dev/null
R=../../oracle/_ref/sph_ref_strict
# one integrate() on zeroed derivatives, with neighbour lists
$R snapshot --config hello            --n 500 --solver asym --jitter 7 --neighbours --in hello_in.snap   --out hello_out.snap --no-lut --lut lut.snap
$R snapshot --config collision_preset --n 500 --jitter 3 --neighbours           --in collision_in.snap --out collision_out.snap --no-lut
$R snapshot --config preset           --n 500 --neighbours                      --in preset_in.snap    --out preset_out.snap --no-lut
$R snapshot --config fluid            --n 500 --jitter 5 --neighbours           --in fluid_in.snap     --out fluid_out.snap --no-lut
# ideal gas ball (IdealGasEos, Eos.cpp:42-45)
$R snapshot --config gas              --n 500 --jitter 9 --neighbours           --in gas_in.snap       --out gas_out.snap --no-lut
# three full time steps (PredictorCorrector / EulerExplicit + MultiCriterion)
$R snapshot --config collision_preset --n 500 --jitter 3 --steps 3 --out collision_pc3.snap --no-lut
$R snapshot --config hello --n 500 --solver asym --jitter 7 --steps 3 --out hello_pc3.snap --no-lut
$R snapshot --config fluid --n 500 --jitter 5 --steps 3 --integrator euler --out fluid_euler3.snap --no-lut
$R snapshot --config gas --n 500 --jitter 9 --steps 3 --out gas_pc3.snap --no-lut
# the library-default SymmetricSolver on the same input must agree with the asymmetric path (Solvers.cpp:178-216)
$R snapshot --config hello --n 500 --solver sym --jitter 7 --out hello_sym_out.snap --no-lut
# the XSph term (SPH_USE_XSPH, epsilon 0.5): one integrate() and three PredictorCorrector steps
$R snapshot --config collision_preset --n 500 --jitter 3 --xsph 0.5 --neighbours --in xsph_in.snap --out xsph_out.snap --no-lut
$R snapshot --config collision_preset --n 500 --jitter 3 --xsph 0.5 --steps 3 --out xsph_pc3.snap --no-lut
# the delta-SPH terms (SPH_USE_DELTASPH): one integrate() and three PredictorCorrector steps of the solid (delta 0.1, alpha 0.05),
# three steps of the fluid with the library defaults (the in-state of the fluid run is fluid_in.snap)
$R snapshot --config collision_preset --n 500 --jitter 3 --deltasph --deltasph-delta 0.1 --deltasph-alpha 0.05 --neighbours --in deltasph_in.snap --out deltasph_out.snap --no-lut
$R snapshot --config collision_preset --n 500 --jitter 3 --deltasph --deltasph-delta 0.1 --deltasph-alpha 0.05 --steps 3 --out deltasph_pc3.snap --no-lut
$R snapshot --config fluid --n 500 --jitter 5 --deltasph --steps 3 --out deltasph_fluid_pc3.snap --no-lut
# the artificial stress (SPH_AV_USE_STRESS, factor 0.2): one integrate() and three PredictorCorrector steps
$R snapshot --config collision_preset --n 500 --jitter 3 --stress-av --stress-av-factor 0.2 --neighbours --in stressav_in.snap --out stressav_out.snap --no-lut
$R snapshot --config collision_preset --n 500 --jitter 3 --stress-av --stress-av-factor 0.2 --steps 3 --out stressav_pc3.snap --no-lut
# FrozenParticles boundary condition: the impactor (flag 1) frozen and everything within 0.3 h of / outside a sphere of 90 km
$R snapshot --config collision_preset --n 500 --jitter 3 --frozen-flag 1 --frozen-domain 9e4 --frozen-radius 0.3 --out frozen_out.snap --no-lut
# self-gravity (IGravity::build + evalSelfGravity on a zeroed buffer): brute force and Barnes-Hut, softened and point-like
$R gravity --config hello --n 500 --jitter 7 --gravity brute --out gravity_brute.snap
$R gravity --config hello --n 500 --jitter 7 --gravity bh --theta 0.5 --order 3 --no-lut --out gravity_bh.snap
$R gravity --config hello --n 500 --jitter 7 --gravity bh --theta 0.8 --order 2 --point --out gravity_bh_point.snap
$R gravity --config hello --n 500 --jitter 7 --gravity brute --point --out gravity_brute_point.snap
ls -la *.snap
