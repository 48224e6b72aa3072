"""Host-side setup builders (the mirror of scripts/*/setup/*Setup.jl) against the oracles -- no GPU needed."""
import numpy as np

from conftest import relerr
from oracle import kseg2d_oracle as K2
from oracle import kseg_oracle as K1
from oracle import ns_oracle as NS


def test_fluid_setup_bases_ics_and_layout_match_oracle(pkg):
    cfg = NS.NSConfig(nx=64, sensors_per_axis=8, variance=0.08)
    ops = NS.NSOperators(cfg)
    setup = pkg.setups.FluidSetup(nx=64, sensors_per_axis=8, variance=0.08)
    assert setup.oversampling == cfg.oversampling == 20                       # floor(16 * 64 * 0.02)
    assert setup.sensor_positions == cfg.sensor_positions
    assert np.array_equal(setup.gaussians, NS.prepare_gaussians(cfg, ops, 1))
    assert np.array_equal(setup.gaussians_actuators, NS.prepare_gaussians(cfg, ops, 2))
    assert relerr(setup.ic(3, np.random.default_rng(5)), NS.ic(cfg, ops, 3, np.random.default_rng(5))) < 1e-13
    assert relerr(setup.ic(2), NS.ic(cfg, ops, 2, None)) < 1e-13
    # Julia column-major flattening of a (ny, nx) matrix: element [j, i] at j + ny * i
    a = np.arange(12.0).reshape(3, 4)
    assert np.array_equal(setup._julia_flat(a), NS.to_julia_memory(a))
    assert setup._julia_flat(a)[1] == a[1, 0]


def test_fluid_canned_configs(pkg):
    for make, spa, var in ((pkg.setups.FluidSetup.fluid8, 8, 0.08), (pkg.setups.FluidSetup.fluid16, 16, 0.04)):
        s = make()
        assert (s.nx, s.sensors_per_axis, s.variance, s.oversampling) == (128, spa, var, 40)
        assert s.gaussians.shape == (spa * spa, 128, 128)
        assert np.allclose(s.gaussians.reshape(spa * spa, -1).sum(1), 1.0) and np.allclose(s.gaussians_actuators.max(axis=(1, 2)), 1.0)


def test_keller_segel_setups_match_oracles(pkg):
    c1 = K1.kseg10_16_config()
    s1 = pkg.setups.KellerSegelSetup()
    assert np.array_equal(s1.gaussians, K1.prepare_rectangles(c1))
    assert np.array_equal(s1.gaussians_actuators, K1.prepare_rectangles(c1)[np.asarray(c1.actuators_to_sensors) - 1])
    c2 = K2.KSeg2DConfig()
    s2 = pkg.setups.KellerSegel2DSetup()
    assert np.array_equal(s2.gaussians, K2.prepare_boxes(c2)) and s2.gaussians.shape == (256, 128, 128)
    assert np.all(s2.gaussians.reshape(256, -1).sum(1) == 25)
    # x fastest in the C ABI's flattening of the Julia (nx, ny) field
    a = np.arange(6.0).reshape(2, 3)
    assert np.array_equal(s2._julia_flat(a), a.T.reshape(-1))


def test_shard_ranges_cover_uneven_batches(pkg):
    assert [pkg.parallel.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
