"""CPU tests (no GPU): the host-side mirror of the reference interface, the parameter packing, the C-ABI library
(loads, exports every symbol of include/rwmpt.h, argument validation that needs no device) and the Philox
known-answer vectors."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import rwm_pt_pytorch_b200 as P
from rwm_pt_pytorch_b200 import _lib
from rwm_pt_pytorch_b200.algorithms import (RandomWalkMH_GPU_Optimized, RandomWalkMetropolis,
                                            ParallelTemperingRWM_GPU_Optimized)
from rwm_pt_pytorch_b200.interfaces import MCMCSimulation_GPU
from rwm_pt_pytorch_b200.proposal_distributions import NormalProposal, LaplaceProposal, UniformRadiusProposal
from rwm_pt_pytorch_b200 import target_distributions as td
from tests._util import (load_golden, golden_names, make_product_targets, target_key_of, specs_equal,
                         philox4x32_10_py, PHILOX_KAT)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU = torch.device("cpu")


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "rwmpt.h")).read()
    declared = set(re.findall(r"\b(rwmpt_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rwmpt_version() == 100
    assert lib.rwmpt_sizeof_run_args() == C.sizeof(_lib.RunArgs)


def test_abi_argument_validation_without_device():
    lib = _lib.load()
    assert lib.rwmpt_rwm_run(None, None) == _lib.EINVAL
    assert b"NULL" in lib.rwmpt_last_error()
    a = _lib.RunArgs()
    a.target = _lib.TargetT(99, 4, 1, 16)
    assert lib.rwmpt_pt_run(C.byref(a), None) == _lib.EINVAL
    assert b"unknown target family" in lib.rwmpt_last_error()
    a.target = _lib.TargetT(_lib.T_EVEN_ROSENBROCK, 5, 1, 32)
    assert lib.rwmpt_pt_run(C.byref(a), None) == _lib.EINVAL
    with pytest.raises(ValueError):
        _lib.check(lib.rwmpt_pt_run(C.byref(a), None))
    e = C.c_int32()
    assert lib.rwmpt_pick_lanes(20, 8, 1024, 0, C.byref(e)) == 4 and e.value == 5
    assert lib.rwmpt_pick_lanes(100000, 1, 1, 0, C.byref(e)) == _lib.ENOTSUP
    with pytest.raises(NotImplementedError):
        _lib.check(lib.rwmpt_pick_lanes(100000, 1, 1, 0, C.byref(e)))


def test_swap_round_count_matches_reference_rule():
    lib = _lib.load()
    for off, n, burn, se in [(0, 1200, 200, 10), (0, 300, 0, 3), (0, 700, 100, 5), (100, 50, 120, 7), (0, 9, 0, 10)]:
        want = sum(1 for s in range(off + 1, off + n + 1) if s % se == 0 and s > burn)
        assert lib.rwmpt_count_swap_rounds(off, n, burn, se) == want


def test_philox_known_answers_python():
    for ctr, key, out in PHILOX_KAT:
        assert philox4x32_10_py(ctr, key) == out


def test_product_targets_carry_the_reference_constants():
    """Constructing the product's targets like the reference does yields bit-identical parameters."""
    targets = make_product_targets()
    for name in golden_names("logp_"):
        spec, _ = load_golden(name)
        t = targets[target_key_of(name)]
        assert specs_equal(t.spec(), spec), name
        p = t.pack()
        assert p.dtype == torch.float32 and p.numel() >= _lib.PARAM_HEADER


def test_target_names_and_errors():
    assert td.RoughCarpetDistributionTorch(3, device=CPU).get_name() == "RoughCarpetTorch"
    assert td.RoughCarpetDistributionTorch(3, device=CPU, mode_centers=[-4.0, 0.0, 4.0]).get_name() == "RoughCarpetTorchCustom"
    assert td.ThreeMixtureDistributionTorch(3, scaling=True, device=CPU).get_name() == "ThreeMixtureTorchScaled"
    assert td.NealFunnelTorch(7, device=CPU).get_name() == "NealFunnelTorch_D7"
    assert td.HybridRosenbrockTorch(3, 5, device=CPU).dim == 11
    with pytest.raises(ValueError):
        td.EvenRosenbrockTorch(5, device=CPU)
    with pytest.raises(ValueError):
        td.FullRosenbrockTorch(1, device=CPU)
    with pytest.raises(ValueError):
        td.RoughCarpetDistributionTorch(3, device=CPU, mode_weights=[0.5, 0.5, 0.5])
    dense = td.MultivariateNormalTorch(2, cov=[[1.0, 0.5], [0.5, 1.0]], device=CPU)        # dense covariance: its own functor
    assert dense.family_id == _lib.T_MVN_DENSE and dense.pack().numel() == _lib.PARAM_HEADER + 2 + 4
    assert td.MultivariateNormalTorch(2, device=CPU).family_id == _lib.T_MVN_DIAG
    with pytest.raises(NotImplementedError):
        td.MultivariateNormalTorch(200, cov=(np.eye(200) + 0.01).tolist(), device=CPU)      # gather limit: dim <= 128
    with pytest.raises(ValueError):
        td.SuperFunnelTorch(2, 2, [], [])                                                   # the reference's validation (:131-134)
    sf = td.SuperFunnelTorch(2, 3, [torch.zeros(4, 3), torch.ones(5, 3)], [torch.zeros(4), torch.ones(5)], device=CPU)
    assert sf.dim == 2 + 6 + 1 + 3 + 2 and sf.get_name() == "SuperFunnelTorch_J2_K3"
    assert sf.pack().numel() == _lib.PARAM_HEADER + 9 * 5 and sf.pack()[9] == 9


def test_proposal_plugins_match_reference_scales_and_errors():
    for tag in ("b1p0", "b0p25"):
        _, g = load_golden(f"prop_normal_{tag}")
        p = NormalProposal(9, float(g["var"]), float(g["beta"]), CPU, torch.float32)
        assert p.std_dev.item() == g["std"] and np.float32(p.chain_scale(float(g["beta"]))) == g["std"]
        _, g = load_golden(f"prop_laplace_{tag}")
        p = LaplaceProposal(9, torch.tensor(g["var_vec"]), float(g["beta"]), CPU, torch.float32)
        np.testing.assert_array_equal(p.scale_vector.numpy(), g["scale"])
        assert p.chain_scale(float(g["beta"])) == 1.0
        _, g = load_golden(f"prop_uniform_{tag}")
        p = UniformRadiusProposal(9, float(g["radius"]), float(g["beta"]), CPU, torch.float32)
        assert p.effective_radius.item() == g["eff_radius"] and np.float32(p.chain_scale(float(g["beta"]))) == g["eff_radius"]
    assert [NormalProposal(2, 1.0, 1.0, CPU, torch.float32).get_name(), LaplaceProposal(2, torch.ones(2), 1.0, CPU, torch.float32).get_name(),
            UniformRadiusProposal(2, 1.0, 1.0, CPU, torch.float32).get_name()] == ["Normal", "Laplace", "UniformRadius"]
    with pytest.raises(ValueError):
        NormalProposal(2, -1.0, 1.0, CPU, torch.float32)
    with pytest.raises(ValueError):
        LaplaceProposal(2, torch.ones(3), 1.0, CPU, torch.float32)
    with pytest.raises(ValueError):
        LaplaceProposal(2, torch.tensor([1.0, -1.0]), 1.0, CPU, torch.float32)
    with pytest.raises(ValueError):
        UniformRadiusProposal(2, 0.0, 1.0, CPU, torch.float32)


def test_sampler_facades_construct_on_cpu_and_refuse_to_fall_back():
    t = td.RoughCarpetDistributionTorch(20, device=CPU)
    assert RandomWalkMetropolis is RandomWalkMH_GPU_Optimized
    a = RandomWalkMH_GPU_Optimized(20, 0.5, t, device="cpu", burn_in=10)
    assert a.get_name() == "RWM_GPU_FUSED_Normal" and a.current_state is None and a.acceptance_rate == 0.0
    assert np.all(a.chain[0] == 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        a.generate_samples(5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        t.log_density(torch.zeros(20))
    with pytest.raises(ValueError):
        RandomWalkMH_GPU_Optimized(20, None, t, device="cpu")
    b = RandomWalkMH_GPU_Optimized(20, target_dist=t, device="cpu",
                                   proposal_distribution=LaplaceProposal(20, torch.ones(20), 1.0, CPU, torch.float32))
    assert b.get_name() == "RWM_GPU_FUSED_Laplace"
    pt = ParallelTemperingRWM_GPU_Optimized(20, 0.9, t, geom_temp_spacing=True, swap_every=10, device="cpu")
    assert pt.beta_ladder == [1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125, 0.015625, 0.01] and pt.num_chains == 8
    assert pt.get_name() == "PT_RWM_GPU_ULTRA_FUSED"
    np.testing.assert_array_equal(pt._scales, np.sqrt(np.asarray([np.float32(0.9 / b) for b in pt.beta_ladder], np.float32)))
    assert tuple(pt.proposal_covs_chol.shape) == (8, 20, 20)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pt.generate_samples(5)
    with pytest.raises(TypeError):
        ParallelTemperingRWM_GPU_Optimized(20, 0.9, object(), device="cpu")


def test_initial_state_rule_batched():
    np.random.seed(0)
    g = RandomWalkMH_GPU_Optimized(6, 0.5, td.IIDGammaTorch(6, device=CPU), device="cpu", num_chains=64)
    assert g._x0.shape == (64, 6) and abs(g._x0.mean() - 5) < 0.01
    b = RandomWalkMH_GPU_Optimized(6, 0.5, td.IIDBetaTorch(6, device=CPU), device="cpu", num_chains=64)
    assert b._x0.dtype == np.float32 and b._x0.min() > 0.2 and b._x0.max() < 0.8
    e = RandomWalkMH_GPU_Optimized(6, 0.5, td.EvenRosenbrockTorch(6, device=CPU), device="cpu", num_chains=4)
    assert np.abs(e._x0).max() < 1e-7 and np.abs(e._x0).max() > 0
    v = RandomWalkMH_GPU_Optimized(6, np.linspace(0.1, 1.0, 4), td.EvenRosenbrockTorch(6, device=CPU), device="cpu", num_chains=4)
    np.testing.assert_array_equal(v._scales, np.sqrt(np.linspace(0.1, 1.0, 4).astype(np.float32)))


def test_simulation_harness_dispatch_and_proposal_factory():
    t = td.EvenRosenbrockTorch(4, device=CPU)
    sim = MCMCSimulation_GPU(4, sigma=0.3, num_iterations=10, algorithm=RandomWalkMH_GPU_Optimized, target_dist=t,
                             device="cpu", burn_in=2, seed=3)
    assert isinstance(sim.algorithm, RandomWalkMH_GPU_Optimized) and sim.algorithm.pre_allocate_steps == 10
    assert not sim.has_run()
    with pytest.raises(ValueError, match="not been run"):
        sim.acceptance_rate()
    sim = MCMCSimulation_GPU(4, proposal_config={'name': 'Laplace', 'params': {'base_variance_vector': 0.2}},
                             algorithm=RandomWalkMH_GPU_Optimized, target_dist=t, device="cpu")
    assert sim.algorithm.get_name() == "RWM_GPU_FUSED_Laplace"
    sim = MCMCSimulation_GPU(4, proposal_config={'name': 'UniformRadius', 'params': {'base_radius': 1.0}},
                             algorithm=RandomWalkMH_GPU_Optimized, target_dist=t, device="cpu", beta_ladder=[0.5])
    assert abs(sim.algorithm.proposal_dist.effective_radius.item() - 1.0 / np.sqrt(np.float32(0.5))) < 1e-6
    with pytest.raises(ValueError, match="Unknown proposal"):
        MCMCSimulation_GPU(4, proposal_config={'name': 'Cauchy'}, algorithm=RandomWalkMH_GPU_Optimized, target_dist=t, device="cpu")
    with pytest.raises(ValueError):
        MCMCSimulation_GPU(4, algorithm=RandomWalkMH_GPU_Optimized, target_dist=t, device="cpu")
    sim = MCMCSimulation_GPU(4, sigma=0.3, algorithm=ParallelTemperingRWM_GPU_Optimized, target_dist=t, device="cpu",
                             beta_ladder=[1.0, 0.5], swap_every=5)
    assert sim.algorithm.num_chains == 2 and sim.algorithm.swap_every == 5


def test_reference_aliases():
    import sys
    P.install_reference_aliases()
    try:
        import algorithms, interfaces, proposal_distributions, target_distributions  # noqa: F401,E401
        assert algorithms.RandomWalkMH_GPU_Optimized is RandomWalkMH_GPU_Optimized
        assert interfaces.MCMCSimulation_GPU is MCMCSimulation_GPU
    finally:
        for n in ("algorithms", "interfaces", "proposal_distributions", "target_distributions"):
            sys.modules.pop(n, None)


def test_experiment_target_factory_matches_reference_defaults():
    from rwm_pt_pytorch_b200.experiments import get_target_distribution
    t = get_target_distribution("RoughCarpet", 20, device="cpu")
    assert t.get_name() == "RoughCarpetTorchCustom" and t.modes.tolist() == [-4.0, 0.0, 4.0]     # experiment_RWM_GPU.py:36
    assert get_target_distribution("HybridRosenbrock", 0, device="cpu").dim == 11
    assert get_target_distribution("Hypercube", 3, device="cpu").left_boundary.item() == -1.0
    assert get_target_distribution("ThreeMixtureScaled", 4, device="cpu").get_name() == "ThreeMixtureTorchScaled"
    with pytest.raises(ValueError):
        get_target_distribution("Nope", 3, device="cpu")
    sf = get_target_distribution("SuperFunnel", 3, device="cpu")       # synthetic data as experiment_RWM_GPU.py:95-120
    assert sf.dim == 5 + 15 + 1 + 3 + 2 and sf.family_id == 12
