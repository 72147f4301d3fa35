"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/quadsim.h declares, the
ctypes mirrors match the header's struct sizes, and host-side config logic behaves like the reference's constructor."""
import ctypes as C
import os
import re

import pytest

from quad_swarm_rl_stable_baselines3_b200 import _capi
from quad_swarm_rl_stable_baselines3_b200.config import QsConfigC, QsStatsC, QuadSimConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _capi.build()
    return _capi.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "quadsim.h")).read()
    declared = set(re.findall(r"\b(qs_[a-z_]+)\s*\(", hdr))
    declared -= {"qs_status"}
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_policy_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "quadpolicy.h")).read()
    declared = set(re.findall(r"\b(qp_[a-z_]+)\s*\(", hdr))
    assert declared == set(_capi.POLICY_EXPORTS), declared ^ set(_capi.POLICY_EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    from quad_swarm_rl_stable_baselines3_b200.fused_policy import QpConfigC
    assert lib.qp_config_size() == C.sizeof(QpConfigC)


def test_policy_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from quad_swarm_rl_stable_baselines3_b200.fused_policy import QpConfigC
    h = C.c_void_p()
    cfg = QpConfigC(1, 18, 6, 6, 256, 4)
    assert lib.qp_create(C.byref(cfg), 0, C.byref(h)) < 0 and not h.value
    assert b"CUDA" in lib.qp_last_error(None)


def test_struct_layouts(lib):
    assert lib.qs_config_size() == C.sizeof(QsConfigC)
    assert lib.qs_stats_size() == C.sizeof(QsStatsC)
    assert lib.qs_api_version() == 2


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    cfg = QuadSimConfig(num_envs=2).to_c()
    rc = lib.qs_create(C.byref(cfg), 0, C.byref(h))
    assert rc < 0 and not h.value
    assert b"CUDA" in lib.qs_last_error(None)
    from quad_swarm_rl_stable_baselines3_b200.sim import QuadSwarmSim
    with pytest.raises(RuntimeError, match="CUDA"):
        QuadSwarmSim(QuadSimConfig(num_envs=2))


def test_bad_config_rejected(lib):
    c = QuadSimConfig(num_envs=2).to_c()
    c.num_agents = 40
    h = C.c_void_p()
    assert lib.qs_create(C.byref(c), 0, C.byref(h)) == -1
    assert b"num_agents" in lib.qs_last_error(None)
    assert lib.qs_create(None, 0, C.byref(h)) == -4


def test_config_derivations():
    c = QuadSimConfig(num_envs=3, num_agents=8)
    assert c.obs_dim == 54 and c.ep_len == 1500 and c.visible == 6
    c3 = QuadSimConfig(num_agents=8, quads_mode="mix", use_obstacles=True, obs_repr="xyz_vxyz_R_omega_floor",
                       neighbor_visible_num=2)
    assert c3.obs_dim == 40 and c3.num_obstacles == 12 and c3.to_c().spawn_box == 0.1   # cfg3
    assert QuadSimConfig(num_agents=32).obs_dim == 54                            # cfg4
    assert QuadSimConfig(num_agents=4, neighbor_visible_num=-1).visible == 3
    with pytest.raises(ValueError):
        QuadSimConfig(num_agents=4, neighbor_visible_num=6).to_c()
    with pytest.raises(ValueError):
        QuadSimConfig(quads_mode="o_swap_goals").to_c()              # not a scenario the device runs (nor the reference: DESIGN.md 9 f2)
    with pytest.raises(ValueError):
        QuadSimConfig(quads_mode="run_away", num_agents=1, neighbor_visible_num=0).to_c()     # run_away.py:20: randint(1, 1)
    assert QuadSimConfig(quads_mode="run_away").to_c().scenario == 14
    with pytest.raises(AssertionError):
        QuadSimConfig(rew_coeff=dict(typo=1.0)).to_c()
    cc = c.to_c()
    assert abs(cc.collision_hitbox_radius * cc.arm - 0.09192388155425119) < 1e-15
    assert cc.svd_period == 100


def test_fork_config_validation():
    """fork-mode configuration: the reference's QuadrotorEnvConfig defaults and what the device rejects."""
    c = QuadSimConfig.fork_default(num_envs=3)
    assert (c.num_agents, c.obs_dim, c.act_dim, c.ep_len) == (4, 6 + 2 * 3, 2, 3000)     # global_cfg.py:63-71, 30 s @ 100 Hz
    cc = c.to_c()
    assert cc.env_mode == 1 and cc.scenario == 4 and cc.fork.substeps == 8 and cc.apply_collision_force == 0
    assert abs(cc.fork.pid[2][4] - 2.0) < 1e-12 and cc.fork.pid[9][3] == -1.0          # z anti-windup 2.0; rate PIDs unsaturated
    with pytest.raises(ValueError):
        QuadSimConfig(env_mode="fork").to_c()                  # upstream obs_repr in fork mode
    with pytest.raises(ValueError):
        QuadSimConfig(quads_mode="dynamic_repulsive").to_c()   # fork scenario in upstream mode
    with pytest.raises(ValueError):
        QuadSimConfig.fork_default(use_obstacles=True).to_c()


def test_step_kernel_prologues_issue_every_load_before_the_first_consumer():
    """Static SASS check of the built objects (profiles/tools/prologue_check.py): in every shipped step-kernel variant all
    prologue loads are issued before the first instruction that consumes an in-flight load.  A consumer in between serialises two
    HBM round trips at the top of every warp (cost 3-5 % twice in round 1)."""
    import importlib.util
    import shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    objdir = os.path.join(root, "build", "obj")
    if shutil.which("cuobjdump") is None or not os.path.exists(os.path.join(objdir, "kernels_kg8.o")):
        pytest.skip("needs cuobjdump and the objects of an in-tree build")
    spec = importlib.util.spec_from_file_location("prologue_check", os.path.join(root, "profiles", "tools", "prologue_check.py"))
    pc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pc)
    for obj, sub in pc.KERNELS:
        r = pc.analyse(os.path.join(objdir, obj), sub)
        assert r is not None, sub
        assert r["loads"] >= 12, r          # (first_use may be None: no consumer at all inside the window is the best case)
        assert r["late_loads"] == [], r
