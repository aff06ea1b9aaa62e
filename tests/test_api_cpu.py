"""Host-side logic of the drop-in API (no GPU): placeholders, registration, schedule records, list contracts,
auto-reset ordering, error behaviour, the C ABI's symbol table, and the data-parallel plumbing over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import actorcritic_b200 as ac
from actorcritic_b200 import _lib, agents, kfac, kfac_utils, multi_env, nn, objectives, parallel, spaces
from actorcritic_b200 import engine as eng
from actorcritic_b200.envs.atari.model import AtariModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(c3=32, n=4):
    return AtariModel(spaces.Box(0, 255, (84, 84, 4), np.uint8), spaces.Discrete(n), c3, random_seed=0)


def test_placeholders_follow_the_reference():
    m = _model()
    # model.py:97-105,172-186: Discrete(n<=255) -> uint8, Box -> its dtype; batch dims [None, None] / [None]
    assert m.observations_placeholder.dtype == np.uint8 and m.observations_placeholder.shape == (None, None, 84, 84, 4)
    assert m.bootstrap_observations_placeholder.shape == (None, 84, 84, 4)
    assert m.actions_placeholder.dtype == np.uint8 and m.actions_placeholder.shape == (None, None)
    assert m.rewards_placeholder.dtype == np.float32 and m.terminals_placeholder.dtype == np.bool_
    assert AtariModel(spaces.Box(0, 255, (84, 84, 4), np.uint8), spaces.Discrete(300), 32).actions_placeholder.dtype == np.uint16
    with pytest.raises(TypeError):
        AtariModel(spaces.Box(0, 255, (84, 84, 4), np.uint8), spaces.Box(0, 1, (2,), np.float32))


def test_variable_names_shapes_and_init_gains():
    m = _model(c3=64, n=6)
    v = m.get_variables()
    assert {k: a.shape for k, a in v.items()} == eng.param_shapes(6, 64)
    # envs/atari/model.py:132-135: orthogonal with gains sqrt(2) / 0.01 / 1.0, zero biases
    w = v["fc4/weights"]
    assert np.allclose(w.T @ w, 2.0 * np.eye(512), atol=1e-4)
    assert np.allclose(v["fc_policy/weights"].T @ v["fc_policy/weights"], 1e-4 * np.eye(6), atol=1e-7)
    assert all(np.all(v[k] == 0) for k in v if k.endswith("bias"))
    flat = eng.flatten_params(v, 6, 64)
    assert flat.size == 1686693 + (6 - 4) * 513       # SURVEY appendix B (A=4: 1 686 693)
    back = eng.unflatten_params(flat, 6, 64)
    assert all(np.array_equal(back[k], v[k]) for k in v)


def test_kfac_registration_matches_reference():
    m = _model()
    lc = kfac.LayerCollection()
    m.register_layers(lc)
    m.register_predictive_distributions(lc, random_seed=3)
    assert [r["strides"] for r in lc.conv2d] == [[1, 4, 4, 1], [1, 2, 2, 1], [1, 1, 1, 1]]
    assert len(lc.fully_connected) == 3 and lc.num_blocks == 6
    assert len(lc.input_factor_groups()) == 5            # the heads share one input factor -> 11 factors in total
    assert lc.categorical[0]["seed"] == 3 and lc.normal[0]["var"] == 1.0
    opt = kfac.KfacOptimizer(learning_rate=0.25, layer_collection=lc)
    cov, inv = opt.make_vars_and_create_op_thunks()
    assert len(cov) == 11 and len(inv) == 11
    with pytest.raises(NotImplementedError):
        ac.model.ActorCriticModel.register_layers(m, lc)


def test_optimizer_records_and_schedule_hyperparameters():
    m = _model()
    lc = kfac.LayerCollection()
    m.register_layers(lc)
    gs = ac.GlobalStep()
    lr = nn.linear_decay(0.25, 0.025, gs, 1e7 / 640)
    cold = nn.ClipGlobalNormOptimizer(nn.MomentumOptimizer(learning_rate=0.0003, momentum=0.9), clip_norm=0.5)
    opt = kfac_utils.ColdStartPeriodicInvUpdateKfacOpt(
        num_cold_updates=30, cold_optimizer=cold, invert_every=10, learning_rate=lr, cov_ema_decay=0.99, damping=0.01,
        layer_collection=lc, momentum=0.9, norm_constraint=0.0001, cov_devices=["/gpu:0"], inv_devices=["/gpu:0"])
    o = opt.engine_overrides()
    assert o["acktr"] and o["num_cold_updates"] == 30 and o["invert_every"] == 10 and o["cold_lr"] == 0.0003
    assert o["lr_start"] == 0.25 and o["lr_end"] == 0.025 and o["lr_decay_steps"] == 15625.0
    a2c = nn.ClipGlobalNormOptimizer(nn.RMSPropOptimizer(learning_rate=nn.linear_decay(7e-4, 7e-5, gs, 1000)), clip_norm=0.5)
    o = a2c.engine_overrides()
    assert not o["acktr"] and o["rms_decay"] == 0.9 and o["rms_epsilon"] == 1e-10 and o["clip_norm"] == 0.5
    # nn.py:130-132 formula
    assert lr.value_at(0) == 0.25 and abs(lr.value_at(15625) - 0.025) < 1e-12 and abs(lr.value_at(1e9) - 0.025) < 1e-12
    assert abs(lr.value_at(7812.5) - 0.1375) < 1e-12


def test_objective_error_behaviour():
    m = _model()
    obj = objectives.A2CObjective(m, discount_factor=0.99, entropy_regularization_strength=0.01)
    with pytest.raises(TypeError):          # objectives.py:52-53 with the default None kwargs
        obj.optimize_separate(None, None)
    with pytest.raises(TypeError):
        obj.optimize_shared(object())
    op = obj.optimize_shared(nn.ClipGlobalNormOptimizer(nn.RMSPropOptimizer(7e-4), 0.5), baseline_loss_weight=0.5,
                             global_step=ac.GlobalStep())
    assert op.kind == "optimize"


def test_no_cuda_means_no_session_and_no_engine():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.AcxError):
        ac.Session()
    with pytest.raises(_lib.AcxError):
        eng.Engine(eng.EngineConfig(num_envs=2, num_steps=2))
    from actorcritic_b200.envs.atari.raw_env import RawFrameMultiEnv
    with pytest.raises(_lib.AcxError):
        RawFrameMultiEnv([object()])


def test_transpose_list_docstring_example():
    # agents.py:235-247
    assert agents.transpose_list([[1, 2], [3, 4], [5, 6]]) == [[1, 3, 5], [2, 4, 6]]


class _ScriptedEnv:
    def __init__(self, terminal_at):
        self.i, self.terminal_at, self.log = 0, set(terminal_at), []
        self.observation_space = spaces.Box(0, 255, (2,), np.uint8)
        self.action_space = spaces.Discrete(4)

    def reset(self):
        self.log.append(("reset", self.i))
        self.i += 1
        return self.i - 1

    def step(self, action):
        self.log.append(("step", self.i, action))
        self.i += 1
        return self.i - 1, 1.0, (self.i - 1) in self.terminal_at, {}


def test_multi_env_auto_reset_ordering_and_none_actions():
    envs = [_ScriptedEnv({2}), _ScriptedEnv(set())]
    me = multi_env.MultiEnv(envs)
    assert me.reset() == [0, 0]
    obs, rew, term, info = me.step([1, 2])
    assert obs == [1, 1] and term == [False, False]
    obs, rew, term, info = me.step([3, None])          # multi_env.py:74-75
    assert obs == [2, None] and term == [True, None] and rew[1] is None
    obs, _, term, _ = me.step([0, 1])
    # env 0 was terminal: reset first (observation 3 discarded), then step with the stale action (multi_env.py:127-132)
    assert envs[0].log[-2:] == [("reset", 3), ("step", 4, 0)] and obs == [4, 2]
    me.close()


class _FakeModel:
    def __init__(self):
        self.calls = []

    def sample_actions(self, observations, session):
        self.calls.append(observations)
        return [o[0] % 4 for o in observations]


def test_multi_env_agent_interact_contract():
    me = multi_env.MultiEnv([_ScriptedEnv(set()) for _ in range(3)])
    model = _FakeModel()
    agent = agents.MultiEnvAgent(me, model, num_steps=5)
    obs, act, rew, term, nxt, infos = agent.interact(session=None)
    assert np.shape(obs) == (3, 5) and np.shape(act) == (3, 5) and np.shape(rew) == (3, 5) and np.shape(term) == (3, 5)
    assert len(nxt) == 3 and np.shape(infos) == (3, 5)
    assert obs[0] == [0, 1, 2, 3, 4] and nxt == [5, 5, 5]
    assert model.calls[0] == [[0], [0], [0]]            # [env] -> [env, 1] batch (agents.py:207)
    obs2, *_ = agent.interact(session=None)
    assert obs2[0][0] == 5                               # next_observations reused between calls (agents.py:198-200,219)
    me.close()


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "acx.h")).read()
    declared = set(re.findall(r"\b(acx_[a-z0-9_]+)\s*\(", header))
    declared -= {"acx_learner_arena_bytes_"}      # (none; placeholder for macro-like false positives)
    assert "acx_learner_phase1" in declared and "acx_preprocess_stack_u8" in declared and "acx_gemm" in declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\sT\s+(acx_[a-z0-9_]+)", out))
    assert declared <= exported, sorted(declared - exported)
    assert declared <= set(_lib.SIGNATURES), sorted(declared - set(_lib.SIGNATURES))
    _lib.load()     # loads and binds every signature without a GPU; no compute call is made here


def test_public_header_is_plain_c():
    """include/acx.h is the drop-in boundary: it must compile as C99 and as C++ on its own (no torch / CUDA types)."""
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    header = os.path.join(ROOT, "include", "acx.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", header])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", header])
    includes = re.findall(r"#\s*include\s*[<\"]([^>\"]+)", open(header).read())
    assert set(includes) <= {"stddef.h", "stdint.h"}, includes


def test_library_is_blackwell_native_and_links_no_vendor_math():
    """The built library carries sm_100a code whose contraction kernels use tcgen05 (UTCHMMA) fed by TMA (UTMALDG) and read
    their accumulators from tensor memory (LDTM), the large-batch K-PRE kernel stages frames with bulk copies (UBLKCP), and
    nothing links cuBLAS / cuDNN / cuSOLVER.  Needs the CUDA binary tools only (no GPU)."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    elf = subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH], text=True)
    assert "sm_100a" in elf, elf
    sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True)
    per_kernel, name = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_kernel[name] = set()
        elif name is not None:
            for op in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP"):
                if op in line:
                    per_kernel[name].add(op)
    def ops_of(fragment):
        found = [v for k, v in per_kernel.items() if fragment in k]
        assert found, fragment
        return set().union(*found)
    assert {"UTCHMMA", "UTMALDG", "LDTM"} <= ops_of("gemm_tc_kernel")
    assert {"UTCHMMA", "UTMALDG", "LDTM"} <= ops_of("conv_tc_kernel")
    assert "UBLKCP" in ops_of("preprocess_persistent_kernel")
    needed = subprocess.check_output(["ldd", _lib.LIB_PATH], text=True)
    for vendor in ("cublas", "cudnn", "cusolver", "nccl", "torch"):
        assert vendor not in needed.lower(), needed


def test_shard_range_and_batch():
    assert [parallel.shard_range(256, r, 8) for r in (0, 7)] == [(0, 32), (224, 256)]
    with pytest.raises(ValueError):
        parallel.shard_range(30, 0, 4)
    b = dict(observations=np.arange(8)[:, None], rewards=np.arange(8.0)[:, None])
    s = parallel.shard_batch(b, 1, 2)
    assert s["observations"][:, 0].tolist() == [4, 5, 6, 7]


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from actorcritic_b200 import parallel
rank, world, _ = parallel.init_from_env("gloo")
rng = np.random.default_rng(0)
x = rng.standard_normal((8, 5, 3))                  # "rows" of 8 environments x 5 steps, 3 features
lo, hi = parallel.shard_range(8, rank, world)
mine = x[lo:hi].reshape(-1, 3)
# per-rank statistics exactly like the engine's bucket: a gradient-like mean and a factor-like second moment
bucket = torch.from_numpy(np.concatenate([mine.mean(0), (mine.T @ mine / mine.shape[0]).ravel(), [mine.sum() / mine.shape[0]]]))
parallel.allreduce_mean_(bucket)
full = x.reshape(-1, 3)
want = np.concatenate([full.mean(0), (full.T @ full / full.shape[0]).ravel(), [full.sum() / full.shape[0]]])
assert np.allclose(bucket.numpy(), want, atol=1e-12), (bucket.numpy(), want)
dist.barrier()
dist.destroy_process_group()
open(os.path.join(%(out)r, "rank%%d.ok" %% rank), "w").write("ok")
"""


def test_bucket_allreduce_equals_full_batch_statistics_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "out": str(tmp_path)})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()
