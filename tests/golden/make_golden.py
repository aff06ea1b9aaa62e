"""Generate the golden vectors under tests/golden/ by running the REFERENCE'S OWN Python code
(/root/reference, unmodified) under the import shims of ref_shims.py, plus the real cv2.

Run in the build container only:   python tests/golden/make_golden.py
(/root/reference does not exist on the GPU box; the tests read only the committed .npz files.)

What each file pins:
  preprocess.npz  AtariPreprocessFrameWrapper.observation (wrappers.py:30-33) and
                  AtariFrameskipWrapper.step's 2-frame max (:52-67) on synth.raw_frames(...)
  framestack.npz  FrameStackWrapper (wrappers.py:201-235) inside MultiEnv's _AutoResetWrapper
                  (multi_env.py:121-137) driven through a scripted terminal sequence
  returns.npz     objectives._discount / _discount_bootstrap (objectives.py:178-214)
  schedule.npz    ColdStartPeriodicInvUpdateKfacOpt.apply_gradients (kfac_utils.py:38-53), the reference's own class
                  on a recording kfac.KfacOptimizer skeleton: which of {cold step, covariance update, inverse
                  update, K-FAC apply} each update runs and how global_step moves, for three (num_cold, every) pairs
  network.npz     AtariModel (envs/atari/model.py) + A2CObjective (objectives.py:100-154) +
                  the shared loss of optimize_shared (:78): logits, values, bootstrap values, the three
                  loss scalars and d(shared loss)/d(all 12 variables) by autograd through the
                  reference's own graph-building code; and what register_layers /
                  register_predictive_distributions register.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shims  # noqa: E402
import synth  # noqa: E402

gym, tf, kfac = ref_shims.install()

import actorcritic.envs.atari.wrappers as ref_wrappers  # noqa: E402
import actorcritic.multi_env as ref_multi_env  # noqa: E402
import actorcritic.objectives as ref_objectives  # noqa: E402
from actorcritic.agents import transpose_list  # noqa: E402
from actorcritic.envs.atari.model import AtariModel  # noqa: E402

from oracle import network as onet  # noqa: E402


class ScriptedRawEnv(gym.Env):
    """Stands in for the emulator: yields pre-generated raw frames and scripted terminals."""

    def __init__(self, frames, terminal_at=()):
        self.frames = frames
        self.i = 0
        self.terminal_at = set(terminal_at)
        self.observation_space = gym.spaces.Box(low=0, high=255, shape=(210, 160, 3), dtype=np.uint8)
        self.action_space = gym.spaces.Discrete(4)
        self.log = []

    def _next(self):
        f = self.frames[self.i % len(self.frames)]
        self.i += 1
        return f

    def step(self, action):
        idx = self.i
        f = self._next()
        self.log.append(("step", idx))
        return f, 1.0, idx in self.terminal_at, {}

    def reset(self, **kwargs):
        idx = self.i
        f = self._next()
        self.log.append(("reset", idx))
        return f


def gen_preprocess():
    frames = synth.raw_frames(11, 8, "mixed")
    env = ScriptedRawEnv(frames)
    pre = ref_wrappers.AtariPreprocessFrameWrapper(env)
    out = np.stack([pre.observation(f) for f in frames])                      # [8,84,84,1]
    # frameskip=4 over the 8 frames: two agent steps; a terminal on the first sub-step of a third
    env2 = ScriptedRawEnv(frames, terminal_at={5})
    skip = ref_wrappers.AtariFrameskipWrapper(env2, frameskip=4)
    m1, r1, t1, _ = skip.step(0)       # frames 0..3 -> max(f2, f3)
    m2, r2, t2, _ = skip.step(0)       # frames 4,5 (terminal at 5) -> max(f4, f5)
    env3 = ScriptedRawEnv(frames, terminal_at={0})
    m3, r3, t3, _ = ref_wrappers.AtariFrameskipWrapper(env3, frameskip=4).step(0)   # single frame -> f0
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), seed=11, count=8, observation=out,
                        skip_max_23=m1, skip_max_45=m2, skip_single_0=m3,
                        skip_rewards=np.array([r1, r2, r3]), skip_terminals=np.array([t1, t2, t3]))


def gen_framestack():
    """Each env: raw frames -> (max over the last two of a 2-frame skip window) -> preprocess ->
    FrameStackWrapper in the main process -> _AutoResetWrapper, as a2c_acktr.py:167-171 /
    multi_env.py:25 stack them.  Terminals are scripted on agent steps."""
    num_envs, num_steps = 2, 8
    term_steps = [{2, 3}, {6}]           # agent-step indices at which env e terminates
    outs, terms = [], []
    envs = []
    for e in range(num_envs):
        frames = synth.raw_frames(100 + e, 24, "mixed")

        class AgentStepEnv(gym.Env):
            """2 raw frames per agent step (so the 2-frame max is exercised); reset consumes 1."""

            def __init__(self, frames, term):
                self.frames, self.term, self.i, self.t = frames, term, 0, 0
                self.observation_space = gym.spaces.Box(low=0, high=255, shape=(210, 160, 3), dtype=np.uint8)
                self.action_space = gym.spaces.Discrete(4)
                self.trace = []

            def step(self, action):
                a, b = self.frames[self.i], self.frames[self.i + 1]
                self.trace.append(("step", self.i, self.i + 1))
                self.i += 2
                done = self.t in self.term
                self.t += 1
                return np.amax((a, b), axis=0), 0.0, done, {}

            def reset(self, **kwargs):
                f = self.frames[self.i]
                self.trace.append(("reset", self.i, -1))
                self.i += 1
                return f
        raw = AgentStepEnv(frames, term_steps[e])
        env = ref_wrappers.AtariPreprocessFrameWrapper(raw)
        env = ref_wrappers.FrameStackWrapper(env, 4)
        env = ref_multi_env._AutoResetWrapper(env)
        envs.append((raw, env))
    first = np.stack([env.reset().copy() for _, env in envs])
    for t in range(num_steps):
        step_out, step_term = [], []
        for raw, env in envs:
            o, r, d, _ = env.step(0)
            step_out.append(o.copy())
            step_term.append(d)
        outs.append(np.stack(step_out))
        terms.append(step_term)
    traces = np.array([[list(x[1:]) + [0 if x[0] == "step" else 1] for x in raw.trace] for raw, _ in envs], dtype=object)
    np.savez_compressed(os.path.join(HERE, "framestack.npz"), seeds=np.array([100, 101]), frames_per_env=24,
                        reset_observation=first, observations=np.stack(outs), terminals=np.array(terms),
                        trace_env0=np.array([[a, b, k] for a, b, k in [(x[1], x[2], int(x[0] == "reset")) for x in envs[0][0].trace]]),
                        trace_env1=np.array([[a, b, k] for a, b, k in [(x[1], x[2], int(x[0] == "reset")) for x in envs[1][0].trace]]))
    del traces


def gen_returns():
    ref_shims.COMPUTE_DTYPE = torch.float32
    cases = {}
    rng = np.random.default_rng(5)
    specs = [(32, 20, 0.05), (16, 5, 0.2), (4, 7, 0.0), (4, 7, 1.0), (3, 1, 0.5), (8, 20, 0.3)]
    for ci, (e, t, p) in enumerate(specs):
        rewards = rng.choice(np.array([-1, 0, 0, 0, 1], np.float32), (e, t)).astype(np.float32)
        if ci == 5:
            rewards = rng.standard_normal((e, t)).astype(np.float32)
        terminals = rng.random((e, t)) < p
        if ci == 0:
            terminals[0, :] = False
            terminals[1, 0] = True
            terminals[2, t - 1] = True
        boot = rng.standard_normal(e).astype(np.float32)
        disc = ref_objectives._discount(torch.as_tensor(rewards), torch.as_tensor(terminals), 0.99)
        bootd = ref_objectives._discount_bootstrap(torch.as_tensor(boot), torch.as_tensor(terminals), 0.99)
        cases["rewards_%d" % ci] = rewards
        cases["terminals_%d" % ci] = terminals
        cases["bootstrap_%d" % ci] = boot
        cases["discounted_rewards_%d" % ci] = disc.numpy().astype(np.float32)
        cases["discounted_bootstrap_%d" % ci] = bootd.numpy().astype(np.float32)
    cases["num_cases"] = len(specs)
    cases["gamma"] = 0.99
    np.savez_compressed(os.path.join(HERE, "returns.npz"), **cases)
    ref_shims.COMPUTE_DTYPE = torch.float64


def gen_network():
    ref_shims.COMPUTE_DTYPE = torch.float64
    e, t, a, c3 = 2, 3, 4, 32
    batch = synth.rollout(21, e, t, a, terminal_prob=0.3)
    params = onet.perturbed_params(a, c3, seed=3)
    ref_shims.VARIABLES.clear()
    for k, v in params.items():
        ref_shims.VARIABLES[k] = torch.tensor(v, dtype=torch.float64, requires_grad=True)
    ref_shims.FEEDS.clear()
    ref_shims.FEEDS.update(observations=batch["observations"], bootstrap_observations=batch["bootstrap_observations"],
                           actions=batch["actions"], rewards=batch["rewards"], terminals=batch["terminals"])
    obs_space = gym.spaces.Box(low=0, high=255, shape=(84, 84, 4), dtype=np.uint8)
    act_space = gym.spaces.Discrete(a)
    model = AtariModel(obs_space, act_space, conv3_num_filters=c3, name="AtariModel")
    objective = ref_objectives.A2CObjective(model, discount_factor=0.99, entropy_regularization_strength=0.01)
    shared = objective.policy_loss + 0.5 * objective.baseline_loss          # objectives.py:78
    shared.backward()
    out = dict(seed=21, num_envs=e, num_steps=t, num_actions=a, c3=c3, param_seed=3,
               logits=model.policy._distribution.logits.detach().numpy(),
               values=model.baseline.value.detach().numpy(),
               bootstrap_values=model.bootstrap_values.detach().numpy(),
               log_prob=model.policy.log_prob.detach().numpy(),
               entropy=model.policy.entropy.detach().numpy(),
               mode=model.policy.mode.numpy() if model.policy.mode.dim() else np.array(model.policy.mode),
               policy_loss=float(objective.policy_loss), baseline_loss=float(objective.baseline_loss),
               mean_entropy=float(objective.mean_entropy), shared_loss=float(shared))
    for k, v in ref_shims.VARIABLES.items():
        g = v.grad.numpy()
        key = "grad_" + k.replace("/", "_")
        if g.size > 100000:                      # fc4 weights: keep a strided sample + the Frobenius norm
            out[key + "_sample"] = g.reshape(-1)[::37].copy()
            out[key + "_norm"] = float(np.sqrt((g * g).sum()))
        else:
            out[key] = g
    lc = kfac.LayerCollection()
    model.register_layers(lc)
    model.register_predictive_distributions(lc)
    reg = []
    for kind, kw in lc.calls:
        if kind in ("conv2d", "fully_connected"):
            reg.append("%s inputs=%s outputs=%s strides=%s padding=%s w=%s" % (
                kind, tuple(kw["inputs"].shape), tuple(kw["outputs"].shape), kw.get("strides"), kw.get("padding"),
                tuple(kw["params"][0].shape)))
        elif kind == "categorical":
            reg.append("categorical logits=%s" % (tuple(kw["logits"].shape),))
        else:
            reg.append("normal mean=%s var=%s" % (tuple(kw["mean"].shape), kw["var"]))
    shared_inputs = lc.calls[4][1]["inputs"] is lc.calls[5][1]["inputs"]
    out["registrations"] = np.array(reg)
    out["heads_share_inputs"] = bool(shared_inputs)
    np.savez_compressed(os.path.join(HERE, "network.npz"), **out)


def gen_schedule():
    """Runs the UNMODIFIED kfac_utils.ColdStartPeriodicInvUpdateKfacOpt.apply_gradients once per update, eagerly (the
    shim's tf.cond / control_dependencies execute in program order; predicates read the current global_step), and records
    per update: global_step before / after and whether the cold optimizer, the covariance thunks, the inverse thunks and
    the base KfacOptimizer.apply_gradients ran.  Rows: [gs_before, cold, cov, inv, kfac_apply, gs_after]."""
    import actorcritic.kfac_utils as ref_kfac_utils
    out = {}
    for tag, (num_cold, every, updates) in {"reference": (30, 10, 60), "short": (4, 2, 16), "odd": (5, 3, 24)}.items():
        gs = ref_shims.StepVariable(0)
        opt = ref_kfac_utils.ColdStartPeriodicInvUpdateKfacOpt(
            num_cold_updates=num_cold, cold_optimizer=ref_shims.RecordingOptimizer("cold"), invert_every=every,
            learning_rate=0.25, cov_ema_decay=0.99, damping=0.01, layer_collection=None, momentum=0.9, norm_constraint=1e-4)
        rows = []
        for _ in range(updates):
            del ref_shims.EVENTS[:]
            before = gs.value
            opt.apply_gradients([], global_step=gs)
            names = [e[0] for e in ref_shims.EVENTS]
            # order as coded: (cold | cov) -> [inv] -> kfac_apply
            assert names[-1] == "kfac_apply" and names[0] in ("cold", "cov"), names
            assert names == [n for n in ("cold", "cov", "inv", "kfac_apply") if n in names], names
            rows.append([before, int("cold" in names), int("cov" in names), int("inv" in names),
                         int("kfac_apply" in names), gs.value])
        out[tag] = np.array(rows, np.int64)
        out[tag + "_config"] = np.array([num_cold, every], np.int64)
    np.savez_compressed(os.path.join(HERE, "schedule.npz"), **out)


def gen_transpose():
    v = [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]
    np.savez_compressed(os.path.join(HERE, "transpose.npz"), inp=np.array(v), out=np.array(transpose_list(v)))


if __name__ == "__main__":
    gen_preprocess()
    gen_framestack()
    gen_returns()
    gen_network()
    gen_schedule()
    gen_transpose()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
