"""Oracle for one whole learner update (the reference's `session.run(optimize_op, feed_dict)`,
a2c_acktr.py:117-126) in both optimiser configurations:

  * ACKTR: ColdStartPeriodicInvUpdateKfacOpt schedule AS CODED (kfac_utils.py:38-53, SURVEY 3.4 /
    D.2: the cold optimizer and the always-run K-FAC apply both increment global_step).
  * A2C  : RMSProp (TF-1 defaults: decay 0.9, momentum 0, eps 1e-10, ms initialised to ones) behind
    ClipGlobalNormOptimizer(0.5) (a2c_acktr.py:250-251, nn.py:185-189).

dtype float64 is the parity oracle; float32 with all host threads is the `cpu_baseline` /
`--impl reference` leg of bench.py (it deliberately keeps the reference's costs: materialised
im2col patch matrices, the [E,T,T] discount matrices, two forward towers, a separate Fisher
backward, dense inverses).
"""
import dataclasses

import numpy as np
import torch

from . import kfac as K
from . import network as net
from . import returns as R


@dataclasses.dataclass
class A2CConfig:
    learning_rate_start: float = 7e-4          # a2c_acktr.py:71
    learning_rate_end: float = 7e-5
    decay_steps: float = 1e7 / (16 * 5)
    rms_decay: float = 0.9
    rms_epsilon: float = 1e-10
    clip_norm: float = 0.5


class OracleLearner:
    def __init__(self, params, num_actions=4, c3=32, acktr=True, cfg=None, gamma=0.99, beta=0.01,
                 value_weight=0.5, dtype=torch.float64, reference_cost=False):
        self.dtype = dtype
        self.params = net.to_torch(params, dtype)
        self.acktr = acktr
        self.cfg = cfg if cfg is not None else (K.KfacConfig() if acktr else A2CConfig())
        # the discount factor reaches the reference's py_func as a float32 tensor (objectives.py:198)
        self.gamma, self.beta, self.value_weight = float(np.float32(gamma)), beta, value_weight
        self.global_step = 0
        self.reference_cost = reference_cost
        if acktr:
            self.kfac = K.KfacState(self.params, self.cfg, dtype)
            self.cold_accum = {name: torch.zeros_like(net.join_vmat(name, self.params)) for name in net.LAYERS}
        else:
            self.rms = {name: torch.ones_like(net.join_vmat(name, self.params)) for name in net.LAYERS}

    # ------------------------------------------------------------------ gradient side
    def compute(self, batch, y_hat=None, eps=None, need_fisher=True, masks=None):
        """Forward both towers, targets, losses, loss gradients, and (optionally) the Fisher-sample
        backward + the 11 batch factors.  Nothing is mutated."""
        obs = np.asarray(batch["observations"])
        e_count, t_count = obs.shape[:2]
        n = e_count * t_count
        fwd = net.forward(self.params, obs.reshape((n,) + obs.shape[2:]))                 # model.py:113
        boot = net.forward(self.params, np.asarray(batch["bootstrap_observations"]), build_policy=False)  # :116
        if self.reference_cost:   # objectives.py:178-214 literally (discount matrices)
            np_dtype = np.float64 if self.dtype == torch.float64 else np.float32
            targets = torch.as_tensor(R.targets_matrix_form(
                batch["rewards"], batch["terminals"], boot["value"].detach().numpy(), self.gamma, np_dtype)).to(self.dtype)
        else:
            targets = net.targets_torch(batch["rewards"], batch["terminals"], boot["value"], self.gamma)
        targets = targets.reshape(n)
        actions = np.asarray(batch["actions"]).reshape(n)
        losses = net.a2c_loss(fwd["logits"], fwd["value"], actions, targets, self.beta, self.value_weight)
        dz, dv = net.output_grads(fwd["logits"], fwd["value"], actions, targets, self.beta, self.value_weight)
        grads, pre_grads = net.backward(self.params, fwd, dz, dv, masks)
        out = dict(fwd=fwd, bootstrap_values=boot["value"], targets=targets, losses=losses, grads=grads,
                   pre_grads=pre_grads, dlogits=dz, dvalue=dv)
        if need_fisher:
            fz, fv = net.fisher_output_grads(fwd["logits"], fwd["value"], y_hat, eps)
            _, fisher_pre = net.backward(self.params, fwd, fz, fv, masks)
            out["fisher_pre_grads"] = fisher_pre
            out["new_a"], out["new_g"] = K.batch_factors(fwd, fisher_pre)
        return out

    # ------------------------------------------------------------------ optimiser side
    def learning_rate(self):
        c = self.cfg
        return K.linear_decay(c.learning_rate_start, c.learning_rate_end, self.global_step, c.decay_steps)

    def update(self, batch, y_hat=None, eps=None, masks=None):
        """One learner update.  masks: see network.backward (ReLU derivative taken at the given masks)."""
        if not self.acktr:
            return self._update_a2c(batch, masks)
        cfg = self.cfg
        plan = K.schedule_events(self.global_step, cfg.num_cold_updates, cfg.invert_every)
        cold = plan["cold"]
        info = self.compute(batch, y_hat, eps, need_fisher=not cold, masks=masks)
        grads = info["grads"]
        lr = self.learning_rate()
        if cold:                                                     # kfac_utils.py:42-43
            clipped, norm = K.clip_by_global_norm(grads, cfg.clip_norm)
            for name in net.LAYERS:
                self.cold_accum[name] = cfg.cold_momentum * self.cold_accum[name] + clipped[name]
                new = net.join_vmat(name, self.params) - cfg.cold_learning_rate * self.cold_accum[name]
                w, b = net.split_vmat(name, new, self.params)
                self.params[name + "/weights"], self.params[name + "/bias"] = w, b
            self.global_step += 1
            info["grad_norm"] = norm
        else:                                                        # :44
            self.kfac.update_covs(info["new_a"], info["new_g"])
        if plan["inv"]:                                              # :47-50
            self.kfac.update_inverses()
            info["inverted"] = True
        # :52-53 - always.  Gradients were computed before the cold step (same session.run).
        coeff, s, precon = self.kfac.step(self.params, grads, lr)
        self.global_step += 1
        assert self.global_step == plan["gs_after"]
        info.update(clip_coeff=coeff, fisher_norm=s, precon=precon, lr=lr)
        return info

    def _update_a2c(self, batch, masks=None):
        cfg = self.cfg
        info = self.compute(batch, need_fisher=False, masks=masks)
        lr = self.learning_rate()
        clipped, norm = K.clip_by_global_norm(info["grads"], cfg.clip_norm)
        for name in net.LAYERS:
            g = clipped[name]
            self.rms[name] = cfg.rms_decay * self.rms[name] + (1 - cfg.rms_decay) * g * g
            new = net.join_vmat(name, self.params) - lr * g / torch.sqrt(self.rms[name] + cfg.rms_epsilon)
            w, b = net.split_vmat(name, new, self.params)
            self.params[name + "/weights"], self.params[name + "/bias"] = w, b
        self.global_step += 1
        info.update(grad_norm=norm, lr=lr)
        return info

    def params_numpy(self):
        return {k: v.detach().numpy().copy() for k, v in self.params.items()}
