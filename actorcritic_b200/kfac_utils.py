"""actorcritic/kfac_utils.py: ColdStartPeriodicInvUpdateKfacOpt - the schedule object.  `minimize` returns the
optimize-op token; the schedule itself (cold steps that also count global_step, covariances every update after the
cold phase, inverses every `invert_every`, K-FAC apply always; kfac_utils.py:38-53) runs in learner.cu phase2."""
from . import kfac


class ColdStartPeriodicInvUpdateKfacOpt(kfac.KfacOptimizer):
    def __init__(self, num_cold_updates, cold_optimizer, invert_every, **kwargs):
        self._num_cold_updates = num_cold_updates
        self._cold_optimizer = cold_optimizer
        self._invert_every = invert_every
        super().__init__(**kwargs)

    def engine_overrides(self):
        """Hyper-parameters this optimizer contributes to the engine configuration."""
        from . import nn
        cold = self._cold_optimizer
        clip = None
        if isinstance(cold, nn.ClipGlobalNormOptimizer):
            clip, cold = cold.clip_norm, cold.optimizer
        if not isinstance(cold, nn.MomentumOptimizer):
            raise NotImplementedError("the cold optimizer on the hot path is ClipGlobalNorm(Momentum) (a2c_acktr.py:240-241)")
        lr = self.learning_rate
        if not isinstance(lr, nn.LinearDecay):
            lr = nn.LinearDecay(float(lr), float(lr), None, 1.0)
        return dict(acktr=True, num_cold_updates=int(self._num_cold_updates), invert_every=int(self._invert_every),
                    cold_lr=float(cold.learning_rate), cold_momentum=float(cold.momentum),
                    clip_norm=float(clip) if clip is not None else 3.4e38,
                    lr_start=lr.start_value, lr_end=lr.end_value, lr_decay_steps=float(lr.total_steps),
                    cov_ema_decay=float(self.cov_ema_decay), damping=float(self.damping), momentum=float(self.momentum),
                    norm_constraint=float(self.norm_constraint) if self.norm_constraint is not None else 3.4e38)
