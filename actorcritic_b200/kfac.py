"""The part of the `tensorflow/kfac` 0.1.x surface the reference calls (SURVEY 8(b) last-but-two row), as plain
metadata records: the arithmetic those registrations imply (Kronecker factors, EMA, damping, inverses,
preconditioning, KL clip - SURVEY A.5) is implemented by libacx's learner engine for exactly the six blocks
AtariModel registers (envs/atari/model.py:219-246)."""


class LayerCollection:
    def __init__(self):
        self.conv2d = []
        self.fully_connected = []
        self.categorical = []
        self.normal = []

    def register_conv2d(self, params, strides, padding, inputs, outputs, **kw):
        if padding != "VALID":
            raise NotImplementedError("only VALID padding is on the hot path (envs/atari/model.py:227-237)")
        self.conv2d.append(dict(params=params, strides=list(strides), padding=padding, inputs=inputs, outputs=outputs))

    def register_fully_connected(self, params, inputs, outputs, **kw):
        self.fully_connected.append(dict(params=params, inputs=inputs, outputs=outputs))

    def register_categorical_predictive_distribution(self, logits, seed=None, **kw):
        self.categorical.append(dict(logits=logits, seed=seed))

    def register_normal_predictive_distribution(self, mean, var=0.5, seed=None, **kw):
        if float(var) != 1.0:
            raise NotImplementedError("the value head's predictive distribution has var=1.0 (baselines.py:66)")
        self.normal.append(dict(mean=mean, var=float(var), seed=seed))

    @property
    def num_blocks(self):
        return len(self.conv2d) + len(self.fully_connected)

    def input_factor_groups(self):
        """Blocks registered with the same `inputs` share one input factor (the two heads: model.py:243,246)."""
        groups = {}
        for rec in self.conv2d + self.fully_connected:
            groups.setdefault(id(rec["inputs"]), []).append(rec)
        return list(groups.values())


class KfacOptimizer:
    """Hyper-parameter record of kfac.KfacOptimizer as constructed at a2c_acktr.py:243-247."""

    def __init__(self, learning_rate, cov_ema_decay=0.95, damping=0.001, layer_collection=None, momentum=0.9,
                 norm_constraint=None, cov_devices=None, inv_devices=None, name="KFAC", **unused):
        if layer_collection is None:
            raise ValueError("layer_collection is required")
        self.learning_rate = learning_rate
        self.cov_ema_decay = cov_ema_decay
        self.damping = damping
        self.layer_collection = layer_collection
        self.momentum = momentum
        self.norm_constraint = norm_constraint
        self.cov_devices, self.inv_devices = cov_devices, inv_devices   # placement only, no numerical effect

    def make_vars_and_create_op_thunks(self):
        """kfac_utils.py:39: (covariance update thunks, inverse update thunks) - one per factor here, metadata only."""
        groups = self.layer_collection.input_factor_groups()
        n_factors = len(groups) + self.layer_collection.num_blocks
        return [("cov", i) for i in range(n_factors)], [("inv", i) for i in range(n_factors)]
