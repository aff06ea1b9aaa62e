"""actorcritic_b200 - a B200-native (sm_100a) implementation of the ACKTR / A2C learner hot path of
jrobine/actor-critic behind the reference's public Python surface: `agents`, `multi_env`, `model`, `objectives`,
`kfac_utils`, `policies`, `baselines`, `nn`, `envs.atari.model`, `envs.atari.wrappers`.
All arithmetic runs in libacx.so (include/acx.h); there is no CPU fallback."""
from . import agents, baselines, checkpoint, kfac, kfac_utils, model, multi_env, nn, objectives, policies, spaces, summary  # noqa: F401
from .session import GlobalStep, Session  # noqa: F401

__all__ = ["agents", "baselines", "checkpoint", "kfac", "kfac_utils", "model", "multi_env", "nn", "objectives", "policies", "spaces", "summary",
           "Session", "GlobalStep"]
