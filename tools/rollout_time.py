"""Time MultiEnvAgent.interact (T x (K-PRE + acting forward + sample), one CUDA graph) on the device-resident synthetic
environment, as bench.py's `rollout` key does.  usage: rollout_time.py [envs] [steps] [--profile]
--profile: one more rollout between cudaProfilerStart / Stop (for `ncu --profile-from-start off`: the launch list of a rollout)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from actorcritic_b200 import engine as eng  # noqa: E402
from actorcritic_b200.agents import MultiEnvAgent  # noqa: E402
from actorcritic_b200.envs.atari.device_env import DeviceAtariMultiEnv  # noqa: E402

profile = "--profile" in sys.argv
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
envs = int(argv[0]) if len(argv) > 0 else 32
steps = int(argv[1]) if len(argv) > 1 else 20
e = eng.Engine(eng.EngineConfig(num_envs=envs, num_steps=steps, conv3_filters=32, seed=1))
e.set_params(eng.orthogonal_init(4, 32, seed=0))


class _EngineModel:
    engine = e


env = DeviceAtariMultiEnv(envs, pool_frames=steps, seed=0, device=e.device)
agent = MultiEnvAgent(env, _EngineModel(), steps)
with torch.cuda.stream(e.stream):
    for _ in range(4):
        agent.interact(None)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        agent.interact(None)
    b.record()
    torch.cuda.synchronize()
knobs = {k: v for k, v in os.environ.items() if k.startswith("ACX_")}
print("ms_per_rollout %.4f  (%.1f us per step)" % (a.elapsed_time(b) / 20, 1e3 * a.elapsed_time(b) / 20 / steps), knobs)
if profile:
    with torch.cuda.stream(e.stream):
        torch.cuda.profiler.start()
        agent.interact(None)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
