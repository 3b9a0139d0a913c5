import sys, torch
sys.path.insert(0, ".")
from egogen_b200.runtime import build_policy
from egogen_b200.crowd_env import default_cfg
dev = torch.device("cuda:0")
pol, _ = build_policy(default_cfg(), dev)
B = 256
obs = {"state": torch.randn(B, 2, 402, device=dev), "egosensing": torch.rand(B, 2, 32, device=dev), "dist": torch.rand(B, 1, device=dev), "time": torch.rand(B, 1, device=dev)}
for _ in range(3): pol.net_forward(obs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): pol.net_forward(obs)
e1.record(); torch.cuda.synchronize()
print("policy forward (actor+critic) B=256:", e0.elapsed_time(e1) / 20, "ms")
# one PPO minibatch update (forward + loss + backward + clip + AdamW), the unit the 4 updates of a bench iteration repeat
from types import SimpleNamespace
mb = SimpleNamespace(obs=obs, act=torch.randn(B, 128, device=dev), logp_old=torch.randn(B, device=dev) - 180.0,
                     adv=torch.randn(B, device=dev), returns=torch.randn(B, device=dev))
for _ in range(3): pol.learn_minibatch(mb)
torch.cuda.synchronize()
e0.record()
for _ in range(20): pol.learn_minibatch(mb)
e1.record(); torch.cuda.synchronize()
print("policy learn_minibatch B=256:", e0.elapsed_time(e1) / 20, "ms")
