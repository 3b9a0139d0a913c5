"""How far are the CUDA policy gradients from an exact reference? Compares eg_ppo_loss_backward with the autograd oracle in
fp32 (what the test uses) and in float64 (the referee), tensor by tensor: worst element / max|g| and norm ratio."""
import copy
import sys

import torch

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gpu_ppo as T
from egogen_b200.ppo_policy import Batch
from oracle import ppo as oppo

dev = torch.device("cuda:0")
for B in (256, 32):
    pol, (oa, oc, os_) = T._make(dev)
    obs = T._obs(B, 5)
    g = torch.Generator().manual_seed(6)
    with torch.no_grad():
        mu, lv = oa(os_(obs))
        sig = torch.exp(lv.clamp(-2.5, 2.5)) ** 0.5
        act = mu + sig * torch.randn(B, 128, generator=g)
        from torch.distributions import Independent, Normal
        logp_old = Independent(Normal(mu, sig), 1).log_prob(act) + torch.randn(B, generator=g) * 0.08
    adv = torch.randn(B, generator=g); ret = torch.randn(B, generator=g)
    d = lambda m: copy.deepcopy(m).double()
    oa64, oc64, os64 = d(oa), d(oc), d(os_)
    oppo.learn_minibatch(oa, oc, os_, obs, act, logp_old, adv, ret)
    oppo.learn_minibatch(oa64, oc64, os64, {k: v.double() for k, v in obs.items()}, act.double(), logp_old.double(), adv.double(), ret.double())
    mb = Batch(obs={k: v.to(dev) for k, v in obs.items()}, act=act.to(dev), logp_old=logp_old.to(dev), adv=adv.to(dev), returns=ret.to(dev))
    pol.loss_backward(mb)
    worst = {"gpu_vs_f64": [0, 0], "gpu_vs_f32": [0, 0], "f32_vs_f64": [0, 0]}
    ps = list(pol.actor.named_parameters()) + list(pol.critic.named_parameters()) + list(pol.shared_net.named_parameters())
    q32 = list(oa.parameters()) + list(oc.parameters()) + list(os_.parameters())
    q64 = list(oa64.parameters()) + list(oc64.parameters()) + list(os64.parameters())
    for (name, p), a32, a64 in zip(ps, q32, q64):
        gg, g32, g64 = p.grad.cpu().double(), a32.grad.double(), a64.grad
        sc, nr = g64.abs().max().item() + 1e-12, g64.norm().item() + 1e-12
        for key, x, y in (("gpu_vs_f64", gg, g64), ("gpu_vs_f32", gg, g32), ("f32_vs_f64", g32, g64)):
            e, n = (x - y).abs().max().item() / sc, (x - y).norm().item() / nr
            if e > worst[key][0]: worst[key][0] = e; worst[key].append(("elem", name, e))
            if n > worst[key][1]: worst[key][1] = n
    print(f"B={B}:", {k: (f"{v[0]:.2e}", f"{v[1]:.2e}") for k, v in worst.items()})
