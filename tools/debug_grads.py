import sys, torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import test_gpu_ppo as t
from egogen_b200.ppo_policy import Batch
from oracle import ppo as oppo
dev = torch.device('cuda:0')
for B in (256,):
    pol, (oa, oc, os_) = t._make(dev)
    obs = t._obs(B, 5)
    g = torch.Generator().manual_seed(6)
    with torch.no_grad():
        mu, lv = oa(os_(obs)); sig = torch.exp(lv.clamp(-2.5, 2.5)) ** 0.5
        act = mu + sig * torch.randn(B, 128, generator=g)
        from torch.distributions import Independent, Normal
        logp_old = Independent(Normal(mu, sig), 1).log_prob(act) + torch.randn(B, generator=g) * 0.08
    adv = torch.randn(B, generator=g); ret = torch.randn(B, generator=g)
    ref = oppo.learn_minibatch(oa, oc, os_, obs, act, logp_old, adv, ret)
    mb = Batch(obs={k: v.to(dev) for k, v in obs.items()}, act=act.to(dev), logp_old=logp_old.to(dev), adv=adv.to(dev), returns=ret.to(dev))
    pol.loss_backward(mb)
    print('stats', pol._stats.cpu().tolist(), {k: ref[k] for k in ('clip','vf','ent','kld','approx_kl')})
    names = [n for n,_ in list(pol.actor.named_parameters()) + list(pol.critic.named_parameters()) + list(pol.shared_net.named_parameters())]
    for name, p, q in zip(names, pol._ordered_params(), list(oa.parameters()) + list(oc.parameters()) + list(os_.parameters())):
        gg, gr = p.grad.cpu(), q.grad
        print(f"{name:40s} scale {gr.abs().max().item():.3e} maxdiff {(gg-gr).abs().max().item():.3e} rel {(gg-gr).abs().max().item()/(gr.abs().max().item()+1e-30):.2e}")
