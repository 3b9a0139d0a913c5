"""PPO oracle (TEST INFRASTRUCTURE): tianshou 0.5.0 GAE restated in numpy float64 and one minibatch of
GAMMAPPOPolicy.learn (motion/crowd_ppo/ppo_policy.py:189-252) with torch autograd on the oracle nets.
tianshou is not vendored / not installed => parity unpinned for the GAE piece. The minibatch update
IS pinned: tests/golden/gen_ppo_golden.py runs the reference's own GAMMAPPOPolicy.forward / learn over a tianshou shim and
tests/test_oracle_golden.py::test_ppo_oracle_matches_reference_policy_learn reproduces its losses, clipped gradients and
post-AdamW parameters.
"""
import numpy as np
import torch
from torch import nn
from torch.distributions import Independent, Normal

EPS = np.finfo(np.float32).eps.item()


def gae_return(v_s, v_s_, rew, end_flag, gamma, gae_lambda):
    """tianshou _gae_return (numba kernel) on 1-D arrays in buffer order."""
    returns = np.zeros(rew.shape)
    delta = rew + v_s_ * gamma - v_s
    discount = (1.0 - end_flag) * (gamma * gae_lambda)
    gae = 0.0
    for i in range(len(rew) - 1, -1, -1):
        gae = delta[i] + discount[i] * gae
        returns[i] = gae
    return returns


def compute_episodic_return(v_s, v_next, rew, terminated, truncated, unfinished_last, gamma=0.99, gae_lambda=0.95):
    """BasePolicy.compute_episodic_return on env-major flat arrays: v_next (float32) is masked by
    value_mask = ~terminated; end_flag = terminated | truncated | unfinished-last-index."""
    v_s_ = v_next * (~terminated)
    end_flag = np.logical_or(terminated, truncated)
    end_flag = np.logical_or(end_flag, unfinished_last)
    adv = gae_return(v_s, v_s_, rew, end_flag.astype(np.float64), gamma, gae_lambda)
    return adv + v_s, adv          # (returns, advantages), float64


def learn_minibatch(actor, critic, shared, obs, act, logp_old, adv, returns, eps_clip=0.1, vf_coef=1.0, ent_coef=0.01,
                    norm_adv=True):
    """Forward + loss + backward (grads left in .grad). Returns dict of scalars."""
    for m in (actor, critic, shared):
        for p in m.parameters():
            p.grad = None
    hx = shared(obs)
    z_mu, z_logvar = actor(hx)
    z_logvar = z_logvar.clamp(actor.min_logvar, actor.max_logvar)
    z_var = torch.exp(z_logvar)
    dist = Independent(Normal(z_mu, z_var ** 0.5), 1)
    if norm_adv:
        mean, std = adv.mean(), adv.std()
        adv = (adv - mean) / (std + EPS)
    log_prob_new = dist.log_prob(act)
    ratio = (log_prob_new - logp_old).exp().float()
    surr1 = ratio * adv
    surr2 = ratio.clamp(1.0 - eps_clip, 1.0 + eps_clip) * adv
    clip_loss = -torch.min(surr1, surr2).mean()
    value = critic(shared(obs)).flatten()
    vf_loss = (returns - value).pow(2).mean()
    kld = 0.5 * torch.mean(z_mu.pow(2))
    ent_loss = dist.entropy().mean()
    loss = clip_loss + vf_coef * vf_loss - ent_coef * ent_loss
    loss.backward()
    return dict(loss=loss.item(), clip=clip_loss.item(), vf=vf_loss.item(), ent=ent_loss.item(), kld=kld.item(),
                approx_kl=(logp_old - log_prob_new).mean().item(), logp=log_prob_new.detach(), value=value.detach(),
                mu=z_mu.detach(), logvar=z_logvar.detach())


def clip_and_adamw(actor, critic, shared, max_grad_norm=0.1, lr=3e-4, wd=0.01, optim=None):
    """ppo_policy.py:243-247 + main_ppo.py:134: clip over ActorCritic(actor, critic) only (quirk 1), AdamW on all."""
    params_ac = list(actor.parameters()) + list(critic.parameters())
    nn.utils.clip_grad_norm_(params_ac, max_norm=max_grad_norm)
    if optim is None:
        optim = torch.optim.AdamW(params_ac + list(shared.parameters()), lr=lr, weight_decay=wd)
    optim.step()
    return optim
