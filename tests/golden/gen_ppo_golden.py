"""Golden vectors for the PPO update from the REFERENCE'S OWN GAMMAPPOPolicy (motion/crowd_ppo/ppo_policy.py) - build
container only. forward (:142-179) and learn (:182-265) run unmodified on the reference's GAMMAActor / GAMMACritic /
GAMMAPolicyBase with the optimiser and distribution of main_ppo.py:134-137. tianshou is absent; what the policy needs
from it is supplied here: a `Batch` with attribute access and split(), and a PPOPolicy base whose __init__ stores the
constructor arguments under tianshou's attribute names (`_weight_vf`, `_weight_ent`, `_grad_norm`, `_eps`, `_actor_critic`
= ActorCritic(actor, critic) WITHOUT the shared net, which is why the gradient clip skips the GRU encoders). GAE
(`compute_episodic_return`) lives entirely inside tianshou and stays unpinned.

Run:  python tests/golden/gen_ppo_golden.py   ->  tests/golden/ppo_golden.npz
"""
import os
import sys
import types

import numpy as np
import torch
from torch import nn
from torch.distributions import Independent, Normal

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = os.environ.get("EGOGEN_REFERENCE", "/root/reference")

from egogen_b200.assets import fill_params_                   # noqa: E402
from oracle.ppo import EPS                                    # noqa: E402


class Batch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __len__(self):
        return len(self["act"])

    def split(self, size, shuffle=True, merge_last=False):
        n = len(self)
        assert size >= n, "single-minibatch fixture"
        idx = np.random.permutation(n) if shuffle else np.arange(n)
        take = lambda v: {k: take(x) for k, x in v.items()} if isinstance(v, dict) else v[idx]
        yield Batch({k: take(v) for k, v in self.items() if v is not None})


class _TsActorCritic(nn.Module):
    def __init__(self, actor, critic):
        super().__init__()
        self.actor, self.critic = actor, critic


class PPOPolicy(nn.Module):
    def __init__(self, actor, critic, optim, dist_fn, vf_coef=0.5, ent_coef=0.01, max_grad_norm=None, deterministic_eval=False,
                 **kwargs):
        super().__init__()
        self.optim, self.dist_fn = optim, dist_fn
        self._weight_vf, self._weight_ent, self._grad_norm = vf_coef, ent_coef, max_grad_norm
        self._eps = EPS
        self._deterministic_eval = deterministic_eval
        self._actor_critic = _TsActorCritic(actor, critic)


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


stub("tianshou")
stub("tianshou.data", Batch=Batch, ReplayBuffer=object, to_torch_as=lambda x, y: torch.as_tensor(x).to(y))
stub("tianshou.policy", PPOPolicy=PPOPolicy)
for m in ["smplx", "torchgeometry", "tensorboardX", "matplotlib", "matplotlib.pyplot", "omegaconf"]:
    stub(m)
sys.modules["tensorboardX"].SummaryWriter = object
sys.path.insert(0, os.path.join(REF, "motion"))
from models import models_policy_ppo as ref_policy           # noqa: E402
from crowd_ppo import ppo_policy as ref_ppo                   # noqa: E402

torch.set_num_threads(4)
cfg = {"h_dim": 512, "z_dim": 128, "n_blocks": 2, "actfun": "lrelu", "body_repr": "ssm2_67_condi_marker_map",
       "min_logvar": -2.5, "max_logvar": 2.5}
actor, critic, shared = ref_policy.GAMMAActor(cfg), ref_policy.GAMMACritic(cfg), ref_policy.GAMMAPolicyBase(cfg)
fill_params_(actor, seed=21); fill_params_(critic, seed=22); fill_params_(shared, seed=23)
with torch.no_grad():
    for mod in actor.pnet.modules():                          # main_ppo.py:127-131: small initial actions
        if isinstance(mod, nn.Linear):
            mod.weight.mul_(0.01)
actor_critic = ref_policy.ActorCritic(actor, critic, shared)
optim = torch.optim.AdamW(actor_critic.parameters(), lr=3e-4, weight_decay=0.01)        # main_ppo.py:134
policy = ref_ppo.GAMMAPPOPolicy(actor, critic, shared, optim, lambda *l: Independent(Normal(*l), 1), eps_clip=0.1,
                                vf_coef=1.0, ent_coef=0.01, max_grad_norm=0.1, advantage_normalization=True,
                                recompute_advantage=False, value_clip=False, dual_clip=None)
g = torch.Generator().manual_seed(404)
B = 12
obs = {"state": torch.randn(B, 2, 402, generator=g) * 0.5, "egosensing": torch.rand(B, 2, 32, generator=g) * 2 - 1,
       "dist": torch.rand(B, 1, generator=g), "time": 1 - torch.randint(0, 13, (B, 1), generator=g) / 13.0}
policy.train()
torch.manual_seed(5)
with torch.no_grad():
    fw = policy(Batch(obs=obs, info={}))
act = fw.act
logp_old = fw.dist.log_prob(act) + torch.randn(B, generator=g) * 0.05
adv = torch.randn(B, generator=g)
returns = torch.randn(B, generator=g)
with torch.no_grad():
    v_s = critic(shared(obs)).flatten()
batch = Batch(obs=obs, act=act, logp_old=logp_old, adv=adv.clone(), returns=returns, v_s=v_s, info={},
              z_mu=fw.z_mu, z_var=fw.z_var, z_logvar=fw.z_logvar)
np.random.seed(0)
res = policy.learn(batch, batch_size=256, repeat=1)
pn = lambda mod: np.array([p.detach().norm().item() for p in mod.parameters()])
gn = lambda mod: np.array([p.grad.norm().item() for p in mod.parameters()])
out = dict(state=obs["state"], egosensing=obs["egosensing"], dist=obs["dist"], time=obs["time"], act=act, logp_fw=fw.dist.log_prob(act),
           z_mu=fw.z_mu, z_logvar=fw.z_logvar, logp_old=logp_old, adv=adv, returns=returns,
           loss=res["loss"][0], clip=res["loss/clip"][0], vf=res["loss/vf"][0], ent=res["loss/ent"][0], kld=res["loss/kld"][0],
           actor_gradnorm=gn(actor), critic_gradnorm=gn(critic), shared_gradnorm=gn(shared),
           actor_after=pn(actor), critic_after=pn(critic), shared_after=pn(shared),
           actor_out_w_after=actor.pnet.out_fc.weight.detach()[:8, :16].clone())
np.savez_compressed(os.path.join(HERE, "ppo_golden.npz"), **{k: np.asarray(v) for k, v in out.items()})
print("wrote ppo_golden.npz", {k: np.asarray(v).shape for k, v in out.items()}, res)
