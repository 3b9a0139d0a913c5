"""Policy networks - host-side mirror of the reference's motion/models/models_policy_ppo.py
(MLPBlock :24-39, GAMMAPolicyBase :233-306, GAMMAActor :309-330, GAMMACritic :334-350,
ActorCritic :353-358). The modules hold parameters under the reference's state_dict names; forward /
backward run in the CUDA library through GAMMAPPOPolicy (egogen_b200/ppo_policy.py)."""
from __future__ import annotations

from torch import nn


class _MLP(nn.Module):
    def __init__(self, in_dim, h_dims):
        super().__init__()
        self.layers = nn.ModuleList()
        d = in_dim
        for h in h_dims:
            self.layers.append(nn.Linear(d, h))
            d = h


class MLPBlock(nn.Module):
    def __init__(self, h_dim, out_dim, n_blocks, actfun="lrelu", residual=True):
        super().__init__()
        if actfun != "lrelu" or not residual:
            raise NotImplementedError("crowd_ppo uses residual lrelu blocks (MPVAEPolicy_samp_collision.yaml)")
        self.layers = nn.ModuleList([_MLP(h_dim, (h_dim, h_dim)) for _ in range(n_blocks)])
        self.out_fc = nn.Linear(h_dim, out_dim)


def _get(config, key, default=None):
    return config.get(key, default) if isinstance(config, dict) else getattr(config, key, default)


class GAMMAPolicyBase(nn.Module):
    """marker (GRU 402->512) and ego-sensing (GRU 32->512) encoders; dist/time positional encodings have no
    parameters."""

    def __init__(self, config):
        super().__init__()
        self.h_dim, self.z_dim, self.n_blocks = _get(config, "h_dim"), _get(config, "z_dim"), _get(config, "n_blocks")
        if _get(config, "body_repr") not in ("ssm2_67_condi_marker", "ssm2_67_condi_marker_map"):
            raise NotImplementedError("crowd_ppo uses body_repr ssm2_67_condi_marker_map (in_dim 402)")
        self.in_dim = 67 * 3 * 2
        self.x_enc = nn.GRU(self.in_dim, self.h_dim)
        self.ego_enc = nn.GRU(32, self.h_dim)


class GAMMAActor(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.h_dim, self.z_dim, self.n_blocks = _get(config, "h_dim"), _get(config, "z_dim"), _get(config, "n_blocks")
        self.min_logvar, self.max_logvar = _get(config, "min_logvar", -1), _get(config, "max_logvar", 3)
        self.pnet = MLPBlock(self.h_dim * 2 + 128, self.z_dim * 2, self.n_blocks, _get(config, "actfun"))


class GAMMACritic(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.h_dim, self.z_dim, self.n_blocks = _get(config, "h_dim"), _get(config, "z_dim"), _get(config, "n_blocks")
        self.vnet = MLPBlock(self.h_dim * 2 + 128, 1, self.n_blocks, _get(config, "actfun"))


class ActorCritic(nn.Module):
    def __init__(self, actor, critic, shared_net=None):
        super().__init__()
        self.actor, self.critic = actor, critic
        if shared_net is not None:
            self.shared_net = shared_net
