"""Network oracles (TEST INFRASTRUCTURE): the C-VAE marker predictor, marker->body regressor,
VPoser v1 encoder and the PPO policy nets, restated as PyTorch-CPU modules whose parameter names
match the reference's state_dict keys, so reference-initialised weights load unchanged.

Follows:
  motion/models/baseops.py:615-641                (MLP)
  motion/models/models_GAMMA_primitive.py:36-133  (GAMMAPrimitiveVAE.decode / sample_prior)
  motion/models/models_GAMMA_primitive.py:160-301 (ResNetBlock, MoshRegressor)
  motion/models/models_GAMMA_primitive.py:307-360 (GAMMAPrimitiveCombo.sample_prior)
  motion/models/models_policy_ppo.py:24-39,233-350 (MLPBlock, GAMMAPolicyBase/Actor/Critic)
  human_body_prior VPoser v1.0 encoder            (third-party, absent => parity unpinned)
Pinned against the reference classes by tests/golden/nets_*.npz (tests/golden/gen_golden.py).
"""
import torch
from torch import nn
import torch.nn.functional as F

from . import tgm


class MLP(nn.Module):
    def __init__(self, in_dim, h_dims=(128, 128), activation="tanh"):
        super().__init__()
        self.activation = {"tanh": torch.tanh, "relu": torch.relu,
                           "lrelu": nn.LeakyReLU()}[activation]
        self.out_dim = h_dims[-1]
        self.layers = nn.ModuleList()
        d = in_dim
        for h in h_dims:
            self.layers.append(nn.Linear(d, h))
            d = h

    def forward(self, x):
        for fc in self.layers:
            x = self.activation(fc(x))
        return x


class PredictorOracle(nn.Module):
    """GAMMAPrimitiveVAE with cfg MPVAE_samp20_2frame_rollout.yml (h 256, z 128, hdims [512,256],
    use_drnn_mlp, residual, body_repr ssm2_67). Encoder branches kept so state_dicts load."""

    def __init__(self, in_dim=201, h_dim=256, z_dim=128, hdims_mlp=(512, 256)):
        super().__init__()
        self.in_dim, self.h_dim, self.z_dim = in_dim, h_dim, z_dim
        self.x_enc = nn.GRU(in_dim, h_dim)
        self.e_rnn = nn.GRU(in_dim, h_dim)
        self.e_mlp = MLP(2 * h_dim, list(hdims_mlp), "tanh")
        self.e_mu = nn.Linear(self.e_mlp.out_dim, z_dim)
        self.e_logvar = nn.Linear(self.e_mlp.out_dim, z_dim)
        self.drnn_mlp = MLP(h_dim, list(hdims_mlp) + [h_dim], "tanh")
        self.d_rnn = nn.GRUCell(in_dim + z_dim + h_dim, h_dim)
        self.d_mlp = MLP(h_dim, list(hdims_mlp), "tanh")
        self.d_out = nn.Linear(self.d_mlp.out_dim, in_dim)

    def decode(self, x, z, t_pred):
        _, hx = self.x_enc(x)
        hx = hx[0]
        h_rnn = self.drnn_mlp(hx)
        y = []
        y_i = None
        for i in range(t_pred):
            y_p = x[-1][:, :self.in_dim] if i == 0 else y_i
            h_rnn = self.d_rnn(torch.cat([hx, z, y_p], dim=-1), h_rnn)
            y_i = self.d_out(self.d_mlp(h_rnn)) + y_p
            y.append(y_i)
        return torch.stack(y)

    def sample_prior(self, x, z):
        return self.decode(x, z, 20 - x.shape[0])


class ResNetBlock(nn.Module):
    def __init__(self, in_dim, h_dim, out_dim, n_blocks, actfun="relu"):
        super().__init__()
        self.in_fc = nn.Linear(in_dim, h_dim)
        self.layers = nn.ModuleList([MLP(h_dim, (h_dim, h_dim), actfun) for _ in range(n_blocks)])
        self.out_fc = nn.Linear(h_dim, out_dim)

    def forward(self, x):
        h = self.in_fc(x)
        for layer in self.layers:
            h = layer(h) + h
        return self.out_fc(h)


class RegressorOracle(nn.Module):
    """MoshRegressor with cfg MoshRegressor_v3_male.yml (h 128, 10 blocks, 3 recurrences, relu, use_cont)."""

    def __init__(self, h_dim=128, n_blocks=10, n_recur=3):
        super().__init__()
        self.in_dim, self.n_recur = 201, n_recur
        self.body_dim = 3 + 6 + 21 * 6 + 24
        self.pnet = ResNetBlock(self.in_dim + self.body_dim + 10, h_dim, self.body_dim, n_blocks, "relu")

    def forward_cont(self, marker_ref, betas):
        xr = marker_ref.reshape(-1, self.in_dim)
        xb = torch.zeros(xr.shape[0], self.body_dim)
        for _ in range(self.n_recur):
            xb = self.pnet(torch.cat([xr, xb, betas], dim=-1)) + xb
        return xb

    @staticmethod
    def cont2aa(xb):
        """models_GAMMA_primitive.py:208-219."""
        n = xb.shape[0]
        aa = tgm.cont2aa(xb[:, 3:3 + 22 * 6].contiguous().view(n, -1, 6)).reshape(n, -1)
        return torch.cat([xb[:, :3], aa[:, :3], aa[:, 3:], xb[:, 135:147], xb[:, 147:]], dim=-1)

    def forward(self, marker_ref, betas):
        return self.cont2aa(self.forward_cont(marker_ref, betas))


class ComboOracle(nn.Module):
    """GAMMAPrimitiveCombo.sample_prior (models_GAMMA_primitive.py:334-360)."""

    def __init__(self):
        super().__init__()
        self.predictor = PredictorOracle()
        self.regressor = RegressorOracle()

    def sample_prior(self, X, betas, z):
        Y = self.predictor.sample_prior(X, z)
        nt, nb = Y.shape[:2]
        Yb = self.regressor(Y.reshape(nt * nb, -1), betas.reshape(nt * nb, -1)).view(nt, nb, -1)
        return Y, Yb


class VPoserEncoderOracle(nn.Module):
    """human_body_prior VPoser v1.0 encoder, eval mode, ``.loc`` only."""

    def __init__(self, num_neurons=512, latentD=32, n_features=63):
        super().__init__()
        self.bodyprior_enc_bn1 = nn.BatchNorm1d(n_features)
        self.bodyprior_enc_fc1 = nn.Linear(n_features, num_neurons)
        self.bodyprior_enc_bn2 = nn.BatchNorm1d(num_neurons)
        self.bodyprior_enc_fc2 = nn.Linear(num_neurons, num_neurons)
        self.bodyprior_enc_mu = nn.Linear(num_neurons, latentD)
        self.bodyprior_enc_logvar = nn.Linear(num_neurons, latentD)
        self.eval()

    def encode_loc(self, pin):
        x = pin.view(pin.size(0), -1)
        x = self.bodyprior_enc_bn1(x)
        x = F.leaky_relu(self.bodyprior_enc_fc1(x), negative_slope=0.2)
        x = self.bodyprior_enc_bn2(x)
        x = F.leaky_relu(self.bodyprior_enc_fc2(x), negative_slope=0.2)
        return self.bodyprior_enc_mu(x)


class MLPBlock(nn.Module):
    def __init__(self, h_dim, out_dim, n_blocks, actfun="lrelu"):
        super().__init__()
        self.layers = nn.ModuleList([MLP(h_dim, (h_dim, h_dim), actfun) for _ in range(n_blocks)])
        self.out_fc = nn.Linear(h_dim, out_dim)

    def forward(self, x):
        h = x
        for layer in self.layers:
            h = layer(h) + h
        return self.out_fc(h)


class PolicyBaseOracle(nn.Module):
    """GAMMAPolicyBase (models_policy_ppo.py:233-306), body_repr ssm2_67_condi_marker_map -> in 402."""

    def __init__(self, h_dim=512, in_dim=402):
        super().__init__()
        self.x_enc = nn.GRU(in_dim, h_dim)
        self.ego_enc = nn.GRU(32, h_dim)

    @staticmethod
    def positional_encoding(inp, L):
        freq_bands = 2.0 ** torch.linspace(0.0, L - 1, L)
        out = []
        for freq in freq_bands:
            for fn in (torch.sin, torch.cos):
                out.append(fn(inp * freq))
        return torch.cat(out, -1)

    def forward(self, obs):
        x_in = obs["state"].permute([1, 0, 2])
        nb = x_in.shape[1]
        _, hx = self.x_enc(x_in)
        _, he = self.ego_enc(obs["egosensing"].permute([1, 0, 2]))
        dist = self.positional_encoding(obs["dist"].reshape(nb, 1), 32)
        time_feat = self.positional_encoding(obs["time"].reshape(nb, 1), 32)
        return torch.cat([hx[0], he[0], dist, time_feat], dim=-1)


class ActorOracle(nn.Module):
    def __init__(self, h_dim=512, z_dim=128, n_blocks=2):
        super().__init__()
        self.z_dim = z_dim
        self.min_logvar, self.max_logvar = -2.5, 2.5
        self.pnet = MLPBlock(h_dim * 2 + 128, z_dim * 2, n_blocks)

    def forward(self, hx):
        z = self.pnet(hx)
        return z[:, :self.z_dim], z[:, self.z_dim:]


class CriticOracle(nn.Module):
    def __init__(self, h_dim=512, n_blocks=2):
        super().__init__()
        self.vnet = MLPBlock(h_dim * 2 + 128, 1, n_blocks)

    def forward(self, hx):
        return self.vnet(hx)


def init_policy_nets(seed: int = 0):
    """main_ppo.py:104-132: seed, orthogonal(gain sqrt 2)+zero bias on every Linear, then every Linear
    of actor.pnet scaled by 0.01; GRUs keep the default init."""
    import numpy as np
    torch.manual_seed(seed)
    actor, critic, shared = ActorOracle(), CriticOracle(), PolicyBaseOracle()
    # module iteration order of ActorCritic(actor, critic, shared_net)
    for net in (actor, critic, shared):
        for m in net.modules():
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight, gain=np.sqrt(2))
                nn.init.zeros_(m.bias)
    for m in actor.pnet.modules():
        if isinstance(m, nn.Linear):
            nn.init.zeros_(m.bias)
            m.weight.data.copy_(0.01 * m.weight.data)
    return actor, critic, shared
