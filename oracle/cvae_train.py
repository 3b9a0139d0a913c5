"""C-VAE training oracle (TEST INFRASTRUCTURE): GAMMAPrimitiveVAE.forward (encode + reparameterise with a GIVEN eps +
decode) and the losses of GAMMAPrimitiveVAETrainOP.calc_loss / calc_loss_rollout
(motion/models/models_GAMMA_primitive.py:75-110, 400-432, 435-503) on the oracle predictor, gradients by torch autograd."""
import torch
import torch.nn.functional as F

from .smplx_lbs import SMPLXParserOracle


def forward(pred, x, y, eps):
    _, hx = pred.x_enc(x)
    _, hy = pred.e_rnn(y)
    h = pred.e_mlp(torch.cat((hx[0], hy[0]), dim=-1))
    mu, logvar = pred.e_mu(h), pred.e_logvar(h)
    z = mu + eps * torch.exp(0.5 * logvar)
    return pred.decode(x, z, y.shape[0]), mu, logvar


def primitive_loss(pred, X, Y, eps, w_rec=1.0, w_td=3.0, w_kld=1.0, robust=True):
    Y_rec, mu, logvar = forward(pred, X, Y, eps)
    loss_rec = w_rec * F.l1_loss(Y, Y_rec) + w_td * F.l1_loss(Y_rec[1:] - Y_rec[:-1], Y[1:] - Y[:-1])
    kld = 0.5 * torch.mean(-1 - logvar + mu.pow(2) + logvar.exp())
    if robust:
        kld = torch.sqrt(1 + kld ** 2) - 1
    return loss_rec + w_kld * kld, loss_rec, kld, Y_rec


def rollout_loss(pred, ref_markers, ref_jts, eps_list, max_rollout=8, t_his=2):
    n_t, n_b = ref_markers.shape[:2]
    ref_jts = ref_jts.contiguous().view(n_t, n_b, -1, 3)
    t_pred = 20 - t_his
    t, losses = 0, []
    Y_rec = R_prev = T_prev = None
    while t < n_t:
        if t + 20 >= n_t:
            break
        mk, jt = ref_markers[t:t + 20], ref_jts[t:t + 20]
        if t == 0:
            X, Y = mk[:t_his].detach(), mk[t_his:, :, :201].detach()
            R_prev, T_prev = SMPLXParserOracle.new_coordinate_from_joints(jt[0])
        else:
            R_curr, T_curr = SMPLXParserOracle.new_coordinate_from_joints(jt[0])
            Yg = mk[t_his:, :, :201].reshape(t_pred, n_b, -1, 3)
            Y = torch.einsum("bij,tbpj->tbpi", R_curr.permute(0, 2, 1), Yg - T_curr.unsqueeze(0))
            X_prev = Y_rec[-t_his:].reshape(t_his, n_b, -1, 3)
            Xg = torch.einsum("bij,tbpj->tbpi", R_prev, X_prev) + T_prev.unsqueeze(0)
            X = torch.einsum("bij,tbpj->tbpi", R_curr.permute(0, 2, 1), Xg - T_curr.unsqueeze(0))
            Y = Y.contiguous().view(t_pred, n_b, -1).detach()
            X = X.contiguous().view(t_his, n_b, -1).detach()
            R_prev, T_prev = R_curr, T_curr
        loss, _, _, Y_rec = primitive_loss(pred, X, Y, eps_list[len(losses)])
        losses.append(loss)
        t += t_pred
        if len(losses) >= max_rollout:
            break
    return torch.stack(losses).mean()


def combo_loss_one(pred, reg, lbs, X, Y, betas_Y, eps, w_rec=1.0, w_td=3.0, w_kld=1.0, robust=True, w_hpose=0.01,
                   scheduled_sampling=False):
    """GAMMAPrimitiveComboTrainOP.calc_loss_one (models_GAMMA_primitive.py:819-838) with calc_loss_marker (:797-815) and
    calc_loss_regressor (:787-794): the predictor's Y_rec goes through the regressor and SMPL-X (`lbs`, an
    SMPLXParserOracle) and is compared with the ground-truth markers. Returns (loss, [rec, kld, reg, hpose], Yb_rec)."""
    Y_rec, mu, logvar = forward(pred, X, Y, eps)
    nt, nb = Y_rec.shape[:2]
    rec_of = lambda a, b: w_rec * F.l1_loss(a, b) + w_td * F.l1_loss(b[1:] - b[:-1], a[1:] - a[:-1])
    loss_rec = rec_of(Y, Y_rec)
    kld = 0.5 * torch.mean(-1 - logvar + mu.pow(2) + logvar.exp())
    if robust:
        kld = torch.sqrt(1 + kld ** 2) - 1
    loss_marker = loss_rec + w_kld * kld if scheduled_sampling else w_kld * kld
    Yb = reg(Y_rec.contiguous().view(nt * nb, -1), betas_Y.contiguous().view(nt * nb, -1))
    x_pred = lbs.forward_smplx(betas_Y.reshape(nt * nb, 10), "male", Yb, "markers").reshape(nt, nb, -1)
    loss_reg = rec_of(Y, x_pred)
    loss_h = torch.mean(Yb[:, 69:] ** 2)
    loss = loss_marker + loss_reg + w_hpose * loss_h
    return loss, [loss_rec, kld, loss_reg, loss_h], Yb.view(nt, nb, -1)


def regressor_loss(reg, lbs, marker_ref, betas, w_hpose=0.01):
    """GAMMARegressorTrainOP: xb = model(marker_ref, betas); calc_loss (models_GAMMA_primitive.py:617-633).
    Returns (xb [M,93], loss, loss_marker, loss_hpose)."""
    xb = reg(marker_ref, betas)
    pred = lbs.forward_smplx(betas, "male", xb, "markers").reshape(marker_ref.shape)
    loss_marker = F.l1_loss(marker_ref, pred)
    loss_hpose = torch.mean(xb[:, 69:] ** 2)
    return xb, loss_marker + w_hpose * loss_hpose, loss_marker, loss_hpose
