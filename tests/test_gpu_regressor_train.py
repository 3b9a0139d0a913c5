"""GPU parity of the body-regressor training step (GAMMARegressorTrainOP.calc_loss + backward,
models_GAMMA_primitive.py:594-633): loss values, regressed body parameters and every parameter gradient of the hand-written
forward/backward (recurrent ResNet MLP, Gram-Schmidt 6-D rotations, SMPL-X markers) against torch autograd through the
reference-pinned regressor oracle, its tgm cont2aa and the LBS oracle; Adam vs torch.optim.Adam; a short training run."""
import numpy as np
import pytest
import torch

from egogen_b200 import assets
from egogen_b200.assets import fill_params_

pytestmark = pytest.mark.gpu


def _setup(dev, smplx_model, seed, w_gain=0.5):     # gain 1 makes the 10-block residual stack explode to |xb| ~ 500
    from egogen_b200.train_gamma_regressor import GAMMARegressorTrainOP
    from oracle import nets
    from oracle.smplx_lbs import SMPLXParserOracle
    op = GAMMARegressorTrainOP(device=dev)
    op.build_model(seed=0)
    fill_params_(op.model, seed=seed, w_gain=w_gain)
    with torch.no_grad():        # rotations away from the degenerate all-zero 6-D vector (GS is singular there)
        g = torch.Generator().manual_seed(seed)
        b = op.model.pnet.out_fc.bias
        pat = torch.tensor([1.0, 0.0, 0.0, 1.0, 0.0, 0.0]).repeat(22) + torch.randn(132, generator=g) * 0.3
        b[3:135] = pat.to(dev)
    orc = nets.RegressorOracle().train()
    orc.load_state_dict(op.model.state_dict())
    lbs = SMPLXParserOracle(smplx_model, marker=assets.marker_ids())
    return op, orc, lbs


def _inputs(lbs, M, seed):
    g = torch.Generator().manual_seed(seed)
    xb = torch.randn(M, 93, generator=g) * 0.3
    betas = torch.randn(M, 10, generator=g) * 0.5
    with torch.no_grad():
        mk = lbs.forward_smplx(betas, "male", xb, "markers").reshape(M, 67, 3)
    return mk + torch.randn(M, 67, 3, generator=g) * 0.01, betas


def _oracle_loss(orc, lbs, mk, betas, w):
    xb = orc(mk, betas)
    pred = lbs.forward_smplx(betas, "male", xb, "markers").reshape(mk.shape)
    lm = torch.nn.functional.l1_loss(mk, pred)
    lh = torch.mean(xb[:, 69:] ** 2)
    return xb, lm + w * lh, lm, lh


@pytest.mark.parametrize("M,seed", [(12, 3), (70, 4), (1, 6)])
def test_regressor_loss_and_grads(smplx_model, M, seed):
    dev = torch.device("cuda:0")
    op, orc, lbs = _setup(dev, smplx_model, seed)
    mk, betas = _inputs(lbs, M, seed + 10)
    xb, loss, items = op.forward_loss_backward(mk.to(dev), betas.to(dev))
    xb_ref, ref, lm, lh = _oracle_loss(orc, lbs, mk, betas, op.lossconfig["weight_reg_hpose"])
    ref.backward()
    assert torch.isfinite(xb).all()
    err = (xb.cpu() - xb_ref.detach()).abs().max().item()
    assert err < 1e-4 * max(1.0, xb_ref.abs().max().item()), err
    assert abs(loss - ref.item()) < 1e-4 * max(1.0, abs(ref.item())), (loss, ref.item())
    assert abs(items[0] - lm.item()) < 1e-4 and abs(items[1] - lh.item()) < 1e-4 * max(1.0, lh.item())
    bad = []
    for (name, p), q in zip(op.model.named_parameters(), orc.parameters()):
        gg, gr = p.grad.cpu(), q.grad
        assert gr is not None and torch.isfinite(gg).all(), name
        e, n = (gg - gr).norm().item(), gr.norm().item()
        if e > 5e-3 * n + 1e-7:
            bad.append((name, e, n))
    assert not bad, bad


def test_regressor_adam_step_and_training_run(smplx_model, tmp_path):
    from egogen_b200.train_gamma_regressor import GAMMARegressorTrainOP, SyntheticBodyMarkerBatchGen
    from egogen_b200.smplx_parser import get_lbs_model
    dev = torch.device("cuda:0")
    op, orc, lbs = _setup(dev, smplx_model, 7)
    mk, betas = _inputs(lbs, 16, 21)
    op.forward_loss_backward(mk.to(dev), betas.to(dev))
    opt = torch.optim.Adam(orc.parameters(), lr=3e-4)
    for p, q in zip(op.model.parameters(), orc.parameters()):
        q.grad = p.grad.detach().cpu().clone()
    opt.step(); op.optimizer_step(3e-4)
    for (name, p), q in zip(op.model.named_parameters(), orc.parameters()):
        assert torch.allclose(p.detach().cpu(), q.detach(), atol=2e-6), name
    # short run on surrogate-body markers: the marker loss must go down, checkpoints keep the reference's layout
    bm = get_lbs_model("male", dev, marker_vids=assets.marker_ids())
    gen = SyntheticBodyMarkerBatchGen(bm, 32, 8, dev, seed=3)
    run = GAMMARegressorTrainOP(trainconfig={"batch_size": 16, "num_epochs": 6, "num_epochs_fix": 6, "saving_per_X_ep": 6,
                                             "learning_rate": 1e-3, "save_dir": str(tmp_path)}, device=dev)
    hist = run.train(gen, log=lambda s: None)
    assert np.isfinite(hist).all() and hist[-1][0] < 0.8 * hist[0][0], hist
    ck = torch.load(str(tmp_path / "epoch-6.ckp"), map_location="cpu")
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict"} and ck["epoch"] == 6
    from oracle import nets
    nets.RegressorOracle().load_state_dict(ck["model_state_dict"])      # same keys / shapes as the reference class
