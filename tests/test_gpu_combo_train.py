"""GPU parity of the joint predictor + regressor objective (GAMMAPrimitiveComboTrainOP.calc_loss_one,
models_GAMMA_primitive.py:787-838): the loss terms, the regressed bodies and every PREDICTOR gradient (KL term + the SMPL-X
cycle loss flowing back through the regressor into the predicted markers) against torch autograd on the oracles."""
import numpy as np
import pytest
import torch

from egogen_b200 import assets
from egogen_b200.assets import fill_params_

pytestmark = pytest.mark.gpu


def _setup(dev, smplx_model, sched):
    from egogen_b200.train_gamma_combo import GAMMAPrimitiveComboTrainOP
    from oracle import nets
    from oracle.smplx_lbs import SMPLXParserOracle
    op = GAMMAPrimitiveComboTrainOP(trainconfig={"scheduled_sampling": sched}, device=dev)
    op.build_model(seed=0)
    fill_params_(op.model.predictor, seed=41)
    fill_params_(op.model.regressor, seed=5, w_gain=0.5)
    with torch.no_grad():        # rotations away from the degenerate all-zero 6-D vector
        g = torch.Generator().manual_seed(9)
        pat = torch.tensor([1.0, 0.0, 0.0, 1.0, 0.0, 0.0]).repeat(22) + torch.randn(132, generator=g) * 0.3
        op.model.regressor.pnet.out_fc.bias[3:135] = pat.to(dev)
    pred, reg = nets.PredictorOracle().train(), nets.RegressorOracle().train()
    pred.load_state_dict(op.model.predictor.state_dict())
    reg.load_state_dict(op.model.regressor.state_dict())
    return op, pred, reg, SMPLXParserOracle(smplx_model, marker=assets.marker_ids())


@pytest.mark.parametrize("sched,B", [(False, 6), (True, 6), (True, 1)])
def test_combo_loss_one_and_predictor_grads(smplx_model, sched, B):
    from oracle import cvae_train as oc
    dev = torch.device("cuda:0")
    op, pred, reg, lbs = _setup(dev, smplx_model, sched)
    g = torch.Generator().manual_seed(3)
    ref = torch.cumsum(torch.randn(20, B, 201, generator=g) * 0.02, dim=0) + torch.randn(1, B, 201, generator=g) * 0.3
    betas = (torch.randn(1, B, 10, generator=g) * 0.5).expand(20, B, 10).contiguous()
    eps = torch.randn(B, 128, generator=g)
    loss, items, Yb = op.calc_loss_one([betas.to(dev), ref.to(dev)], 0, eps=eps.to(dev))
    lc = op.lossconfig
    ref_loss, ref_items, Yb_ref = oc.combo_loss_one(pred, reg, lbs, ref[:2], ref[2:], betas[2:], eps, lc["weight_rec"],
                                                    lc["weight_td"], lc["weight_kld"], lc["robust_kld"],
                                                    lc["weight_reg_hpose"], scheduled_sampling=sched)
    ref_loss.backward()
    assert (Yb.cpu() - Yb_ref.detach()).abs().max().item() < 2e-4 * max(1.0, Yb_ref.abs().max().item())
    assert abs(loss - ref_loss.item()) < 2e-4 * max(1.0, abs(ref_loss.item())), (loss, ref_loss.item())
    for a, b in zip(items, ref_items):
        assert abs(a - b.item()) < 2e-4 * max(1.0, abs(b.item())), (items, [x.item() for x in ref_items])
    bad = []
    for (name, p), q in zip(op.model.predictor.named_parameters(), pred.parameters()):
        gg, gr = p.grad.cpu(), q.grad
        assert gr is not None and torch.isfinite(gg).all(), name
        e, n = (gg - gr).norm().item(), gr.norm().item()
        if e > 5e-3 * n + 1e-7:
            bad.append((name, e, n))
    assert not bad, bad
    # the regressor is a fixed layer of this objective: its gradient buffer is untouched, and the optimiser step moves
    # the predictor only
    assert float(op._rop.flat_grads.abs().max()) == 0.0
    r0 = op._rop.flat_params.clone(); p0 = op._pop.flat_params.clone()
    op.optimizer_step(1e-4)
    assert torch.equal(op._rop.flat_params, r0) and not torch.equal(op._pop.flat_params, p0)
    assert set(op.model.state_dict()) == {"predictor." + k for k in pred.state_dict()} | {"regressor." + k for k in reg.state_dict()}


def test_cycle_loss_without_grad_and_call_order(smplx_model):
    """want_grad = 0 (the reference's torch.no_grad() evaluation of the regressor terms) reports the same terms and leaves
    the gradient output alone; eg_cvae_backward refuses to run before eg_cvae_forward_train sized the workspace."""
    import ctypes as C
    from egogen_b200 import _lib
    dev = torch.device("cuda:0")
    op, _, _, _ = _setup(dev, smplx_model, True)
    g = torch.Generator().manual_seed(8)
    T, B = 18, 3
    Yr = (torch.randn(T, B, 201, generator=g) * 0.3).to(dev)
    Y = (Yr.cpu() + torch.randn(T, B, 201, generator=g) * 0.02).to(dev)
    betas = (torch.randn(T, B, 10, generator=g) * 0.5).to(dev)
    L, st = _lib.lib(), _lib.stream_ptr(dev)
    outs = []
    for want in (1, 0):
        d = torch.full((T, B, 201), 7.0, device=dev)
        stats = torch.zeros(2, device=dev)
        _lib.check(L.eg_regressor_cycle_backward(op._rop._h, _lib.ptr(Yr), _lib.ptr(betas), _lib.ptr(Y), T, B, 1.0, 3.0, 0.01,
                                                 0.5, want, _lib.ptr(d), None, _lib.ptr(stats), st))
        outs.append((d.cpu(), stats.cpu()))
    assert torch.equal(outs[0][1], outs[1][1]) and float(outs[0][1][0]) > 0
    assert torch.isfinite(outs[0][0]).all() and float(outs[0][0].abs().max()) < 7.0        # overwritten with the gradient
    assert torch.equal(outs[1][0], torch.full((T, B, 201), 7.0))                              # untouched
    assert L.eg_regressor_cycle_backward(op._rop._h, _lib.ptr(Yr), _lib.ptr(betas), _lib.ptr(Y), T, B, 1.0, 3.0, 0.01, 1.0, 1,
                                         None, None, _lib.ptr(stats), st) != 0              # gradient requested, no buffer
    # a fresh predictor handle has no forward activations yet
    from egogen_b200.train_gamma_predictor import GAMMAPrimitiveVAETrainOP
    fresh = GAMMAPrimitiveVAETrainOP(device=dev)
    fresh.build_model(seed=1)
    X = torch.zeros(2, B, 201, device=dev); eps = torch.zeros(B, 128, device=dev); s4 = torch.zeros(4, device=dev)
    rc = L.eg_cvae_backward(fresh._h, _lib.ptr(X), _lib.ptr(Y), _lib.ptr(eps), B, 1.0, 3.0, 1.0, 1, 1.0, 1, _lib.ptr(Yr), None,
                            _lib.ptr(s4), st)
    assert rc != 0 and b"forward" in L.eg_last_error()
