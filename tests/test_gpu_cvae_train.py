"""GPU parity of the C-VAE training step (BASELINE config 3): loss values and every parameter gradient of the
hand-written forward/backward against torch autograd on the reference-pinned oracle predictor; Adam vs torch.optim.Adam."""
import pytest
import torch

from egogen_b200.assets import fill_params_

pytestmark = pytest.mark.gpu


def _setup(dev, seed=41):
    from egogen_b200.train_gamma_predictor import GAMMAPrimitiveVAETrainOP
    from oracle import nets
    op = GAMMAPrimitiveVAETrainOP(device=dev)
    op.build_model(seed=0)
    fill_params_(op.model, seed=seed)
    orc = nets.PredictorOracle().train()
    orc.load_state_dict(op.model.state_dict())
    return op, orc


def _check_grads(op, orc, rtol=3e-3):
    for (name, p), q in zip(op.model.named_parameters(), orc.parameters()):
        gg, gr = p.grad.cpu(), q.grad
        assert gr is not None, name
        assert (gg - gr).norm().item() <= rtol * gr.norm().item() + 1e-7, (name, (gg - gr).norm().item(), gr.norm().item())


def test_single_primitive_loss_and_grads():
    from oracle import cvae_train as oc
    dev = torch.device("cuda:0")
    op, orc = _setup(dev)
    g = torch.Generator().manual_seed(1)
    B = 48
    data = torch.cumsum(torch.randn(20, B, 201, generator=g) * 0.02, dim=0) + torch.randn(1, B, 201, generator=g) * 0.3
    eps = torch.randn(B, 128, generator=g)
    loss, items = op.calc_loss(data.to(dev), 0, eps=eps.to(dev))
    ref, rec, kld, _ = oc.primitive_loss(orc, data[:2], data[2:], eps)
    ref.backward()
    assert abs(loss - ref.item()) < 1e-4 * max(1, abs(ref.item()))
    assert abs(items[1] - rec.item()) < 1e-4 and abs(items[2] - kld.item()) < 1e-4
    _check_grads(op, orc)
    # Adam step on identical gradients
    opt = torch.optim.Adam(orc.parameters(), lr=5e-4)
    for p, q in zip(op.model.parameters(), orc.parameters()):
        p.grad.copy_(q.grad.to(dev))
    opt.step(); op.optimizer_step(5e-4)
    for (name, p), q in zip(op.model.named_parameters(), orc.parameters()):
        assert torch.allclose(p.detach().cpu(), q.detach(), atol=2e-6), name


def test_rollout_loss_and_grads():
    from egogen_b200.train_gamma_predictor import SyntheticPrimitiveBatchGen
    from oracle import cvae_train as oc
    dev = torch.device("cuda:0")
    op, orc = _setup(dev, seed=43)
    gen = SyntheticPrimitiveBatchGen(16, 80, dev, seed=2)           # 80 frames -> 3 chained primitives
    mk, jt = gen.next_batch_with_jts(16)
    g = torch.Generator().manual_seed(5)
    eps = [torch.randn(16, 128, generator=g) for _ in range(8)]
    loss, _ = op.calc_loss_rollout((mk, jt), 0, eps_list=[e.to(dev) for e in eps])
    ref = oc.rollout_loss(orc, mk.cpu(), jt.cpu(), eps)
    ref.backward()
    assert abs(loss - ref.item()) < 2e-4 * max(1, abs(ref.item())), (loss, ref.item())
    _check_grads(op, orc, rtol=5e-3)


def test_train_loop_reduces_loss(tmp_path):
    from egogen_b200.train_gamma_predictor import GAMMAPrimitiveVAETrainOP, SyntheticPrimitiveBatchGen
    dev = torch.device("cuda:0")
    logs = []
    op = GAMMAPrimitiveVAETrainOP(trainconfig={"batch_size": 32, "num_epochs": 3, "num_epochs_fix": 1, "max_rollout": 2,
                                               "saving_per_X_ep": 3, "save_dir": str(tmp_path)}, device=dev)
    op.train(SyntheticPrimitiveBatchGen(64, 60, dev, seed=3), log=logs.append)
    rec = [float(l.split("REC=")[1].split(",")[0]) for l in logs]
    assert len(rec) == 3 and rec[-1] < rec[0]
    ck = torch.load(tmp_path / "epoch-3.ckp", map_location="cpu")
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict"} and ck["epoch"] == 3
    assert "d_rnn.weight_ih" in ck["model_state_dict"] and "x_enc.weight_ih_l0" in ck["model_state_dict"]


def test_resume_restores_adam_state(tmp_path):
    """resume_training (models_GAMMA_primitive.py:517-531): weights, Adam moments and step count continue from the
    last checkpoint, so a 2-epoch run resumed for the third epoch ends where the uninterrupted 3-epoch run ends."""
    from egogen_b200.train_gamma_predictor import GAMMAPrimitiveVAETrainOP, SyntheticPrimitiveBatchGen
    dev = torch.device("cuda:0")
    base = {"batch_size": 32, "num_epochs_fix": 1, "max_rollout": 2, "saving_per_X_ep": 1}

    def run(num_epochs, save_dir, resume):
        torch.manual_seed(5)
        op = GAMMAPrimitiveVAETrainOP(trainconfig=dict(base, num_epochs=num_epochs, save_dir=str(save_dir), resume_training=resume),
                                      device=dev)
        op.train(SyntheticPrimitiveBatchGen(64, 60, dev, seed=3), log=lambda *_: None)
        return op

    full = run(3, tmp_path / "full", False)
    part = run(2, tmp_path / "part", False)
    ck = torch.load(tmp_path / "part" / "epoch-2.ckp", map_location="cpu")
    resumed = GAMMAPrimitiveVAETrainOP(trainconfig=dict(base, num_epochs=3, save_dir=str(tmp_path / "part"), resume_training=True),
                                       device=dev)
    resumed.build_model()
    resumed.load_optimizer_state_dict(ck["optimizer_state_dict"])
    assert resumed._step == part._step > 0
    assert torch.equal(resumed.exp_avg, part.exp_avg) and torch.equal(resumed.exp_avg_sq, part.exp_avg_sq)
    with pytest.raises(FileExistsError):
        GAMMAPrimitiveVAETrainOP(trainconfig=dict(base, num_epochs=3, save_dir=str(tmp_path / "none"), resume_training=True),
                                 device=dev).train(SyntheticPrimitiveBatchGen(64, 60, dev, seed=3), log=lambda *_: None)
    assert full._step == 3 * part._step // 2
