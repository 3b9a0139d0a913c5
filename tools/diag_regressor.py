"""Where does the regressor error sit? GPU sample_prior (current build / env settings) vs the float64 oracle, per column
group of Yb: transl (direct xb columns), axis-angle (through 6D -> rotmat -> quaternion -> aa), hand PCA (direct)."""
import os, sys
import torch
sys.path.insert(0, ".")
from egogen_b200.models_gamma_primitive import GAMMAPrimitiveComboGenOP
from oracle import nets

dev = torch.device("cuda:0")
g = GAMMAPrimitiveComboGenOP(testconfig={"gpu_index": 0}); m = g.build_model(seed=0)
combo = nets.ComboOracle()
combo.predictor.load_state_dict(m.predictor.state_dict()); combo.regressor.load_state_dict(m.regressor.state_dict())
combo = combo.double().eval()
for B in (96, 256):
    gen = torch.Generator().manual_seed(100 + B)
    X = torch.randn(2, B, 201, generator=gen) * 0.3
    z = torch.randn(B, 128, generator=gen)
    betas = torch.randn(B, 10, generator=gen) * 0.5
    Y, Yb = m.sample_prior(X.to(dev), betas.unsqueeze(0).repeat(18, 1, 1).to(dev), z.to(dev))
    with torch.no_grad():
        Yo, Ybo = combo.sample_prior(X.double(), betas.double().unsqueeze(0).repeat(18, 1, 1), z.double())
        # regressor alone on the GPU's own Y (isolates the regressor from decode differences)
        xb_o = combo.regressor.forward_cont(Y.cpu().double().reshape(-1, 201), betas.double().unsqueeze(0).repeat(18, 1, 1).reshape(-1, 10))
        Yb_o2 = combo.regressor.cont2aa(xb_o).reshape(18, B, 93)
    Ybc = Yb.cpu().double()
    for name, ref in (("vs oracle end-to-end", Ybo), ("vs oracle regressor on GPU Y", Yb_o2)):
        d = (Ybc - ref).abs()
        tr, aa, hd = d[..., :3].max().item(), d[..., 3:69].max().item(), d[..., 69:].max().item()
        rel = ((Ybc[..., 69:] - ref[..., 69:]) / ref[..., 69:].abs().clamp_min(1e-3))
        print(f"B={B} {name}: transl {tr:.2e}  aa {aa:.2e}  hand {hd:.2e}   hand signed-rel mean {rel.mean().item():+.2e} std {rel.std().item():.2e}"
              f"  |dY| {(Y.cpu().double() - Yo).abs().max().item():.2e}", flush=True)
    n6 = xb_o[:, 3:135].reshape(-1, 22, 2, 3).norm(dim=-1)
    print(f"   6D column norms: min {n6.min().item():.3f} median {n6.median().item():.3f}", flush=True)
