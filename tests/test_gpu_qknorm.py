"""qk_norm presets (a13): the per-head LayerNorm kernels vs torch F.layer_norm, and the overlay FourM with qk_norm=True vs the
golden outputs of the UNMODIFIED reference (tests/golden/make_golden_qknorm.py)."""
import random

import pytest
import torch
import torch.nn.functional as F

from oracle import fourm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,H,packed", [(256, 6, True), (1000, 12, False), (37, 1, True)])
def test_headnorm_fwd_bwd(R, H, packed):
    from b200fm import ops
    g = torch.Generator().manual_seed(R)
    C = H * 64
    buf = (torch.randn(R, 3 * C if packed else C, generator=g) * 2 + 0.3).to(torch.bfloat16).cuda()
    x = buf[:, C:2 * C] if packed else buf
    gamma = (1 + 0.1 * torch.randn(64, generator=g)).cuda()
    beta = (0.1 * torch.randn(64, generator=g)).cuda()
    y, stats = ops.headnorm_fwd(x, H, gamma, beta, 1e-6)
    xr = x.float().reshape(R, H, 64).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (64,), gr, br, 1e-6)
    # bf16 output: half an ulp of the largest value
    torch.testing.assert_close(y.float().reshape(R, H, 64), yr.detach(), rtol=8e-3, atol=8e-3)
    torch.testing.assert_close(stats[..., 0], xr.detach().mean(-1), rtol=1e-5, atol=1e-5)
    dy = torch.randn(R, C, generator=g).to(torch.bfloat16).cuda()
    yr.backward(dy.float().reshape(R, H, 64))
    dgamma, dbeta = torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda")
    dx = ops.headnorm_bwd(dy, x, gamma, stats, H, dgamma, dbeta)
    torch.testing.assert_close(dx.float().reshape(R, H, 64), xr.grad, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(dgamma, gr.grad, rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(dbeta, br.grad, rtol=1e-3, atol=1e-2)


@pytest.fixture(scope="module")
def tiny_qknorm():
    from b200fm.compat import build_mod7_embeddings, create_model
    gold = H.load_golden("fourm_tiny_qknorm_golden.pt")
    specs = O.mod7_specs()
    sd = H.fill_fourm_buffers(H.golden_state_dict(gold), specs, 384)
    enc, dec, info = build_mod7_embeddings()
    model = create_model(gold["model"], encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info, qk_norm=True)
    model.load_state_dict(sd, strict=True)
    return gold, specs, sd, model.cuda()


def _to_cuda(batch):
    return {m: {k: v.cuda() for k, v in d.items()} for m, d in batch.items()}


def test_qknorm_state_dict_contract(tiny_qknorm):
    gold, specs, sd, model = tiny_qknorm
    msd = model.state_dict()
    assert list(msd.keys()) == list(gold["shapes"].keys())
    assert [k for k, _ in model.named_parameters(remove_duplicate=False)] == gold["param_names"]
    assert type(model.encoder[0].attn).__name__ == "NormAttention" and type(model.decoder[0].cross_attn).__name__ == "NormCrossAttention"


@pytest.mark.parametrize("tag", ["fp32_128", "bf16_128"])
def test_qknorm_forward_matches_reference(tiny_qknorm, tag):
    gold, specs, sd, model = tiny_qknorm
    c = gold["cases"][tag]
    batch = O.synthetic_mod7_batch(2, seed=c["batch_seed"], extra_valid=c["extra_valid"])
    random.seed(c["py_seed"])
    with torch.no_grad():
        loss, mod_loss = model(_to_cuda(batch), num_encoder_tokens=c["N"], num_decoder_tokens=c["M"], loss_type="mod")
    random.seed(c["py_seed"])
    with torch.no_grad():
        logits = model(_to_cuda(batch), num_encoder_tokens=c["N"], num_decoder_tokens=c["M"], return_logits=True)
    # same tolerances as tests/test_gpu_fourm.py (bf16 contractions vs the fp32 / bf16-autocast reference)
    assert abs(float(loss) - float(c["loss"])) <= 5e-3
    for m, v in c["mod_loss"].items():
        assert abs(float(mod_loss[m]) - float(v)) <= 1e-2, m
    for m, v in c["logits_slices"].items():
        torch.testing.assert_close(logits[m][:, :4, :32].float().cpu(), v, rtol=5e-2, atol=2e-2)


def test_qknorm_backward_matches_reference(tiny_qknorm):
    gold, specs, sd, model = tiny_qknorm
    c = gold["cases"]["fp32_128"]
    batch = O.synthetic_mod7_batch(2, seed=c["batch_seed"])
    model.zero_grad(set_to_none=True)
    random.seed(c["py_seed"])
    loss, _ = model(_to_cuda(batch), num_encoder_tokens=c["N"], num_decoder_tokens=c["M"])
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad for k, p in model.named_parameters()}
    for k, ref_norm in c["grads"]["norm"].items():
        g = grads[k]
        assert g is not None, k
        rel = abs(float(g.float().norm()) - ref_norm) / max(ref_norm, 1e-6)
        assert rel <= 3e-2, f"{k}: grad norm {float(g.norm())} vs {ref_norm}"
    for k, sl in c["grads"]["slices"].items():
        got = grads[k].flatten()[:64].float().cpu()
        scale = sl.abs().max().item() + 1e-12
        assert (got - sl).abs().max().item() <= 8e-2 * scale + 1e-7, k
