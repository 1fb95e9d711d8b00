"""FusedAdamW (b200fm_adamw / b200fm_adamw_multi) against torch.optim.AdamW as the reference configures it
(fourm/utils/optim_factory.py:239-240).  fp32 elementwise: tolerance 2e-6 relative (fma contraction / sqrt rounding)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(dev, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(768, 768), (2048, 768), (768,), (1,), (7, 13), (8192 + 5,), (3, 8192), (64, 3, 16, 16)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]


@pytest.mark.parametrize("wd", [0.0, 0.05])
def test_fused_adamw_matches_torch(wd):
    from b200fm import functional as BF
    from b200fm.optim import FusedAdamW
    dev = torch.device("cuda")
    ours, ref = _params(dev, 1), _params(dev, 1)
    shadow = BF.weight_bf16(ours[0])            # a bf16 mirror the optimizer has to keep fresh
    o1 = FusedAdamW([dict(params=ours[:4], weight_decay=wd), dict(params=ours[4:], weight_decay=0.0)], lr=1e-3, betas=(0.9, 0.95))
    o2 = torch.optim.AdamW([dict(params=ref[:4], weight_decay=wd), dict(params=ref[4:], weight_decay=0.0)], lr=1e-3, betas=(0.9, 0.95))
    g = torch.Generator().manual_seed(2)
    for step in range(4):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g).to(dev)
            if step == 2 and a.ndim == 1:
                a.grad = b.grad = None          # a parameter without a gradient this step is skipped by both
                continue
            a.grad, b.grad = gr.clone(), gr.clone()
        o1.step()
        o2.step()
    for a, b in zip(ours, ref):
        torch.testing.assert_close(a.data, b.data, rtol=2e-6, atol=2e-7)
    for a, b in zip(ours, ref):
        sa, sb = o1.state[a], o2.state[b]
        torch.testing.assert_close(sa["exp_avg"], sb["exp_avg"], rtol=2e-6, atol=1e-7)
        torch.testing.assert_close(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=2e-6, atol=1e-7)
    assert torch.equal(BF.weight_bf16(ours[0]), ours[0].data.to(torch.bfloat16))
    assert BF.weight_bf16(ours[0]).data_ptr() == shadow.data_ptr()


@pytest.mark.parametrize("capturable", [False, True])
def test_fused_adamw_tracks_gradient_norm(capturable):
    """track_grad_norm: the AdamW kernels accumulate sum(g^2) of the gradients they consume; grad_norm() equals the global L2 norm the
    training loop logs (utils/native_scaler.py:56-65), with the same parameter update as without tracking."""
    from b200fm.optim import FusedAdamW
    dev = torch.device("cuda")
    a, b = _params(dev, 3), _params(dev, 3)
    oa = FusedAdamW([dict(params=a[:4]), dict(params=a[4:], weight_decay=0.0)], lr=1e-3, betas=(0.9, 0.95), capturable=capturable)
    ob = FusedAdamW([dict(params=b[:4]), dict(params=b[4:], weight_decay=0.0)], lr=1e-3, betas=(0.9, 0.95), capturable=capturable)
    oa.track_grad_norm = True
    g = torch.Generator().manual_seed(4)
    for step in range(3):
        for x, y in zip(a, b):
            gr = torch.randn(x.shape, generator=g).to(dev) * (step + 1)
            x.grad, y.grad = gr.clone(), gr.clone()
        want = torch.linalg.vector_norm(torch.stack(torch._foreach_norm([x.grad for x in a])))
        if capturable:
            oa.prepare_step(); ob.prepare_step()
        oa.step(); ob.step()
        torch.testing.assert_close(oa.grad_norm(), want, rtol=1e-5, atol=0)
    for x, y in zip(a, b):
        assert torch.equal(x.data, y.data)


def test_fused_adamw_grad_scale():
    from b200fm.optim import FusedAdamW
    dev = torch.device("cuda")
    ours, ref = _params(dev, 3), _params(dev, 3)
    o1 = FusedAdamW(ours, lr=2e-3, betas=(0.9, 0.95), weight_decay=0.05)
    o2 = torch.optim.AdamW(ref, lr=2e-3, betas=(0.9, 0.95), weight_decay=0.05)
    g = torch.Generator().manual_seed(4)
    for a, b in zip(ours, ref):
        gr = torch.randn(a.shape, generator=g).to(dev)
        a.grad, b.grad = gr.clone(), gr * 0.25
    o1.step(grad_scale=0.25)
    o2.step()
    for a, b in zip(ours, ref):
        torch.testing.assert_close(a.data, b.data, rtol=2e-6, atol=2e-7)


def test_device_prefetcher_yields_batches_in_order():
    from b200fm.data import DevicePrefetcher
    host = [{"rgb": {"tensor": torch.full((4, 8), float(i)).pin_memory()}, "cap": {"tensor": torch.arange(5) + i}} for i in range(5)]
    seen = []
    for batch in DevicePrefetcher(host, "cuda", depth=2):
        assert batch["rgb"]["tensor"].is_cuda
        seen.append((float(batch["rgb"]["tensor"][0, 0]), int(batch["cap"]["tensor"][0])))
    assert seen == [(float(i), i) for i in range(5)]


def test_weight_caches_follow_fused_adamw():
    """FusedAdamW rewrites weights and their registered bf16 mirrors through raw pointers; every OTHER version-keyed cache (the
    K-padded fc2 shadows of odd SwiGLU widths, the conv-as-linear re-layouts of the ViT tokenizers, reshaped 1x1-conv views)
    must still notice the update: the loss after a step equals the loss computed with all caches dropped."""
    import random
    from functools import partial
    import torch.nn as nn
    from b200fm import functional as BF
    from b200fm.compat import build_mod7_embeddings
    from b200fm.optim import FusedAdamW
    from fourm.models.fm import FourM
    from fourm.models.fm_utils import LayerNorm
    from fourm.vq.models.vit_models import _ConvAsLinear
    from oracle import fourm_oracle as O
    import fourm.vq as vq
    torch.manual_seed(0)
    enc, dec, info = build_mod7_embeddings()
    model = FourM(enc, dec, info, dim=256, encoder_depth=1, decoder_depth=1, num_heads=4, qkv_bias=False, proj_bias=False, mlp_bias=False,
                  norm_layer=partial(LayerNorm, eps=1e-6, bias=False), act_layer=nn.SiLU, gated_mlp=True).cuda()      # H = 682: K-padded fc2
    batch = {m: {k: v.cuda() for k, v in d.items()} for m, d in O.synthetic_mod7_batch(2, seed=3).items()}
    opt = FusedAdamW(model.parameters(), lr=1e-2, betas=(0.9, 0.95), weight_decay=0.0)
    for _ in range(2):
        random.seed(0)
        loss, _ = model(batch, 128, 128)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    with torch.no_grad():
        random.seed(0)
        l1, _ = model(batch, 128, 128)
        BF.clear_weight_cache()
        random.seed(0)
        l2, _ = model(batch, 128, 128)
    assert float(l1) == float(l2), (float(l1), float(l2))

    vae = vq.VQVAE(enc_type="vit_s_enc", dec_type="vit_s_dec", image_size=64, patch_size=16, codebook_size=128, latent_dim=32,
                   post_mlp=True, sync_codebook=False, threshold_ema_dead_code=0.0).cuda().train()
    x = torch.randn(4, 3, 64, 64, device="cuda")
    opt = FusedAdamW(vae.parameters(), lr=1e-2, betas=(0.9, 0.95), weight_decay=0.0)
    n_before = len(BF._shadow)
    for _ in range(3):
        dec, code_loss = vae(x)
        (torch.nn.functional.mse_loss(dec.float(), x) + code_loss.sum()).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    grown = len(BF._shadow) - n_before
    vae.eval()
    with torch.no_grad():
        d1, _ = vae(x)
        BF.clear_weight_cache()
        _ConvAsLinear._cache.clear()
        d2, _ = vae(x)
    assert torch.equal(d1, d2)
    vae.train()
    dec, code_loss = vae(x)                                    # re-creates the bf16 shadows dropped by clear_weight_cache()
    n_steady = len(BF._shadow)
    dec, code_loss = vae(x)                                    # repeated forwards (training bf16 path, inference fp32-faithful path) must
    vae.eval()                                                 # not add cache entries, e.g. for the reshaped conv weights
    with torch.no_grad():
        vae(x)
    assert len(BF._shadow) == n_steady, (len(BF._shadow), n_steady, n_before, grown)


def test_training_loop_reduces_loss_and_tolerates_autocast():
    """End-to-end coherence of forward, backward, gradient clipping and the fused optimizer: over-fitting one fixed batch for 25 steps
    drives the 4M-Tiny loss down by more than 2 nats, also when the caller wraps the forward in torch.autocast(bf16) as
    run_training_4m.py does (:725); and torch.optim.AdamW (the reference's optimizer) drives the same model equally."""
    import random
    from b200fm.compat import build_mod7_embeddings, create_model
    from b200fm.optim import FusedAdamW, param_groups_like_reference
    from oracle import fourm_oracle as O
    batch = {m: {k: v.cuda() for k, v in d.items()} for m, d in O.synthetic_mod7_batch(4, seed=21).items()}

    def run(make_opt, autocast, steps=25):
        torch.manual_seed(0)
        enc, dec, info = build_mod7_embeddings()
        model = create_model("fm_tiny_6e_6d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info).cuda()
        opt = make_opt(model)
        losses = []
        for _ in range(steps):
            random.seed(0)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                loss, _ = model({m: dict(d) for m, d in batch.items()}, 128, 128)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 3.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
            losses.append(float(loss))
        return losses

    fused = run(lambda m: FusedAdamW(param_groups_like_reference(m, 0.05), lr=1e-3, betas=(0.9, 0.95)), autocast=False)
    assert fused[0] > 9.0 and fused[-1] < fused[0] - 2.0, (fused[0], fused[-1])
    amp = run(lambda m: FusedAdamW(param_groups_like_reference(m, 0.05), lr=1e-3, betas=(0.9, 0.95)), autocast=True)
    # same first step; afterwards the two runs are separate (chaotic) trajectories of the same quality
    assert abs(amp[0] - fused[0]) <= 1e-3 and abs(amp[-1] - fused[-1]) <= 0.6, (amp[0], fused[0], amp[-1], fused[-1])
    ref_opt = run(lambda m: torch.optim.AdamW(param_groups_like_reference(m, 0.05), lr=1e-3, betas=(0.9, 0.95)), autocast=False)
    assert abs(ref_opt[-1] - fused[-1]) <= 0.6, (ref_opt[-1], fused[-1])


def test_fused_model_ema_matches_reference_expression():
    """FusedModelEma (one multi-tensor launch) vs the reference's ModelEmaV2.update expression (fourm/utils/timm/model_ema.py:123-127)
    applied entry by entry: fp32 entries bit-identical, integer buffers through the same torch expression; `.module` stays usable."""
    import copy
    from b200fm.optim import FusedModelEma
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.BatchNorm1d(64), torch.nn.Linear(64, 8193)).cuda()
    ema = FusedModelEma(model, decay=0.999)
    ref = copy.deepcopy(model).eval()
    for step in range(3):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.1)
            model[1].running_mean.add_(0.5)
            model[1].num_batches_tracked.add_(1)
            for e, m in zip(ref.state_dict().values(), model.state_dict().values()):
                e.copy_(0.999 * e + (1. - 0.999) * m)
        ema.update(model)
    for (k, a), b in zip(ema.module.state_dict().items(), ref.state_dict().values()):
        assert torch.equal(a, b), k
    assert not ema.module.training and ema.module(torch.randn(4, 37, device="cuda")).shape == (4, 8193)
    ema.set(model)
    for a, b in zip(ema.module.state_dict().values(), model.state_dict().values()):
        assert torch.equal(a, b)


def test_fused_model_ema_matches_reference_golden():
    """FusedModelEma vs the UNMODIFIED reference's ModelEmaV2 (golden from tests/golden/make_golden_ema.py: 3 updates on CPU): every
    state_dict entry bit-identical after every update (fp32 mul / mul / add are IEEE on both sides; the integer buffer goes through
    the same torch expression)."""
    import os
    import sys
    from b200fm.optim import FusedModelEma
    from tests import helpers as H
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_ema as G                      # build_model / perturbations: the recipe the golden was made with
    gold = H.load_golden("ema_golden.pt")
    model = G.build_model().cuda()
    ema = FusedModelEma(model, decay=gold["decay"])
    for step in range(3):
        with torch.no_grad():
            for p, d in zip(model.parameters(), G.perturbations(model, step)):
                p.add_(d.cuda())
            model[1].running_mean.add_(0.5)
            model[1].num_batches_tracked.add_(1)
        ema.update(model)
        for k, v in ema.module.state_dict().items():
            assert torch.equal(v.cpu(), gold["states"][step][k]), (step, k)
