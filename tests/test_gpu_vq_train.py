"""VQ-VAE training side (a25): codebook EMA kernels and one VQVAE training step vs the golden outputs of the UNMODIFIED reference
(tests/golden/make_golden_vqtrain.py) and the oracle restatement (oracle/vq_oracle.py)."""
import pytest
import torch

from oracle import fourm_oracle as O
from oracle import vq_oracle as V
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_cosine_codebook_training_steps_match_reference():
    from fourm.vq.quantizers.quantize_lucid import CosineSimCodebook
    c = H.load_golden("vq_train_golden.pt")["cosine"]
    cb = CosineSimCodebook(dim=32, codebook_size=512, decay=0.9, threshold_ema_dead_code=0).cuda().train()
    cb.embed.copy_(c["embed0"])
    for z, st in zip(c["z"], c["steps"]):
        q, idx = cb(z.cuda())
        assert torch.equal(idx.cpu(), st["idx"])                                      # integer: exact
        torch.testing.assert_close(q.double().sum(-1).cpu(), st["quant_sum"], rtol=1e-6, atol=1e-6)
        # fp32 atomics sum the per-code latents in a different order than the reference's GEMM: 1e-5 relative
        torch.testing.assert_close(cb.embed.cpu(), st["embed"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(cb.cluster_size.cpu(), st["cluster_size"], rtol=1e-6, atol=1e-6)


def test_euclidean_codebook_training_steps_match_reference():
    from fourm.vq.quantizers.quantize_lucid import EuclideanCodebook
    gold = H.load_golden("vq_train_golden.pt")
    e = gold["euclid"]
    cb = EuclideanCodebook(dim=32, codebook_size=300, decay=0.8, threshold_ema_dead_code=0).cuda().train()
    cb.embed.copy_(e["embed0"]); cb.embed_avg.copy_(e["embed0"])
    for z, st in zip(gold["cosine"]["z"], e["steps"]):
        q, idx = cb(z.cuda())
        assert torch.equal(idx.cpu(), st["idx"])
        torch.testing.assert_close(cb.embed.cpu(), st["embed"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(cb.embed_avg.cpu(), st["embed_avg"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(cb.cluster_size.cpu(), st["cluster_size"], rtol=1e-6, atol=1e-6)


def test_ema_statistics_are_additive_and_full_size():
    """What sync_codebook relies on: the packed [bins | embed_sum] of two shards adds up to the statistics of the union (so ONE
    all-reduce of the packed buffer replaces the reference's two).  Also the cfg-5 size: 131072 latents, K = 16384, d = 32."""
    from b200fm import ops
    g = torch.Generator().manual_seed(3)
    n, K, d = 131072, 16384, 32
    z = torch.randn(n, d, generator=g).cuda()
    idx = torch.randint(0, K, (n,), generator=g).cuda()
    full = ops.vq_ema_stats(z, idx, K, True)
    a = ops.vq_ema_stats(z[: n // 3].contiguous(), idx[: n // 3].contiguous(), K, True)
    ops.vq_ema_stats(z[n // 3:].contiguous(), idx[n // 3:].contiguous(), K, True, stats=a)
    torch.testing.assert_close(a, full, rtol=1e-4, atol=1e-4)
    assert float(full[:K].sum()) == n
    zn = torch.nn.functional.normalize(z, dim=-1)
    ref = torch.zeros(K, d, device="cuda").index_add_(0, idx, zn)
    torch.testing.assert_close(full[K:].view(K, d), ref, rtol=1e-4, atol=1e-4)


def test_dead_code_expiry_reseeds_from_the_batch():
    from fourm.vq.quantizers.quantize_lucid import CosineSimCodebook
    torch.manual_seed(0)
    cb = CosineSimCodebook(dim=32, codebook_size=64, decay=0.5, threshold_ema_dead_code=0.4).cuda().train()
    z = torch.randn(1, 40, 32, device="cuda")                                           # at most 40 codes can be hit
    before = cb.embed.clone()
    _, idx = cb(z)
    used = torch.zeros(64, dtype=torch.bool, device="cuda")
    used[idx.flatten()] = True
    zn = torch.nn.functional.normalize(z[0], dim=-1)
    dead = ~used                                                                       # cluster_size stays 0 < 0.4 -> expired
    assert int(dead.sum()) >= 24
    sims = cb.embed[dead] @ zn.t()
    torch.testing.assert_close(sims.max(dim=1).values, torch.ones(int(dead.sum()), device="cuda"), rtol=1e-5, atol=1e-5)   # rows ARE batch latents
    assert not torch.equal(cb.embed[used], before[used])                               # live codes moved by the EMA, not replaced


def _vqvae():
    import fourm.vq as vq
    gold = H.load_golden("vq_train_golden.pt")["vqvae"]
    sd = {}
    for k, shape in gold["shapes"].items():
        if k.endswith("pos_emb"):
            sd[k] = V.sincos_2d_grid(shape[2], shape[3], shape[1])
        elif k.endswith("initted"):
            sd[k] = torch.ones(shape)
        elif k.endswith("cluster_size"):
            sd[k] = torch.zeros(shape)
        elif k.endswith("_codebook.embed"):
            sd[k] = torch.nn.functional.normalize(O.deterministic_tensor("quantize._codebook.embed", shape, 1.0), dim=-1)
        else:
            sd[k] = O.deterministic_tensor(k, shape, 0.05 if len(shape) > 1 else 0.02)
    model = vq.VQVAE(**gold["kw"])
    assert list(model.state_dict().keys()) == list(gold["shapes"].keys())
    assert [k for k, _ in model.named_parameters()] == gold["param_names"]
    model.load_state_dict(sd, strict=True)
    x = torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(7))
    return gold, sd, model.cuda().train(), x


def test_vqvae_training_step_matches_reference():
    gold, sd, model, x = _vqvae()
    xc = x.cuda()
    dec, code_loss = model(xc)
    rec = torch.nn.functional.mse_loss(dec.float(), xc)
    (rec + code_loss.sum()).backward()
    torch.cuda.synchronize()
    # bf16 contractions against the fp32 reference: losses within 2 %, reconstruction slice within bf16 noise of its scale
    assert abs(float(rec) - float(gold["rec_loss"])) <= 2e-2 * float(gold["rec_loss"])
    assert abs(float(code_loss) - float(gold["code_loss"])) <= 2e-2 * float(gold["code_loss"])
    ref = gold["dec_slice"]
    assert (dec[:, :, :8, :8].float().cpu() - ref).abs().max() <= 5e-2 * ref.abs().max()
    # codebook buffers after the step: codes whose assignment is unchanged by the bf16 encoder follow the reference update
    with torch.no_grad():
        _, _, tokens = model.eval().encode(xc)
    same = (tokens.cpu() == gold["tokens_after"]).float().mean()
    assert same >= 0.85, same
    close = ((model.quantize._codebook.embed.cpu() - gold["embed_after"]).abs().max(dim=1).values < 2e-2).float().mean()
    assert close >= 0.9, close
    grads = {k: p.grad for k, p in model.named_parameters()}
    for k, ref_norm in gold["grad_norm"].items():
        g = grads[k]
        assert g is not None, k
        rel = abs(float(g.float().norm()) - ref_norm) / max(ref_norm, 1e-6)
        assert rel <= 5e-2, f"{k}: grad norm {float(g.float().norm())} vs {ref_norm}"
    for k, sl in gold["grad_slices"].items():
        got = grads[k].flatten()[:64].float().cpu()
        assert (got - sl).abs().max().item() <= 1e-1 * sl.abs().max().item() + 1e-7, k
