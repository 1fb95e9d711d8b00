"""FourM on the B200 kernels vs the golden outputs of the UNMODIFIED reference (4M-Tiny mod7, BASELINE configs[0] shape)
and vs the oracle.  Integer/bool artefacts must be exact; floating point is compared with the tolerances stated inline:
the product computes its contractions in bf16 (fp32 accumulate) like the reference under autocast(bf16), so the yardstick
is the reference's own fp32-vs-bf16 gap."""
import random

import pytest
import torch

from oracle import fourm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny():
    from b200fm.compat import build_mod7_embeddings, create_model
    gold = H.load_golden("fourm_tiny_golden.pt")
    specs = O.mod7_specs()
    sd = H.fill_fourm_buffers(H.golden_state_dict(gold), specs, 384)
    enc, dec, info = build_mod7_embeddings()
    model = create_model(gold["model"], encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    return gold, specs, sd, model.cuda()


def _to_cuda(batch):
    return {m: {k: v.cuda() for k, v in d.items()} for m, d in batch.items()}


def test_state_dict_contract(tiny):
    gold, specs, sd, model = tiny
    msd = model.state_dict()
    assert list(msd.keys()) == list(gold["shapes"].keys())
    assert all(tuple(v.shape) == gold["shapes"][k] for k, v in msd.items())
    assert [k for k, _ in model.named_parameters(remove_duplicate=False)] == gold["param_names"]
    # sharing (fm.py:176-180, decoder_embeddings.py:89-91)
    assert model.decoder_embeddings["caption"].mod_emb is model.encoder_embeddings["caption"].mod_emb
    assert model.decoder_embeddings["tok_rgb@224"].to_logits.weight is model.decoder_embeddings["tok_rgb@224"].token_emb.weight


@pytest.mark.parametrize("tag", ["fp32_128", "fp32_trunc", "fp32_pad"])
def test_selection_is_bit_exact(tiny, tag):
    """PLAN + EMBED kernels vs the reference's cat/argsort/gather/mask path: masks, modality ids, targets, decoder attention
    mask are integer-exact; gathered fp32 rows are bit-exact (same operation order)."""
    from b200fm import ops
    gold, specs, sd, model = tiny
    c = gold["cases"][tag]
    batch = _to_cuda(O.synthetic_mod7_batch(2, seed=c["batch_seed"], extra_valid=c["extra_valid"]))
    enc_mods = [m for m in batch if m in model.encoder_embeddings]
    with torch.no_grad():
        x0, emb, ep = model._embed_side(batch, False, c["N"], enc_mods)
        y0, _, dp = model._embed_side(batch, True, c["M"], c["decoder_order"])
        amask = ops.decoder_attention_mask(dp.dam, dp.mod_raw, False, True)
    assert torch.equal(ep.pad_mask[:, None].cpu(), c["enc_mask"])
    assert torch.equal(ep.mod_mask.cpu(), c["enc_mod"])
    assert torch.equal(dp.pad_mask[:, None].cpu(), c["dec_mask"])
    assert torch.equal(dp.mod_mask.cpu(), c["dec_mod"])
    assert torch.equal(dp.target_ids.cpu(), c["target_ids"].long())
    assert torch.equal(amask.cpu(), c["dec_attn_mask"])
    # decoder rows never touch the bf16 patch projection -> bit-exact against the fp32 reference
    assert torch.equal(y0.double().sum(-1).cpu(), c["dec_y0_sum"])
    # encoder rows: exact except the rgb rows, whose patch projection is a bf16 GEMM (autocast contract)
    rgb = ep.mod_mask.cpu() == specs["rgb@224"]["id"]
    got, ref = x0.double().sum(-1).cpu(), c["enc_x0_sum"]
    assert torch.equal(got[~rgb], ref[~rgb])
    torch.testing.assert_close(got[rgb], ref[rgb], rtol=0, atol=0.5)


@pytest.mark.parametrize("tag", ["fp32_128", "bf16_128", "fp32_trunc", "fp32_pad"])
def test_forward_loss_and_logits(tiny, tag):
    gold, specs, sd, model = tiny
    c = gold["cases"][tag]
    batch = O.synthetic_mod7_batch(2, seed=c["batch_seed"], extra_valid=c["extra_valid"])
    random.seed(c["py_seed"])
    with torch.no_grad():
        loss, mod_loss = model(_to_cuda(batch), num_encoder_tokens=c["N"], num_decoder_tokens=c["M"], loss_type="mod")
    random.seed(c["py_seed"])
    with torch.no_grad():
        tl, _ = model(_to_cuda(batch), num_encoder_tokens=c["N"], num_decoder_tokens=c["M"], loss_type="token")
    random.seed(c["py_seed"])
    with torch.no_grad():
        logits = model(_to_cuda(batch), num_encoder_tokens=c["N"], num_decoder_tokens=c["M"], return_logits=True)
    # tolerance: 5e-3 absolute on ~9.5 nats (the reference's own fp32-vs-bf16 autocast gap on this case is 4e-5 .. 3e-4 per modality)
    assert loss.shape == c["loss"].shape and loss.dtype == torch.float32     # [1] when a modality is empty (zeros(1) term), like the reference
    assert abs(float(loss) - float(c["loss"])) <= 5e-3
    assert abs(float(tl) - float(c["token_loss"])) <= 5e-3
    for m, v in c["mod_loss"].items():
        assert abs(float(mod_loss[m]) - float(v)) <= 1e-2, m
        if float(v) == 0.0:                      # empty modality -> zeros(1) (fm.py:593-595)
            assert float(mod_loss[m]) == 0.0
    for m, v in c["logits_slices"].items():
        assert logits[m].shape[:2] == (2, c["M"])
        torch.testing.assert_close(logits[m][:, :4, :32].float().cpu(), v, rtol=5e-2, atol=2e-2)


def test_backward_matches_reference(tiny):
    gold, specs, sd, model = tiny
    c = gold["cases"]["fp32_128"]
    batch = O.synthetic_mod7_batch(2, seed=c["batch_seed"])
    model.zero_grad(set_to_none=True)
    random.seed(c["py_seed"])
    loss, _ = model(_to_cuda(batch), num_encoder_tokens=c["N"], num_decoder_tokens=c["M"])
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad for k, p in model.named_parameters()}
    worst = 0.0
    for k, ref_norm in c["grads"]["norm"].items():
        g = grads[k]
        assert g is not None, k
        rel = abs(float(g.float().norm()) - ref_norm) / max(ref_norm, 1e-6)
        worst = max(worst, rel)
        # bf16 contractions in fwd and bwd: per-tensor gradient norms within 3 % of the fp32 reference
        assert rel <= 3e-2, f"{k}: grad norm {float(g.norm())} vs {ref_norm}"
    for k, sl in c["grads"]["slices"].items():
        got = grads[k].flatten()[:64].float().cpu()
        scale = sl.abs().max().item() + 1e-12
        assert (got - sl).abs().max().item() <= 8e-2 * scale + 1e-7, k


def test_loss_invariant_to_decoder_shuffle_when_no_truncation(tiny):
    """identical loss across Python-random modality shuffles when #valid <= budget."""
    gold, specs, sd, model = tiny
    batch = _to_cuda(O.synthetic_mod7_batch(2, seed=5))
    vals = []
    for s in (0, 1, 2):
        random.seed(s)
        with torch.no_grad():
            loss, _ = model({m: dict(d) for m, d in batch.items()}, 128, 128)
        vals.append(float(loss))
    assert max(vals) - min(vals) <= 2e-3


def test_odd_swiglu_width_matches_oracle():
    """SwiGLU hidden widths that are not multiples of 8 (4M-L: 2730, 4M-XL: 5461; here dim 256 -> 682) run on zero-padded
    weight shadows; forward loss and gradients must match the oracle on the same weights."""
    from functools import partial
    import torch.nn as nn
    from b200fm.compat import build_mod7_embeddings
    from fourm.models.fm import FourM
    from fourm.models.fm_utils import LayerNorm
    torch.manual_seed(0)
    enc, dec, info = build_mod7_embeddings()
    model = FourM(enc, dec, info, dim=256, encoder_depth=2, decoder_depth=2, num_heads=4, qkv_bias=False, proj_bias=False, mlp_bias=False,
                  norm_layer=partial(LayerNorm, eps=1e-6, bias=False), act_layer=nn.SiLU, gated_mlp=True).cuda()
    assert model.encoder[0].mlp.fc1.weight.shape[0] == 682
    batch = O.synthetic_mod7_batch(2, seed=11)
    random.seed(1)
    loss, _ = model(_to_cuda(batch), 128, 128)
    loss.backward()
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    keys = ["encoder.1.mlp.fc2.weight", "decoder.0.mlp.fc1.weight", "decoder.1.mlp.fc3.weight"]
    for k in keys:
        sd[k] = sd[k].clone().requires_grad_(True)
    random.seed(1)
    order = random.sample([m for m in batch if m in model.decoder_embeddings], len(model.decoder_embeddings))
    ref, _ = O.fourm_forward(sd, O.model_cfg(256, 4, 2, 2), O.mod7_specs(), batch, 128, 128, order)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 5e-3
    params = dict(model.named_parameters())
    for k in keys:
        g, r = params[k].grad.float().cpu(), sd[k].grad
        assert g.shape == r.shape
        assert abs(float(g.norm()) - float(r.norm())) <= 3e-2 * float(r.norm()), k


def test_large_config_shapes_cfg3():
    """BASELINE configs[2] shape: 4M-L (dim 1024, 16 heads, SwiGLU 2730) with 256 + 256 tokens: one fwd+bwd step runs and is finite."""
    from b200fm.compat import build_mod7_embeddings, create_model
    from b200fm.synthetic import budgets_for, mod7_batch
    torch.manual_seed(0)
    enc, dec, info = build_mod7_embeddings()
    model = create_model("fm_large_24e_24d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info).cuda()
    a, b, c, d = budgets_for(256)
    batch = _to_cuda(mod7_batch(4, a, b, c, d, seed=3))
    random.seed(0)
    loss, mod_loss = model(batch, 256, 256)
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss) and 8.0 < float(loss) < 12.0
    g = model.encoder[3].mlp.fc2.weight.grad
    assert g is not None and g.shape == (1024, 2730) and torch.isfinite(g).all() and float(g.abs().sum()) > 0


@pytest.mark.parametrize("N,M,causal", [(600, 196, False), (1900, 64, True), (257, 290, True)])
def test_generation_shaped_stack_calls(tiny, N, M, causal):
    """a19: the calls fourm/models/generate.py makes on the model (forward_encoder on a data-dependent N > 256 context in eval
    mode, generate.py:407-445; forward_decoder with sa_mask all-False (MaskGIT / ROAR) or causal (AR loop, :886-913) and
    cross-attention over all N context keys) against the oracle stacks on the same weights, fp32."""
    gold, specs, sd, model = tiny
    cfg = O.PRESETS[gold["model"]]
    g = torch.Generator().manual_seed(N + M)
    B, D = 1, 384
    x = torch.randn(B, N, D, generator=g)
    y = torch.randn(B, M, D, generator=g)
    enc_mask = torch.zeros(B, 1, N, dtype=torch.bool)
    enc_mask[:, :, N - 37:] = True                                        # padded tail of the context
    sa_mask = torch.triu(torch.ones(M, M, dtype=torch.bool), diagonal=1)[None] if causal else torch.zeros(B, M, M, dtype=torch.bool)
    model.eval()
    with torch.no_grad():
        ctx = model.forward_encoder(x.cuda(), enc_mask.cuda())
        out = model.forward_decoder(y.cuda(), ctx, enc_mask.cuda(), sa_mask.cuda())
    model.train()
    ctx_ref = O.run_encoder(x, sd, cfg, enc_mask)
    out_ref = O.run_decoder(y, ctx_ref, sd, cfg, enc_mask, sa_mask)
    # 6+6 layers of bf16 contractions against an fp32 oracle on unit-variance activations (LayerNorm'd outputs, O(1) entries)
    torch.testing.assert_close(ctx.float().cpu(), ctx_ref, rtol=5e-2, atol=5e-2)
    torch.testing.assert_close(out.float().cpu(), out_ref, rtol=5e-2, atol=5e-2)


@pytest.mark.parametrize("tag", ["plain", "bottleneck"])
def test_sequence_feature_embedding_module(tag):
    """a6: SequenceEmbEncoderEmbedding (T5-XXL features of 4M-21) -- the emb_proj GEMM + the selection / embedding kernels in
    identity mode (segment kind SEQ_EMB) against the golden outputs and gradients of the unmodified reference module."""
    from fourm.models.encoder_embeddings import SequenceEmbEncoderEmbedding
    c = H.load_golden("seqemb_golden.pt")["cases"][tag]
    g = torch.Generator().manual_seed(41)                        # == tests/golden/make_golden_seqemb.inputs()
    feats = torch.randn(3, 77, 4096, generator=g)
    mask = torch.rand(3, 77, generator=g) < 0.35
    mask[0] = False
    mask[1, 5:] = True
    wx = torch.randn(3, 77, 384, generator=g)
    we = torch.randn(3, 77, 384, generator=g)
    m = SequenceEmbEncoderEmbedding(max_length=77, dim_tokens=384, orig_emb_dim=4096, **c["kw"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == c["shapes"]
    sd = {k: (O.sincos_1d(512, 384) if k == "pos_emb" else O.deterministic_tensor("seqemb." + k, shape, 0.02)) for k, shape in c["shapes"].items()}
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    d = m(dict(tensor=feats.cuda(), input_mask=mask.cuda()))
    (d["x"].float() * wx.cuda()).sum().add((d["emb"] * we.cuda()).sum()).backward()
    # positional / modality part: fp32 gather, exact; projected features: bf16 GEMM over K = 4096 (values ~1.3, ulp 2^-8)
    assert torch.equal(d["emb"].detach().cpu()[:, :6, :48], c["emb_slice"])
    assert torch.equal(d["emb"].detach().double().sum(-1).cpu(), c["emb_sum"])
    torch.testing.assert_close(d["x"].detach().float().cpu()[:, :6, :48], c["x_slice"], rtol=2e-2, atol=2e-2)
    for k, p in m.named_parameters():
        ref = c["grad_norm"][k]
        assert abs(float(p.grad.float().norm()) - ref) <= 3e-2 * max(ref, 1e-6), k
        sl = c["grad_slices"][k]
        assert (p.grad.flatten()[:64].float().cpu() - sl).abs().max() <= 8e-2 * sl.abs().max() + 1e-7, k


def test_embedding_modules_forward_contract(tiny):
    """The reference's per-module API that fourm/models/generate.py drives (`encoder_embeddings[mod](d)` -> d['x'], d['emb'];
    `decoder_embeddings[mod].forward_embed(d)` -> + d['ids']): every position of the modality, no selection.  Against the
    oracle's restatement of encoder_embeddings.py:87-121, 184-211, 280-309 and decoder_embeddings.py:98-139, 226-255.
    Token / positional / modality rows are fp32 gathers (exact); the rgb patch projection is a bf16 GEMM."""
    gold, specs, sd, model = tiny
    batch = O.synthetic_mod7_batch(2, seed=11)
    cb = _to_cuda(batch)
    with torch.no_grad():
        for mod, d in batch.items():
            s = specs[mod]
            if mod in model.encoder_embeddings:
                out = model.encoder_embeddings[mod](dict(cb[mod]))
                pfx = f"encoder_embeddings.{mod}."
                if s["kind"] == "seq":
                    x, emb = O.embed_sequence(d["tensor"], d["input_mask"], sd[pfx + "token_emb.weight"], sd[pfx + "pos_emb"], sd[pfx + "mod_emb"], None)
                elif s["kind"] == "tok_img":
                    x, emb = O.embed_image_tokens(d["tensor"], sd[pfx + "token_emb.weight"], sd[pfx + "pos_emb"], sd[pfx + "mod_emb"])
                else:
                    x, emb = O.embed_image_pixels(d["tensor"], sd[pfx + "proj.weight"], sd[pfx + "pos_emb"], sd[pfx + "mod_emb"], s["patch_size"])
                assert out["x"].shape == x.shape and out["emb"].shape == emb.shape
                assert torch.equal(out["emb"].cpu(), emb.expand_as(out["emb"].cpu())), mod
                if s["kind"] == "img":
                    torch.testing.assert_close(out["x"].float().cpu(), x, rtol=2e-2, atol=2e-2)
                else:
                    assert torch.equal(out["x"].float().cpu(), x), mod
            if mod in model.decoder_embeddings:
                out = model.decoder_embeddings[mod].forward_embed(dict(cb[mod]))
                pfx = f"decoder_embeddings.{mod}."
                if s["kind"] == "seq":
                    x, emb = O.embed_sequence(d["tensor"], d["target_mask"], sd[pfx + "token_emb.weight"], sd[pfx + "pos_emb"], sd[pfx + "mod_emb"],
                                              s["max_length"])
                else:
                    x, emb = O.embed_image_tokens(d["tensor"], sd[pfx + "token_emb.weight"], sd[pfx + "pos_emb"], sd[pfx + "mod_emb"])
                assert torch.equal(out["ids"].cpu(), d["tensor"].reshape(d["tensor"].shape[0], -1))     # 'b h w -> b (h w)' for image tokens
                assert torch.equal(out["emb"].cpu(), emb.expand_as(out["emb"].cpu())), mod
                assert torch.equal(out["x"].float().cpu(), x), mod


@pytest.mark.parametrize("tag", ["fp32_128", "fp32_trunc", "fp32_pad"])
def test_reference_shaped_selection_helpers(tiny, tag):
    """The reference-shaped entry points that generate.py and downstream code call on the model -- per-module forward,
    `forward_mask_encoder`, `forward_mask_decoder` (fm.py:338-438), `forward_logits` -- reproduce the golden artefacts that
    make_golden.py recorded from exactly these calls on the unmodified reference."""
    gold, specs, sd, model = tiny
    c = gold["cases"][tag]
    b = _to_cuda(O.synthetic_mod7_batch(2, seed=c["batch_seed"], extra_valid=c["extra_valid"]))
    random.seed(c["py_seed"])
    with torch.no_grad():
        enc_d = {m: model.encoder_embeddings[m](dict(d)) for m, d in b.items() if m in model.encoder_embeddings}
        et, ee, em, emod = model.forward_mask_encoder(enc_d, c["N"])
        dec_d = {m: model.decoder_embeddings[m].forward_embed(dict(d)) for m, d in b.items() if m in model.decoder_embeddings}
        dt, de, dm, tgt, damask, dmod = model.forward_mask_decoder(dec_d, c["M"])
        logits = model.forward_logits(torch.zeros(2, c["M"], 384, device="cuda"), dec_d, dmod)
    assert torch.equal(em.cpu(), c["enc_mask"]) and torch.equal(emod.cpu(), c["enc_mod"])
    assert torch.equal(dm.cpu(), c["dec_mask"]) and torch.equal(dmod.cpu(), c["dec_mod"])
    assert torch.equal(tgt.cpu().long(), c["target_ids"].long())
    assert torch.equal(damask.cpu(), c["dec_attn_mask"])
    assert torch.equal((dt + de).double().sum(-1).cpu(), c["dec_y0_sum"])
    rgb = emod.cpu() == specs["rgb@224"]["id"]
    got, ref = (et.float() + ee).double().sum(-1).cpu(), c["enc_x0_sum"]
    assert torch.equal(got[~rgb], ref[~rgb])
    torch.testing.assert_close(got[rgb], ref[rgb], rtol=0, atol=0.5)            # bf16 patch projection (autocast contract)
    for m, lg in logits.items():
        assert lg.shape == (int((dmod == specs[m]["id"]).sum()), specs[m]["vocab"])


@pytest.mark.parametrize("kw", [dict(num_register_tokens=4), dict(decoder_causal_mask=True), dict(decoder_sep_mask=False),
                                dict(decoder_causal_mask=True, decoder_sep_mask=False, num_register_tokens=2)])
def test_constructor_variants_match_oracle(kw):
    """FourM options outside the shipped presets (fm.py:96-104): register tokens prepended to the encoder sequence
    (fm.py:372-383), causal decoder mask and no modality separation (fm.py:440-475) -- forward loss and a few gradients against
    the oracle on the same weights."""
    from functools import partial
    import torch.nn as nn
    from b200fm.compat import build_mod7_embeddings
    from fourm.models.fm import FourM
    from fourm.models.fm_utils import LayerNorm
    torch.manual_seed(0)
    enc, dec, info = build_mod7_embeddings()
    model = FourM(enc, dec, info, dim=256, encoder_depth=2, decoder_depth=2, num_heads=4, qkv_bias=False, proj_bias=False, mlp_bias=False,
                  norm_layer=partial(LayerNorm, eps=1e-6, bias=False), act_layer=nn.SiLU, gated_mlp=True, **kw).cuda()
    if kw.get("num_register_tokens"):
        with torch.no_grad():
            model.register_tokens.normal_(std=0.5)
    batch = O.synthetic_mod7_batch(2, seed=13)
    random.seed(2)
    loss, _ = model(_to_cuda(batch), 128, 128)
    loss.backward()
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    keys = ["encoder.0.attn.qkv.weight", "decoder.1.self_attn.proj.weight", "decoder.0.cross_attn.kv.weight"] + \
        (["register_tokens"] if kw.get("num_register_tokens") else [])
    for k in keys:
        sd[k] = sd[k].clone().requires_grad_(True)
    random.seed(2)
    order = random.sample([m for m in batch if m in model.decoder_embeddings], len(model.decoder_embeddings))
    cfg = O.model_cfg(256, 4, 2, 2, causal=kw.get("decoder_causal_mask", False), sep=kw.get("decoder_sep_mask", True),
                      num_register_tokens=kw.get("num_register_tokens", 0))
    ref, _ = O.fourm_forward(sd, cfg, O.mod7_specs(), batch, 128, 128, order)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 5e-3, (float(loss), float(ref))
    params = dict(model.named_parameters())
    for k in keys:
        g, r = params[k].grad.float().cpu(), sd[k].grad
        assert g.shape == r.shape
        assert abs(float(g.norm()) - float(r.norm())) <= 3e-2 * float(r.norm()) + 1e-6, k
