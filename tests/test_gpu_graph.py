"""CUDA-graph train step (b200fm.graph.GraphedTrainStep) and its building blocks vs the eager path on the same weights / batches:
device-side row counts in the masked-token head (incl. an EMPTY modality and ragged counts), the decoder shuffle as device data
(token selection must stay bit-exact with the reference's goldens), device-side AdamW scalars, and whole-step replays."""
import random

import pytest
import torch

from oracle import fourm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _to_cuda(batch):
    return {m: {k: v.cuda() for k, v in d.items()} for m, d in batch.items()}


def _tiny(seed=0):
    from b200fm.compat import build_mod7_embeddings, create_model
    torch.manual_seed(seed)
    enc, dec, info = build_mod7_embeddings()
    return create_model("fm_tiny_6e_6d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info).cuda()


def test_dyn_gemm_matches_static_shapes():
    """b200fm_gemm_bf16_dyn: NT / NN with a device-side row count and TN with a device-side contraction length give the results of
    the plain launch on the truncated problem (rows beyond the count untouched), including count 0 and split-K wgrad shapes."""
    from b200fm import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    R, D, V = 1024, 256, 1000
    for n in (0, 1, 130, 517, 1024):
        nd = torch.tensor([n], dtype=torch.int32, device="cuda")
        a = torch.randn(R, D, device="cuda", generator=g).bfloat16()
        a[n:] = 0                                                     # the gather kernel zero-fills behind the count
        w = torch.randn(V, D, device="cuda", generator=g).bfloat16()
        out = torch.full((R, V), 7.0, device="cuda")
        ops.gemm(a, w, epilogue=ops.EPI_F32, out=out, dyn=nd)
        ref = a[:n].float() @ w.float().t()
        torch.testing.assert_close(out[:n], ref, rtol=2e-2, atol=2e-1)
        assert bool((out[n:] == 7.0).all())
        dl = torch.randn(R, V, device="cuda", generator=g).bfloat16()
        dl[n:] = 0
        dh = torch.full((R, D), 3.0, device="cuda").bfloat16()
        ops.gemm(dl, w, layout=ops.LAYOUT_NN, epilogue=ops.EPI_BF16, out=dh, dyn=nd)
        torch.testing.assert_close(dh[:n].float(), dl[:n].float() @ w.float(), rtol=3e-2, atol=1.0)
        assert bool((dh[n:] == 3.0).all())
        dw = torch.full((V, D), 5.0, device="cuda")
        garbage = torch.full((R, D), float("nan"), device="cuda").bfloat16()
        h = a.clone()
        k64 = (n + 63) // 64 * 64
        h[k64:] = garbage[k64:]                                       # rows beyond the last contracted block may hold anything
        ops.gemm(dl, h, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32, out=dw, dyn=nd)
        torch.testing.assert_close(dw, dl[:n].float().t() @ a[:n].float(), rtol=3e-2, atol=1.0)
    # long contraction, few output tiles -> the split-K plan (atomics into a pre-zeroed output) with a short device-side K
    K, Mo, No = 16384, 256, 512
    x = torch.randn(K, Mo, device="cuda", generator=g).bfloat16()
    y = torch.randn(K, No, device="cuda", generator=g).bfloat16()
    for n in (0, 700, 16384):
        nd = torch.tensor([n], dtype=torch.int32, device="cuda")
        xx, yy = x.clone(), y.clone()
        xx[n:(n + 63) // 64 * 64] = 0
        out = ops.gemm(xx, yy, layout=ops.LAYOUT_TN, epilogue=ops.EPI_F32, dyn=nd)
        torch.testing.assert_close(out, xx[:n].float().t() @ yy[:n].float(), rtol=3e-2, atol=2.0)


@pytest.mark.parametrize("drop", [None, "tok_semseg@224"])
def test_static_head_matches_dynamic_head(drop):
    """FourM.static_head (device-side counts, no host sync) vs the default head: same loss, same per-modality losses, same
    gradients.  `drop`: that modality has NO target rows (empty -> loss 0, like the reference's zeros(1) term)."""
    model = _tiny()
    batch = O.synthetic_mod7_batch(2, seed=21)
    if drop is not None:
        batch[drop]["target_mask"][:] = True
        batch[drop]["decoder_attention_mask"][:] = 0
    outs = []
    for static in (False, True):
        model.static_head = static
        model.zero_grad(set_to_none=True)
        random.seed(3)
        loss, mod_loss = model(_to_cuda(batch), 128, 128 if drop is None else 100)
        loss.backward()
        torch.cuda.synchronize()
        outs.append((float(loss), {k: float(v) for k, v in mod_loss.items()}, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}))
    (l0, m0, g0), (l1, m1, g1) = outs
    assert abs(l0 - l1) <= 1e-5 and all(abs(m0[k] - m1[k]) <= 1e-5 for k in m0)
    if drop is not None:
        assert m1[drop] == 0.0
    # the dynamic head skips an empty modality (its to_logits weight gets no gradient, like the reference's `if n == 0` branch); the
    # static head launches with count 0 and produces an all-zero gradient for it
    assert set(g0) <= set(g1)
    for k in g1:
        if k in g0:
            torch.testing.assert_close(g1[k], g0[k], rtol=1e-3, atol=1e-6, msg=k)
        else:
            assert float(g1[k].abs().max()) == 0.0, k
    model.static_head = False


def test_device_side_decoder_order_is_bit_exact():
    """The decoder shuffle as device data (b200fm_select_plan_ordered) selects exactly the reference's tokens (golden case with
    truncation, where the order matters)."""
    from b200fm import ops
    from b200fm.compat import build_mod7_embeddings, create_model
    gold = H.load_golden("fourm_tiny_golden.pt")
    c = gold["cases"]["fp32_trunc"]
    enc, dec, info = build_mod7_embeddings()
    model = create_model(gold["model"], encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info)
    model.load_state_dict(H.fill_fourm_buffers(H.golden_state_dict(gold), O.mod7_specs(), 384))
    model = model.cuda()
    batch = _to_cuda(O.synthetic_mod7_batch(2, seed=c["batch_seed"], extra_valid=c["extra_valid"]))
    dec_mods = [m for m in batch if m in model.decoder_embeddings]
    perm = torch.tensor([dec_mods.index(m) for m in c["decoder_order"]], dtype=torch.int32, device="cuda")
    with torch.no_grad():
        y0, _, dp = model._embed_side(batch, True, c["M"], dec_mods, order_dev=perm)
        amask = ops.decoder_attention_mask(dp.dam, dp.mod_raw, False, True)
    assert torch.equal(dp.pad_mask[:, None].cpu(), c["dec_mask"])
    assert torch.equal(dp.mod_mask.cpu(), c["dec_mod"])
    assert torch.equal(dp.target_ids.cpu(), c["target_ids"].long())
    assert torch.equal(amask.cpu(), c["dec_attn_mask"])
    assert torch.equal(y0.double().sum(-1).cpu(), c["dec_y0_sum"])


def test_graphed_train_step_matches_eager_steps():
    """6 optimizer steps on identical weights, batches and Python-random state, three ways:
         eager        Python-issued launches, host-side head counts (the default path), FusedAdamW
         graph-eager  GraphedTrainStep that never captures (same code path as the capture: static head, device-side order / AdamW scalars)
         graph        GraphedTrainStep: 2 eager calls, one capture, 4 replays
       graph vs graph-eager differ only by fp32 atomics order (split-K, embedding scatter); eager differs by the head variant as well.
       Element-wise parameter comparison after Adam steps is meaningless for near-zero gradients (a sign flip moves a weight by 2 lr), so
       the parameters are compared through the size of the total update."""
    from b200fm.graph import GraphedTrainStep
    from b200fm.optim import FusedAdamW, param_groups_like_reference
    batches = [O.synthetic_mod7_batch(2, seed=40 + i) for i in range(3)]
    results = {}
    for mode in ("eager", "graph-eager", "graph"):
        model = _tiny(seed=1)
        init = {k: v.detach().clone() for k, v in model.state_dict().items()}
        opt = FusedAdamW(param_groups_like_reference(model, 0.05), lr=1e-3, betas=(0.9, 0.95), capturable=(mode != "eager"))
        gstep = None if mode == "eager" else GraphedTrainStep(model, opt, 128, 128, eager_steps=(2 if mode == "graph" else 1000))
        random.seed(11)
        losses = []
        for it in range(6):
            for g in opt.param_groups:
                g["lr"] = 1e-3 * (1.0 - 0.1 * it)                      # a schedule: the graph must pick up the new value every replay
            b = _to_cuda(batches[it % 3])
            if gstep is not None:
                loss, mod_loss, gnorm = gstep(b)
            else:
                loss, mod_loss = model(b, num_encoder_tokens=128, num_decoder_tokens=128)
                loss.backward()
                opt.step()
                grads = [p.grad for p in model.parameters() if p.grad is not None]
                gnorm = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(grads)))
                opt.zero_grad(set_to_none=True)
            losses.append((float(loss), float(gnorm)))
        torch.cuda.synchronize()
        if mode == "graph":
            assert gstep.graph is not None and gstep.replays == 4 and gstep.kernel_calls_per_step > 100
        if mode == "graph-eager":
            assert gstep.graph is None
        upd = torch.cat([(v.float() - init[k].float()).flatten() for k, v in model.state_dict().items() if v.is_floating_point()])
        results[mode] = (losses, upd)
    for mode, (losses, _) in results.items():
        print(mode, [round(l, 5) for l, _ in losses], [round(n, 4) for _, n in losses])
    le, lge, lg = results["eager"][0], results["graph-eager"][0], results["graph"][0]
    for it in range(6):
        assert abs(lg[it][0] - lge[it][0]) <= 5e-4, (it, lg, lge)                        # replay == the code it captured
        assert abs(lg[it][1] - lge[it][1]) <= 5e-3 * lge[it][1]
        assert abs(lge[it][0] - le[it][0]) <= (5e-4 if it < 3 else 3e-3), (it, lge, le)  # == the default path
    assert lg[-1][0] < lg[0][0] - 0.3
    ue, uge, ug = results["eager"][1], results["graph-eager"][1], results["graph"][1]
    assert float((ug - uge).norm() / uge.norm()) <= 0.2 and float((uge - ue).norm() / ue.norm()) <= 0.3
    assert abs(float(ug.norm() / ue.norm()) - 1.0) <= 0.02
