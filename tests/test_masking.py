"""On-GPU batch assembly: the masking oracle vs the reference's goldens (CPU), the device kernel vs both (GPU), uint8 RGB path."""
import numpy as np
import pytest
import torch

from oracle import masking_oracle as MO
from tests import helpers as H


def test_oracle_matches_reference_image_mask():
    gold = H.load_golden("masking_golden.pt")
    for c in gold["cases"]:
        im, tm, dam = MO.image_mask(c["noise"].numpy(), c["n_in"], c["n_tgt"])
        assert np.array_equal(im, c["input_mask"].numpy()) and np.array_equal(tm, c["target_mask"].numpy())
        assert np.array_equal(dam, c["dam"].numpy())


@pytest.mark.gpu
def test_device_masks_bit_exact_vs_reference_and_oracle():
    from b200fm import ops
    gold = H.load_golden("masking_golden.pt")
    for c in gold["cases"]:
        nb = torch.tensor([c["n_in"]], dtype=torch.int32, device="cuda")
        tb = None if c["n_tgt"] is None else torch.tensor([c["n_tgt"]], dtype=torch.int32, device="cuda")
        im, tm, dam = ops.mask_images(c["noise"].cuda()[None].contiguous(), nb, tb)
        assert torch.equal(im[0].cpu(), c["input_mask"]) and torch.equal(tm[0].cpu(), c["target_mask"]) and torch.equal(dam[0].cpu(), c["dam"].int())
    # a whole batch at once, ragged budgets, against the oracle
    g = torch.Generator().manual_seed(1)
    B, n, L = 64, 6, 196
    noise = torch.rand(B, n, L, generator=g)
    n_in = torch.randint(0, 60, (B, n), generator=g, dtype=torch.int32)
    n_tg = torch.randint(0, 150, (B, n), generator=g, dtype=torch.int32)
    im, tm, dam = ops.mask_images(noise.cuda(), n_in.cuda(), n_tg.cuda())
    for b in range(0, B, 7):
        for m in range(n):
            a, t, d = MO.image_mask(noise[b, m].numpy(), int(n_in[b, m]), int(n_tg[b, m]))
            assert np.array_equal(im[b, m].cpu().numpy(), a) and np.array_equal(tm[b, m].cpu().numpy(), t) and np.array_equal(dam[b, m].cpu().numpy(), d)
    assert int((~im).sum()) == int(n_in.sum())


@pytest.mark.gpu
def test_budgets_and_device_masking_feed_the_model():
    """DeviceImageMasking output is a valid mod_dict for FourM.forward; budgets follow the reference's recipe (sum == num_tokens unless
    clamped, never above max_tokens); uint8 RGB gives the same rows as the fp32 RGB the loader would have produced."""
    import random
    from b200fm import masking
    from b200fm.compat import build_mod7_embeddings, create_model
    from oracle import fourm_oracle as O
    torch.manual_seed(0)
    enc, dec, info = build_mod7_embeddings()
    model = create_model("fm_tiny_6e_6d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info).cuda()
    B = 4
    img_mods = ['tok_rgb@224', 'tok_depth@224', 'tok_normal@224', 'tok_semseg@224', 'tok_clip@224']
    alphas = torch.full((B, 5), 0.5)
    gen = torch.Generator(device="cuda").manual_seed(3)
    nb = masking.sample_budgets(alphas, 100, [196] * 5, generator=gen)
    tb = masking.sample_budgets(alphas, 110, [196] * 5, generator=gen)
    assert nb.shape == (B, 5) and bool((nb.sum(1) == 100).all()) and bool((nb <= 196).all()) and bool((nb >= 0).all())
    tb = torch.minimum(tb, 196 - nb)
    base = O.synthetic_mod7_batch(B, seed=9)
    toks = {m: base[m]["tensor"].cuda() for m in img_mods}
    masked = masking.DeviceImageMasking(img_mods, 196)(toks, nb, tb, generator=gen)
    batch = {m: {k: v.cuda() for k, v in d.items()} for m, d in base.items()}
    batch.update(masked)
    for i, m in enumerate(img_mods):
        assert int((~masked[m]["input_mask"]).sum()) == int(nb[:, i].sum()) and int((~masked[m]["target_mask"]).sum()) == int(tb[:, i].sum())
        assert not bool(((~masked[m]["input_mask"]) & (~masked[m]["target_mask"])).any())              # inputs and targets are disjoint
    # uint8 RGB: same patch rows as the float image the loader would have produced
    gcpu = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (B, 3, 224, 224), generator=gcpu, dtype=torch.uint8)
    f32 = torch.from_numpy(MO.normalise_rgb_u8(u8.numpy()))
    emb = model.encoder_embeddings["rgb@224"]
    with torch.no_grad():
        a = emb.project_patches(u8.cuda())
        b = emb.project_patches(f32.cuda())
    assert torch.equal(a, b)
    batch["rgb@224"]["tensor"] = u8.cuda()
    random.seed(0)
    loss, _ = model(batch, num_encoder_tokens=128, num_decoder_tokens=128)
    loss.backward()
    assert torch.isfinite(loss)
