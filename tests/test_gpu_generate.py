"""GenerationSampler of the overlay (CFG passes batched, KV-cached AR loop, no deepcopy) vs the golden trajectory of the UNMODIFIED
reference's GenerationSampler (tests/golden/make_golden_gen.py: 4M-Tiny, CPU fp32, RGB -> depth (MaskGIT) -> normals (ROAR) ->
caption (AR), all guided).  Every schedule step is replayed from the reference's state before that step (teacher forcing), so one
bf16 near-tie cannot cascade into the following steps:
  * ROAR / AR at temperature 0: the decoded positions are identical and the tokens are the reference's, except where the reference's
    own best and second-best logit are closer than the bf16 noise (the golden stores that gap for every arg-max);
  * MaskGIT at temperature 1 / top-p 0.8: the same host random numbers feed torch.multinomial, so all but a few samples agree."""
import pytest
import torch

from oracle import fourm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gen():
    from b200fm.compat import build_mod7_embeddings, create_model
    from fourm.models.generate import GenerationSampler
    gold = H.load_golden("gen_tiny_golden.pt")
    tiny = H.load_golden("fourm_tiny_golden.pt")
    enc, dec, info = build_mod7_embeddings()
    model = create_model(gold["model"], encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info)
    model.load_state_dict(H.fill_fourm_buffers(H.golden_state_dict(tiny), O.mod7_specs(), 384))
    model = model.cuda().eval()
    model.modality_info = {k: dict(v) for k, v in model.modality_info.items()}
    model.modality_info['caption']['max_tokens'] = gold["ar_max_tokens"]
    sampler = GenerationSampler(model)
    sampler.rng_device = "cpu"
    sample, schedule = H.generation_case(model.modality_info, device="cuda")
    return gold, model, sampler, sample, schedule


def _state(sample, gold_state):
    st = {m: dict(d) for m, d in sample.items()}
    if gold_state is not None:
        for m, d in gold_state.items():
            st[m] = {k: v.cuda() for k, v in d.items()}
            st[m]['decoder_attention_mask'] = torch.zeros_like(st[m]['target_mask'])
    return st


def test_schedule_and_sample_construction_match_reference(gen):
    gold, model, sampler, sample, schedule = gen
    assert len(schedule) == len(gold["schedule"])
    for a, b in zip(schedule, gold["schedule"]):
        assert a['target_domain'] == b['target_domain'] and a['scheme'] == b['scheme'] and a['cfg_cond_domains'] == b['cfg_cond_domains']
        assert (a['num_tokens'] is None and b['num_tokens'] is None) or int(a['num_tokens']) == int(b['num_tokens'])
        assert abs(float(a['temperature']) - float(b['temperature'])) < 1e-12 and abs(float(a['cfg_scale']) - float(b['cfg_scale'])) < 1e-12


def test_image_steps_teacher_forced(gen):
    gold, model, sampler, sample, schedule = gen
    tok = H.StubTextTokenizer()
    for i, info in enumerate(schedule):
        if info['scheme'] == 'autoregressive':
            continue
        tgt = info['target_domain']
        before = _state(sample, gold["states"][i - 1] if i > 0 else None)
        ref_before = before[tgt]['input_mask'].clone()
        out = sampler._one_step({m: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()} for m, d in before.items()},
                                info, i, 0.0, gold["top_p"], tok, gold["seed"], write_all=False)
        ref = gold["states"][i][tgt]
        new_ref = (ref_before.cpu() & ~ref['input_mask'])
        new_got = (ref_before & ~out[tgt]['input_mask']).cpu()
        assert int(new_got.sum()) == int(new_ref.sum()) == int(info['num_tokens']) * new_ref.shape[0]
        if info['scheme'] == 'roar':
            assert torch.equal(new_got, new_ref)                       # the random order comes from the same host generator
            # map every decoded slot to the reference's top-2 logit gap of that decision (slots are in ROAR order)
            dec = {tgt: model.decoder_embeddings[tgt].forward_embed(dict(before[tgt]))}
            pos = sampler._mask_decoder(dec, tgt, 'roar', int(info['num_tokens']), seed=gold["seed"] + i)[4].cpu()
            gap = gold["gaps"][i][0].reshape(pos.shape)
            got_tok = torch.gather(out[tgt]['tensor'].cpu(), 1, pos)
            ref_tok = torch.gather(ref['tensor'], 1, pos)
            bad = got_tok != ref_tok
            print(f"step {i} ROAR {tgt}: {int(bad.sum())}/{bad.numel()} tokens differ; reference top-2 gaps at the differences "
                  f"{[round(x, 4) for x in gap[bad].tolist()]}; median gap {float(gap.median()):.4f}")
            assert not bad.any() or float(gap[bad].max()) < 0.04         # only genuine near-ties of the reference may flip
            assert int(bad.sum()) <= 0.15 * bad.numel()
        else:
            both = new_got & new_ref
            agree = float(both.sum()) / float(new_ref.sum())
            same_tok = float((out[tgt]['tensor'].cpu()[both] == ref['tensor'][both]).float().mean())
            print(f"step {i} MaskGIT {tgt}: {agree:.3f} of the selected positions agree, {same_tok:.3f} of their tokens")
            assert agree >= 0.9 and same_tok >= 0.95


@pytest.mark.parametrize("cached", [True, False])
def test_autoregressive_teacher_forced(gen, cached):
    """AR at temperature 0 with guidance: the golden's arg-max picks are forced as the prefix, our arg-max at every position must equal
    them unless the reference's top-2 gap is a near-tie.  cached=True is the KV-cached decoder, False the full re-computation."""
    gold, model, sampler, sample, schedule = gen
    i = len(schedule) - 1
    info = schedule[i]
    assert info['scheme'] == 'autoregressive'
    picks = [p.cuda() for p in gold["picks"][i]]
    gaps = torch.stack(gold["gaps"][i])                                # [T, B]
    mine = []
    real = torch.argmax

    def forced(x, *a, **k):
        if x.dim() == 2 and x.shape[-1] > 1000 and len(mine) < len(picks):
            mine.append(real(x, dim=-1).clone())
            return picks[len(mine) - 1].reshape(-1, 1) if k.get("keepdim") else picks[len(mine) - 1]
        return real(x, *a, **k)
    sampler.kv_cache = cached
    before = _state(sample, gold["states"][i - 1])
    torch.argmax = forced
    try:
        out = sampler._one_step(before, info, i, 0.0, gold["top_p"], H.StubTextTokenizer(), gold["seed"], write_all=False)
    finally:
        torch.argmax = real
        sampler.kv_cache = True
    assert len(mine) == len(picks)
    got, want = torch.stack(mine).cpu(), torch.stack([p.cpu() for p in picks])
    bad = got != want
    print(f"AR cached={cached}: {int(bad.sum())}/{bad.numel()} arg-max decisions differ; gaps at the differences {gaps[bad].tolist()}; "
          f"median gap {float(gaps.median()):.4f}")
    assert int(bad.sum()) <= 4 and (not bad.any() or float(gaps[bad].max()) < 0.04)
    # with the forced prefix the merged result is the reference's
    assert torch.equal(out['caption']['tensor'].cpu(), gold["states"][i]['caption']['tensor'])


def test_generate_end_to_end_runs_and_is_deterministic(gen):
    gold, model, sampler, sample, schedule = gen
    tok = H.StubTextTokenizer()
    a = sampler.generate(sample, schedule, top_k=0.0, top_p=gold["top_p"], text_tokenizer=tok, seed=gold["seed"])
    b = sampler.generate(sample, schedule, top_k=0.0, top_p=gold["top_p"], text_tokenizer=tok, seed=gold["seed"])
    for m in H.GEN_TARGETS:
        assert torch.equal(a[m]['tensor'], b[m]['tensor'])
        assert not bool(a[m]['input_mask'].any()) or m == 'caption'
    assert sample['tok_depth@224']['input_mask'].all()                 # the caller's dict is untouched
    # unbatched CFG (two passes like the reference) gives the same tokens as the batched pass on the deterministic steps
    sampler.batch_cfg = False
    try:
        c = sampler.generate(sample, schedule, top_k=0.0, top_p=gold["top_p"], text_tokenizer=tok, seed=gold["seed"])
    finally:
        sampler.batch_cfg = True
    agree = float((a['tok_normal@224']['tensor'] == c['tok_normal@224']['tensor']).float().mean())
    assert agree >= 0.9


def test_fused_nucleus_sampling_matches_reference_rule():
    """b200fm_sample_top_p vs the torch restatement of the reference's rule (generate.py:332-371): (1) the sampled token always lies in
    the reference's nucleus, (2) with u swept over [0, 1) the empirical distribution equals softmax(filtered / T) (total variation),
    (3) top_p = 0 keeps everything, a peaked row returns its arg-max."""
    from b200fm import ops
    from fourm.models.generate import GenerationSampler
    g = torch.Generator(device="cuda").manual_seed(0)
    flt = GenerationSampler.top_k_top_p_filtering
    for V, top_p, T in ((1000, 0.8, 1.0), (30000, 0.8, 0.7), (133, 0.5, 1.3), (8192, 0.0, 1.0), (50000, 0.95, 1.0)):
        logits = torch.randn(4, V, device="cuda", generator=g) * 2.5
        filt = flt(None, logits.clone(), 0.0, top_p)
        ref_p = torch.softmax(filt / T, dim=-1)
        n = 4000
        u = (torch.arange(n, device="cuda", dtype=torch.float32) + 0.5) / n            # a stratified sweep of the unit interval
        for r in range(2):
            row = logits[r:r + 1].expand(n, V).contiguous()
            tok = ops.sample_top_p(row, top_p, T, u)
            assert bool((ref_p[r][tok] > 0).all()), "sampled a token outside the reference's nucleus"
            emp = torch.bincount(tok, minlength=V).float() / n
            tv = 0.5 * float((emp - ref_p[r]).abs().sum())
            assert tv <= 0.03 + 1.5 * (float((ref_p[r] > 0).sum()) / n) ** 0.5 * 0.5, (V, top_p, tv)
    peaked = torch.full((2, 500), -5.0, device="cuda")
    peaked[0, 17] = 30.0
    peaked[1, 499] = 30.0
    tok = ops.sample_top_p(peaked, 0.8, 1.0, torch.tensor([0.3, 0.999], device="cuda"))
    assert tok.tolist() == [17, 499]
