"""bench.py --workload gen : BASELINE.json configs[3] -- 4M-21 XL generation latency (Demo4MSampler RGB -> all, 224x224, single GPU).

Model: `fm_xlarge_24e_24d_swiglu_nobias` with the 4M-21 domain lists (cfgs/default/4m/data/cc12m+coyo+c4/main/mix_mod21_...yaml:7-8),
random-init weights (no checkpoints without network), B = 1, input rgb@224 ~ N(0, 1).  Schedule = the reference demo's RGB->all
defaults (fourm/demo_4M_sampler.py:29-78 DEFAULTS_RGB2X in DEFAULT_ORDER, cfg_grow_conditioning=True, top_p 0.8, top_k 0): eight image
modalities by one guided ROAR step each, then caption / det / human_poses / sam_instance / color_palette / metadata autoregressively.
With random weights the AR modalities never emit EOS, so every sequence runs to max_tokens (the worst case; stated in `config`).
A "step" of this workload is one full `GenerationSampler.generate` call; the metric is its latency."""
import json
import os
import time

import torch

DEFAULT_ORDER = ['tok_clip@224', 'tok_dinov2@224', 'tok_imagebind@224', 'tok_depth@224', 'tok_normal@224', 'tok_semseg@224',
                 'tok_canny_edge@224', 'tok_sam_edge@224', 'tok_rgb@224', 'caption', 'det', 'human_poses', 'sam_instance', 'color_palette',
                 'metadata']      # demo_4M_sampler.py:27-31


def rgb2x_defaults(domain):
    """DEFAULTS_RGB2X (demo_4M_sampler.py:38-78) -> (tokens, scheme, steps, token schedule, temp, temp schedule, cfg scale, cfg schedule)."""
    if domain in ('tok_dinov2@224', 'tok_imagebind@224'):
        return 256, 'roar', 1, 'linear', 0.01, 'constant', 2.0, 'constant'
    if domain.startswith('tok_'):
        return 196, 'roar', 1, 'linear', 0.01, 'constant', 2.0, 'constant'
    table = {'caption': (256, 0.3), 'det': (256, 0.3), 'human_poses': (275, 0.1), 'sam_instance': (256, 0.01), 'color_palette': (23, 0.1),
             'metadata': (40, 0.1)}
    n, temp = table[domain]
    return n, 'autoregressive', None, None, temp, 'constant', 1.0, 'constant'


class _Tokenizer:
    """The special-token ids the generation code needs from the reference's WordPiece tokenizer ([PAD] 0, [EOS] 3, sentinels [S_k] = 4 + k)."""

    def __init__(self):
        self.vocab = {"[PAD]": 0, "[UNK]": 1, "[SOS]": 2, "[EOS]": 3, **{f"[S_{i}]": 4 + i for i in range(100)}}

    def get_vocab(self):
        return dict(self.vocab)

    def token_to_id(self, t):
        return self.vocab.get(t)


def build_case(model_name, device, seed=0):
    from b200fm.compat import MOD21_IN, MOD21_OUT, build_embeddings, create_model
    from fourm.models import generate as G
    enc, dec, info = build_embeddings(MOD21_IN, MOD21_OUT)
    torch.manual_seed(seed)
    with torch.device(device):
        model = create_model(model_name, encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info)
    model = model.to(device).eval()
    targets = [d for d in DEFAULT_ORDER if d in model.decoder_embeddings and d != 'tok_rgb@224']     # pixel RGB is the input (:298-301)
    cols = list(zip(*[rgb2x_defaults(d) for d in targets]))
    schedule = G.build_chained_generation_schedules(
        cond_domains=['rgb@224'], target_domains=targets, tokens_per_target=list(cols[0]), autoregression_schemes=list(cols[1]),
        decoding_steps=list(cols[2]), token_decoding_schedules=list(cols[3]), temps=list(cols[4]), temp_schedules=list(cols[5]),
        cfg_scales=list(cols[6]), cfg_schedules=list(cols[7]), cfg_grow_conditioning=True)
    g = torch.Generator().manual_seed(seed)
    rgb_host = torch.randn(1, 3, 224, 224, generator=g).pin_memory()

    def make_sample(rgb_dev):
        s = {'rgb@224': {'tensor': rgb_dev}}
        s = G.init_full_input_modality(s, info, 'rgb@224', device)
        for t, n in zip(targets, cols[0]):
            s = G.init_empty_target_modality(s, info, t, 1, n, device)
        return s
    return model, G.GenerationSampler(model), schedule, targets, rgb_host, make_sample


def run(args, ClockSampler, measured_peaks, cpu_threads, reference_tree):
    from b200fm import lib, ops
    assert torch.cuda.is_available(), "bench.py (B200 arm) needs a GPU; there is no CPU fallback"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lib.load()
    model_name = args.model or "fm_xlarge_24e_24d_swiglu_nobias"
    model, sampler, schedule, targets, rgb_host, make_sample = build_case(model_name, dev)
    tok = _Tokenizer()
    n_params = sum(p.numel() for p in model.parameters())
    ar_tokens = sum(model.modality_info[t]['max_tokens'] for t in targets if model.modality_info[t]['type'] == 'seq')

    def one(e2e):
        rgb = rgb_host.to(dev, non_blocking=True) if e2e else rgb_dev
        out = sampler.generate(make_sample(rgb), schedule, top_k=0.0, top_p=0.8, text_tokenizer=tok, seed=0)
        if e2e:
            return {m: out[m]['tensor'].cpu() for m in targets}            # device -> host read of every generated modality
        return out

    rgb_dev = rgb_host.to(dev)
    for _ in range(max(1, min(args.warmup, 2))):
        one(False)
    torch.cuda.synchronize()
    sampler_clk = ClockSampler(dev.index or 0)
    sampler_clk.start()
    steps = max(1, args.steps)
    c0 = lib.CALLS["n"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = one(False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = lib.CALLS["n"] - c0
    t0 = time.perf_counter()
    for _ in range(steps):
        host_out = one(True)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    clocks = sampler_clk.stop()
    # per-step breakdown: the same schedule, one step at a time, CUDA events around every step
    from fourm.models.generate import _deep_clone
    state = _deep_clone(make_sample(rgb_dev))
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(len(schedule) + 1)]
    marks[0].record()
    with torch.no_grad():
        for i, info in enumerate(schedule):
            state = sampler._one_step(state, info, i, 0.0, 0.8, tok, 0, write_all=False)
            marks[i + 1].record()
    torch.cuda.synchronize()
    breakdown = {info['target_domain']: round(marks[i].elapsed_time(marks[i + 1]), 1) for i, info in enumerate(schedule)}
    # weight-streaming roofline of the dominant kernel of the AR loop: the M = 1 linears of one decode step, launched back to back
    # over all 24 decoder blocks (2.8 GB of distinct bf16 weights per pass >> 126 MB L2), CUDA events around the whole pass
    from b200fm import functional as BF
    x1 = torch.randn(1, model.dim, device=dev).to(torch.bfloat16)
    shapes = []
    for blk in model.decoder:
        sa, xa, mlp = blk.self_attn, blk.cross_attn, blk.mlp
        shapes += [(BF.weight_bf16(sa.qkv.weight), None), (BF.weight_bf16(sa.proj.weight), None), (BF.weight_bf16(xa.q.weight), None),
                   (BF.weight_bf16(xa.proj.weight), None), (BF.weight_bf16(mlp.fc1.weight, mlp.fc3.weight), "swiglu")]
    def gemv_pass():
        for w, kind in shapes:
            if kind == "swiglu":
                ops.gemm(x1, w, epilogue=ops.EPI_SWIGLU)
            else:
                ops.gemm(x1, w, epilogue=ops.EPI_BF16)
    gemv_pass()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(3):
        gemv_pass()
    g1.record()
    torch.cuda.synchronize()
    gemm_ms = g0.elapsed_time(g1) / 3
    gemm_bytes = sum(w.numel() * 2 for w, _ in shapes)
    all_ms = gemm_ms
    peaks = measured_peaks()
    achieved = gemm_bytes / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else 0.0
    n_tok = {m: int(host_out[m].shape[1]) for m in targets}
    line = dict(metric="generation_latency", value=ms / 1e3, unit="s", n_gpus=1, steps=steps, warmup=max(1, min(args.warmup, 2)), ms_per_step=ms,
                higher_is_better=False, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload="4M-21 XL generation, Demo4MSampler RGB->all defaults (8 guided ROAR modalities + 6 autoregressive "
                                     "modalities run to max_tokens: random weights never emit EOS), BASELINE.json configs[3]",
                            model=model_name, params_m=round(n_params / 1e6, 1), batch=1, image_size=224, schedule_steps=len(schedule),
                            ar_tokens=ar_tokens, targets=targets, top_p=0.8, precision="bf16 contractions, fp32 accumulate (reference: fp32/TF32)",
                            l2_policy="weights (5.6 GB bf16) exceed the 126 MB L2: every decode step streams them from HBM"),
                e2e=dict(value=ms_e2e / 1e3, unit="s", h2d_bytes_per_step=rgb_host.numel() * 4,
                         d2h_bytes_per_step=int(sum(v.numel() * v.element_size() for v in host_out.values())), ms_per_step=ms_e2e),
                gpu_launches=launches, tokens_generated=n_tok, ms_per_target=breakdown,
                roofline=dict(bound="hbm", kernel="gemv_kernel<2, *> (csrc/gemv.cu): the M = 1 linears of one K/V-cached decode step over the 24 decoder blocks", achieved=achieved,
                              peak=peaks["hbm"], unit="GB/s", frac=achieved / peaks["hbm"], traffic=None, peak_source=peaks["src"],
                              gemv_ms_per_decode_pass=gemm_ms, weight_bytes_per_pass=gemm_bytes, launches_per_pass=len(shapes)),
                clocks=clocks)
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_sample(cpu_threads())
    print(json.dumps(line))


def cpu_sample(threads):
    """The reference algorithm of one UNGUIDED ROAR pass (196 tokens, forward_enc_dec_roar_batched, generate.py:745-764) + 4 autoregressive
    decoder passes over a 32-token prefix (generate.py:886-901) on the host, through the oracle's functional 4M-XL stacks (fp32).
    A bounded sample: the whole RGB->all schedule on CPU would take hours."""
    from oracle import fourm_oracle as O
    torch.set_num_threads(threads)
    cfg = O.model_cfg(2048, 32, 24, 24)
    D = 2048
    g = torch.Generator().manual_seed(0)
    sd = {}

    def w(*shape):
        return torch.randn(*shape, generator=g) * 0.02
    H = int(2 * 4 * D / 3)
    for i in range(24):
        for side, names in (("encoder", ("norm1", "norm2")), ("decoder", ("norm1", "norm2", "query_norm", "context_norm"))):
            p = f"{side}.{i}."
            for n in names:
                sd[p + n + ".weight"], sd[p + n + ".bias"] = torch.ones(D), torch.zeros(D)
            sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc3.weight"], sd[p + "mlp.fc2.weight"] = w(H, D), w(H, D), w(D, H)
        sd[f"encoder.{i}.attn.qkv.weight"], sd[f"encoder.{i}.attn.proj.weight"] = w(3 * D, D), w(D, D)
        p = f"decoder.{i}."
        sd[p + "self_attn.qkv.weight"], sd[p + "self_attn.proj.weight"] = w(3 * D, D), w(D, D)
        sd[p + "cross_attn.q.weight"], sd[p + "cross_attn.kv.weight"], sd[p + "cross_attn.proj.weight"] = w(D, D), w(2 * D, D), w(D, D)
    for n in ("encoder_norm", "decoder_norm"):
        sd[n + ".weight"], sd[n + ".bias"] = torch.ones(D), torch.zeros(D)
    wc, bc, wl = w(D, D), torch.zeros(D), w(8192, D)
    x = torch.randn(1, 196, D, generator=g)
    enc_mask = torch.zeros(1, 1, 196, dtype=torch.bool)
    t0 = time.perf_counter()
    with torch.no_grad():
        ctx = torch.nn.functional.linear(O.run_encoder(x, sd, cfg, enc_mask), wc, bc)
        y = O.run_decoder(torch.randn(1, 196, D, generator=g), ctx, sd, cfg, enc_mask, None)
        torch.nn.functional.linear(y, wl)
        roar = time.perf_counter() - t0
        t1 = time.perf_counter()
        for cur in range(32, 36):
            causal = torch.ones(cur, cur, dtype=torch.bool).triu(1)[None]
            O.run_decoder(torch.randn(1, cur, D, generator=g), ctx, sd, cfg, enc_mask, causal)
        ar = (time.perf_counter() - t1) / 4
    return dict(value=roar, unit="s per unguided 196-token ROAR pass", cores=threads, kind="port",
                ar_pass_s=ar, sample=f"1 encoder+decoder ROAR pass (196+196 tokens) and 4 AR decoder passes (prefix 32..35) of 4M-XL, fp32, oracle "
                                     f"port, {threads} threads of {os.cpu_count()}; the full schedule = 8 guided ROAR passes (2 each) + ~1100 AR "
                                     "passes with prefixes up to 275 tokens")


def run_reference(args, cpu_threads, reference_tree):
    c = cpu_sample(cpu_threads())
    # the schedule's cost on the CPU, extrapolated from the bounded sample (the AR pass grows with the prefix; 32 tokens is a low estimate)
    est = 16 * c["value"] + 1106 * c["ar_pass_s"]
    line = dict(metric="generation_latency", value=est, unit="s", n_gpus=args.gpus, steps=1, warmup=0, ms_per_step=est * 1e3, higher_is_better=False,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload="4M-21 XL generation RGB->all on host CPU: EXTRAPOLATED from a bounded sample (16 ROAR passes + 1106 AR passes)",
                            parallelism="cpu"),
                cpu_baseline=c, e2e=dict(value=est, unit="s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))
