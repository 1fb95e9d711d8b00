"""KV-cached decoder pass for autoregressive generation.

The reference's AR loop (fourm/models/generate.py:886-913 / :984-1022) re-runs the WHOLE decoder over the whole prefix for every
new token -- O(L^2) decoder work, and per step it re-projects the encoder context through every layer's `context_norm` + `kv`
(`fm_utils.py:364`), which alone is 2 x N x D x 2D FLOPs per layer per token.  `CachedDecoder` does what the arithmetic allows:

  * cross-attention keys / values of the (fixed) context: computed ONCE per layer for the whole AR call;
  * self-attention keys / values: one row per generated token appended to a per-layer cache; each step runs the decoder on the
    single new position (queries of length 1 against the cache, positions beyond the current one masked).

Token t's hidden state is a function of tokens <= t only (causal mask), so the cached pass computes exactly what the reference's
full re-computation yields for its last position.  Evaluation only (no autograd); plain modules only (`supported(model)`):
anything else (LoRA-wrapped linears, drop-path in training mode, ...) makes the caller fall back to the reference algorithm.
"""
import torch
import torch.nn as nn

from . import functional as BF
from . import ops


def _plain(m):
    return type(m) is nn.Linear


def _norm_ok(n):
    from fourm.models.fm_utils import LayerNorm
    return isinstance(n, LayerNorm) or (type(n) is nn.LayerNorm and n.elementwise_affine)


def supported(model):
    from fourm.models.fm_utils import Attention, CrossAttention, GatedMlp, Mlp, NormAttention, NormCrossAttention
    for blk in model.decoder:
        sa, xa, mlp = blk.self_attn, blk.cross_attn, blk.mlp
        if type(sa) not in (Attention, NormAttention) or type(xa) not in (CrossAttention, NormCrossAttention):
            return False
        if not (_plain(sa.qkv) and _plain(sa.proj) and _plain(xa.q) and _plain(xa.kv) and _plain(xa.proj)):
            return False
        if sa.allow_zero_attn or xa.allow_zero_attn or sa.qkv.weight.shape[1] // sa.num_heads != 64:
            return False
        if type(mlp) is GatedMlp:
            if not (_plain(mlp.fc1) and _plain(mlp.fc2) and _plain(mlp.fc3) and type(mlp.act) is nn.SiLU):
                return False
        elif type(mlp) is Mlp:
            if not (_plain(mlp.fc1) and _plain(mlp.fc2) and type(mlp.act) is nn.GELU and getattr(mlp.act, "approximate", "none") == "none"):
                return False
        else:
            return False
        if not all(_norm_ok(n) for n in (blk.norm1, blk.query_norm, blk.context_norm, blk.norm2)):
            return False
    return _norm_ok(model.decoder_norm)


def _ln(norm, x2):
    """fp32 [R, D] -> bf16 [R, D]"""
    return ops.layernorm_fwd(x2, norm.weight, norm.bias, norm.eps, out_bf16=True, save_stats=False)[0]


def _lin(lin, x_bf16):
    return ops.gemm(x_bf16, BF.weight_bf16(lin.weight), epilogue=ops.EPI_BF16, bias=lin.bias, n_out=lin.weight.shape[0])


DECODE_ATTN_MAX_KEYS = 1024     # one-query kernel up to here (B x H small CTAs walk the keys); longer contexts: the tile kernels


def _attend1(q, k, v, B, H, Nk, mask, scale):
    """Attention of ONE query row per sequence over Nk cached keys."""
    if Nk <= DECODE_ATTN_MAX_KEYS:
        return ops.attention_decode(q, k, v, B, H, Nk, mask, scale)
    return ops.attention_fwd(q, k, v, B, H, 1, Nk, mask, scale)[0]


def _lin_resid(lin, x_bf16, resid):
    """resid (fp32) + bf16(lin(x)): the residual add as the GEMM's epilogue (no cast / add kernels between the linears, so the whole
    decode step stays one chain of programmatically dependent launches)."""
    return ops.gemm(x_bf16, BF.weight_bf16(lin.weight), epilogue=ops.EPI_RESID, bias=lin.bias, resid=resid, n_out=lin.weight.shape[0])


def _gated_mlp(mlp, h, resid=None):
    """fc2(silu(fc1 h) * fc3 h) on bf16 rows, with the zero-padded operands the training path uses for widths that are not
    multiples of 8 (4M-L 2730, 4M-XL 5461).  resid: fp32 rows to add the result to (returns fp32)."""
    w1, w3, w2 = mlp.fc1.weight, mlp.fc3.weight, mlp.fc2.weight
    H = w1.shape[0]
    Hp = (H + 7) // 8 * 8
    bias13 = None
    if mlp.fc1.bias is not None:
        bias13 = torch.zeros(2 * Hp, device=h.device, dtype=torch.float32)
        bias13[:H] = mlp.fc1.bias
        bias13[Hp:Hp + H] = mlp.fc3.bias
    _, g = ops.gemm(h, BF.weight_bf16(w1, w3), epilogue=ops.EPI_SWIGLU, bias=bias13)
    w2b = BF.weight_bf16(w2) if Hp == H else BF.weight_bf16_padk(w2, Hp)
    if resid is not None:
        return ops.gemm(g, w2b, epilogue=ops.EPI_RESID, bias=mlp.fc2.bias, resid=resid, n_out=w2.shape[0])
    return ops.gemm(g, w2b, epilogue=ops.EPI_BF16, bias=mlp.fc2.bias, n_out=w2.shape[0])


class CachedDecoder:
    """context fp32 [B, N, D], encoder_mask bool [B, 1, N] (True = masked) as `FourM.forward_decoder` takes them; max_len = the
    longest sequence that will be decoded.

    One decode step is ~12 small launches per layer (300 for 24 layers) on a few rows: the HOST is what a step waits for.  The step
    is therefore captured in a CUDA graph on its second call and replayed afterwards (the position is a device scalar, the K/V
    caches, masks and the input row are static buffers); B200FM_GEN_GRAPH=0 keeps Python-issued launches."""

    def __init__(self, model, context, encoder_mask, max_len, use_graph=None):
        import os
        self.model = model
        B, N, D = context.shape
        self.B, self.N, self.D, self.L = B, N, D, int(max_len)
        dev = context.device
        self.enc_mask = encoder_mask.contiguous()
        c2 = context.reshape(B * N, D).float().contiguous()
        self.kv_ctx = []
        for blk in model.decoder:
            xa = blk.cross_attn
            kv = _lin(xa.kv, _ln(blk.context_norm, c2))                          # [B*N, 2D] bf16, once per AR call
            if hasattr(xa, "k_norm"):
                kv = torch.cat([BF.head_norm(kv[:, :D], xa.num_heads, xa.k_norm), kv[:, D:]], dim=1).contiguous()
            self.kv_ctx.append(kv)
        self.kv_self = [torch.zeros(B, self.L, 2 * D, device=dev, dtype=torch.bfloat16) for _ in model.decoder]
        self.sa_mask = torch.ones(B, 1, self.L, dtype=torch.bool, device=dev)     # True = not yet generated
        self.pos = 0
        self.pos_dev = torch.zeros(1, dtype=torch.int64, device=dev)              # the same position as device data (graph replays)
        self.use_graph = (os.environ.get("B200FM_GEN_GRAPH", "1") != "0") if use_graph is None else use_graph
        self.graph = None
        self.y_static = torch.zeros(B, D, device=dev, dtype=torch.float32)
        self.h_static = None

    def _step_impl(self, y):
        B, D, L = self.B, self.D, self.L
        self.sa_mask.index_fill_(2, self.pos_dev, False)
        x = y.float().contiguous()
        for li, blk in enumerate(self.model.decoder):
            sa, xa = blk.self_attn, blk.cross_attn
            H = sa.num_heads
            qkv = _lin(sa.qkv, _ln(blk.norm1, x))                                 # [B, 3D]
            q, k = qkv[:, :D], qkv[:, D:2 * D]
            if hasattr(sa, "q_norm"):
                q, k = BF.head_norm(q, H, sa.q_norm), BF.head_norm(k, H, sa.k_norm)
            cache = self.kv_self[li]
            if hasattr(sa, "q_norm"):                                             # normalised keys are their own tensor
                ops.kv_append(k, cache, self.pos_dev, 0)
                ops.kv_append(qkv[:, 2 * D:], cache, self.pos_dev, D)
            else:
                ops.kv_append(qkv[:, D:], cache, self.pos_dev, 0)                 # k | v are adjacent in the qkv row
            c2 = cache.view(B * L, 2 * D)
            o = _attend1(q, c2[:, :D], c2[:, D:], B, H, L, self.sa_mask, sa.scale)
            x = _lin_resid(sa.proj, o, x)                                         # x + proj(o): the add is the GEMM's epilogue
            q = _lin(xa.q, _ln(blk.query_norm, x))
            if hasattr(xa, "q_norm"):
                q = BF.head_norm(q, H, xa.q_norm)
            kv = self.kv_ctx[li]
            o = _attend1(q, kv[:, :D], kv[:, D:], B, H, self.N, self.enc_mask, xa.scale)
            x = _lin_resid(xa.proj, o, x)
            h = _ln(blk.norm2, x)
            mlp = blk.mlp
            if hasattr(mlp, "fc3"):
                x = _gated_mlp(mlp, h, resid=x)
            else:
                x = _lin_resid(mlp.fc2, BF.MlpActFn.apply(h, mlp.fc1.weight, mlp.fc1.bias, "gelu"), x)
        self.pos_dev.add_(1)
        return _ln(self.model.decoder_norm, x)

    @torch.no_grad()
    def step(self, y):
        """y fp32 [B, D]: embedding of the token at position self.pos (token_emb + positional / modality embedding).
        Returns decoder_norm(hidden) bf16 [B, D] for that position (valid until the next call) and advances the cache."""
        if self.pos >= self.L:
            raise IndexError(f"CachedDecoder: position {self.pos} beyond the cache length {self.L}")
        self.pos += 1
        if not self.use_graph:
            return self._step_impl(y)
        if self.graph is None and self.pos == 1:
            return self._step_impl(y)                       # first call: eager (warms caches / allocator)
        self.y_static.copy_(y)
        if self.graph is None:
            torch.cuda.synchronize(y.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self.h_static = self._step_impl(self.y_static)
            self.graph = g
        self.graph.replay()
        return self.h_static
