"""Glue to the parts of the reference that stay as they are (`fourm.utils` model registry, `fourm.data.modality_info`,
huggingface_hub mixin).  When the reference tree is importable these names ARE the reference's objects, so
`fourm.utils.create_model(...)` finds the B200 models; when it is not (e.g. on a bare GPU box) small local equivalents are
used so the overlay is self-contained."""
import hashlib
import sys
from functools import partial

try:
    from huggingface_hub import PyTorchModelHubMixin
except Exception:      # pragma: no cover
    class PyTorchModelHubMixin:      # minimal stand-in: keeps the class hierarchy importable
        pass


def generate_uint15_hash(seed_str: str) -> int:
    """Modality id = sha256(name) mod 2**15 (reference fourm/utils/misc.py:39-41)."""
    return int(hashlib.sha256(seed_str.encode('utf-8')).hexdigest(), 16) % (2 ** 15)


_local_entrypoints = {}


def _local_register_model(fn):
    mod = sys.modules[fn.__module__]
    if hasattr(mod, '__all__'):
        mod.__all__.append(fn.__name__)
    else:
        mod.__all__ = [fn.__name__]
    _local_entrypoints[fn.__name__] = fn
    return fn


def _try_reference_registry():
    before = set(sys.modules)
    try:
        from fourm.utils.timm.registry import register_model as ref_register      # noqa: WPS433
        return ref_register
    except Exception:
        for k in set(sys.modules) - before:        # drop half-imported reference modules
            if k.startswith("fourm.utils"):
                sys.modules.pop(k, None)
        return None


_ref_register = _try_reference_registry()


def register_model(fn):
    """Register in the reference's timm-style registry when available (so `fourm.utils.create_model` works), and always
    in the local one."""
    _local_register_model(fn)
    if _ref_register is not None:
        mod = sys.modules[fn.__module__]
        names = list(mod.__all__)
        _ref_register(fn)
        mod.__all__[:] = names
    return fn


def create_model(model_name, **kwargs):
    """Local equivalent of fourm.utils.create_model (reference utils/timm/model_builder.py:27-74)."""
    if model_name not in _local_entrypoints:
        import fourm.models.fm  # noqa: F401  (registers the presets)
    return _local_entrypoints[model_name](**kwargs)


class _LazyModalityInfo(dict):
    """MODALITY_INFO: the reference's registry when importable, else the locally restated entries below."""
    _loaded = False

    def _load(self):
        if self._loaded:
            return
        self._loaded = True
        try:
            from fourm.data.modality_info import MODALITY_INFO as ref      # noqa: WPS433
            self.update(ref)
        except Exception:
            self.update(local_modality_info())

    def __getitem__(self, k):
        self._load()
        return super().__getitem__(k)

    def __contains__(self, k):
        self._load()
        return super().__contains__(k)

    def keys(self):
        self._load()
        return super().keys()

    def items(self):
        self._load()
        return super().items()

    def get(self, k, default=None):
        self._load()
        return super().get(k, default)


def local_modality_info():
    """Restatement of the hot-path fields of fourm/data/modality_info.py:32-383 (vocab sizes, max lengths, patch sizes, types, ids,
    embedding factories) for the 4M-7 and 4M-21 modalities -- used only where the reference tree is not importable."""
    from fourm.models.decoder_embeddings import ImageTokenDecoderEmbedding, SequenceDecoderEmbedding
    from fourm.models.encoder_embeddings import (ImageEncoderEmbedding, ImageTokenEncoderEmbedding, SequenceEmbEncoderEmbedding,
                                                 SequenceEncoderEmbedding)
    info = {}
    for res in (224, 448):
        info[f'rgb@{res}'] = dict(input_size=res, patch_size=16, encoder_embedding=partial(ImageEncoderEmbedding, num_channels=3),
                                  decoder_embedding=None, min_tokens=0, max_tokens=None, type='img', num_channels=3, path='rgb')
    # tokenised image-like modalities: (name, vocab, input size, patch size)
    tok_img = [('tok_rgb@224', 16384, 224, 16), ('tok_depth@224', 8192, 224, 16), ('tok_normal@224', 8192, 224, 16),
               ('tok_semseg@224', 4096, 224, 16), ('tok_clip@224', 8192, 224, 16), ('tok_canny_edge@224', 8192, 224, 16),
               ('tok_sam_edge@224', 8192, 224, 16), ('tok_dinov2@224', 8192, 224, 14), ('tok_imagebind@224', 8192, 224, 14),
               ('tok_rgb@448', 16384, 448, 16), ('tok_depth@448', 8192, 448, 16), ('tok_normal@448', 8192, 448, 16),
               ('tok_semseg@448', 4096, 448, 16), ('tok_clip@448', 8192, 448, 16)]
    for name, vocab, size, patch in tok_img:
        info[name] = dict(input_size=size, patch_size=patch, vocab_size=vocab,
                          encoder_embedding=partial(ImageTokenEncoderEmbedding, vocab_size=vocab),
                          decoder_embedding=partial(ImageTokenDecoderEmbedding, vocab_size=vocab),
                          min_tokens=0, max_tokens=None, type='img', pretokenized=True)
    for name in ('tok_dinov2_global', 'tok_imagebind_global'):        # 16 global tokens, LEARNED positional table (modality_info.py:277-303)
        info[name] = dict(vocab_size=8192, patch_size=56,
                          encoder_embedding=partial(ImageTokenEncoderEmbedding, vocab_size=8192, sincos_pos_emb=False),
                          decoder_embedding=partial(ImageTokenDecoderEmbedding, vocab_size=8192, sincos_pos_emb=False),
                          min_tokens=0, max_tokens=16, type='img', pretokenized=True)
    # token sequences sharing the 30k WordPiece vocabulary: (name, embedding max_length, max_tokens)
    for name, max_len, max_tok in (('caption', 256, 256), ('det', 256, 256), ('metadata', 40, 40), ('human_poses', 263, 275),
                                   ('color_palette', 23, 23), ('sam_instance', 290, 290)):
        info[name] = dict(vocab_size=30_000,
                          encoder_embedding=partial(SequenceEncoderEmbedding, vocab_size=30_000, max_length=max_len, padding_idx=0),
                          decoder_embedding=partial(SequenceDecoderEmbedding, vocab_size=30_000, max_length=max_len, padding_idx=0),
                          min_tokens=0, max_tokens=max_tok, type='seq')
    info['t5_caption'] = dict(encoder_embedding=partial(SequenceEmbEncoderEmbedding, max_length=77, padding_idx=0), decoder_embedding=None,
                              min_tokens=0, max_tokens=77, type='seq_emb')
    for name, d in info.items():
        d['id'] = generate_uint15_hash(name)
    return info


MOD21_IN = ('caption-t5_caption-det-metadata-rgb@224-tok_rgb@224-tok_normal@224-tok_depth@224-tok_semseg@224-tok_clip@224-human_poses-'
            'tok_dinov2@224-tok_dinov2_global-tok_imagebind@224-tok_imagebind_global-tok_sam_edge@224-tok_canny_edge@224-color_palette-'
            'sam_instance').split('-')      # cfgs/default/4m/data/cc12m+coyo+c4/main/mix_mod21_all2allmix_rgb2all_capT5bias_C4.yaml:7-8
MOD21_OUT = [m for m in MOD21_IN if m not in ('t5_caption', 'rgb@224')]


def build_embeddings(domains_in, domains_out, image_size=224, patch_size=16):
    """Embedding dicts + modality_info for arbitrary domain lists, the way run_training_4m.py:359-377 / run_generation.py build them:
    image-like modalities get (patch_size, image_size) from their MODALITY_INFO entry when it names them."""
    info = MODALITY_INFO
    enc, dec = {}, {}
    for side, doms, out in (('encoder_embedding', domains_in, enc), ('decoder_embedding', domains_out, dec)):
        for mod in doms:
            e = info[mod].get(side)
            if e is None:
                continue
            if info[mod]['type'] == 'img':
                out[mod] = e(patch_size=info[mod].get('patch_size', patch_size), image_size=info[mod].get('input_size', image_size))
            else:
                out[mod] = e()
    minfo = {m: info[m] for m in dict.fromkeys(list(domains_in) + list(domains_out))}
    return enc, dec, minfo


MODALITY_INFO = _LazyModalityInfo()


def build_mod7_embeddings(image_size=224, patch_size=16, domains_in=None, domains_out=None):
    """The embedding dicts `run_training_4m.py:359-377` builds for the mod-7 all-to-all mixture."""
    info = MODALITY_INFO
    all7 = ['rgb@224', 'tok_rgb@224', 'tok_depth@224', 'tok_normal@224', 'tok_semseg@224', 'tok_clip@224', 'caption', 'det']
    domains_in = domains_in or all7
    domains_out = domains_out or all7[1:]
    enc, dec = {}, {}
    for mod in domains_in:
        e = info[mod].get('encoder_embedding')
        if e is not None:
            enc[mod] = e(patch_size=patch_size, image_size=image_size) if info[mod]['type'] == 'img' else e()
    for mod in domains_out:
        e = info[mod].get('decoder_embedding')
        if e is not None:
            dec[mod] = e(patch_size=patch_size, image_size=image_size) if info[mod]['type'] == 'img' else e()
    minfo = {m: info[m] for m in dict.fromkeys(list(domains_in) + list(domains_out))}
    return enc, dec, minfo
