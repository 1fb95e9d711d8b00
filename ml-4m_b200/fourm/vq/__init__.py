"""Overlay of `fourm.vq` (apple/ml-4m): the tokenization path (VQ encoder + quantizer) and the VQ-VAE training model (ViT decoder,
codebook EMA) are B200-native; the diffusion decoders (DiVAE / VQControlNet, schedulers, UNets) keep resolving to the reference
tree when it is on sys.path."""
import os
import pkgutil

import torch

__path__ = pkgutil.extend_path(__path__, __name__)

from .vqvae import VQ, VQVAE  # noqa: E402


def __getattr__(name):
    """DiVAE / VQControlNet live in the reference's vqvae.py (they need `diffusers`); import them lazily from there."""
    if name in ("DiVAE", "VQControlNet"):
        import importlib.util
        for p in __path__[1:]:
            f = os.path.join(p, "vqvae.py")
            if os.path.exists(f):
                spec = importlib.util.spec_from_file_location("fourm.vq._reference_vqvae", f, submodule_search_locations=None)
                mod = importlib.util.module_from_spec(spec)
                mod.__package__ = "fourm.vq"
                spec.loader.exec_module(mod)
                return getattr(mod, name)
        raise ImportError(f"{name} is provided by the reference tree (not on sys.path); the B200 overlay ships VQ only")
    raise AttributeError(name)


def get_image_tokenizer(tokenizer_id: str, tokenizers_root: str = './tokenizer_ckpts', encoder_only: bool = False,
                        device: str = 'cuda', verbose: bool = True, return_None_on_fail: bool = False):
    """Load a pretrained image tokenizer (reference fourm/vq/__init__.py:8-79): `{root}/{id}.pth` holding `args` (Namespace)
    and `model` (state_dict).  Returns (model.eval(), args)."""
    path = os.path.join(tokenizers_root, f'{tokenizer_id}.pth')
    if return_None_on_fail and not os.path.exists(path):
        return None
    if verbose:
        print(f'Loading tokenizer {tokenizer_id} ... ', end='')
    ckpt = torch.load(path, map_location='cpu', weights_only=False)
    args = ckpt['args']
    # legacy argument names (reference :38-59)
    if hasattr(args, 'quantizer_type') and not hasattr(args, 'quant_type'):
        args.quant_type = args.quantizer_type
    if hasattr(args, 'encoder_type') and not hasattr(args, 'enc_type'):
        args.enc_type = args.encoder_type
    if hasattr(args, 'input_size') and not hasattr(args, 'image_size'):
        args.image_size = args.input_size
    kw = {k: v for k, v in vars(args).items()}
    kw['sync_codebook'] = False
    kw['ckpt_path'] = None
    if encoder_only:
        model = VQ(**kw)
        sd = {k: v for k, v in ckpt['model'].items() if not (k.startswith('decoder') or k.startswith('post_quant'))}
        msg = model.load_state_dict(sd, strict=False)
    else:
        model_type = getattr(args, 'model_type', 'VQVAE')
        model = (VQVAE if model_type == 'VQVAE' else __getattr__(model_type))(**kw)
        msg = model.load_state_dict(ckpt['model'], strict=False)
    if verbose:
        print(msg)
    return model.to(device).eval(), args
