# Overlay of the reference's `fourm.models` package: modules defined here (fm, fm_utils, encoder_embeddings,
# decoder_embeddings) shadow the reference's; everything else (generate.py, fm_vit.py, lora_utils.py, ...) keeps resolving
# to the reference tree when it is on sys.path (overlay mechanism v9).
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
