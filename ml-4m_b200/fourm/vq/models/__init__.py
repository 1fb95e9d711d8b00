# Overlay of `fourm.vq.models`: vit_models is B200-native; unet / uvit / controlnet / mlp_models / lm_models keep resolving to
# the reference tree when it is on sys.path.
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
