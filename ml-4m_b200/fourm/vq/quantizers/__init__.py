# Overlay of `fourm.vq.quantizers`: quantize_lucid is B200-native; quantize_memcodes keeps resolving to the reference.
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
