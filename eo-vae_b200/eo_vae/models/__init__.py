from .model import Decoder, Encoder
from .new_autoencoder import EOFluxVAE

__all__ = ['Encoder', 'Decoder', 'EOFluxVAE']
