"""recoder_b200 — B200-native implementation of the Recoder mini-batch training hot path."""
__version__ = '0.4.0+b200'
