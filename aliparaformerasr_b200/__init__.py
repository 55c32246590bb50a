"""B200-native drop-in for the AliParaformerAsr offline hot path."""
__version__ = "0.1.0"
