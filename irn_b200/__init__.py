"""irn_b200: B200-native pseudo-label hot path of jiwoon-ahn/irn."""
__version__ = "0.1.0"
