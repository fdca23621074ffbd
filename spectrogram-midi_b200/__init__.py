"""B200-native audio-analysis front end for Aegis Engine."""
__version__ = "0.1.0"
