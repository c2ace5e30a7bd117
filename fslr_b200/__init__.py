"""B200-native read-clustering step of fslr."""
__version__ = "0.1.0"
