"""B200-native batched bin-a-cube physics + VectorEnv (drop-in for gym_so100's env step/reset)."""
__version__ = "0.1.0"
